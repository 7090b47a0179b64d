// adt_internal.h — shared between the translation units of libadt_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/adt_b200.h"

#define ADT_COPY_STREAMS 3

struct adt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;                        // all user-visible work is ordered here
    cudaStream_t copy_stream[ADT_COPY_STREAMS] = {};      // H2D / kernel / D2H pipelining of *_host calls
    cudaEvent_t copy_done[ADT_COPY_STREAMS] = {};
    cudaEvent_t fence = nullptr;
    uint64_t launches = 0;                                // kernels of this library launched so far
    std::string last_error;
};

int adt_set_error(adt_ctx* ctx, int status, const char* fmt, ...);
int adt_cuda_fail(adt_ctx* ctx, cudaError_t e, const char* what);

#define ADT_CK(ctx, call)                                              \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return adt_cuda_fail(ctx, e__, #call); \
    } while (0)
