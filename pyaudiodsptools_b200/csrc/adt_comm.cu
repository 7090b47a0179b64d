// adt_comm.cu — channel sharding across the GPUs of one box.
//
// The reference has no distributed anything; its unit of independence is "one
// device object per mono channel" (Example2.py:13-22).  Sharding by channel
// therefore needs no data-path collective: the only exchanges are a scatter
// of input rows from a root rank and a gather of output rows (grouped
// ncclSend/ncclRecv over NVLink/NVSwitch) plus a broadcast for small
// parameter blobs.  libnccl is dlopen()ed on first use so that the library
// loads (and every non-comm entry point works) on machines without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <new>

#include "adt_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define LOAD(field, sym)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
    if (!api.field) return;
        LOAD(GetUniqueId, "ncclGetUniqueId")
        LOAD(CommInitRank, "ncclCommInitRank")
        LOAD(CommDestroy, "ncclCommDestroy")
        LOAD(Send, "ncclSend")
        LOAD(Recv, "ncclRecv")
        LOAD(GroupStart, "ncclGroupStart")
        LOAD(GroupEnd, "ncclGroupEnd")
        LOAD(Broadcast, "ncclBroadcast")
        LOAD(AllReduce, "ncclAllReduce")
        LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
        api.ok = true;
    });
    return api;
}

}  // namespace

struct adt_comm {
    adt_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int* d_flag = nullptr;
};

#define NK(ctx, call)                                                                                  \
    do {                                                                                               \
        ncclResult_t r__ = (call);                                                                     \
        if (r__ != ncclSuccess) return adt_set_error(ctx, ADT_ERR_NCCL, "%s: %s", #call, nccl().GetErrorString(r__)); \
    } while (0)

static_assert(sizeof(ncclUniqueId) == ADT_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");

extern "C" int adt_comm_unique_id(unsigned char id[ADT_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) return ADT_ERR_INVALID;
    if (!nccl().ok) return ADT_ERR_NCCL;
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != ncclSuccess) return ADT_ERR_NCCL;
    memcpy(id, &u, sizeof u);
    return ADT_OK;
}

extern "C" int adt_comm_create(adt_ctx* ctx, const unsigned char id[ADT_NCCL_UNIQUE_ID_BYTES], int32_t rank,
                               int32_t world, adt_comm** out) {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return ADT_ERR_INVALID;
    *out = nullptr;
    if (!nccl().ok) return adt_set_error(ctx, ADT_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    adt_comm* c = new (std::nothrow) adt_comm();
    if (!c) return ADT_ERR_NOMEM;
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return adt_set_error(ctx, ADT_ERR_NCCL, "ncclCommInitRank: %s", nccl().GetErrorString(r));
    }
    cudaError_t e = cudaMalloc((void**)&c->d_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->d_flag, 0, sizeof(int));
    if (e != cudaSuccess) {
        nccl().CommDestroy(c->comm);
        delete c;
        return adt_cuda_fail(ctx, e, "adt_comm_create");
    }
    *out = c;
    return ADT_OK;
}

extern "C" int adt_comm_destroy(adt_comm* c) {
    if (!c) return ADT_ERR_INVALID;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    cudaFree(c->d_flag);
    delete c;
    return ADT_OK;
}

extern "C" int adt_comm_scatter_rows(adt_comm* c, const float* full, float* shard, int64_t rows_per_rank, int64_t pitch,
                                     int32_t root) {
    if (!c || !shard || rows_per_rank < 0 || pitch < 0 || root < 0 || root >= c->world) return ADT_ERR_INVALID;
    if (c->rank == root && !full) return ADT_ERR_INVALID;
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const size_t count = (size_t)rows_per_rank * pitch;
    if (count == 0) return ADT_OK;
    NK(ctx, nccl().GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->world; ++r) {
            if (r == root) continue;
            NK(ctx, nccl().Send(full + (size_t)r * count, count, ncclFloat, r, c->comm, ctx->stream));
        }
    } else {
        NK(ctx, nccl().Recv(shard, count, ncclFloat, root, c->comm, ctx->stream));
    }
    NK(ctx, nccl().GroupEnd());
    if (c->rank == root)
        ADT_CK(ctx, cudaMemcpyAsync(shard, full + (size_t)root * count, count * sizeof(float), cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    return ADT_OK;
}

extern "C" int adt_comm_gather_rows(adt_comm* c, const float* shard, float* full, int64_t rows_per_rank, int64_t pitch,
                                    int32_t root) {
    if (!c || !shard || rows_per_rank < 0 || pitch < 0 || root < 0 || root >= c->world) return ADT_ERR_INVALID;
    if (c->rank == root && !full) return ADT_ERR_INVALID;
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const size_t count = (size_t)rows_per_rank * pitch;
    if (count == 0) return ADT_OK;
    NK(ctx, nccl().GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->world; ++r) {
            if (r == root) continue;
            NK(ctx, nccl().Recv(full + (size_t)r * count, count, ncclFloat, r, c->comm, ctx->stream));
        }
    } else {
        NK(ctx, nccl().Send(shard, count, ncclFloat, root, c->comm, ctx->stream));
    }
    NK(ctx, nccl().GroupEnd());
    if (c->rank == root)
        ADT_CK(ctx, cudaMemcpyAsync(full + (size_t)root * count, shard, count * sizeof(float), cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    return ADT_OK;
}

extern "C" int adt_comm_broadcast(adt_comm* c, void* buf, size_t bytes, int32_t root) {
    if (!c || (!buf && bytes) || root < 0 || root >= c->world) return ADT_ERR_INVALID;
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    if (bytes == 0) return ADT_OK;
    NK(ctx, nccl().Broadcast(buf, buf, bytes, ncclUint8, root, c->comm, ctx->stream));
    return ADT_OK;
}

extern "C" int adt_comm_barrier(adt_comm* c) {
    if (!c) return ADT_ERR_INVALID;
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    NK(ctx, nccl().AllReduce(c->d_flag, c->d_flag, 1, ncclInt32, ncclSum, c->comm, ctx->stream));
    ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
