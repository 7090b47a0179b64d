// adt_biquad.cu — the streaming biquad of the reference's CreateEQ3Band
// (pyAudioDspTools/EffectEQ3Band.py:90-118, 121-149, 152-180), one band per
// adt_biquad object, batched over channels.
//
// The recurrence is sequential in time, so the unit of parallelism is the
// channel: one thread per channel.  To keep HBM accesses coalesced a warp (one
// CTA) owns 32 channels and walks time in tiles of 32 samples, with the next
// tile's loads in flight while the current one is filtered: the tile is loaded
// row-wise (32 consecutive samples of one channel = one 128-byte line per warp
// instruction), transposed through shared memory (pitch 33, conflict free),
// filtered column-wise in registers, and stored back row-wise.
//
// Arithmetic is the reference's: every product and sum is an individually
// rounded float64 operation in the reference's left-to-right order (no FMA
// contraction), and with float32 data each output is rounded to float32
// before it is fed back (numpy.insert keeps the array dtype, :109-113).  That
// makes the result bit-identical to the reference, not merely close.
#include <cuda_runtime.h>

#include <new>

#include "adt_internal.h"

namespace {

struct BiquadCoef {
    double c[5];
};

template <typename T>
__device__ __forceinline__ T round_to(double v);
template <>
__device__ __forceinline__ float round_to<float>(double v) { return __double2float_rn(v); }
template <>
__device__ __forceinline__ double round_to<double>(double v) { return v; }

// state: [n_channels][5] doubles = x[n-1], x[n-2], x[n-3], y[n-1], y[n-2]
// One warp per CTA (32 channels); the global loads of tile k+1 are issued into registers before tile k is
// filtered, so HBM latency hides behind the fp64 recurrence (the real bound: ~40 dependent cycles/sample).
template <typename T>
__global__ void __launch_bounds__(32) biquad_kernel(const T* __restrict__ x, T* __restrict__ y, long long pitch,
                                                    long long n, int n_channels, BiquadCoef k,
                                                    double* __restrict__ state) {
    __shared__ T tl[32][33];
    const int lane = threadIdx.x;
    const int c0 = blockIdx.x * 32;
    if (c0 >= n_channels) return;
    const int ch = c0 + lane;
    const bool live = ch < n_channels;
    double x1 = 0, x2 = 0, x3 = 0, y1 = 0, y2 = 0;
    if (live) {
        const double* s = state + (long long)ch * 5;
        x1 = s[0]; x2 = s[1]; x3 = s[2]; y1 = s[3]; y2 = s[4];
    }
    const int rows = min(32, n_channels - c0);
    const T* xr = x + (long long)c0 * pitch + lane;
    T* yr = y + (long long)c0 * pitch + lane;
    T nxt[32];
    // prologue: tile 0 into registers (row i of the tile = 32 consecutive samples of channel c0 + i)
#pragma unroll
    for (int i = 0; i < 32; ++i) nxt[i] = (i < rows && lane < n) ? xr[(long long)i * pitch] : T(0);
    for (long long base = 0; base < n; base += 32) {
        const int w = (int)min((long long)32, n - base);
#pragma unroll
        for (int i = 0; i < 32; ++i) tl[i][lane] = nxt[i];
        __syncwarp();
        // prefetch the next tile while this one is filtered
        const long long nb = base + 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) nxt[i] = (i < rows && nb + lane < n) ? xr[(long long)i * pitch + nb] : T(0);
        if (live) {
            for (int j = 0; j < w; ++j) {
                const double xin = (double)tl[lane][j];
                // ((((c0*x1) + (c1*x2)) + (c2*x3)) - (c3*y1)) - (c4*y2), EffectEQ3Band.py:112
                double acc = __dmul_rn(k.c[0], x1);
                acc = __dadd_rn(acc, __dmul_rn(k.c[1], x2));
                acc = __dadd_rn(acc, __dmul_rn(k.c[2], x3));
                acc = __dsub_rn(acc, __dmul_rn(k.c[3], y1));
                acc = __dsub_rn(acc, __dmul_rn(k.c[4], y2));
                const T out = round_to<T>(acc);
                tl[lane][j] = out;
                x3 = x2; x2 = x1; x1 = xin;
                y2 = y1; y1 = (double)out;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < rows && lane < w) yr[(long long)i * pitch + base] = tl[i][lane];
        __syncwarp();
    }
    if (live) {
        double* s = state + (long long)ch * 5;
        s[0] = x1; s[1] = x2; s[2] = x3; s[3] = y1; s[4] = y2;
    }
}

}  // namespace

struct adt_biquad {
    adt_ctx* ctx = nullptr;
    BiquadCoef k{};
    int n_channels = 0;
    int f64 = 0;
    double* d_state = nullptr;
    void* d_x = nullptr;  // staging for apply_host
    void* d_y = nullptr;
    size_t cap = 0;
};

extern "C" int adt_biquad_create(adt_ctx* ctx, const double coef[5], int32_t n_channels, int32_t f64,
                                 adt_biquad** out) {
    if (!ctx || !coef || !out || n_channels <= 0) return ADT_ERR_INVALID;
    *out = nullptr;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    adt_biquad* b = new (std::nothrow) adt_biquad();
    if (!b) return ADT_ERR_NOMEM;
    b->ctx = ctx;
    for (int i = 0; i < 5; ++i) b->k.c[i] = coef[i];
    b->n_channels = n_channels;
    b->f64 = f64 ? 1 : 0;
    cudaError_t e = cudaMalloc((void**)&b->d_state, (size_t)n_channels * 5 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(b->d_state, 0, (size_t)n_channels * 5 * sizeof(double));
    if (e != cudaSuccess) {
        cudaFree(b->d_state);
        delete b;
        return adt_cuda_fail(ctx, e, "adt_biquad_create");
    }
    *out = b;
    return ADT_OK;
}

extern "C" int adt_biquad_destroy(adt_biquad* b) {
    if (!b) return ADT_ERR_INVALID;
    cudaSetDevice(b->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(b->d_state);
    cudaFree(b->d_x);
    cudaFree(b->d_y);
    delete b;
    return ADT_OK;
}

extern "C" int adt_biquad_reset(adt_biquad* b) {
    if (!b) return ADT_ERR_INVALID;
    ADT_CK(b->ctx, cudaSetDevice(b->ctx->device));
    ADT_CK(b->ctx, cudaMemsetAsync(b->d_state, 0, (size_t)b->n_channels * 5 * sizeof(double), b->ctx->stream));
    return ADT_OK;
}

extern "C" int adt_biquad_apply_dev(adt_biquad* b, const void* x, void* y, int64_t pitch, int64_t n) {
    if (!b || !x || !y || n < 0 || pitch < n) return ADT_ERR_INVALID;
    adt_ctx* ctx = b->ctx;
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((b->n_channels + 31) / 32);
    if (b->f64)
        biquad_kernel<double><<<grid, 32, 0, ctx->stream>>>((const double*)x, (double*)y, pitch, n, b->n_channels, b->k,
                                                             b->d_state);
    else
        biquad_kernel<float><<<grid, 32, 0, ctx->stream>>>((const float*)x, (float*)y, pitch, n, b->n_channels, b->k,
                                                            b->d_state);
    ADT_CK(ctx, cudaGetLastError());
    ctx->launches++;
    return ADT_OK;
}

extern "C" int adt_biquad_apply_host(adt_biquad* b, const void* x, void* y, int64_t pitch, int64_t n) {
    if (!b || !x || !y || n < 0 || pitch < n) return ADT_ERR_INVALID;
    adt_ctx* ctx = b->ctx;
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const size_t es = b->f64 ? sizeof(double) : sizeof(float);
    const size_t need = (size_t)b->n_channels * n * es;
    if (b->cap < need) {
        ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(b->d_x);
        cudaFree(b->d_y);
        b->d_x = b->d_y = nullptr;
        b->cap = 0;
        ADT_CK(ctx, cudaMalloc(&b->d_x, need));
        ADT_CK(ctx, cudaMalloc(&b->d_y, need));
        b->cap = need;
    }
    ADT_CK(ctx, cudaMemcpy2DAsync(b->d_x, n * es, x, pitch * es, n * es, b->n_channels, cudaMemcpyHostToDevice,
                                  ctx->stream));
    int rc = adt_biquad_apply_dev(b, b->d_x, b->d_y, n, n);
    if (rc) return rc;
    ADT_CK(ctx, cudaMemcpy2DAsync(y, pitch * es, b->d_y, n * es, n * es, b->n_channels, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
