// adt_consumers.cu — the reference's in-repo consumers of the FFT filter path:
// standalone pointwise wave-shapers (EffectSaturator.py:41-48, EffectSoftClipper.py:37-44) and the
// feedback delay (EffectDelay.py:31-74), batched over channels.  The shapers can also be attached to an
// adt_fir as a store epilogue (adt_fir_set_epilogue, adt_api.cu).
#include <cuda_runtime.h>

#include <new>
#include <vector>

#include "adt_internal.h"
#include "shape.cuh"

using namespace adt;

namespace {

// pointwise, float4 per thread (n4 = n / 4 vectors; the host handles alignment and the tail)
__global__ void __launch_bounds__(256) shape_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                    ShapeParams sp, int vec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        float4* y4 = reinterpret_cast<float4*>(y);
        const long long n4 = n >> 2;
        for (long long i = i0; i < n4; i += stride) {
            float4 v = x4[i];
            v.x = shape_apply<false>(sp, v.x); v.y = shape_apply<false>(sp, v.y);
            v.z = shape_apply<false>(sp, v.z); v.w = shape_apply<false>(sp, v.w);
            y4[i] = v;
        }
        for (long long i = (n4 << 2) + i0; i < n; i += stride) y[i] = shape_apply<false>(sp, x[i]);
    } else {
        for (long long i = i0; i < n; i += stride) y[i] = shape_apply<false>(sp, x[i]);
    }
}

// delay line: ring[c][(head + off) % m]
// pass k: ring[head + d*(k+1) + j] += x[j] * ramp[k]     (EffectDelay.py:60-64; one launch per k keeps the
//                                                         reference's accumulation order even when n > d)
__global__ void __launch_bounds__(256) delay_add_kernel(const float* __restrict__ x, float* __restrict__ ring,
                                                        long long n, long long m, long long pos0, float r,
                                                        int n_channels) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= n || c >= n_channels) return;
    float* rc = ring + (long long)c * m;
    const long long p = (pos0 + j) % m;
    rc[p] = __fadd_rn(rc[p], __fmul_rn(x[(long long)c * n + j], r));
}
// output: y = x + ring[head + j] (or just the ring when wet), then the consumed slots are zeroed (:66-72)
__global__ void __launch_bounds__(256) delay_out_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        float* __restrict__ ring, long long n, long long m,
                                                        long long head, int wet, int n_channels) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= n || c >= n_channels) return;
    float* rc = ring + (long long)c * m;
    const long long p = (head + j) % m;
    const float b = rc[p];
    y[(long long)c * n + j] = wet ? b : __fadd_rn(x[(long long)c * n + j], b);
    rc[p] = 0.0f;
}

}  // namespace

static int shape_from_args(int kind, const float* params, ShapeParams* sp) {
    if (kind < 1 || kind > 2 || !params) return ADT_ERR_INVALID;
    sp->kind = kind;
    sp->mode = kind == 1 ? (int)params[4] : 0;
    sp->p0 = params[0];
    sp->p1 = params[1];
    sp->p2 = params[2];
    sp->p3 = params[3];
    if (kind == 1 && sp->mode != 1 && sp->mode != 2) return ADT_ERR_INVALID;
    return ADT_OK;
}

extern "C" int adt_shape_apply_dev(adt_ctx* ctx, int kind, const float* params, const float* x_dev, float* y_dev,
                                   int64_t n) {
    if (!ctx || !x_dev || !y_dev || n < 0) return ADT_ERR_INVALID;
    ShapeParams sp;
    int rc = shape_from_args(kind, params, &sp);
    if (rc) return adt_set_error(ctx, rc, "bad shaper kind/params");
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const int vec = (((uintptr_t)x_dev | (uintptr_t)y_dev) & 15) == 0;
    long long blocks = ((vec ? n / 4 : n) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    shape_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(x_dev, y_dev, n, sp, vec);
    ADT_CK(ctx, cudaGetLastError());
    ctx->launches++;
    return ADT_OK;
}

extern "C" int adt_shape_apply_host(adt_ctx* ctx, int kind, const float* params, const float* x, float* y, int64_t n) {
    if (!ctx || !x || !y || n < 0) return ADT_ERR_INVALID;
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    float *dx = nullptr, *dy = nullptr;
    ADT_CK(ctx, cudaMallocAsync((void**)&dx, n * sizeof(float), ctx->stream));
    ADT_CK(ctx, cudaMallocAsync((void**)&dy, n * sizeof(float), ctx->stream));
    ADT_CK(ctx, cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int rc = adt_shape_apply_dev(ctx, kind, params, dx, dy, n);
    if (rc == ADT_OK) {
        cudaError_t e = cudaMemcpyAsync(y, dy, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = adt_cuda_fail(ctx, e, "cudaMemcpyAsync");
    }
    cudaFreeAsync(dx, ctx->stream);
    cudaFreeAsync(dy, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == ADT_OK && e != cudaSuccess) rc = adt_cuda_fail(ctx, e, "cudaStreamSynchronize");
    return rc;
}

// ---- delay ---------------------------------------------------------------------------------------
struct adt_delay {
    adt_ctx* ctx = nullptr;
    int64_t d = 0, m = 0, head = 0;
    int loops = 0, wet = 0, n_channels = 0;
    std::vector<float> ramp;
    float* d_ring = nullptr;
    float *d_x = nullptr, *d_y = nullptr;
    size_t cap = 0;
};

extern "C" int adt_delay_create(adt_ctx* ctx, int64_t delay_samples, int32_t feedback_loops, const float* ramp,
                                int32_t wet, int32_t n_channels, adt_delay** out) {
    if (!ctx || !out || delay_samples < 1 || feedback_loops < 0 || n_channels < 1 || (feedback_loops && !ramp))
        return ADT_ERR_INVALID;
    *out = nullptr;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    adt_delay* dl = new (std::nothrow) adt_delay();
    if (!dl) return ADT_ERR_NOMEM;
    dl->ctx = ctx;
    dl->d = delay_samples;
    dl->loops = feedback_loops;
    dl->wet = wet ? 1 : 0;
    dl->n_channels = n_channels;
    dl->m = delay_samples * (feedback_loops + 2);       // EffectDelay.py:34 max_samples
    dl->ramp.assign(ramp, ramp + feedback_loops);
    const size_t bytes = (size_t)n_channels * dl->m * sizeof(float);
    cudaError_t e = cudaMalloc((void**)&dl->d_ring, bytes);
    if (e == cudaSuccess) e = cudaMemset(dl->d_ring, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(dl->d_ring);
        delete dl;
        return adt_cuda_fail(ctx, e, "adt_delay_create");
    }
    *out = dl;
    return ADT_OK;
}

extern "C" int adt_delay_destroy(adt_delay* dl) {
    if (!dl) return ADT_ERR_INVALID;
    cudaSetDevice(dl->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(dl->d_ring);
    cudaFree(dl->d_x);
    cudaFree(dl->d_y);
    delete dl;
    return ADT_OK;
}

extern "C" int adt_delay_apply_dev(adt_delay* dl, const float* x_dev, float* y_dev, int64_t n) {
    if (!dl || !x_dev || !y_dev || n < 0) return ADT_ERR_INVALID;
    adt_ctx* ctx = dl->ctx;
    // the reference writes delay_buffer[d*(k+1) : d*(k+1)+n] inside a buffer of d*(loops+2) samples
    if (n > 2 * dl->d) return adt_set_error(ctx, ADT_ERR_INVALID, "chunk of %lld samples exceeds 2 x delay (%lld)",
                                            (long long)n, (long long)dl->d);
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)dl->n_channels);
    for (int k = 0; k < dl->loops; ++k) {
        delay_add_kernel<<<grid, 256, 0, ctx->stream>>>(x_dev, dl->d_ring, n, dl->m, dl->head + dl->d * (k + 1),
                                                       dl->ramp[k], dl->n_channels);
        ctx->launches++;
    }
    delay_out_kernel<<<grid, 256, 0, ctx->stream>>>(x_dev, y_dev, dl->d_ring, n, dl->m, dl->head, dl->wet,
                                                   dl->n_channels);
    ctx->launches++;
    ADT_CK(ctx, cudaGetLastError());
    dl->head = (dl->head + n) % dl->m;
    return ADT_OK;
}

extern "C" int adt_delay_apply_host(adt_delay* dl, const float* x, float* y, int64_t n) {
    if (!dl || !x || !y || n < 0) return ADT_ERR_INVALID;
    adt_ctx* ctx = dl->ctx;
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const size_t need = (size_t)dl->n_channels * n * sizeof(float);
    if (dl->cap < need) {
        ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dl->d_x);
        cudaFree(dl->d_y);
        dl->d_x = dl->d_y = nullptr;
        dl->cap = 0;
        ADT_CK(ctx, cudaMalloc((void**)&dl->d_x, need));
        ADT_CK(ctx, cudaMalloc((void**)&dl->d_y, need));
        dl->cap = need;
    }
    ADT_CK(ctx, cudaMemcpyAsync(dl->d_x, x, need, cudaMemcpyHostToDevice, ctx->stream));
    int rc = adt_delay_apply_dev(dl, dl->d_x, dl->d_y, n);
    if (rc) return rc;
    ADT_CK(ctx, cudaMemcpyAsync(y, dl->d_y, need, cudaMemcpyDeviceToHost, ctx->stream));
    ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
