// Microbenchmark: issue/throughput of FFMA vs FFMA2 / FADD vs FADD2 on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32x2 fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float2 seed, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {  // scalar FFMA x2
                    a[i].x = fmaf(a[i].x, m.x, c.x);
                    a[i].y = fmaf(a[i].y, m.y, c.y);
                } else if (MODE == 1) {  // packed FFMA2
                    a[i] = __ffma2_rn(a[i], m, c);
                } else if (MODE == 2) {  // scalar FADD x2
                    a[i].x = a[i].x + c.x;
                    a[i].y = a[i].y + c.y;
                } else if (MODE == 3) {  // packed FADD2
                    a[i] = __fadd2_rn(a[i], c);
                } else if (MODE == 4) {  // FFMA2 with swizzle + broadcast operands
                    a[i] = __ffma2_rn(make_float2(a[i].y, -a[i].x), make_float2(m.x, m.x), c);
                }
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float2* d, int warps_per_sm) {
    const int blocks = 148 * warps_per_sm / 8, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, make_float2(1.f, 2.f), iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, make_float2(1.f, 2.f), iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)blocks * 256 * iters * 4 * 8 * 2;  // scalar-equivalent fp32 ops (fma = 1)
    printf("%-28s warps/SM=%2d  %.3f ms  %.2f T scalar-ops/s  (%.1f ops/clk/SM @1.9GHz)\n", name, warps_per_sm, ms,
           lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
    float2* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2) * 8);
    for (int w : {8, 16, 32}) {
        run<0>("FFMA (scalar, 2 per pair)", d, w);
        run<1>("FFMA2 (packed)", d, w);
        run<2>("FADD (scalar, 2 per pair)", d, w);
        run<3>("FADD2 (packed)", d, w);
        run<4>("FFMA2 swizzled+bcast", d, w);
    }
    return 0;
}
