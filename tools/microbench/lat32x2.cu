// Dependent-issue latency of the packed FP32x2 instructions the FIR kernel is made of (sm_100a): a chain of N
// dependent operations in one warp, and the same with 2 / 4 independent chains interleaved (ILP).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int MODE, int ILP>
__global__ void k(float2* out, float2 a, float2 b) {
    float2 v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = make_float2(a.x + threadIdx.x + i, a.y - i);
    long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < N; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) v[i] = __ffma2_rn(v[i], make_float2(0.999f, 0.999f), b);           // imm * pair + pair
            if (MODE == 1) v[i] = __fadd2_rn(v[i], b);                                         // pair + pair
            if (MODE == 2) v[i] = __ffma2_rn(v[i], make_float2(b.x, b.x), a);                  // pair * scalar.F32 + pair
            if (MODE == 3) v[i] = __fadd2_rn(a, make_float2(v[i].y, -v[i].x));                 // pair + swizzled (x -i)
            if (MODE == 4) v[i].x = fmaf(v[i].x, 0.999f, b.x);                                 // scalar FFMA
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) out[64].x = (float)(t1 - t0) / N;
}
template <int MODE, int ILP>
void run(const char* name, float2* d) {
    k<MODE, ILP><<<1, 32>>>(d, make_float2(0.5f, 0.25f), make_float2(1e-3f, 2e-3f));
    cudaDeviceSynchronize();
    k<MODE, ILP><<<1, 32>>>(d, make_float2(0.5f, 0.25f), make_float2(1e-3f, 2e-3f));
    float2 h;
    cudaMemcpy(&h, d + 64, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-44s ILP=%d  %6.2f cycles per iteration  (%.2f per instruction)\n", name, ILP, h.x, h.x / ILP);
}
#define ALL(M, NAME) run<M, 1>(NAME, d); run<M, 2>(NAME, d); run<M, 4>(NAME, d); run<M, 8>(NAME, d);
int main() {
    float2* d;
    cudaMalloc(&d, 128 * sizeof(float2));
    ALL(0, "FFMA2 imm*pair+pair")
    ALL(1, "FADD2 pair+pair")
    ALL(2, "FFMA2 pair*scalar.F32+pair")
    ALL(3, "FADD2 pair+swizzled(.LO_HI.NP)")
    ALL(4, "FFMA scalar")
    return 0;
}
