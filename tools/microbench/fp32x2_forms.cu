// Microbenchmark: throughput of the FFMA2/FADD2 operand forms used by the FIR kernel (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>

#define NA 6
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, const float2* in, int iters) {
    float2 a[NA], b[NA], c[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        a[i] = in[threadIdx.x + 256 * i];
        b[i] = in[threadIdx.x + 256 * (i + NA)];
        c[i] = in[threadIdx.x + 256 * (i + 2 * NA)];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int j = (i + 1) % NA, l = (i + 2) % NA;
            if (MODE == 0) {          // FFMA2: three distinct register pairs, results feed other slots (no reuse)
                a[i] = __ffma2_rn(a[i], b[j], c[l]);
                b[i] = __ffma2_rn(b[i], c[j], a[l]);
                c[i] = __ffma2_rn(c[i], a[j], b[l]);
            } else if (MODE == 1) {   // FFMA2: scalar broadcast (.F32) * pair + pair   (cmul / fma_s form)
                a[i] = __ffma2_rn(make_float2(b[j].x, b[j].x), c[l], a[i]);
                b[i] = __ffma2_rn(make_float2(c[j].y, c[j].y), a[l], b[i]);
                c[i] = __ffma2_rn(make_float2(a[j].x, a[j].x), b[l], c[i]);
            } else if (MODE == 2) {   // FFMA2: immediate * pair + pair  (constant-twiddle butterfly form)
                a[i] = __ffma2_rn(make_float2(0.92387953f, 0.92387953f), b[j], a[i]);
                b[i] = __ffma2_rn(make_float2(0.38268343f, 0.38268343f), c[j], b[i]);
                c[i] = __ffma2_rn(make_float2(-0.70710678f, -0.70710678f), a[j], c[i]);
            } else if (MODE == 3) {   // FFMA2: immediate * swizzled/negated pair + pair
                a[i] = __ffma2_rn(make_float2(0.92387953f, 0.92387953f), make_float2(-b[j].y, b[j].x), a[i]);
                b[i] = __ffma2_rn(make_float2(0.38268343f, 0.38268343f), make_float2(-c[j].y, c[j].x), b[i]);
                c[i] = __ffma2_rn(make_float2(-0.70710678f, -0.70710678f), make_float2(-a[j].y, a[j].x), c[i]);
            } else if (MODE == 4) {   // FADD2: two distinct pairs
                a[i] = __fadd2_rn(a[i], b[j]);
                b[i] = __fadd2_rn(b[i], c[j]);
                c[i] = __fadd2_rn(c[i], a[j]);
            } else if (MODE == 5) {   // FADD2 with swizzle+negate on one operand (the -i butterfly)
                a[i] = __fadd2_rn(a[i], make_float2(b[j].y, -b[j].x));
                b[i] = __fadd2_rn(b[i], make_float2(c[j].y, -c[j].x));
                c[i] = __fadd2_rn(c[i], make_float2(a[j].y, -a[j].x));
            } else if (MODE == 6) {   // scalar FFMA x2, three distinct registers each
                a[i].x = fmaf(a[i].x, b[j].x, c[l].x); a[i].y = fmaf(a[i].y, b[j].y, c[l].y);
                b[i].x = fmaf(b[i].x, c[j].x, a[l].x); b[i].y = fmaf(b[i].y, c[j].y, a[l].y);
                c[i].x = fmaf(c[i].x, a[j].x, b[l].x); c[i].y = fmaf(c[i].y, a[j].y, b[l].y);
            } else if (MODE == 7) {   // FFMA2: pair * 2.0 - pair   (b' = 2a - a')
                a[i] = __ffma2_rn(b[j], make_float2(2.f, 2.f), make_float2(-a[i].x, -a[i].y));
                b[i] = __ffma2_rn(c[j], make_float2(2.f, 2.f), make_float2(-b[i].x, -b[i].y));
                c[i] = __ffma2_rn(a[j], make_float2(2.f, 2.f), make_float2(-c[i].x, -c[i].y));
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < NA; ++i) { s.x += a[i].x + b[i].x + c[i].x; s.y += a[i].y + b[i].y + c[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float2* d, const float2* in, int warps_per_sm) {
    const int blocks = 148 * warps_per_sm / 8, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, in, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double packed = (double)blocks * 8 * iters * NA * 3;          // warp-level packed instructions
    const double cyc = ms * 1e-3 * 1.915e9 * 148 * 4;                   // SMSP-cycles available
    printf("%-44s warps/SM=%2d  %.3f ms  %.2f SMSP-cycles per packed instr (2 scalar for mode 6)\n", name, warps_per_sm,
           ms, cyc / packed);
}

int main() {
    float2 *d, *in;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2) * 8);
    cudaMalloc(&in, 256 * 3 * NA * sizeof(float2));
    cudaMemset(in, 0, 256 * 3 * NA * sizeof(float2));
    for (int w : {16, 32}) {
        run<0>("FFMA2 pair*pair+pair (3 distinct pairs)", d, in, w);
        run<1>("FFMA2 scalar.F32*pair+pair", d, in, w);
        run<2>("FFMA2 imm*pair+pair", d, in, w);
        run<3>("FFMA2 imm*swizzled(-y,x)+pair", d, in, w);
        run<4>("FADD2 pair+pair", d, in, w);
        run<5>("FADD2 pair+swizzled(y,-x)", d, in, w);
        run<6>("2x scalar FFMA (3 distinct regs)", d, in, w);
        run<7>("FFMA2 pair*2-pair", d, in, w);
    }
    return 0;
}
