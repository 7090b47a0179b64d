// Dependent-chain latencies that bound the bit-exact biquad recurrence (adt_biquad.cu) on sm_100a:
// DMUL, DADD, DFMA, the float64 -> float32 -> float64 rounding round trip, an integer emulation of that
// rounding, and the full per-sample step.  One warp, clock64() around a chain of N dependent operations.
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
__device__ __forceinline__ double round32_int(double v) {
    // round-to-nearest-even of a double to 24 significant bits by integer arithmetic on the bit pattern
    // (valid for float32-normal magnitudes; denormal / overflow ranges need the F2F path)
    long long b = __double_as_longlong(v);
    b += 0x0FFFFFFFLL + ((b >> 29) & 1);
    b &= ~0x1FFFFFFFLL;
    return __longlong_as_double(b);
}

template <int MODE>
__global__ void k(double* out, double a, double b, int lanes) {
    __shared__ double sm[32][33];
    __shared__ float smf[32][33];
    const int wid = threadIdx.x >> 5;      // several warps (one per SM partition, then two ...) run the same chain
    const int ln = threadIdx.x & 31;
    if (wid == 0) for (int j = 0; j < 33; ++j) { sm[ln][j] = 0.7 + j * 1e-3; smf[ln][j] = 0.1f * j; }
    __syncthreads();
    if (ln >= lanes) return;
    double y1 = a * ln + 0.1, y2 = 0.3, acc = 0.7, z4 = 0.9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) y1 = __dmul_rn(y1, b);
        if (MODE == 1) y1 = __dadd_rn(y1, b);
        if (MODE == 2) y1 = __fma_rn(y1, b, a);
        if (MODE == 3) y1 = (double)__double2float_rn(y1 + 0.0) ;          // F2F.F32.F64 + F2F.F64.F32 (+ a DADD so it is not folded)
        if (MODE == 4) y1 = round32_int(__dadd_rn(y1, b));
        if (MODE == 5) {   // the biquad step: chain = DMUL, DADD, DADD, round trip
            double t = __dsub_rn(acc, __dmul_rn(b, y1));
            t = __dsub_rn(t, __dmul_rn(a, y2));
            y2 = y1;
            y1 = (double)__double2float_rn(t);
        }
        if (MODE == 6) {   // same with the integer rounding
            double t = __dsub_rn(acc, __dmul_rn(b, y1));
            t = __dsub_rn(t, __dmul_rn(a, y2));
            y2 = y1;
            y1 = round32_int(t);
        }
        if (MODE == 7) y1 = (double)(float)((float)y1 * 1.0001f);            // FMUL + conversions, for comparison
        if (MODE == 10 || MODE == 11 || MODE == 12) {   // the biquad step as the kernels run it: operands through shared memory
            const int j = i & 31;
            double ff = MODE == 12 ? (double)smf[ln][j] * b : sm[ln][j];        // 12: + float->double of an input
            double t = __dsub_rn(ff, __dmul_rn(b, y1));
            t = __dsub_rn(t, __dmul_rn(a, y2));
            y2 = y1;
            const float f = __double2float_rn(t);
            if (MODE != 10) smf[ln][j] = f;                                                // 11, 12: + store of the output
            y1 = (double)f;
        }
        if (MODE == 8) {   // four INDEPENDENT DMUL chains: per-iteration time = 4 x issue interval if the pipe is narrow
            y1 = __dmul_rn(y1, b); y2 = __dmul_rn(y2, b); acc = __dmul_rn(acc, b); z4 = __dmul_rn(z4, b);
        }
        if (MODE == 9) {   // four independent rounding round trips
            y1 = (double)__double2float_rn(__dadd_rn(y1, b)); y2 = (double)__double2float_rn(__dadd_rn(y2, b));
            acc = (double)__double2float_rn(__dadd_rn(acc, b)); z4 = (double)__double2float_rn(__dadd_rn(z4, b));
        }
    }
    long long t1 = clock64();
    out[ln] = y1 + y2 + acc + z4;
    if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / N;
}

template <int MODE>
void run(const char* name, double* d) {
    for (int lanes : {32, 8, 1}) {
        k<MODE><<<1, 32>>>(d, 0.999, 0.5, lanes);
        cudaDeviceSynchronize();
        k<MODE><<<1, 32>>>(d, 0.999, 0.5, lanes);
        double h;
        cudaMemcpy(&h, d + 64, sizeof h, cudaMemcpyDeviceToHost);
        printf("%-52s lanes=%2d  %7.1f cycles per iteration\n", name, lanes, h);
    }
}

template <int MODE>
void run_warps(const char* name, double* d) {
    for (int warps : {1, 2, 3, 4, 6, 8}) {
        k<MODE><<<1, 32 * warps>>>(d, 0.999, 0.5, 32);
        cudaDeviceSynchronize();
        k<MODE><<<1, 32 * warps>>>(d, 0.999, 0.5, 32);
        double h;
        cudaMemcpy(&h, d + 64, sizeof h, cudaMemcpyDeviceToHost);
        printf("%-52s warps on one SM=%d  %7.1f cycles per iteration (warp 0)\n", name, warps, h);
    }
}

int main() {
    double* d;
    cudaMalloc(&d, 128 * sizeof(double));
    run_warps<0>("DMUL chain", d);
    run_warps<3>("DADD + F2F.F32.F64 + F2F.F64.F32", d);
    run_warps<5>("biquad step (DMUL,DADD,DADD,F2F,F2F)", d);
    run_warps<8>("4 independent DMUL chains", d);
    run_warps<9>("4 independent DADD + F2F + F2F chains", d);
    run<0>("DMUL chain", d);
    run<1>("DADD chain", d);
    run<2>("DFMA chain", d);
    run<3>("DADD + F2F.F32.F64 + F2F.F64.F32", d);
    run<4>("DADD + integer round-to-f32", d);
    run<5>("biquad step (DMUL,DADD,DADD,F2F,F2F)", d);
    run<6>("biquad step with integer rounding", d);
    run<7>("F2F + FMUL + F2F", d);
    run<10>("biquad step, ff operand loaded from shared memory", d);
    run<11>("biquad step, ff from shared memory + output stored", d);
    run<12>("biquad step, float input converted + output stored", d);
    run<8>("4 independent DMUL chains (per iteration)", d);
    run<9>("4 independent DADD + F2F + F2F chains (per iteration)", d);
    return 0;
}
