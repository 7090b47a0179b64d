// fir_kernel.cuh — the fused overlap-save FIR kernel (sm_100a).
//
// One CTA = one (block b, channel pair p): it reads the N-sample window of two
// planar float32 rows, runs FFT_N -> mask -> IFFT_N entirely in registers and
// one shared-memory tile, and writes the `hop` valid output samples of both
// rows.  Drop-in for the per-chunk body of
//   pyAudioDspTools/EffectFFTFilter.py:143-151 and EffectEQ3BandFFT.py:175-211.
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace adt {

struct FirKernelArgs {
    const float* x;        // [n_rows][in_pitch]
    float* y;              // [n_rows][out_pitch]
    const void* mask;      // kernel-order mask (float or float2 per bin), 1/N folded in
    const cf* tw1;         // [M1]
    const cf* tw2;         // [32]
    int n_rows;            // channels
    FirGeom g;
};

template <class C, class MaskT, int MIN_CTAS>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_block_kernel(const FirKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    const int t = threadIdx.x;
    const int blk = blockIdx.x;
    const int row_a = 2 * blockIdx.y, row_b = row_a + 1;
    const bool has_b = row_b < a.n_rows;

    const float* xa = a.x + (long long)row_a * a.g.in_pitch;
    const float* xb = has_b ? a.x + (long long)row_b * a.g.in_pitch : nullptr;
    float* ya = a.y + (long long)row_a * a.g.out_pitch;
    float* yb = has_b ? a.y + (long long)row_b * a.g.out_pitch : nullptr;

    const long long m0 = (long long)blk * a.g.hop;
    const long long ws = m0 - a.g.back + a.g.in_shift;

    cf v[32];
    load_window<C>(v, t, xa, xb, ws, a.g.n_in);
    fwd_stage1<C>(v, t, a.tw1, tile);
    __syncthreads();
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();  // stage 3 reads only rows written by this warp
    mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask), tile);
    __syncwarp();
    inv_stage2<C>(v, t, a.tw2, tile);
    __syncthreads();
    inv_stage1<C>(v, t, a.tw1, tile);
    store_slice<C>(v, t, ya, yb, m0, a.g);
}

template <class C>
constexpr size_t fir_smem_bytes() {
    return (size_t)C::TILE * sizeof(cf);
}

}  // namespace adt
