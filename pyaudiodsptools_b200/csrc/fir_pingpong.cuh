// fir_pingpong.cuh — the fused FIR block with an ENFORCED complementary schedule (sm_100a).
//
// ncu on fir_block_kernel (profiles/r02_fir_block_kernel_ncu_summary.txt): the FP32 pipe and the L1/shared data
// pipe are co-limiters, each ~2/3 busy, because the two CTAs resident on an SM drift through their FP-heavy
// and exchange-heavy phases at random relative offsets (a persistent loop, which locks them in phase, is 7 %
// slower still).  Here one CTA holds BOTH items of an SM: two groups of T threads, each with its own tile and
// its own work item, walk the same phase sequence
//
//     L | F1 | X1 | F2 | X2 | F3 | X3 | F4 | X4 | F5 | S        F = arithmetic only (registers)
//                                                              X = store to the tile, barrier, load from it
// but a group may only run an F phase while it holds the single "FP token", passed back and forth with named
// barriers (bar.sync / bar.arrive on two ids, the FlashAttention-3 ping-pong pattern): while one group computes,
// the other is necessarily in a load / exchange / store phase, so the two pipes are busy at the same time by
// construction instead of by chance.  The loop is persistent (dynamic queue per CTA) so the stagger survives
// item boundaries.  Arithmetic and index algebra are the phase functions of fft_core.cuh split at the
// register / shared-memory boundary; results are bit-identical to fir_block_kernel.
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"
#include "fir_kernel.cuh"

namespace adt {

#if defined(__CUDACC__)
// named barriers: 1 + g = group-local (T threads), 3 + g = token for group g (2T threads: T wait, T arrive)
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <class C>
struct PingPong {
    int g;   // group 0 / 1
    __device__ __forceinline__ void group_sync() const { bar_sync(1 + g, C::T); }
    __device__ __forceinline__ void acquire() const { bar_sync(3 + g, 2 * C::T); }        // wait for the token
    __device__ __forceinline__ void release() const { bar_arrive(3 + (g ^ 1), 2 * C::T); }  // hand it over
};

// ---- the phases of fft_core.cuh, split into register-only and shared-memory parts -------------------------
template <class C>
__device__ __forceinline__ void pp_fwd1_compute(cf* v, int t, const cf* __restrict__ tw1) {
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        dft<C::N1, -1>(b);
        apply_powers<C::N1, false, true>(b, tw1[t + u * C::T]);
    });
}
template <class C>
__device__ __forceinline__ void pp_fwd1_store(const cf* v, int t, cf* tile) {
    const int lane = t & 31;
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        const cf* b = v + u * C::N1;
        const int row0 = (t + u * C::T) >> 5;
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            tile[(k1 * C::N2 + row0) * C::PITCH + lane] = b[brev<C::N1>(k1)];
        });
    });
}
template <class C, bool BREV_OUT>
__device__ __forceinline__ void pp_cols(cf* v, int t, cf* tile, bool store) {   // stage-2 column access, both directions
    const int lane = t & 31, warp = t >> 5;
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N2;
        cf* col = tile + ((warp + u * C::WARPS) * C::N2) * C::PITCH + lane;
        static_for<0, C::N2>([&](auto K) {
            constexpr int k = decltype(K)::value;
            if (store)
                col[k * C::PITCH] = b[BREV_OUT ? brev<C::N2>(k) : k];
            else
                b[k] = col[k * C::PITCH];
        });
    });
}
template <class C>
__device__ __forceinline__ void pp_fwd2_compute(cf* v, int t, const cf* __restrict__ tw2) {
    static_for<0, C::B2>([&](auto U) { dft<C::N2, -1>(v + decltype(U)::value * C::N2); });
    apply_powers<C::N2, false, true, C::B2>(v, tw2[t & 31]);
}
template <class C, class MaskT>
__device__ __forceinline__ void pp_mid3_compute(cf* v, int t, const MaskT* __restrict__ mask) {
    dft<32, -1>(v);
    cf y[32];
    masked_idft32<MaskT>(v, y, [&](auto K) { return mask[decltype(K)::value * C::T + t]; });
    static_for<0, 32>([&](auto K) { constexpr int r2 = decltype(K)::value; v[r2] = y[brev<32>(r2)]; });
}
template <class C>
__device__ __forceinline__ void pp_row(cf* v, int t, cf* tile, bool store) {
    cf* row = tile + C::stage3_row(t) * C::PITCH;
    static_for<0, 32>([&](auto K) {
        constexpr int r2 = decltype(K)::value;
        if (store)
            row[r2] = v[r2];
        else
            v[r2] = row[r2];
    });
}
template <class C>
__device__ __forceinline__ void pp_inv1_load(cf* v, int t, const cf* tile) {
    const int lane = t & 31;
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        const int row0 = (t + u * C::T) >> 5;
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            b[k1] = tile[(k1 * C::N2 + row0) * C::PITCH + lane];
        });
    });
}
template <class C>
__device__ __forceinline__ void pp_inv1_compute(cf* v, int t, const cf* __restrict__ tw1) {
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        twiddle_idft<C::N1, 1>(v + u * C::N1, tw1[t + u * C::T]);
    });
}

// CTA = 2 groups x C::T threads, one CTA per SM (the register file and the two tiles fill it), persistent.
// ex.work_counter: next unclaimed item (preset to 2 * gridDim.x by the host).
template <class C, class MaskT>
__global__ void __launch_bounds__(2 * C::T, 1) fir_pingpong_kernel(const FirKernelArgs a, const FirExtra ex) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int next_item_s[2];
    PingPong<C> pp;
    pp.g = threadIdx.x / C::T;
    const int t = threadIdx.x % C::T;
    cf* tile = reinterpret_cast<cf*>(smem_raw) + pp.g * C::TILE;
    const MaskT* mask = reinterpret_cast<const MaskT*>(a.mask);
    long long item = 2LL * blockIdx.x + pp.g;
    // Both groups run the token protocol in lock-step for the same number of rounds even when one of them has
    // run out of items (it then skips the work but still passes the token), so nobody waits for a token forever.
    if (pp.g == 1) pp.release();   // prime: group 0 owns the token first
    bool more = true;
    while (more) {
        const bool work = item < a.n_items;
        FirItem<float> it;
        cf v[32];
        if (work) {
            it = fir_item<float>(a, item);
            load_window<C, IoF32>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
            fir_prefetch_l2<C::N, C::T, float>(a, item, t);
        }
        pp.acquire();
        if (work) pp_fwd1_compute<C>(v, t, a.tw1);
        pp.release();
        if (work) pp_fwd1_store<C>(v, t, tile);
        if (t == 0) next_item_s[pp.g] = atomicAdd(ex.work_counter, 1u);
        pp.group_sync();
        if (work) pp_cols<C, false>(v, t, tile, false);
        pp.acquire();
        if (work) pp_fwd2_compute<C>(v, t, a.tw2);
        pp.release();
        if (work) pp_cols<C, true>(v, t, tile, true);
        __syncwarp();
        if (work) pp_row<C>(v, t, tile, false);
        pp.acquire();
        if (work) pp_mid3_compute<C, MaskT>(v, t, mask);
        pp.release();
        if (work) pp_row<C>(v, t, tile, true);
        __syncwarp();
        if (work) pp_cols<C, false>(v, t, tile, false);
        pp.acquire();
        if (work) twiddle_idft<C::N2, C::B2>(v, a.tw2[t & 31]);
        pp.release();
        if (work) pp_cols<C, true>(v, t, tile, true);
        const long long next_item = next_item_s[pp.g];
        pp.group_sync();
        if (work) pp_inv1_load<C>(v, t, tile);
        pp.acquire();
        if (work) pp_inv1_compute<C>(v, t, a.tw1);
        // the loop ends for BOTH groups in the same round: a group continues while either has work left.
        // next items are claimed in increasing order, so "my next item exists or the other group's does" is
        // decided from the two claimed indices, which both groups can read after this point.
        pp.release();
        if (work) store_slice<C, IoF32, false>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
        // agree on termination: both groups read both claimed indices (written before the group barriers above;
        // a CTA-wide barrier makes the other group's value visible and keeps the rounds aligned)
        bar_sync(5, 2 * C::T);
        more = next_item_s[0] < a.n_items || next_item_s[1] < a.n_items;
        bar_sync(5, 2 * C::T);   // nobody overwrites next_item_s before everyone has read it
        item = next_item;
    }
    if (pp.g == 0) pp.acquire();   // consume the last release of group 1, so no barrier is left half-arrived
}
#endif  // __CUDACC__

}  // namespace adt
