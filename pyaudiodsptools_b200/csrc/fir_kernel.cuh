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
#include "fft_core16.cuh"
#include "shape.cuh"

namespace adt {

ADT_HD float fir_shape(const FirShape& sh, float v) {
    if (sh.kind == 0) return v;
    ShapeParams sp;
    sp.kind = sh.kind; sp.mode = sh.mode; sp.p0 = sh.p0; sp.p1 = sh.p1; sp.p2 = sh.p2; sp.p3 = sh.p3;
    return shape_apply<true>(sp, v);   // fused epilogue: fast division (shape.cuh)
}

struct FirKernelArgs {
    const void* x;         // [n_rows][in_pitch]  float32 (IoF32) or int16 (IoI16)
    void* y;               // [n_rows][out_pitch]
    const void* mask;      // kernel-order mask / coefS table (float or float2 per bin), 1/N folded in
    const cf* coef_x;      // 16-point variant with N3 = 32 only: cross coefficients Hd * W32^(+-j)
    const cf* tw1;         // [M1]
    const cf* tw2;         // [32]
    int n_rows;            // channels
    int blocks_per_row;    // ceil(n_out / hop)
    long long n_items;     // blocks_per_row * ceil(n_rows / 2)
    int prefetch_ahead;    // L2-prefetch the window of item + prefetch_ahead (0 = off)
    FirGeom g;
};

// Rarely used extras travel in a SECOND kernel parameter: growing FirKernelArgs itself by even one
// pointer makes nvcc 12.9 emit a different (measured 4 % slower) code shape for the default kernels.
struct FirExtra {
    unsigned int* work_counter;  // persistent variant: next unclaimed item (preset to gridDim.x by the host)
    FirShape shape;              // SHAPED kernels: wave-shaper applied to every output sample before the store
    // split launch (fir_split_kernel): this launch covers `blk_count` time blocks per row, namely
    // blk = e + blk_offset for e < blk_skip_from, and e + blk_offset + blk_skip_len after that
    int blk_count, blk_offset, blk_skip_from, blk_skip_len;
};

// Work item -> (time block, channel pair).  Time block is the fast index so CTAs that run
// concurrently read overlapping windows of the same rows (the overlap is served by L2).
template <class E>
struct FirItem {
    const E *xa, *xb;
    E *ya, *yb;
    long long m0, ws;
};
template <class E>
__device__ __forceinline__ FirItem<E> fir_item(const FirKernelArgs& a, long long item) {
    FirItem<E> it;
    const int blk = (int)(item % a.blocks_per_row);
    const int row_a = 2 * (int)(item / a.blocks_per_row), row_b = row_a + 1;
    const bool has_b = row_b < a.n_rows;
    const E* x = static_cast<const E*>(a.x);
    E* y = static_cast<E*>(a.y);
    it.xa = x + (long long)row_a * a.g.in_pitch;
    it.xb = has_b ? x + (long long)row_b * a.g.in_pitch : nullptr;
    it.ya = y + (long long)row_a * a.g.out_pitch;
    it.yb = has_b ? y + (long long)row_b * a.g.out_pitch : nullptr;
    it.m0 = (long long)blk * a.g.hop;
    it.ws = it.m0 - a.g.back + a.g.in_shift;
    return it;
}

// Warm L2 with the window of the item `ahead` positions later in launch order (about half a wave of
// resident CTAs): by the time that CTA starts, its 64 loads per thread hit L2 instead of HBM.
// One thread hands each row's window to the bulk-copy (TMA) engine as a single
// cp.async.bulk.prefetch.L2 (SASS: UBLKPF) — no per-line prefetch instructions in the hot loop.
template <int N, int T, class E>
__device__ __forceinline__ void fir_prefetch_l2(const FirKernelArgs& a, long long item, int t) {
    if (a.prefetch_ahead <= 0 || t >= 2) return;
    const long long nxt = item + a.prefetch_ahead;
    if (nxt >= a.n_items) return;
    const FirItem<E> it = fir_item<E>(a, nxt);
    const E* base = t ? it.xb : it.xa;
    if (!base) return;
    // clamp the window to the valid samples and to 16-byte granularity
    constexpr long long G = 16 / (long long)sizeof(E);
    long long lo = it.ws < 0 ? 0 : it.ws, hi = it.ws + N;
    if (hi > a.g.n_in) hi = a.g.n_in;
    const unsigned long long addr = (unsigned long long)(base + lo);
    const long long skip = (long long)((16 - (addr & 15)) & 15) / (long long)sizeof(E);   // up to the next 16-byte boundary
    lo += skip;
    const long long cnt = (hi - lo) / G * G;
    if (cnt <= 0) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + lo), "r"((unsigned)(cnt * sizeof(E))) : "memory");
}

// One CTA per work item (the default).  Persistent CTA loops were measured slower at N = 8192 on B200
// (static stride -15 %, dynamic queue -5 %); see fir_persist_kernel below and DESIGN.md §5.4.
template <class C, class MaskT, int MIN_CTAS, class IO = IoF32, bool SHAPED = false, bool ACCUM = false>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_block_kernel(const FirKernelArgs a, const FirExtra ex) {
    typedef typename IO::elem E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    const int t = threadIdx.x;
    const long long item = blockIdx.x;
    const FirItem<E> it = fir_item<E>(a, item);
    cf v[32];
    load_window<C, IO>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
    fir_prefetch_l2<C::N, C::T, E>(a, item, t);
    fwd_stage1<C>(v, t, a.tw1, tile);
    __syncthreads();
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();  // stage 3 reads only rows written by this warp
    mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask), tile);
    __syncwarp();
    inv_stage2<C>(v, t, a.tw2, tile);
    __syncthreads();
    inv_stage1<C>(v, t, a.tw1, tile);
    store_slice<C, IO, SHAPED, ACCUM>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// MASK IN SHARED MEMORY (A/B, ADT_FIR_SMASK=1; real masks only): the l1tex hit rate of fir_block_kernel is 6 % —
// the streaming windows push the 32 KB mask table out of L1, so the 32 mask loads per thread in the middle of
// the DFT_32 -> mask -> IDFT_32 phase are L2 hits.  Here one thread hands the whole table to the bulk-copy (TMA)
// engine at CTA start (cp.async.bulk.shared::cluster.global, completion on an mbarrier); it lands during stages
// 1-2 and stage 3 reads it with conflict-free LDS.  Tile 67.6 KB + mask 32 KB = 99.6 KB -> still 2 CTAs/SM.
template <class C, int MIN_CTAS>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_smask_kernel(const FirKernelArgs a, const FirExtra ex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    float* smask = reinterpret_cast<float*>(smem_raw + (size_t)C::TILE * sizeof(cf));
    __shared__ __align__(8) unsigned long long bar;
    const int t = threadIdx.x;
    const long long item = blockIdx.x;
    const FirItem<float> it = fir_item<float>(a, item);
    constexpr unsigned MASK_BYTES = C::N * sizeof(float);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(MASK_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smask)), "l"(a.mask), "r"(MASK_BYTES), "r"(smem_u32(&bar)) : "memory");
    }
    cf v[32];
    load_window<C, IoF32>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
    fir_prefetch_l2<C::N, C::T, float>(a, item, t);
    fwd_stage1<C>(v, t, a.tw1, tile);
    __syncthreads();                 // also publishes the initialised mbarrier to every thread
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    mid_stage3<C, float>(v, t, smask, tile);
    __syncwarp();
    inv_stage2<C>(v, t, a.tw2, tile);
    __syncthreads();
    inv_stage1<C>(v, t, a.tw1, tile);
    store_slice<C, IoF32, false>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
}

// TWO VIRTUAL THREADS PER THREAD (A/B, ADT_FIR_VT=2): the CTA has T/2 threads and every phase runs twice, for
// virtual thread ids t and t + T/2.  All exchanges are in place and separated by the same barriers, so the
// result is identical; what changes is the schedule: 3 CTAs of 128 threads per SM instead of 2 of 256 (three
// independent phase positions per SM), up to 168 registers per thread, and two independent butterfly streams
// inside each thread whose shared-memory and FP instructions the compiler can interleave.
template <class C, class MaskT, int MIN_CTAS>
__global__ void __launch_bounds__(C::T / 2, MIN_CTAS) fir_vt2_kernel(const FirKernelArgs a, const FirExtra ex) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    constexpr int H = C::T / 2;
    const int t0 = threadIdx.x, t1 = threadIdx.x + H;
    const long long item = blockIdx.x;
    const FirItem<float> it = fir_item<float>(a, item);
    cf v[32], w[32];
    load_window<C, IoF32>(v, t0, it.xa, it.xb, it.ws, a.g.n_in);
    load_window<C, IoF32>(w, t1, it.xa, it.xb, it.ws, a.g.n_in);
    fir_prefetch_l2<C::N, C::T, float>(a, item, t0);
    fwd_stage1<C>(v, t0, a.tw1, tile);
    fwd_stage1<C>(w, t1, a.tw1, tile);
    __syncthreads();
    fwd_stage2<C>(v, t0, a.tw2, tile);
    fwd_stage2<C>(w, t1, a.tw2, tile);
    __syncwarp();  // virtual warps w and w + WARPS/2 live in the same physical warp: stage 3 reads its own rows
    mid_stage3<C, MaskT>(v, t0, reinterpret_cast<const MaskT*>(a.mask), tile);
    mid_stage3<C, MaskT>(w, t1, reinterpret_cast<const MaskT*>(a.mask), tile);
    __syncwarp();
    inv_stage2<C>(v, t0, a.tw2, tile);
    inv_stage2<C>(w, t1, a.tw2, tile);
    __syncthreads();
    inv_stage1<C>(v, t0, a.tw1, tile);
    store_slice<C, IoF32, false>(v, t0, it.ya, it.yb, it.m0, a.g, ex.shape);
    inv_stage1<C>(w, t1, a.tw1, tile);
    store_slice<C, IoF32, false>(w, t1, it.ya, it.yb, it.m0, a.g, ex.shape);
}

// TMA-FED variant (A/B, ADT_FIR_TMA=1): the two row windows of an interior item are brought
// into the (not yet used) tile by the bulk-copy engine — one cp.async.bulk.shared::cluster.global per row,
// completion on an mbarrier (SASS: UBLKCP + SYNCS) — and stage 1 reads its 32 points from shared memory
// instead of issuing 64 LDG.32.  Costs two extra block barriers (mbarrier init visible; all reads done before
// stage 1 overwrites the tile) and keeps the same number of L1 wavefronts (64 LDS.32 replace 64 LDG.32).
// Edge items (window crossing the ends of the row, odd last row, rows not 16-byte aligned) use the LDG path.

template <class C, class MaskT, int MIN_CTAS>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_tma_kernel(const FirKernelArgs a, const FirExtra ex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    __shared__ __align__(8) unsigned long long bar;
    const int t = threadIdx.x;
    const long long item = blockIdx.x;
    const FirItem<float> it = fir_item<float>(a, item);
    cf v[32];
    const bool bulk = it.xb && it.ws >= 0 && it.ws + C::N <= a.g.n_in &&
                      ((((unsigned long long)(it.xa + it.ws)) | ((unsigned long long)(it.xb + it.ws))) & 15ull) == 0;
    if (bulk) {
        float* sa = reinterpret_cast<float*>(smem_raw);
        float* sb = sa + C::N;
        constexpr unsigned BYTES = C::N * sizeof(float);
        if (t == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(2 * BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sa)), "l"(it.xa + it.ws), "r"(BYTES), "r"(smem_u32(&bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sb)), "l"(it.xb + it.ws), "r"(BYTES), "r"(smem_u32(&bar)) : "memory");
        }
        fir_prefetch_l2<C::N, C::T, float>(a, item, t);
        __syncthreads();                                  // the initialised barrier is visible to every waiter
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
        static_for<0, C::B1>([&](auto U) {
            static_for<0, C::N1>([&](auto K) {
                constexpr int u = decltype(U)::value, n1 = decltype(K)::value;
                constexpr int s = n1 * C::M1 + u * C::T;
                v[u * C::N1 + n1] = mk(sa[s + t], sb[s + t]);
            });
        });
        __syncthreads();                                  // every point is in registers before stage 1 reuses the tile
    } else {
        load_window<C, IoF32>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
        fir_prefetch_l2<C::N, C::T, float>(a, item, t);
    }
    fwd_stage1<C>(v, t, a.tw1, tile);
    __syncthreads();
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();
    mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask), tile);
    __syncwarp();
    inv_stage2<C>(v, t, a.tw2, tile);
    __syncthreads();
    inv_stage1<C>(v, t, a.tw1, tile);
    store_slice<C, IoF32, false>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
}

// PERSISTENT variant with a DYNAMIC work queue: the grid is one wave of resident CTAs; each CTA claims
// items from an atomic counter (a static stride would pace the kernel by the slowest SM — measured 15 %
// slower).  No barrier is needed between items because every tile exchange is in place, so warps flow
// into the next item while slower warps of the CTA finish the current one, and there is no CTA
// launch / drain gap.  The next item index is claimed before barrier 1 and read between the barriers.
template <class C, class MaskT, int MIN_CTAS, class IO = IoF32>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_persist_kernel(const FirKernelArgs a, const FirExtra ex) {
    typedef typename IO::elem E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    __shared__ unsigned int next_item_s;
    const int t = threadIdx.x;
    long long item = blockIdx.x;
    while (item < a.n_items) {
        const FirItem<E> it = fir_item<E>(a, item);
        cf v[32];
        load_window<C, IO>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
        fir_prefetch_l2<C::N, C::T, E>(a, item, t);
        fwd_stage1<C>(v, t, a.tw1, tile);
        if (t == 0) next_item_s = atomicAdd(ex.work_counter, 1u);
        __syncthreads();
        fwd_stage2<C>(v, t, a.tw2, tile);
        __syncwarp();
        mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask), tile);
        __syncwarp();
        inv_stage2<C>(v, t, a.tw2, tile);
        const unsigned int next_item = next_item_s;
        __syncthreads();
        inv_stage1<C>(v, t, a.tw1, tile);
        store_slice<C, IO, false>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
        item = next_item;
    }
}

// SPLIT launch (opt-in, ADT_FIR_SPLIT=1): an interior-only kernel (INTERIOR = true) handles every block whose
// window lies inside the input and whose hop lies inside the output — it contains no bounds code — and the
// same kernel with INTERIOR = false handles the one or two edge blocks per row.
template <class E>
__device__ __forceinline__ FirItem<E> fir_item_split(const FirKernelArgs& a, const FirExtra& ex, long long item) {
    FirItem<E> it;
    int blk = (int)(item % ex.blk_count);
    const int row_a = 2 * (int)(item / ex.blk_count), row_b = row_a + 1;
    blk += ex.blk_offset + (blk >= ex.blk_skip_from ? ex.blk_skip_len : 0);
    const bool has_b = row_b < a.n_rows;
    const E* x = static_cast<const E*>(a.x);
    E* y = static_cast<E*>(a.y);
    it.xa = x + (long long)row_a * a.g.in_pitch;
    it.xb = has_b ? x + (long long)row_b * a.g.in_pitch : nullptr;
    it.ya = y + (long long)row_a * a.g.out_pitch;
    it.yb = has_b ? y + (long long)row_b * a.g.out_pitch : nullptr;
    it.m0 = (long long)blk * a.g.hop;
    it.ws = it.m0 - a.g.back + a.g.in_shift;
    return it;
}

template <class C, class MaskT, int MIN_CTAS, bool INTERIOR>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_split_kernel(const FirKernelArgs a, const FirExtra ex) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    const int t = threadIdx.x;
    const long long item = blockIdx.x;
    const FirItem<float> it = fir_item_split<float>(a, ex, item);
    cf v[32];
    if constexpr (INTERIOR)
        load_window_interior<C, IoF32>(v, t, it.xa, it.xb, it.ws);
    else
        load_window<C, IoF32>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
    if (a.prefetch_ahead > 0 && t < 2) {   // same bulk L2 prefetch, in this launch's own item numbering
        const long long nxt = item + a.prefetch_ahead;
        if (nxt < (long long)ex.blk_count * ((a.n_rows + 1) / 2)) {
            const FirItem<float> nx = fir_item_split<float>(a, ex, nxt);
            const float* base = t ? nx.xb : nx.xa;
            long long lo = nx.ws < 0 ? 0 : nx.ws, hi = nx.ws + C::N;
            if (hi > a.g.n_in) hi = a.g.n_in;
            lo = (lo + 3) & ~3LL;
            const long long cnt = (hi - lo) & ~3LL;
            if (base && cnt > 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + lo), "r"((unsigned)(cnt * 4)) : "memory");
        }
    }
    fwd_stage1<C>(v, t, a.tw1, tile);
    __syncthreads();
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();
    mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask), tile);
    __syncwarp();
    inv_stage2<C>(v, t, a.tw2, tile);
    __syncthreads();
    inv_stage1<C>(v, t, a.tw1, tile);
    store_slice<C, IoF32, false>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
}

// 16 points per thread (fft_core16.cuh): 512 threads at <= 64 registers -> 32 warps per SM.
template <class C, class MaskT, int MIN_CTAS, class IO = IoF32, bool SHAPED = false>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir16_block_kernel(const FirKernelArgs a, const FirExtra ex) {
    typedef typename IO::elem E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    const int t = threadIdx.x;
    const long long item = blockIdx.x;
    const FirItem<E> it = fir_item<E>(a, item);
    cf v[16];
    load_window16<C, IO>(v, t, it.xa, it.xb, it.ws, a.g.n_in);
    fir_prefetch_l2<C::N, C::T, E>(a, item, t);
    fwd16_stage1<C>(v, t, a.tw1, tile);
    __syncthreads();
    fwd16_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();
    mid16_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask), a.coef_x, tile);
    __syncwarp();
    inv16_stage2<C>(v, t, a.tw2, tile);
    __syncthreads();
    inv16_stage1<C>(v, t, a.tw1, tile);
    store_slice16<C, IO, SHAPED>(v, t, it.ya, it.yb, it.m0, a.g, ex.shape);
}

}  // namespace adt
