// fir_cluster.cuh — the fused overlap-save FIR block for transforms that do not fit one CTA's shared memory:
// a thread-block CLUSTER of CS CTAs computes one N-point transform pair, N = N1 * N2 * 32, exchanging the
// stage-1 <-> stage-2 transposition through DISTRIBUTED SHARED MEMORY (sm_90+/sm_100a).
//
//   N = 32768 (N1 = 32, N2 = 32, CS = 4)   the chunk-16384 filters and the EQ at large chunks
//   N = 16384 (N1 = 16, N2 = 32, CS = 2)   A/B against the one-CTA N = 16384 kernel
//
// Same drop-in scope as fir_kernel.cuh (pyAudioDspTools/EffectFFTFilter.py:143-151, EffectEQ3BandFFT.py:175-211).
//
// Work split.  Every CTA has T = N1*N2/CS threads with 32 points each and one [T][33] float2 tile:
//   stage 1   CTA c owns the columns r in [c*M1/CS, (c+1)*M1/CS) of the [N1][M1] view of the window (M1 = N2*32):
//             DFT_N1 down each column, times W_N^(r*k1), and element (k1, r) is PUSHED into the tile of the CTA
//             that owns sub-transform k1 (CTA k1 / (N1/CS)) with st.async — a remote shared-memory store that
//             reports its bytes to an mbarrier in the destination CTA;
//   stage 2/3 CTA c owns the N1/CS sub-transforms k1 in [c*N1/CS, ...): an M1-point transform each, done exactly
//             like the single-CTA kernel (DFT_N2 down tile columns, DFT_32 along tile rows, mask, and back);
//   inverse   the results of inverse stage 2 are pushed back to the CTA that owns column r, which runs the
//             inverse DFT_N1 and stores its slice of the output.
// Synchronisation is mbarrier-only after one cluster barrier at kernel start (mbarrier initialisation visible):
//   full   (per CTA) transaction count = the 64 KB the tile receives; phase 0 = forward push, phase 1 = inverse push
//   freed  (per CTA) CS arrivals: "every CTA of the cluster has read its stage-2 data into registers", after which
//          the tiles may be overwritten by the inverse push.  Arrive and wait are split, the inverse DFT_N2 runs
//          in between.
// No cluster-wide barrier sits on the data path (barrier.cluster.wait also invalidates L1, which would evict the
// mask and twiddle tables).
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"
#include "fir_kernel.cuh"

namespace adt {

template <int N1_, int N2_, int CS_>
struct FirClusterCfg {
    static constexpr int N1 = N1_, N2 = N2_, N3 = 32, CS = CS_;
    static constexpr int N = N1 * N2 * 32;
    static constexpr int T = N1 * N2 / CS;    // threads per CTA = tile rows per CTA
    static constexpr int M1 = N2 * 32;        // columns of the stage-1 view = size of one sub-transform
    static constexpr int MC = M1 / CS;        // columns owned by one CTA in stage 1
    static constexpr int K1L = N1 / CS;       // sub-transforms owned by one CTA in stages 2/3
    static constexpr int N2L = MC / 32;       // tile rows per k1 in the inverse stage-1 layout
    static constexpr int B1 = 32 / N1;        // stage-1 columns per thread
    static constexpr int B2 = 32 / N2;        // stage-2 butterflies per thread
    static constexpr int PITCH = 33;
    static constexpr int TILE = T * PITCH;
    static constexpr int WARPS = T / 32;
    static_assert(MC == T * B1, "every column of the CTA's range has exactly one owner thread");
    static_assert(K1L == WARPS * B2, "stage 2: every local sub-transform has one warp");
    static_assert(N2L * N1 == T, "inverse stage-1 layout fills the tile");
    ADT_HD static constexpr int stage3_row(int t) {
        return (((t >> 5) + ((t & 31) / N2) * WARPS) * N2) + ((t & 31) % N2);
    }
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// remote (or local) 8-byte store whose completion is counted on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_cf(unsigned addr, cf v, unsigned bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "f"(v.x), "f"(v.y), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ cf tw1_at(const cf* __restrict__ tw1, int r) { return tw1[r]; }

// ---- forward stage 1: DFT_N1 down this thread's columns, twiddle, push to the owner of each sub-transform ----
// tcol = thread index + first column of the CTA (a multiple of 32, so the lane is unchanged).
template <class C>
__device__ __forceinline__ void cl_fwd_stage1(cf* v, int tcol, const cf* __restrict__ tw1, const unsigned* tile_at,
                                              const unsigned* full_at) {
    const int lane = tcol & 31;
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        dft<C::N1, -1>(b);
        const int r = tcol + u * C::T;
        apply_powers<C::N1, false, true>(b, tw1[r]);
        const int row0 = r >> 5;  // n2
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            constexpr int d = k1 / C::K1L, kl = k1 % C::K1L;
            st_async_cf(tile_at[d] + (unsigned)(((kl * C::N2 + row0) * C::PITCH + lane) * sizeof(cf)), b[brev<C::N1>(k1)],
                        full_at[d]);
        });
    });
}

// ---- inverse stage 2, split in three so the "tile is free" hand-shake overlaps the arithmetic ----------------
template <class C>
__device__ __forceinline__ void cl_inv_stage2_load(cf* v, int t, const cf* tile) {
    const int lane = t & 31, warp = t >> 5;
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N2;
        const cf* col = tile + ((warp + u * C::WARPS) * C::N2) * C::PITCH + lane;
        static_for<0, C::N2>([&](auto K) { constexpr int k2 = decltype(K)::value; b[k2] = col[k2 * C::PITCH]; });
    });
}
// After twiddle_idft the value for (k1, column n2*32 + lane) sits in b[brev(n2)]; it goes to the CTA that owns
// the column, in the layout inverse stage 1 reads: row k1*N2L + (local column >> 5), col lane.
template <class C>
__device__ __forceinline__ void cl_inv_stage2_push(const cf* v, int t, unsigned rank, const unsigned* tile_at,
                                                   const unsigned* full_at) {
    const int lane = t & 31, warp = t >> 5;
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        const cf* b = v + u * C::N2;
        const int k1 = (int)rank * C::K1L + warp + u * C::WARPS;
        static_for<0, C::N2>([&](auto K) {
            constexpr int n2 = decltype(K)::value;
            constexpr int d = (n2 * 32) / C::MC, rl = (n2 * 32) % C::MC;   // owner CTA and local column of lane 0
            st_async_cf(tile_at[d] + (unsigned)(((k1 * C::N2L + (rl >> 5)) * C::PITCH + lane) * sizeof(cf)),
                        b[brev<C::N2>(n2)], full_at[d]);
        });
    });
}
template <class C>
__device__ __forceinline__ void cl_inv_stage1(cf* v, int t, int tcol, const cf* __restrict__ tw1, const cf* tile) {
    const int lane = t & 31;
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        const int rl = t + u * C::T;           // local column
        const int row0 = rl >> 5;
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            b[k1] = tile[(k1 * C::N2L + row0) * C::PITCH + lane];
        });
        twiddle_idft<C::N1, 1>(b, tw1[tcol + u * C::T]);
    });
}

template <class C, class MaskT, int MIN_CTAS, bool ACCUM>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_cluster_kernel(const FirKernelArgs a, const FirExtra ex) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    __shared__ __align__(8) unsigned long long bars[2];   // [0] full (tx count), [1] freed (CS arrivals)
    const int t = threadIdx.x;
    const unsigned rank = cluster_rank();
    const long long item = blockIdx.x / C::CS;
    const FirItem<float> it = fir_item<float>(a, item);
    const int tcol = t + (int)rank * C::MC;
    const unsigned full_l = smem_u32(&bars[0]), freed_l = smem_u32(&bars[1]);
    constexpr unsigned TILE_BYTES = (unsigned)(C::T * 32 * sizeof(cf));   // what one push phase delivers to a tile
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_l));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(freed_l), "r"((unsigned)C::CS));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_l), "r"(TILE_BYTES) : "memory");
    }
    unsigned tile_at[C::CS], full_at[C::CS];
    static_for<0, C::CS>([&](auto D) {
        constexpr int d = decltype(D)::value;
        tile_at[d] = map_to_rank(smem_u32(tile), (unsigned)d);
        full_at[d] = map_to_rank(full_l, (unsigned)d);
    });
    cf v[32];
    load_window<C, IoF32>(v, tcol, it.xa, it.xb, it.ws, a.g.n_in);
    if (rank == 0) fir_prefetch_l2<C::N, C::T, float>(a, item, t);
    // every CTA of the cluster is running and its barriers are initialised before anything is pushed into it
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    cl_fwd_stage1<C>(v, tcol, a.tw1, tile_at, full_at);
    mbar_wait(full_l, 0);                                  // all CS*T threads' pushes have landed in this tile
    if (t == 0)   // re-arm for the inverse push (phase 1) — early remote complete_tx only drive the count negative
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_l), "r"(TILE_BYTES) : "memory");
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();  // stage 3 reads only rows written by this warp
    mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask) + (size_t)rank * 32 * C::T, tile);
    __syncwarp();
    cl_inv_stage2_load<C>(v, t, tile);
    __syncthreads();                                       // this CTA's tile has been read completely
    if (t < C::CS)                                         // tell every CTA of the cluster (including this one)
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_rank(freed_l, (unsigned)t)) : "memory");
    twiddle_idft<C::N2, C::B2>(v, a.tw2[t & 31]);
    mbar_wait(freed_l, 0);                                 // every tile of the cluster may be overwritten now
    cl_inv_stage2_push<C>(v, t, rank, tile_at, full_at);
    mbar_wait(full_l, 1);
    cl_inv_stage1<C>(v, t, tcol, a.tw1, tile);
    store_slice<C, IoF32, false, ACCUM>(v, tcol, it.ya, it.yb, it.m0, a.g, ex.shape);
}

// ---------------------------------------------------------------------------------------------------------
// 2-CTA cluster with a BULK distributed-shared-memory exchange (fir_cluster2b_kernel).
//
// The st.async version above issues 32 remote 8-byte stores per thread and exchange and is bound by their
// throughput (~11 B/clk/SM).  Here the half of a CTA's stage-1 results that belongs to the
// other CTA is first written — in exactly the destination's tile layout — into a 33 KB staging buffer in the
// CTA's own shared memory (ordinary STS, the same number of stores as writing the tile), and then handed to
// the bulk-copy (TMA) engine: 8 x cp.async.bulk.shared::cluster.shared::cta of 4224 bytes (16 tile rows each),
// completing on the destination CTA's mbarrier.  The threads are free as soon as the copies are issued, and the
// other CTA resident on the SM computes while the engine moves the data.
// Tile 67.6 KB + staging 33.8 KB = 101.4 KB per CTA -> still 2 CTAs per SM.
template <class C>
struct Cluster2Layout {
    static_assert(C::CS == 2, "bulk exchange is written for 2-CTA clusters");
    static constexpr int CHUNK_ROWS = C::N2L;                                   // rows of one k1 block that one CTA fills
    static constexpr unsigned CHUNK_BYTES = CHUNK_ROWS * C::PITCH * sizeof(cf);   // 16 x 264 = 4224
    static constexpr int CHUNKS = C::K1L;                                       // 8
    static constexpr unsigned STAGE_BYTES = CHUNKS * CHUNK_BYTES;               // 33 792
    static_assert(CHUNK_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
};

__device__ __forceinline__ void bulk_push(unsigned dst_cluster, unsigned src_cta, unsigned bytes, unsigned bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}

template <class C, class MaskT, int MIN_CTAS, bool ACCUM>
__global__ void __launch_bounds__(C::T, MIN_CTAS) fir_cluster2b_kernel(const FirKernelArgs a, const FirExtra ex) {
    typedef Cluster2Layout<C> L;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cf* tile = reinterpret_cast<cf*>(smem_raw);
    cf* stage = reinterpret_cast<cf*>(smem_raw + (size_t)C::TILE * sizeof(cf));
    __shared__ __align__(8) unsigned long long bars[2];   // [0] full (tx count), [1] freed (CS arrivals)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned rank = cluster_rank(), peer = rank ^ 1u;
    const long long item = blockIdx.x / C::CS;
    const FirItem<float> it = fir_item<float>(a, item);
    const int tcol = t + (int)rank * C::MC;
    const unsigned full_l = smem_u32(&bars[0]), freed_l = smem_u32(&bars[1]);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_l));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(freed_l), "r"((unsigned)C::CS));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_l), "r"(L::STAGE_BYTES) : "memory");
    }
    const unsigned peer_tile = map_to_rank(smem_u32(tile), peer), peer_full = map_to_rank(full_l, peer);
    cf v[32];
    load_window<C, IoF32>(v, tcol, it.xa, it.xb, it.ws, a.g.n_in);
    if (rank == 0) fir_prefetch_l2<C::N, C::T, float>(a, item, t);
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

    // ---- forward stage 1: own sub-transforms go to the tile, the peer's to the staging buffer ----------------
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        dft<C::N1, -1>(b);
        const int r = tcol + u * C::T;
        apply_powers<C::N1, false, true>(b, tw1_at(a.tw1, r));
        const int row0 = r >> 5;                       // global n2 row, in [rank*N2L, (rank+1)*N2L)
        const int row0l = row0 - (int)rank * C::N2L;   // row inside a staging chunk
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            constexpr int d = k1 / C::K1L, kl = k1 % C::K1L;
            cf* dst = ((unsigned)d == rank) ? tile + (kl * C::N2 + row0) * C::PITCH + lane
                                            : stage + (kl * L::CHUNK_ROWS + row0l) * C::PITCH + lane;
            *dst = b[brev<C::N1>(k1)];
        });
    });
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes visible to the bulk-copy engine
    __syncthreads();
    if (t < L::CHUNKS)   // chunk t of the staging buffer = rows [t*N2 + rank*N2L, +N2L) of the peer's tile
        bulk_push(peer_tile + (unsigned)((t * C::N2 + (int)rank * C::N2L) * C::PITCH * sizeof(cf)),
                  smem_u32(stage) + (unsigned)t * L::CHUNK_BYTES, L::CHUNK_BYTES, peer_full);
    mbar_wait(full_l, 0);                                  // the peer's half has landed in this tile
    if (t == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_l), "r"(L::STAGE_BYTES) : "memory");
    fwd_stage2<C>(v, t, a.tw2, tile);
    __syncwarp();
    mid_stage3<C, MaskT>(v, t, reinterpret_cast<const MaskT*>(a.mask) + (size_t)rank * 32 * C::T, tile);
    __syncwarp();
    cl_inv_stage2_load<C>(v, t, tile);
    __syncthreads();                                       // this CTA's tile has been read completely
    if (t < C::CS)
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_rank(freed_l, (unsigned)t)) : "memory");
    twiddle_idft<C::N2, C::B2>(v, a.tw2[lane]);
    // The peer's arrival on `freed` also tells that it has consumed the forward push, i.e. that the bulk copies
    // have finished reading this CTA's staging buffer: it may be refilled now.
    mbar_wait(freed_l, 0);
    // ---- inverse push: value (k1 = rank*K1L + warp, column n2*32 + lane) -> owner of the column ---------------
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        const cf* b = v + u * C::N2;
        const int kl = warp + u * C::WARPS;            // local sub-transform
        const int k1 = (int)rank * C::K1L + kl;
        static_for<0, C::N2>([&](auto K) {
            constexpr int n2 = decltype(K)::value;
            constexpr int d = (n2 * 32) / C::MC, rl = (n2 * 32) % C::MC;
            cf* dst = ((unsigned)d == rank) ? tile + (k1 * C::N2L + (rl >> 5)) * C::PITCH + lane
                                            : stage + (kl * L::CHUNK_ROWS + (rl >> 5)) * C::PITCH + lane;
            *dst = b[brev<C::N2>(n2)];
        });
    });
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t < L::CHUNKS)   // chunk t = rows [(rank*K1L + t)*N2L, +N2L) of the peer's tile (inverse stage-1 layout)
        bulk_push(peer_tile + (unsigned)((((int)rank * C::K1L + t) * C::N2L) * C::PITCH * sizeof(cf)),
                  smem_u32(stage) + (unsigned)t * L::CHUNK_BYTES, L::CHUNK_BYTES, peer_full);
    mbar_wait(full_l, 1);
    cl_inv_stage1<C>(v, t, tcol, a.tw1, tile);
    store_slice<C, IoF32, false, ACCUM>(v, tcol, it.ya, it.yb, it.m0, a.g, ex.shape);
}
#endif  // __CUDACC__

}  // namespace adt
