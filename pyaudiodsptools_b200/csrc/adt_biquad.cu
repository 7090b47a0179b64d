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

#include <cstdlib>
#include <new>

#include "adt_internal.h"

#ifndef ADT_BIQUAD_PIPE_DEFAULT
#define ADT_BIQUAD_PIPE_DEFAULT 1   /* 1: helper + chain warp per band (biquad3p_kernel, 27.4 ms); 0: one warp per band (biquad3_kernel, 31.7 ms) */
#endif
#ifndef ADT_BIQUAD_ROUND_INT_DEFAULT
#define ADT_BIQUAD_ROUND_INT_DEFAULT 0   /* set from the measurement in tools/microbench/lat64.cu */
#endif

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

// ---- input tiles through an asynchronous ring -------------------------------------------------------------------
// Measured (tools/biquad_clock_probe.py, both chain kernels at 32.9 ms to within 0.2 %): the recurrence kernels
// were bound by the LATENCY of the one-tile-ahead global loads — 32 rows in 32 different 2 MB pages per tile,
// ~2.4 us per round trip — not by arithmetic.  The loader warp therefore keeps BQ_RING - 1 tiles in flight with
// cp.async (LDGSTS: global -> shared without registers); rows beyond the channel count and samples beyond n are
// zero-filled by the copy itself (src-size 0).
#define BQ_RING 8
template <int BYTES>
__device__ __forceinline__ void bq_cp_async(void* dst_smem, const void* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(src), "n"(BYTES), "r"(valid ? BYTES : 0) : "memory");
}
// tile `tile` (32 samples x 32 channels) -> slot[tile % BQ_RING][channel][sample]; always commits one group
template <typename T>
__device__ __forceinline__ void bq_issue_tile(T (*slot)[32][33], long long tile, const T* x, long long row0, long long pitch,
                                              long long n, int rows, int lane) {
    T (&dst)[32][33] = slot[tile % BQ_RING];
    const long long smp = tile * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const bool ok = i < rows && smp < n;
        bq_cp_async<sizeof(T)>(&dst[i][lane], ok ? x + (row0 + i) * pitch + smp : x, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bq_wait_tile() { asm volatile("cp.async.wait_group %0;" ::"n"(BQ_RING - 1) : "memory"); }

// state: [n_channels][5] doubles = x[n-1], x[n-2], x[n-3], y[n-1], y[n-2]
// One warp per CTA (32 channels); the global loads of tile k+1 are issued into registers before tile k is
// filtered, so HBM latency hides behind the fp64 recurrence (the real bound: ~40 dependent cycles/sample).
template <typename T>
__global__ void __launch_bounds__(32) biquad_kernel(const T* __restrict__ x, T* __restrict__ y, long long pitch,
                                                    long long n, int n_channels, BiquadCoef k,
                                                    double* __restrict__ state) {
    extern __shared__ __align__(16) unsigned char bq_smem[];
    typedef T Tile[32][33];
    Tile* ring = reinterpret_cast<Tile*>(bq_smem);             // BQ_RING input tiles
    Tile& tl = ring[BQ_RING];                                  // output tile
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
    T* yr = y + (long long)c0 * pitch + lane;
    const long long n_tiles = (n + 31) / 32;
    for (int p = 0; p < BQ_RING - 1; ++p) bq_issue_tile<T>(ring, p, x, c0, pitch, n, rows, lane);
    for (long long tile = 0; tile < n_tiles; ++tile) {
        const long long base = tile * 32;
        const int w = (int)min((long long)32, n - base);
        bq_issue_tile<T>(ring, tile + BQ_RING - 1, x, c0, pitch, n, rows, lane);
        bq_wait_tile();
        __syncwarp();
        Tile& src = ring[tile % BQ_RING];
        if (live) {
            for (int j = 0; j < w; ++j) {
                const double xin = (double)src[lane][j];
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
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
        double* s = state + (long long)ch * 5;
        s[0] = x1; s[1] = x2; s[2] = x3; s[3] = y1; s[4] = y2;
    }
}
#define ADT_BQ1_SMEM(T) ((BQ_RING + 1) * 32 * 33 * sizeof(T))


// ---- low -> mid -> high chain in ONE launch -----------------------------------------------------------
// The recurrence leaves exactly one unit of parallelism per (channel, band): a dependent chain
// DMUL -> DADD -> DADD -> round-to-float32 per sample.  Three separate band launches walk that chain three
// times back to back; here the three bands of a 32-channel tile run CONCURRENTLY as a software pipeline over
// time tiles: warp b filters tile (step - b) with band b, handing tiles to the next warp through double-
// buffered shared memory, so the chain is walked once (+2 tiles of fill).  Arithmetic per band is the same
// individually rounded sequence as biquad_kernel, so the result stays bit-identical to
// applyhighband(applymidband(applylowband(x))).
//
// ROUND_INT: round the float64 accumulator to float32 precision by integer arithmetic on the bit pattern
// (round-to-nearest-even on the 29 dropped mantissa bits; the carry propagates into the exponent exactly as
// IEEE rounding does) instead of the F2F.F32.F64 / F2F.F64.F32 pair.  Identical results for every value whose
// float32 image is a normal number; values below 2^-126 (float32 denormals), infinities and NaNs take the
// conversion path, so the result is bit-exact everywhere.  Measured: tools/microbench/lat64.cu.
struct Biquad3Args {
    BiquadCoef k[3];
    double* state[3];
};

template <typename T, bool ROUND_INT>
struct BiquadRound;
template <bool ROUND_INT>
struct BiquadRound<double, ROUND_INT> {
    static __device__ __forceinline__ double run(double acc, double* out) { *out = acc; return acc; }
};
template <>
struct BiquadRound<float, false> {
    static __device__ __forceinline__ double run(double acc, float* out) {
        const float f = __double2float_rn(acc);
        *out = f;
        return (double)f;
    }
};
template <>
struct BiquadRound<float, true> {
    static __device__ __forceinline__ double run(double acc, float* out) {
        const long long a = __double_as_longlong(acc);
        long long b = a + (0x0FFFFFFFLL + ((a >> 29) & 1));               // the chain: shift, and, add, and
        b &= ~0x1FFFFFFFLL;
        double r = __longlong_as_double(b);
        const unsigned e = ((unsigned)(a >> 52)) & 0x7ffu;                // biased float64 exponent (off the chain)
        if (__builtin_expect(e - 897u >= 1150u - 897u, 0))                // outside [2^-126, 2^127): zero, float32-
            r = (double)__double2float_rn(acc);                           // denormal, huge, inf, nan -> conversion path
        *out = __double2float_rn(r);      // exact (r is float32-representable) and off the dependent chain
        return r;
    }
};

template <typename T, bool ROUND_INT>
__global__ void __launch_bounds__(96) biquad3_kernel(const T* __restrict__ x, T* __restrict__ y, long long pitch,
                                                     long long n, int n_channels, Biquad3Args a) {
    extern __shared__ __align__(16) unsigned char bq_smem[];
    typedef T Tile[32][33];
    Tile* tiles = reinterpret_cast<Tile*>(bq_smem);   // [1..2] low->mid, [3..4] mid->high, [5] output ([0] unused)
    Tile* ring = tiles + 6;                           // BQ_RING input tiles filled by cp.async (band 0's warp)
    const int lane = threadIdx.x & 31, band = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32;
    const int ch = c0 + lane;
    const bool live = ch < n_channels;
    const int rows = min(32, n_channels - c0);
    const BiquadCoef k = a.k[band];
    double x1 = 0, x2 = 0, x3 = 0, y1 = 0, y2 = 0;
    if (live) {
        const double* s = a.state[band] + (long long)ch * 5;
        x1 = s[0]; x2 = s[1]; x3 = s[2]; y1 = s[3]; y2 = s[4];
    }
    const long long n_tiles = (n + 31) / 32;
    T* yr = y + (long long)c0 * pitch + lane;
    if (band == 0)
        for (int p = 0; p < BQ_RING - 1; ++p) bq_issue_tile<T>(ring, p, x, c0, pitch, n, rows, lane);
    for (long long step = 0; step < n_tiles + 2; ++step) {
        const long long tile = step - band;
        const bool active = tile >= 0 && tile < n_tiles;
        const long long base = tile * 32;
        const int w = active ? (int)min((long long)32, n - base) : 0;
        Tile& src = band == 0 ? ring[(active ? tile : 0) % BQ_RING] : tiles[(band == 1 ? 1 : 3) + (int)(tile & 1)];
        Tile& dst = band == 2 ? tiles[5] : tiles[(band == 0 ? 1 : 3) + (int)(tile & 1)];
        if (band == 0 && active) {
            bq_issue_tile<T>(ring, tile + BQ_RING - 1, x, c0, pitch, n, rows, lane);   // keep BQ_RING - 1 tiles in flight
            bq_wait_tile();                                                            // this step's tile has landed
            __syncwarp();
        }
        if (active && live) {
            // Full tiles run a fixed 32-iteration loop, unrolled so that the loads, the float -> double
            // conversions and the feed-forward sum of later samples are issued while the feedback chain
            // (DMUL -> DSUB -> DSUB -> round) of earlier ones is still in flight.
            auto one = [&](int j) {
                const double xin = (double)src[lane][j];
                double ff = __dmul_rn(k.c[0], x1);
                ff = __dadd_rn(ff, __dmul_rn(k.c[1], x2));
                ff = __dadd_rn(ff, __dmul_rn(k.c[2], x3));
                double acc = __dsub_rn(ff, __dmul_rn(k.c[3], y1));
                acc = __dsub_rn(acc, __dmul_rn(k.c[4], y2));
                T out;
                const double fb = BiquadRound<T, ROUND_INT>::run(acc, &out);
                dst[lane][j] = out;
                x3 = x2; x2 = x1; x1 = xin;
                y2 = y1; y1 = fb;
            };
            if (w == 32) {
                // Full tile in two phases, because a warp issues in order and nothing else runs on its SM
                // partition: (A) every feed-forward sum of the tile — independent of the outputs, so its
                // conversions and products pipeline at full rate; (B) the feedback recurrence alone, whose
                // per-sample cost is then exactly the dependent chain DMUL -> DSUB -> DSUB -> round.
                double ff[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double xin = (double)src[lane][j];
                    double f = __dmul_rn(k.c[0], x1);
                    f = __dadd_rn(f, __dmul_rn(k.c[1], x2));
                    ff[j] = __dadd_rn(f, __dmul_rn(k.c[2], x3));
                    x3 = x2; x2 = x1; x1 = xin;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    double acc = __dsub_rn(ff[j], __dmul_rn(k.c[3], y1));
                    acc = __dsub_rn(acc, __dmul_rn(k.c[4], y2));
                    T out;
                    const double fb = BiquadRound<T, ROUND_INT>::run(acc, &out);
                    dst[lane][j] = out;
                    y2 = y1; y1 = fb;
                }
            } else {
                for (int j = 0; j < w; ++j) one(j);
            }
        }
        if (band == 2 && active) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < rows && lane < w) yr[(long long)i * pitch + base] = dst[i][lane];
        }
        __syncthreads();   // hand the tiles over: band b's output of this step is band b+1's input of the next
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
        double* s = a.state[band] + (long long)ch * 5;
        s[0] = x1; s[1] = x2; s[2] = x3; s[3] = y1; s[4] = y2;
    }
}
#define ADT_BQ3_SMEM(T) ((6 + BQ_RING) * 32 * 33 * sizeof(T))



// ---- the chain with HELPER warps (biquad3p_kernel) ------------------------------------------------------------
// A warp issues in order and is alone on its SM partition, so in biquad3_kernel every conversion or feed-forward
// product the compiler places next to its consumer stalls the feedback chain: ~135 cycles per sample against
// the 62 of the bare chain (tools/microbench/lat64.cu).  Here each band has TWO warps: a helper that turns a tile
// of inputs into the feed-forward sums ff[j] = (c0*x[j-1] + c1*x[j-2]) + c2*x[j-3] — throughput-bound,
// independent of the outputs — and a chain warp that only runs acc = (ff - c3*y1) - c4*y2, rounds and feeds
// back.  Six warps form a pipeline over 32-sample tiles (helper of band b on tile s-2b, chain on tile s-2b-1);
// tiles travel as float64 in double-buffered shared memory, one block barrier per step.  Same individually
// rounded operations in the same order as biquad_kernel -> bit-identical.
#define ADT_BQ3P_SMEM(T) (10 * 32 * 33 * sizeof(double) + (1 + BQ_RING) * 32 * 33 * sizeof(T))

template <typename T>
__global__ void __launch_bounds__(192) biquad3p_kernel(const T* __restrict__ x, T* __restrict__ y, long long pitch,
                                                       long long n, int n_channels, Biquad3Args a) {
    extern __shared__ __align__(16) unsigned char bq_smem[];
    typedef double TileD[32][33];
    typedef T TileT[32][33];
    TileD* ffb = reinterpret_cast<TileD*>(bq_smem);            // [band*2 + parity] feed-forward sums
    TileD* yb = ffb + 6;                                       // [band*2 + parity] band outputs (bands 0, 1)
    TileT& out_t = *reinterpret_cast<TileT*>(bq_smem + 10 * sizeof(TileD));
    TileT* ring = &out_t + 1;                                  // BQ_RING input tiles filled by cp.async (warp 0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int band = warp >> 1;
    const bool is_chain = warp & 1;
    const int c0 = blockIdx.x * 32;
    const int ch = c0 + lane;
    const bool live = ch < n_channels;
    const int rows = min(32, n_channels - c0);
    const BiquadCoef k = a.k[band];
    double s0 = 0, s1 = 0, s2 = 0;     // helper: x[n-1], x[n-2], x[n-3];  chain: y[n-1], y[n-2]
    if (live) {
        const double* st = a.state[band] + (long long)ch * 5;
        if (is_chain) { s0 = st[3]; s1 = st[4]; } else { s0 = st[0]; s1 = st[1]; s2 = st[2]; }
    }
    const long long n_tiles = (n + 31) / 32;
    T* yr = y + (long long)c0 * pitch + lane;
    if (warp == 0)
        for (int p = 0; p < BQ_RING - 1; ++p) bq_issue_tile<T>(ring, p, x, c0, pitch, n, rows, lane);
    for (long long step = 0; step < n_tiles + 5; ++step) {
        const long long tile = step - warp;                    // helper of band b: step - 2b, chain: step - 2b - 1
        const bool active = tile >= 0 && tile < n_tiles;
        const long long base = tile * 32;
        const int w = active ? (int)min((long long)32, n - base) : 0;
        const int par = (int)(tile & 1);
        if (active && !is_chain) {
            // ---- helper: inputs (global for band 0, the previous band's float64 outputs otherwise) -> ff ----
            TileD& ff = ffb[band * 2 + par];
            if (band == 0) {
                bq_issue_tile<T>(ring, tile + BQ_RING - 1, x, c0, pitch, n, rows, lane);
                bq_wait_tile();
                __syncwarp();
            }
            const TileT& in0 = ring[tile % BQ_RING];
            const TileD& inb = yb[(band > 0 ? band - 1 : 0) * 2 + par];
            auto in_at = [&](int j) -> double { return band == 0 ? (double)in0[lane][j] : inb[lane][j]; };
            if (w == 32) {
                double xi[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) xi[j] = in_at(j);
                asm volatile("" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    double f = __dmul_rn(k.c[0], s0);
                    f = __dadd_rn(f, __dmul_rn(k.c[1], s1));
                    ff[lane][j] = __dadd_rn(f, __dmul_rn(k.c[2], s2));
                    s2 = s1; s1 = s0; s0 = xi[j];
                }
            } else {
                for (int j = 0; j < w; ++j) {
                    const double xin = in_at(j);
                    double f = __dmul_rn(k.c[0], s0);
                    f = __dadd_rn(f, __dmul_rn(k.c[1], s1));
                    ff[lane][j] = __dadd_rn(f, __dmul_rn(k.c[2], s2));
                    s2 = s1; s1 = s0; s0 = xin;
                }
            }
        }
        if (active && is_chain) {
            // ---- chain: only the feedback recurrence ----
            TileD& ff = ffb[band * 2 + par];
            auto one = [&](int j) {
                double acc = __dsub_rn(ff[lane][j], __dmul_rn(k.c[3], s0));
                acc = __dsub_rn(acc, __dmul_rn(k.c[4], s1));
                T out;
                const double fb = BiquadRound<T, false>::run(acc, &out);
                if (band == 2)
                    out_t[lane][j] = out;
                else
                    yb[band * 2 + par][lane][j] = fb;
                s1 = s0; s0 = fb;
            };
            if (w == 32) {
                // all 32 operands into registers first (the compiler otherwise issues each LDS right before its
                // consumer, and the in-order warp then waits a shared-memory latency per sample on top of the chain)
                double f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = ff[lane][j];
                asm volatile("" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    double acc = __dsub_rn(f[j], __dmul_rn(k.c[3], s0));
                    acc = __dsub_rn(acc, __dmul_rn(k.c[4], s1));
                    T out;
                    const double fb = BiquadRound<T, false>::run(acc, &out);
                    if (band == 2)
                        out_t[lane][j] = out;
                    else
                        yb[band * 2 + par][lane][j] = fb;
                    s1 = s0; s0 = fb;
                }
            } else {
                for (int j = 0; j < w; ++j) one(j);
            }
            if (band == 2) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < rows && lane < w) yr[(long long)i * pitch + base] = out_t[i][lane];
            }
        }
        __syncthreads();   // every tile moves one pipeline stage
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
        double* st = a.state[band] + (long long)ch * 5;
        if (is_chain) { st[3] = s0; st[4] = s1; } else { st[0] = s0; st[1] = s1; st[2] = s2; }
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
    static bool attr1 = false;
    if (!attr1) {   // the double ring is 76 KB: dynamic shared memory above 48 KB is opt-in
        ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ADT_BQ1_SMEM(double)));
        ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ADT_BQ1_SMEM(float)));
        attr1 = true;
    }
    if (b->f64)
        biquad_kernel<double><<<grid, 32, ADT_BQ1_SMEM(double), ctx->stream>>>((const double*)x, (double*)y, pitch, n, b->n_channels, b->k,
                                                             b->d_state);
    else
        biquad_kernel<float><<<grid, 32, ADT_BQ1_SMEM(float), ctx->stream>>>((const float*)x, (float*)y, pitch, n, b->n_channels, b->k,
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

// low -> mid -> high in one launch; each band keeps its own coefficients and state (the three adt_biquad
// objects of one CreateEQ3Band), so chained and per-band calls can be mixed freely.
extern "C" int adt_biquad_chain_apply_dev(adt_biquad* low, adt_biquad* mid, adt_biquad* high, const void* x, void* y,
                                          int64_t pitch, int64_t n) {
    if (!low || !mid || !high || !x || !y || n < 0 || pitch < n) return ADT_ERR_INVALID;
    adt_ctx* ctx = low->ctx;
    if (mid->ctx != ctx || high->ctx != ctx || mid->n_channels != low->n_channels || high->n_channels != low->n_channels ||
        mid->f64 != low->f64 || high->f64 != low->f64)
        return adt_set_error(ctx, ADT_ERR_INVALID, "the three bands of a chain must share context, channel count and dtype");
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    Biquad3Args a;
    adt_biquad* b[3] = {low, mid, high};
    for (int i = 0; i < 3; ++i) {
        a.k[i] = b[i]->k;
        a.state[i] = b[i]->d_state;
    }
    const unsigned grid = (unsigned)((low->n_channels + 31) / 32);
    static const int pipe = getenv("ADT_BIQUAD_PIPE") ? atoi(getenv("ADT_BIQUAD_PIPE")) : ADT_BIQUAD_PIPE_DEFAULT;
    if (pipe) {   // helper + chain warp per band (biquad3p_kernel)
        static bool attr_p = false;
        if (!attr_p) {
            ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad3p_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)ADT_BQ3P_SMEM(double)));
            ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad3p_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)ADT_BQ3P_SMEM(float)));
            attr_p = true;
        }
        if (low->f64)
            biquad3p_kernel<double><<<grid, 192, ADT_BQ3P_SMEM(double), ctx->stream>>>((const double*)x, (double*)y, pitch, n,
                                                                                      low->n_channels, a);
        else
            biquad3p_kernel<float><<<grid, 192, ADT_BQ3P_SMEM(float), ctx->stream>>>((const float*)x, (float*)y, pitch, n,
                                                                                    low->n_channels, a);
        ADT_CK(ctx, cudaGetLastError());
        ctx->launches++;
        return ADT_OK;
    }
    static const int round_int = getenv("ADT_BIQUAD_ROUND_INT") ? atoi(getenv("ADT_BIQUAD_ROUND_INT")) : ADT_BIQUAD_ROUND_INT_DEFAULT;
    static bool attr_done = false;
    if (!attr_done) {
        ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad3_kernel<double, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADT_BQ3_SMEM(double)));
        ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad3_kernel<float, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADT_BQ3_SMEM(float)));
        ADT_CK(ctx, cudaFuncSetAttribute((const void*)biquad3_kernel<float, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADT_BQ3_SMEM(float)));
        attr_done = true;
    }
    if (low->f64)
        biquad3_kernel<double, false><<<grid, 96, ADT_BQ3_SMEM(double), ctx->stream>>>((const double*)x, (double*)y, pitch, n,
                                                                                      low->n_channels, a);
    else if (round_int)
        biquad3_kernel<float, true><<<grid, 96, ADT_BQ3_SMEM(float), ctx->stream>>>((const float*)x, (float*)y, pitch, n,
                                                                                   low->n_channels, a);
    else
        biquad3_kernel<float, false><<<grid, 96, ADT_BQ3_SMEM(float), ctx->stream>>>((const float*)x, (float*)y, pitch, n,
                                                                                    low->n_channels, a);
    ADT_CK(ctx, cudaGetLastError());
    ctx->launches++;
    return ADT_OK;
}

extern "C" int adt_biquad_chain_apply_host(adt_biquad* low, adt_biquad* mid, adt_biquad* high, const void* x, void* y,
                                           int64_t pitch, int64_t n) {
    if (!low || !mid || !high || !x || !y || n < 0 || pitch < n) return ADT_ERR_INVALID;
    adt_ctx* ctx = low->ctx;
    if (n == 0) return ADT_OK;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    const size_t es = low->f64 ? sizeof(double) : sizeof(float);
    const size_t need = (size_t)low->n_channels * n * es;
    if (low->cap < need) {
        ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(low->d_x);
        cudaFree(low->d_y);
        low->d_x = low->d_y = nullptr;
        low->cap = 0;
        ADT_CK(ctx, cudaMalloc(&low->d_x, need));
        ADT_CK(ctx, cudaMalloc(&low->d_y, need));
        low->cap = need;
    }
    ADT_CK(ctx, cudaMemcpy2DAsync(low->d_x, n * es, x, pitch * es, n * es, low->n_channels, cudaMemcpyHostToDevice,
                                  ctx->stream));
    int rc = adt_biquad_chain_apply_dev(low, mid, high, low->d_x, low->d_y, n, n);
    if (rc) return rc;
    ADT_CK(ctx, cudaMemcpy2DAsync(y, pitch * es, low->d_y, n * es, n * es, low->n_channels, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
