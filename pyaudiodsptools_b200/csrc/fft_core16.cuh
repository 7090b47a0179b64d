// fft_core16.cuh — the 16-points-per-thread variant of the fused FIR block
// (same math as fft_core.cuh, twice the resident warps).
//
//   N = 16 * 16 * N3,  N3 = 32 (N = 8192, 512 threads) or 16 (N = 4096, 256 threads)
//   every thread holds 16 complex points (32 registers) in all stages, so the
//   kernel fits 64 registers and two 512-thread CTAs (32 warps) live on an SM.
//
//   stage 1 (thread r = t):            DFT_16 over n1 of x[n1*M1 + r], * W_N^(r*k1)
//   stage 2 (thread (k1, r2)):         DFT_16 over n2 of A[k1][n2*N3 + r2], * W_M1^(r2*k2)
//   stage 3, N3 = 32 (two lanes per tile row (k1,k2), l and l^16):
//        lane half h takes r2 = 2m + h: E = DFT_16(even r2) on lane A, O = DFT_16(odd r2)
//        on lane B.  The DFT_32 recombination, the spectral mask and the first
//        radix-2 step of the inverse DFT_32 collapse into ONE pair exchange:
//            e'[j] = Hs[j]*E[j] + (Hd[j]*W32^j )*O[j]     (lane A)
//            o'[j] = Hs[j]*O[j] + (Hd[j]*W32^-j)*E[j]     (lane B)
//        with Hs = H[k]+H[k+N/2], Hd = H[k]-H[k+N/2] precomputed on the host, i.e.
//        new = coefS*mine + coefX*other with `other` fetched by __shfl_xor(.., 16).
//        Then IDFT_16 of e' gives the even r2 outputs, of o' the odd ones.
//        For a REAL (zero-phase) mask lane B's cross coefficient is the conjugate
//        of lane A's, so both lanes read ONE row-indexed table (48 KB in all, stays
//        L1-resident); complex masks need per-thread tables.
//   stage 3, N3 = 16: plain DFT_16 -> mask -> IDFT_16 per row.
//   then the mirror image of stages 2 and 1.
//
// Tile: 256 rows (k1*16 + k2) x (N3 + 1) float2, all exchanges in place; stage 3 rows
// are owned by the warp that wrote them, so only two block barriers are needed.
#pragma once
#include "fft_core.cuh"

namespace adt {

template <int N3_>
struct Fir16Cfg {
    static constexpr int N1 = 16, N2 = 16, N3 = N3_;
    static constexpr int N = 256 * N3;
    static constexpr int T = 16 * N3;        // threads per CTA
    static constexpr int M1 = 16 * N3;       // N / N1
    static constexpr int PITCH = N3 + 1;
    static constexpr int TILE = 256 * PITCH;
    static constexpr int HALVES = N3 / 16;   // lanes cooperating on one tile row in stage 3
    // stage-3 ownership: (row, half) of thread t
    ADT_HD static constexpr int s3_row(int t) { return N3 == 32 ? ((t >> 5) * 16 + (t & 15)) : t; }
    ADT_HD static constexpr int s3_half(int t) { return N3 == 32 ? ((t >> 4) & 1) : 0; }
};

template <class C, class IO = IoF32>
ADT_HD void load_window16(cf* v, int t, const typename IO::elem* __restrict__ xa,
                          const typename IO::elem* __restrict__ xb, long long ws, long long n_in) {
    const bool interior = (ws >= 0) && (ws + C::N <= n_in);
    if (interior) {
        static_for<0, 16>([&](auto K) {
            constexpr int n1 = decltype(K)::value;
            const long long s = ws + n1 * C::M1 + t;
            v[n1] = mk(IO::load(xa + s), xb ? IO::load(xb + s) : 0.0f);
        });
    } else {
        static_for<0, 16>([&](auto K) {
            constexpr int n1 = decltype(K)::value;
            const long long s = ws + n1 * C::M1 + t;
            const bool ok = (s >= 0) && (s < n_in);
            v[n1] = mk(ok ? IO::load(xa + s) : 0.0f, (ok && xb) ? IO::load(xb + s) : 0.0f);
        });
    }
}

template <class C>
ADT_HD void fwd16_stage1(cf* v, int t, const cf* __restrict__ tw1, cf* tile) {
    dft<16, -1>(v);
    apply_powers<16, false, true>(v, tw1[t]);
    cf* col = tile + (t / C::N3) * C::PITCH + (t % C::N3);
    static_for<0, 16>([&](auto K) {
        constexpr int k1 = decltype(K)::value;
        col[k1 * 16 * C::PITCH] = v[brev<16>(k1)];
    });
}

template <class C>
ADT_HD void fwd16_stage2(cf* v, int t, const cf* __restrict__ tw2, cf* tile) {
    const int r2 = t % C::N3;
    cf* col = tile + (t / C::N3) * 16 * C::PITCH + r2;
    static_for<0, 16>([&](auto K) { constexpr int n2 = decltype(K)::value; v[n2] = col[n2 * C::PITCH]; });
    dft<16, -1>(v);
    apply_powers<16, false, true>(v, tw2[r2]);
    static_for<0, 16>([&](auto K) { constexpr int k2 = decltype(K)::value; col[k2 * C::PITCH] = v[brev<16>(k2)]; });
}

// stage 3a: read this thread's share of its row and transform it (result X[j] at v[brev(j)])
template <class C>
ADT_HD void mid16_load_dft(cf* v, int t, const cf* tile) {
    const cf* row = tile + C::s3_row(t) * C::PITCH + C::s3_half(t);
    static_for<0, 16>([&](auto K) { constexpr int m = decltype(K)::value; v[m] = row[m * C::HALVES]; });
    dft<16, -1>(v);
}

// stage 3b, one bin: u = coefS * mine (+ coefX * other)
template <class MaskT>
ADT_HD cf mid16_combine(cf mine, cf other, MaskT cs, cf cx) {
    return cfma(other, cx, mask_mul(mine, cs));
}

// stage 3c: inverse transform and write back in place
template <class C>
ADT_HD void mid16_idft_store(cf* u, int t, cf* tile) {
    dft<16, +1>(u);
    cf* row = tile + C::s3_row(t) * C::PITCH + C::s3_half(t);
    static_for<0, 16>([&](auto K) { constexpr int m = decltype(K)::value; row[m * C::HALVES] = u[brev<16>(m)]; });
}

#if defined(__CUDACC__)
// whole stage 3 on the device (pair exchange by warp shuffle)
template <class C, class MaskT>
__device__ __forceinline__ void mid16_stage3(cf* v, int t, const MaskT* __restrict__ coef_s,
                                             const cf* __restrict__ coef_x, cf* tile) {
    mid16_load_dft<C>(v, t, tile);
    cf u[16];
    constexpr bool kSharedRows = (C::N3 == 32) && (sizeof(MaskT) == sizeof(float));
    const int ci = kSharedRows ? C::s3_row(t) : t;            // coefficient column
    constexpr int CS = kSharedRows ? 256 : C::T;              // coefficient row stride
    const float conj_sign = C::s3_half(t) ? -1.0f : 1.0f;     // real mask: lane B uses conj(coefX)
    static_for<0, 16>([&](auto K) {
        constexpr int j = decltype(K)::value;
        const cf mine = v[brev<16>(j)];
        const MaskT cs = coef_s[j * CS + ci];
        if constexpr (C::N3 == 32) {
            cf other;
            other.x = __shfl_xor_sync(0xffffffffu, mine.x, 16);
            other.y = __shfl_xor_sync(0xffffffffu, mine.y, 16);
            cf cx = coef_x[j * CS + ci];
            if constexpr (kSharedRows) cx.y *= conj_sign;
            u[j] = mid16_combine<MaskT>(mine, other, cs, cx);
        } else {
            u[j] = mask_mul(mine, cs);
        }
    });
    mid16_idft_store<C>(u, t, tile);
}
#endif

template <class C>
ADT_HD void inv16_stage2(cf* v, int t, const cf* __restrict__ tw2, cf* tile) {
    const int r2 = t % C::N3;
    cf* col = tile + (t / C::N3) * 16 * C::PITCH + r2;
    static_for<0, 16>([&](auto K) { constexpr int k2 = decltype(K)::value; v[k2] = col[k2 * C::PITCH]; });
    twiddle_idft<16, 1>(v, tw2[r2]);
    static_for<0, 16>([&](auto K) { constexpr int n2 = decltype(K)::value; col[n2 * C::PITCH] = v[brev<16>(n2)]; });
}

// on return v[brev(n1)] = z[n1*M1 + t]
template <class C>
ADT_HD void inv16_stage1(cf* v, int t, const cf* __restrict__ tw1, const cf* tile) {
    const cf* col = tile + (t / C::N3) * C::PITCH + (t % C::N3);
    static_for<0, 16>([&](auto K) { constexpr int k1 = decltype(K)::value; v[k1] = col[k1 * 16 * C::PITCH]; });
    twiddle_idft<16, 1>(v, tw1[t]);
}

template <class C, class IO, bool SHAPED>
ADT_HD void store_slice16_impl(const cf* v, int t, typename IO::elem* __restrict__ ya,
                               typename IO::elem* __restrict__ yb, long long m0, const FirGeom& g,
                               const FirShape& shape) {
    const long long room = g.n_out - m0;
    const unsigned lim = (unsigned)(room < (long long)g.hop ? (room < 0 ? 0 : room) : g.hop);
    typename IO::elem* pa = ya + (m0 - g.n0) + t;
    typename IO::elem* pb = yb ? yb + (m0 - g.n0) + t : nullptr;
    const int jt = t - g.n0;
    static_for<0, 16>([&](auto K) {
        constexpr int n1 = decltype(K)::value;
        constexpr int off = n1 * C::M1;
        const cf z = v[brev<16>(n1)];
        const bool ok = (unsigned)(jt + off) < lim;
        if (ok) IO::store(pa + off, (SHAPED ? fir_shape(shape, z.x) : z.x));
        if (ok && pb) IO::store(pb + off, (SHAPED ? fir_shape(shape, z.y) : z.y));
    });
}

template <class C, class IO = IoF32, bool SHAPED = false>
ADT_HD void store_slice16(const cf* v, int t, typename IO::elem* __restrict__ ya, typename IO::elem* __restrict__ yb,
                          long long m0, const FirGeom& g, const FirShape& shape) {
    store_slice16_impl<C, IO, SHAPED>(v, t, ya, yb, m0, g, shape);
}

}  // namespace adt
