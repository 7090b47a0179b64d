// fft_core.cuh — register-resident radix-16/32 FFT building blocks and the
// per-thread phases of the fused overlap-save FIR block
//     window -> FFT_N -> x mask -> IFFT_N -> slice
// for TWO real channels packed as one complex signal (re = channel a,
// im = channel b; valid because the taps are real, so no split/merge pass).
//
// Replaces the per-chunk numpy pipeline of the reference
//     pyAudioDspTools/EffectFFTFilter.py:143-151   (concatenate, fft, *mask, ifft, slice)
//     pyAudioDspTools/EffectEQ3BandFFT.py:175-211  (1 fft + 3 masked iffts + mix)
// by one N-point complex transform pair per (block, channel pair); the
// equivalence is SURVEY.md Appendix A.3/A.4 and DESIGN.md §2.
//
// Decomposition  N = N1 * N2 * 32,  T = N/32 threads, 32 complex points per
// thread held in registers through every stage.  M1 = N/N1 = N2*32.
//
//   stage 1 (thread r, B1=32/N1 butterflies): DFT_N1 over n1 of x[n1*M1 + r],
//            times W_N^(r*k1)                      -> A[k1][r]
//   stage 2 (thread (k1, r2=lane), B2=32/N2):  DFT_N2 over n2 of A[k1][n2*32+r2],
//            times W_M1^(r2*k2)                    -> B[k1][k2][r2]
//   stage 3 (thread = (k1, k2), one row of the tile): DFT_32 over r2
//            -> X[k1 + N1*k2 + T*k3]   (natural order, stride T across k3)
//            Rows are assigned so that warp w owns exactly the rows its own
//            stage-2 butterflies wrote (k1 in {w, w+WARPS, ..}): the exchanges
//            around stage 3 are warp-local and need only __syncwarp().
//   mask multiply in registers, then the mirror image (DIT, conjugate
//   twiddles applied on input) back to y[n1*M1 + r].
//
// Shared memory holds one [T rows][33] float2 tile; every exchange is done
// IN PLACE (each thread writes exactly the addresses it read last), so only
// read-after-write synchronisation is needed: two block barriers (around the
// stage-1 transposition) and two warp barriers (around stage 3).  Row pitch 33
// makes both the row-wise (stage 3) and column-wise (stages 1,2) accesses
// bank-conflict free with immediate offsets.
//
// All functions are __host__ __device__: tests/emu/ runs the very same code
// on the CPU with a loop over thread ids to validate the index algebra
// without a GPU (development check only — never a product path).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define ADT_HD __host__ __device__ __forceinline__
#define ADT_ALIGN8 __align__(8)
#else
#define ADT_HD inline __attribute__((always_inline))
#define ADT_ALIGN8 alignas(8)
#endif

namespace adt {

// A complex number is a (re, im) pair in an aligned 64-bit register pair.  On
// sm_100a the arithmetic below is PACKED: FADD2 / FMUL2 / FFMA2 (add/mul/fma
// .f32x2) do both halves in one issue slot, and ptxas folds the half swaps,
// scalar broadcasts and per-half sign flips written here with make pairs into
// operand modifiers (.LO_HI, .F32, .NP), so e.g. a complex multiply is 2
// instructions and a radix-2 butterfly with a general twiddle is 3 (8 scalar).
// The host versions are the same formulas in scalar float (CPU emulation check).
#if defined(__CUDACC__)
typedef float2 cf;
#else
struct ADT_ALIGN8 cf {
    float x, y;
};
#endif

ADT_HD cf mk(float x, float y) {
    cf r;
    r.x = x;
    r.y = y;
    return r;
}
ADT_HD cf cadd(cf a, cf b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return mk(a.x + b.x, a.y + b.y);
#endif
}
ADT_HD cf csub(cf a, cf b) { return cadd(a, mk(-b.x, -b.y)); }
// acc + s * w   (s real scalar, broadcast)
ADT_HD cf fma_s(float s, cf w, cf acc) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(mk(s, s), w, acc);
#else
    return mk(s * w.x + acc.x, s * w.y + acc.y);
#endif
}
ADT_HD cf mul_s(float s, cf w) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(mk(s, s), w);
#else
    return mk(s * w.x, s * w.y);
#endif
}
// pair * broadcast scalar (+ pair), with the PAIR as the first operand: ptxas folds a half swap and a
// per-half sign flip of the first operand into the instruction (SASS `-R.F32x2.LO_HI.NP`), but only there —
// written the other way round (scalar first, swizzled pair second) it materialises the swizzled pair with a
// MOV + FADD per use (round 1: 90 FADD + ~130 MOV per thread in the headline kernel).
ADT_HD cf pmul_s(cf w, float s) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(w, mk(s, s));
#else
    return mk(w.x * s, w.y * s);
#endif
}
ADT_HD cf pfma_s(cf w, float s, cf acc) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(w, mk(s, s), acc);
#else
    return mk(w.x * s + acc.x, w.y * s + acc.y);
#endif
}
// a*b = b.x*(a.x, a.y) + b.y*(-a.y, a.x)        (2 packed instructions)
ADT_HD cf cmul(cf a, cf b) { return pfma_s(a, b.x, pmul_s(mk(-a.y, a.x), b.y)); }
// a * conj(b) = b.x*(a.x, a.y) + b.y*(a.y, -a.x)
ADT_HD cf cmulc(cf a, cf b) { return pfma_s(a, b.x, pmul_s(mk(a.y, -a.x), b.y)); }
// acc + a*b
ADT_HD cf cfma(cf a, cf b, cf acc) { return pfma_s(a, b.x, pfma_s(mk(-a.y, a.x), b.y, acc)); }

ADT_HD cf mask_mul(cf a, float h) { return mul_s(h, a); }  // zero-phase (real) mask
ADT_HD cf mask_mul(cf a, cf h) { return cmul(a, h); }
// acc + a*h
ADT_HD cf mask_fma(cf a, float h, cf acc) { return fma_s(h, a, acc); }
ADT_HD cf mask_fma(cf a, cf h, cf acc) { return cfma(a, h, acc); }

// cos / sin of 2*pi*q/32 as literals (constexpr trig is not available).
ADT_HD constexpr float cos32(int q) {
    switch (q & 31) {
        case 0: return 1.0f;
        case 1: case 31: return 0.98078528040323043f;
        case 2: case 30: return 0.92387953251128674f;
        case 3: case 29: return 0.83146961230254524f;
        case 4: case 28: return 0.70710678118654752f;
        case 5: case 27: return 0.55557023301960218f;
        case 6: case 26: return 0.38268343236508977f;
        case 7: case 25: return 0.19509032201612825f;
        case 8: case 24: return 0.0f;
        case 9: case 23: return -0.19509032201612825f;
        case 10: case 22: return -0.38268343236508977f;
        case 11: case 21: return -0.55557023301960218f;
        case 12: case 20: return -0.70710678118654752f;
        case 13: case 19: return -0.83146961230254524f;
        case 14: case 18: return -0.92387953251128674f;
        case 15: case 17: return -0.98078528040323043f;
        default: return -1.0f;  // 16
    }
}
ADT_HD constexpr float sin32(int q) { return cos32(q - 8); }

template <int R>
ADT_HD constexpr int brev(int i) {
    int r = 0;
    for (int b = 1, c = R >> 1; c; b <<= 1, c >>= 1)
        if (i & b) r |= c;
    return r;
}

// DIT butterfly  a' = a + W b,  b' = a - W b,  W = exp(DIR * 2*pi*i * Q/32).
// DIR = -1: forward transform, +1: inverse.  Q in [0, 16).
// General twiddle: a' = a + wr*b + wi*(i b) (2 FFMA2), b' = 2a - a' (1 FFMA2).
template <int Q, int DIR>
ADT_HD void bfly(cf& a, cf& b) {
    if constexpr (Q == 0) {
        const cf t = b;
        b = csub(a, t);
        a = cadd(a, t);
    } else if constexpr (Q == 8) {
        // W = i*DIR :  W b = (-DIR*b.y, DIR*b.x)
        const cf t = (DIR < 0) ? mk(b.y, -b.x) : mk(-b.y, b.x);
        b = csub(a, t);
        a = cadd(a, t);
    } else {
        constexpr float wr = cos32(Q);
        constexpr float wi = (DIR < 0) ? -sin32(Q) : sin32(Q);
        // W b = wr*(b.x, b.y) + wi*(-b.y, b.x): constants as broadcast immediates, data swizzled
        const cf a1 = fma_s(wr, b, fma_s(wi, mk(-b.y, b.x), a));
        b = fma_s(2.0f, a, mk(-a1.x, -a1.y));
        a = a1;
    }
}

template <int R, int DIR, int HALF, int I>
ADT_HD void dft_step(cf* v) {
    if constexpr (I < R) {
        if constexpr ((I & HALF) == 0) {
            constexpr int j = I & (HALF - 1);
            constexpr int q = j * 16 / HALF;
            bfly<q, DIR>(v[brev<R>(I)], v[brev<R>(I + HALF)]);
        }
        dft_step<R, DIR, HALF, I + 1>(v);
    }
}
template <int R, int DIR, int HALF>
ADT_HD void dft_stage(cf* v) {
    if constexpr (HALF < R) {
        dft_step<R, DIR, HALF, 0>(v);
        dft_stage<R, DIR, HALF * 2>(v);
    }
}
// In-place R-point DFT (R = 2..32) on a register array.
// Input x[n] at v[n]; output X[k] at v[brev<R>(k)].
template <int R, int DIR>
ADT_HD void dft(cf* v) {
    dft_stage<R, DIR, 1>(v);
}

// compile-time loop: f receives an IC<i>; read the index as decltype(K)::value
template <int I>
struct IC {
    static constexpr int value = I;
};
template <int B, int E, class F>
ADT_HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}
ADT_HD constexpr int hibit(int k) {  // largest power of two <= k
    int h = 1;
    while (h * 2 <= k) h *= 2;
    return h;
}

// Multiply element k (k = 1..R-1) by w^k (or conj(w^k)) for a unit-modulus base w.
// Powers come from a small basis kept in registers — b[j] = w^j (j = 1..3) and
// q[a] = w^(4a) — and every other power is the just-in-time product q[a]*b[j],
// so only R/4 + 2 pairs are live instead of R - 1 (matters at 80 registers).
// Product-tree depth is <= log2(R) + 1, i.e. a few ulp of error on the twiddle.
// BREV: element k lives at v[brev<R>(k)] (output of dft<>) instead of v[k].
// NB butterflies stored back to back (v + u*R) share the same base: powers are formed once.
template <int R, bool CONJ, bool BREV, int NB = 1>
ADT_HD void apply_powers(cf* v, cf w1) {
    static_assert(R >= 8 && R % 4 == 0, "radix");
    constexpr int A = R / 4;
    cf b[4], q[A];
    b[1] = w1;
    b[2] = cmul(w1, w1);
    b[3] = cmul(b[2], w1);
    q[1] = cmul(b[2], b[2]);
    static_for<2, A>([&](auto K) {
        constexpr int a = decltype(K)::value;
        constexpr int hi = (a & (a - 1)) == 0 ? a / 2 : hibit(a);
        q[a] = cmul(q[hi], q[a - hi]);
    });
    static_for<1, R>([&](auto K) {
        constexpr int k = decltype(K)::value;
        constexpr int a = k / 4, j = k % 4;
        constexpr int idx = BREV ? brev<R>(k) : k;
        cf p;
        if constexpr (a == 0)
            p = b[j];
        else if constexpr (j == 0)
            p = q[a];
        else
            p = cmul(q[a], b[j]);
        static_for<0, NB>([&](auto U) {
            constexpr int o = decltype(U)::value * R + idx;
            v[o] = CONJ ? cmulc(v[o], p) : cmul(v[o], p);
        });
    });
}

// acc + x * conj(p)
ADT_HD cf cfmac(cf x, cf p, cf acc) { return pfma_s(x, p.x, pfma_s(mk(x.y, -x.x), p.y, acc)); }

// v[k] *= conj(w1^k) followed by the inverse R-point DFT, with the twiddles FOLDED into the first
// radix-2 stage: that stage pairs (v[j], v[j+R/2]) with unit twiddle, so
//     A = v[j]*conj(p_j);  a' = A + v[j+R/2]*conj(p_{j+R/2});  b' = 2A - a'
// is 5 packed instructions instead of 6 (3 instead of 4 for j = 0).  NB butterflies share the powers.
template <int R, int NB>
ADT_HD void twiddle_idft(cf* v, cf w1) {
    static_assert(R >= 8 && R % 4 == 0, "radix");
    constexpr int A = R / 4;
    cf b[4], q[A];
    b[1] = w1;
    b[2] = cmul(w1, w1);
    b[3] = cmul(b[2], w1);
    q[1] = cmul(b[2], b[2]);
    static_for<2, A>([&](auto K) {
        constexpr int a = decltype(K)::value;
        constexpr int hi = (a & (a - 1)) == 0 ? a / 2 : hibit(a);
        q[a] = cmul(q[hi], q[a - hi]);
    });
    auto power = [&](auto K) {
        constexpr int k = decltype(K)::value;
        constexpr int a = k / 4, j = k % 4;
        if constexpr (a == 0)
            return b[j];
        else if constexpr (j == 0)
            return q[a];
        else
            return cmul(q[a], b[j]);
    };
    static_for<0, R / 2>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const cf p_hi = power(IC<j + R / 2>{});
        cf p_lo = p_hi;
        if constexpr (j > 0) p_lo = power(IC<j>{});
        static_for<0, NB>([&](auto U) {
            constexpr int o = decltype(U)::value * R;
            cf a0 = v[o + j];
            if constexpr (j > 0) a0 = cmulc(a0, p_lo);
            const cf a1 = cfmac(v[o + j + R / 2], p_hi, a0);
            v[o + j + R / 2] = fma_s(2.0f, a0, mk(-a1.x, -a1.y));
            v[o + j] = a1;
        });
    });
    static_for<0, NB>([&](auto U) { dft_stage<R, +1, 2>(v + decltype(U)::value * R); });
}

// y = IDFT_32(mask .* x) with the mask folded into the first radix-2 stage; x[k] is read from
// xs[brev<32>(k)] (the layout dft<32,-1> leaves), the result c[r] is at y[brev<32>(r)].
template <class MaskT, class LoadMask>
ADT_HD void masked_idft32(const cf* xs, cf* y, LoadMask&& load_mask) {
    static_for<0, 16>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const MaskT h_lo = load_mask(IC<j>{}), h_hi = load_mask(IC<j + 16>{});
        const cf a0 = mask_mul(xs[brev<32>(j)], h_lo);
        const cf a1 = mask_fma(xs[brev<32>(j + 16)], h_hi, a0);
        y[j + 16] = fma_s(2.0f, a0, mk(-a1.x, -a1.y));
        y[j] = a1;
    });
    dft_stage<32, +1, 2>(y);
}

// ---------------------------------------------------------------------------
// configuration of one transform size
// ---------------------------------------------------------------------------
template <int N1_, int N2_>
struct FirCfg {
    static constexpr int N1 = N1_, N2 = N2_, N3 = 32;
    static constexpr int N = N1 * N2 * 32;
    static constexpr int T = N1 * N2;        // threads per CTA (= rows of the tile)
    static constexpr int M1 = N2 * 32;       // size of the sub-transform after stage 1
    static constexpr int B1 = 32 / N1;       // stage-1 butterflies per thread
    static constexpr int B2 = 32 / N2;       // stage-2 butterflies per thread
    static constexpr int PITCH = 33;         // float2 per tile row
    static constexpr int TILE = T * PITCH;   // float2 elements of shared memory
    static constexpr int WARPS = T / 32;
    // tile row (k1*N2 + k2) owned by thread t in stage 3: the rows written by t's own warp in stage 2
    ADT_HD static constexpr int stage3_row(int t) {
        return (((t >> 5) + ((t & 31) / N2) * WARPS) * N2) + ((t & 31) % N2);
    }
};

// What one FIR block needs to know (see DESIGN.md §3 for the derivation).
struct FirShape {   // store epilogue (shape.cuh ShapeParams); kind 0 = none
    int kind, mode;
    float p0, p1, p2, p3;
};

struct FirGeom {
    int hop;         // outputs produced per block
    int n0;          // first kept index of the N-point circular result
    int back;        // window of block b starts at stream index b*hop - back + in_shift
    long long in_shift;
    long long n_in;  // valid input samples per row (others read as 0)
    long long n_out; // outputs wanted per row
    long long in_pitch, out_pitch;  // row pitches in floats
};

// Sample formats at the HBM boundary.  F32: the reference's float32 chunks.  I16: 16-bit PCM fused
// into the load/store — x = int16/32768 on load (Utility.py:236-237, MonoWavToNumpyFloat) and
// int16(y*32767) with C truncation on store (Utility.py:306, NumpyFloatToWav); halves HBM bytes.
// Streaming read (window samples are used once per CTA): read-only path, no L1 allocation, so the
// 64 KB windows do not evict the mask / twiddle tables that every CTA re-reads from L1.  Plain (non
// volatile) asm: the loads stay freely schedulable.  Compile-time switch, OFF by default (measured slower).
#ifndef ADT_STORE_PRED
#define ADT_STORE_PRED 0     /* 1: one predicated STG (inline PTX) instead of the branch region around `if (ok) *p = v`: 130 fewer instructions per thread but measured 1.4 % SLOWER on the same box (gpurun_out/r2h) -> off */
#endif
#ifndef ADT_STREAM_LOADS
#define ADT_STREAM_LOADS 0   /* measured on B200: no_allocate loads are 2 % slower than ordinary allocating loads */
#endif
ADT_HD float ld_stream_f32(const float* p) {
#if defined(__CUDA_ARCH__) && ADT_STREAM_LOADS == 1
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#elif defined(__CUDA_ARCH__) && ADT_STREAM_LOADS == 2
    float v;
    asm("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
// Mask table loads (re-read by every CTA): optionally ask L1 to keep them (A/B, ADT_MASK_EVICT_LAST)
#ifndef ADT_MASK_EVICT_LAST
#define ADT_MASK_EVICT_LAST 0
#endif
ADT_HD float ld_mask(const float* p) {
#if defined(__CUDA_ARCH__) && ADT_MASK_EVICT_LAST
    float v;
    asm("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
ADT_HD cf ld_mask(const cf* p) {
#if defined(__CUDA_ARCH__) && ADT_MASK_EVICT_LAST
    cf v;
    asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#else
    return *p;
#endif
}
struct IoF32 {
    typedef float elem;
    ADT_HD static float load(const float* p) { return ld_stream_f32(p); }
    ADT_HD static void store(float* p, float v) { *p = v; }
    // store only when ok: ONE predicated STG instead of the branch region (BSSY / BRA / BSYNC) the compiler
    // builds around `if (ok) *p = v` — 32 such regions per thread in the store phase of the FIR kernel
    ADT_HD static void store_if(bool ok, float* p, float v) {
#if defined(__CUDA_ARCH__) && ADT_STORE_PRED
        // no "memory" clobber: the kernel never reads what it stores, so the compiler stays free to hoist the
        // next butterfly's shared-memory loads above these stores
        asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.global.f32 [%0], %1;\n}" ::"l"(p), "f"(v), "r"((unsigned)ok));
#else
        if (ok) *p = v;
#endif
    }
};
struct IoI16 {
    typedef short elem;
    ADT_HD static float load(const short* p) { return (float)*p * (1.0f / 32768.0f); }
    ADT_HD static void store(short* p, float v) {
        // numpy's float32 -> int16 astype on x86: truncate toward zero to int32, keep the low 16 bits
        *p = (short)(int)(v * 32767.0f);
    }
    ADT_HD static void store_if(bool ok, short* p, float v) {
        if (ok) store(p, v);
    }
};

// ---- phase 0: global -> registers (stage-1 layout) -------------------------
template <class C, class IO = IoF32>
ADT_HD void load_window(cf* v, int t, const typename IO::elem* __restrict__ xa,
                        const typename IO::elem* __restrict__ xb, long long ws, long long n_in) {
    const bool interior = (ws >= 0) && (ws + C::N <= n_in);
    if (interior) {
        static_for<0, C::B1>([&](auto U) {
            static_for<0, C::N1>([&](auto K) {
                constexpr int u = decltype(U)::value, n1 = decltype(K)::value;
                const long long s = ws + n1 * C::M1 + t + u * C::T;
                v[u * C::N1 + n1] = mk(IO::load(xa + s), xb ? IO::load(xb + s) : 0.0f);
            });
        });
    } else {
        static_for<0, C::B1>([&](auto U) {
            static_for<0, C::N1>([&](auto K) {
                constexpr int u = decltype(U)::value, n1 = decltype(K)::value;
                const long long s = ws + n1 * C::M1 + t + u * C::T;
                const bool ok = (s >= 0) && (s < n_in);
                v[u * C::N1 + n1] = mk(ok ? IO::load(xa + s) : 0.0f, (ok && xb) ? IO::load(xb + s) : 0.0f);
            });
        });
    }
}

// interior blocks only: the whole window is inside [0, n_in) — no bounds code in the kernel at all
template <class C, class IO = IoF32>
ADT_HD void load_window_interior(cf* v, int t, const typename IO::elem* __restrict__ xa,
                                 const typename IO::elem* __restrict__ xb, long long ws) {
    static_for<0, C::B1>([&](auto U) {
        static_for<0, C::N1>([&](auto K) {
            constexpr int u = decltype(U)::value, n1 = decltype(K)::value;
            const long long s = ws + n1 * C::M1 + t + u * C::T;
            v[u * C::N1 + n1] = mk(IO::load(xa + s), xb ? IO::load(xb + s) : 0.0f);
        });
    });
}

// ---- phase 1: forward stage 1, write tile ----------------------------------
template <class C>
ADT_HD void fwd_stage1(cf* v, int t, const cf* __restrict__ tw1, cf* tile) {
    const int lane = t & 31;
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        dft<C::N1, -1>(b);
        const int r = t + u * C::T;
        apply_powers<C::N1, false, true>(b, tw1[r]);
        const int row0 = r >> 5;  // n2
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            tile[(k1 * C::N2 + row0) * C::PITCH + lane] = b[brev<C::N1>(k1)];
        });
    });
}

// ---- phase 2: forward stage 2, in place in the tile --------------------------
// The twiddle W_M1^(lane*k2) does not depend on k1, so all B2 butterflies of a thread share it.
template <class C>
ADT_HD void fwd_stage2(cf* v, int t, const cf* __restrict__ tw2, cf* tile) {
    const int lane = t & 31, warp = t >> 5;
    const cf w2 = tw2[lane];  // W_M1^lane; the stage-2 twiddles are its powers
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N2;
        const cf* col = tile + ((warp + u * C::WARPS) * C::N2) * C::PITCH + lane;
        static_for<0, C::N2>([&](auto K) { constexpr int n2 = decltype(K)::value; b[n2] = col[n2 * C::PITCH]; });
        dft<C::N2, -1>(b);
    });
    apply_powers<C::N2, false, true, C::B2>(v, w2);
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N2;
        cf* col = tile + ((warp + u * C::WARPS) * C::N2) * C::PITCH + lane;
        static_for<0, C::N2>([&](auto K) {
            constexpr int k2 = decltype(K)::value;
            col[k2 * C::PITCH] = b[brev<C::N2>(k2)];
        });
    });
}

// ---- phase 3: forward DFT32, mask, inverse DFT32, in place row t ------------
template <class C, class MaskT>
ADT_HD void mid_stage3(cf* v, int t, const MaskT* __restrict__ mask, cf* tile) {
    cf* row = tile + C::stage3_row(t) * C::PITCH;
    static_for<0, 32>([&](auto K) { constexpr int r2 = decltype(K)::value; v[r2] = row[r2]; });
    dft<32, -1>(v);
    cf y[32];
    masked_idft32<MaskT>(v, y, [&](auto K) { return ld_mask(mask + decltype(K)::value * C::T + t); });
    static_for<0, 32>([&](auto K) { constexpr int r2 = decltype(K)::value; row[r2] = y[brev<32>(r2)]; });
}

// ---- phase 4: inverse stage 2 (conj twiddle on input), in place --------------
template <class C>
ADT_HD void inv_stage2(cf* v, int t, const cf* __restrict__ tw2, cf* tile) {
    const int lane = t & 31, warp = t >> 5;
    const cf w2 = tw2[lane];
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N2;
        const cf* col = tile + ((warp + u * C::WARPS) * C::N2) * C::PITCH + lane;
        static_for<0, C::N2>([&](auto K) { constexpr int k2 = decltype(K)::value; b[k2] = col[k2 * C::PITCH]; });
    });
    twiddle_idft<C::N2, C::B2>(v, w2);
    static_for<0, C::B2>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N2;
        cf* col = tile + ((warp + u * C::WARPS) * C::N2) * C::PITCH + lane;
        static_for<0, C::N2>([&](auto K) { constexpr int n2 = decltype(K)::value; col[n2 * C::PITCH] = b[brev<C::N2>(n2)]; });
    });
}

// ---- phase 5: inverse stage 1, results stay in registers --------------------
// On return v[u*N1 + brev(n1)] = z[n1*M1 + t + u*T].
template <class C>
ADT_HD void inv_stage1(cf* v, int t, const cf* __restrict__ tw1, const cf* tile) {
    const int lane = t & 31;
    static_for<0, C::B1>([&](auto U) {
        constexpr int u = decltype(U)::value;
        cf* b = v + u * C::N1;
        const int r = t + u * C::T;
        const int row0 = r >> 5;
        static_for<0, C::N1>([&](auto K) {
            constexpr int k1 = decltype(K)::value;
            b[k1] = tile[(k1 * C::N2 + row0) * C::PITCH + lane];
        });
        twiddle_idft<C::N1, 1>(b, tw1[r]);
    });
}

// store epilogue: identity unless a wave-shaper was attached (uniform branch; shape.cuh has the formulas)
ADT_HD float fir_shape(const FirShape& sh, float v);

// ---- phase 6: registers -> global (only the valid slice) -------------------
// z[n], n = n1*M1 + t + u*T, goes to y[m0 + n - n0] when 0 <= n - n0 < hop and
// the stream index is below n_out.  `lim` = min(hop, n_out - m0) folds both
// upper bounds into one unsigned compare per element.
template <class C, class IO, bool SHAPED, bool ACCUM = false>
ADT_HD void store_slice_impl(const cf* v, int t, typename IO::elem* __restrict__ ya, typename IO::elem* __restrict__ yb,
                             long long m0, const FirGeom& g, const FirShape& shape) {
    const long long room = g.n_out - m0;
    const unsigned lim = (unsigned)(room < (long long)g.hop ? (room < 0 ? 0 : room) : g.hop);
    typename IO::elem* pa = ya + (m0 - g.n0) + t;
    typename IO::elem* pb = yb ? yb + (m0 - g.n0) + t : nullptr;
    const int jt = t - g.n0;
    static_for<0, C::B1>([&](auto U) {
        static_for<0, C::N1>([&](auto K) {
            constexpr int u = decltype(U)::value, n1 = decltype(K)::value;
            constexpr int off = n1 * C::M1 + u * C::T;
            const cf z = v[u * C::N1 + brev<C::N1>(n1)];
            const bool ok = (unsigned)(jt + off) < lim;
            if constexpr (ACCUM) {   // later tap segments of a partitioned filter add to what is already there
                if (ok) pa[off] += z.x;
                if (ok && pb) pb[off] += z.y;
            } else {
                if constexpr (SHAPED) {
                    if (ok) IO::store(pa + off, fir_shape(shape, z.x));
                    if (ok && pb) IO::store(pb + off, fir_shape(shape, z.y));
                } else {
                    IO::store_if(ok, pa + off, z.x);
                    IO::store_if(ok && pb, pb + off, z.y);
                }
            }
        });
    });
}

// SHAPED kernels are separate instantiations, so the default kernels carry no epilogue code at all.
template <class C, class IO = IoF32, bool SHAPED = false, bool ACCUM = false>
ADT_HD void store_slice(const cf* v, int t, typename IO::elem* __restrict__ ya, typename IO::elem* __restrict__ yb,
                        long long m0, const FirGeom& g, const FirShape& shape) {
    store_slice_impl<C, IO, SHAPED, ACCUM>(v, t, ya, yb, m0, g, shape);
}

}  // namespace adt
