// fir_tables.h — host-side construction of the twiddle tables and of the
// kernel-order spectral mask for one FirCfg.  Plain C++ (no CUDA), shared by
// the library (adt_api.cu) and by the CPU emulation check (tests/emu/).
#pragma once
#include <cmath>
#include <vector>

#include "fft_core.cuh"

namespace adt {

// tw1[r] = exp(-2*pi*i * r / N), r in [0, M1): base of the stage-1 twiddles W_N^(r*k1)
template <class C>
inline std::vector<cf> build_tw1() {
    std::vector<cf> tw(C::M1);
    for (int r = 0; r < C::M1; ++r) {
        const double a = -2.0 * M_PI * (double)r / (double)C::N;
        tw[r] = mk((float)std::cos(a), (float)std::sin(a));
    }
    return tw;
}

// tw2[lane] = exp(-2*pi*i * lane / M1): base of the stage-2 twiddles W_M1^(lane*k2)
template <class C>
inline std::vector<cf> build_tw2() {
    std::vector<cf> tw(32);
    for (int lane = 0; lane < 32; ++lane) {
        const double a = -2.0 * M_PI * (double)lane / (double)C::M1;
        tw[lane] = mk((float)std::cos(a), (float)std::sin(a));
    }
    return tw;
}

// Frequency index held by thread t at register k3 after forward stage 3.
template <class C>
inline int freq_index(int t, int k3) {
    const int row = C::stage3_row(t);
    const int k1 = row / C::N2, k2 = row % C::N2;
    return k1 + C::N1 * k2 + C::T * k3;
}

// mask_natural: N complex values (interleaved re,im), natural frequency order,
// WITHOUT the 1/N of the inverse transform.  Output is in kernel order
// [k3*T + t], scaled by 1/N; `real_only` keeps just the real part (1 float
// per bin) for zero-phase masks.
template <class C>
inline std::vector<float> permute_mask(const float* mask_natural, bool real_only) {
    const double inv = 1.0 / (double)C::N;
    std::vector<float> out((size_t)C::N * (real_only ? 1 : 2));
    for (int k3 = 0; k3 < 32; ++k3)
        for (int t = 0; t < C::T; ++t) {
            const int k = freq_index<C>(t, k3);
            const size_t o = (size_t)k3 * C::T + t;
            if (real_only) {
                out[o] = (float)(mask_natural[2 * k] * inv);
            } else {
                out[2 * o] = (float)(mask_natural[2 * k] * inv);
                out[2 * o + 1] = (float)(mask_natural[2 * k + 1] * inv);
            }
        }
    return out;
}

// Cluster transform (fir_cluster.cuh): CTA `rank` owns the sub-transforms k1 in [rank*K1L, (rank+1)*K1L); its
// slice of the mask is stored at [(rank*32 + k3)*T + t] for the bin  k1 + N1*k2 + N1*N2*k3  that thread t holds in
// register k3 after forward stage 3.
template <class C>
inline std::vector<float> permute_mask_cluster(const float* mask_natural, bool real_only) {
    const double inv = 1.0 / (double)C::N;
    std::vector<float> out((size_t)C::N * (real_only ? 1 : 2));
    for (int rank = 0; rank < C::CS; ++rank)
        for (int k3 = 0; k3 < 32; ++k3)
            for (int t = 0; t < C::T; ++t) {
                const int row = C::stage3_row(t);
                const int k1 = rank * C::K1L + row / C::N2, k2 = row % C::N2;
                const int k = k1 + C::N1 * k2 + C::N1 * C::N2 * k3;
                const size_t o = ((size_t)rank * 32 + k3) * C::T + t;
                if (real_only) {
                    out[o] = (float)(mask_natural[2 * k] * inv);
                } else {
                    out[2 * o] = (float)(mask_natural[2 * k] * inv);
                    out[2 * o + 1] = (float)(mask_natural[2 * k + 1] * inv);
                }
            }
    return out;
}

// ---- 16-points-per-thread variant (fft_core16.cuh) ---------------------------
template <class C>
inline std::vector<cf> build16_tw1() {
    std::vector<cf> tw(C::M1);
    for (int r = 0; r < C::M1; ++r) {
        const double a = -2.0 * M_PI * (double)r / (double)C::N;
        tw[r] = mk((float)std::cos(a), (float)std::sin(a));
    }
    return tw;
}
template <class C>
inline std::vector<cf> build16_tw2() {
    std::vector<cf> tw(C::N3);
    for (int r2 = 0; r2 < C::N3; ++r2) {
        const double a = -2.0 * M_PI * (double)r2 / (double)C::M1;
        tw[r2] = mk((float)std::cos(a), (float)std::sin(a));
    }
    return tw;
}
// Stage-3 coefficients.  Complex mask: coefS[j*T + t] (float2), coefX[j*T + t] (float2) per thread.
// Real mask with N3 = 32: both lanes of a row share coefS[j*256 + row] (float) and
// coefX[j*256 + row] = Hd*W32^j (float2; lane B conjugates it).  N3 = 16: coefS only, per thread.
template <class C>
inline void build16_coef(const float* mask_natural, bool real_only, std::vector<float>& coef_s,
                         std::vector<float>& coef_x) {
    const double inv = 1.0 / (double)C::N;
    const bool shared_rows = real_only && C::N3 == 32;
    const int cols = shared_rows ? 256 : C::T;
    coef_s.assign((size_t)16 * cols * (real_only ? 1 : 2), 0.f);
    coef_x.assign(C::N3 == 32 ? (size_t)16 * cols * 2 : 0, 0.f);
    for (int j = 0; j < 16; ++j)
        for (int c = 0; c < cols; ++c) {
            const int row = shared_rows ? c : C::s3_row(c), half = shared_rows ? 0 : C::s3_half(c);
            const int k_lo = row / 16 + 16 * (row % 16) + 256 * j;
            const size_t o = (size_t)j * cols + c;
            double lr = mask_natural[2 * k_lo], li = real_only ? 0.0 : mask_natural[2 * k_lo + 1];
            double sr = lr, si = li;
            if (C::N3 == 32) {
                const int k_hi = k_lo + C::N / 2;
                const double hr = mask_natural[2 * k_hi], hi = real_only ? 0.0 : mask_natural[2 * k_hi + 1];
                sr = lr + hr; si = li + hi;
                const double dr = lr - hr, di = li - hi;
                const double ang = (half ? +1.0 : -1.0) * 2.0 * M_PI * (double)j / 32.0;
                const double wr = std::cos(ang), wi = std::sin(ang);
                coef_x[2 * o] = (float)((dr * wr - di * wi) * inv);
                coef_x[2 * o + 1] = (float)((dr * wi + di * wr) * inv);
            }
            if (real_only) {
                coef_s[o] = (float)(sr * inv);
            } else {
                coef_s[2 * o] = (float)(sr * inv);
                coef_s[2 * o + 1] = (float)(si * inv);
            }
        }
}

}  // namespace adt
