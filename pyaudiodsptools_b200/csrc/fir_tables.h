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

}  // namespace adt
