// CPU emulation of one FIR block of the CUDA kernel (development check).
//
// NOT a product path and not an oracle: it compiles the very same
// __host__ __device__ phase functions the kernel runs (fft_core.cuh) with g++
// and executes them with a loop over thread ids, the shared-memory tile being
// a plain array.  tests/test_emu_fft.py uses it to validate the index algebra
// of the decomposition against numpy without a GPU.
#include <cstring>
#include <vector>

#include "../../pyaudiodsptools_b200/csrc/fft_core16.cuh"
#include "../../pyaudiodsptools_b200/csrc/fir_tables.h"

using namespace adt;

template <class C, class MaskT>
static void run_block(const float* xa, const float* xb, long long n_in, long long ws,
                      const float* mask_natural, bool real_mask, float* z) {
    std::vector<cf> tw1 = build_tw1<C>(), tw2 = build_tw2<C>();
    std::vector<float> mk = permute_mask<C>(mask_natural, real_mask);
    std::vector<cf> tile(C::TILE);
    std::vector<cf> regs((size_t)C::T * 32);
    for (int t = 0; t < C::T; ++t) load_window<C>(&regs[t * 32], t, xa, xb, ws, n_in);
    for (int t = 0; t < C::T; ++t) fwd_stage1<C>(&regs[t * 32], t, tw1.data(), tile.data());
    for (int t = 0; t < C::T; ++t) fwd_stage2<C>(&regs[t * 32], t, tw2.data(), tile.data());
    for (int t = 0; t < C::T; ++t) mid_stage3<C, MaskT>(&regs[t * 32], t, (const MaskT*)mk.data(), tile.data());
    for (int t = 0; t < C::T; ++t) inv_stage2<C>(&regs[t * 32], t, tw2.data(), tile.data());
    for (int t = 0; t < C::T; ++t) inv_stage1<C>(&regs[t * 32], t, tw1.data(), tile.data());
    // dump the whole circular result z[n], n = n1*M1 + r
    for (int t = 0; t < C::T; ++t)
        for (int u = 0; u < C::B1; ++u)
            for (int n1 = 0; n1 < C::N1; ++n1) {
                const int n = n1 * C::M1 + t + u * C::T;
                const cf v = regs[t * 32 + u * C::N1 + brev<C::N1>(n1)];
                z[2 * n] = v.x;
                z[2 * n + 1] = v.y;
            }
}

extern "C" int emu_fir_block(int n, const float* xa, const float* xb, long long n_in, long long ws,
                             const float* mask_natural, int real_mask, float* z) {
#define CASE(N1, N2)                                                                              \
    if (n == N1 * N2 * 32) {                                                                      \
        if (real_mask) run_block<FirCfg<N1, N2>, float>(xa, xb, n_in, ws, mask_natural, true, z); \
        else run_block<FirCfg<N1, N2>, cf>(xa, xb, n_in, ws, mask_natural, false, z);             \
        return 0;                                                                                 \
    }
    CASE(8, 8)
    CASE(16, 8)
    CASE(16, 16)
    CASE(16, 32)
    CASE(32, 32)
#undef CASE
    return -1;
}

// plain R-point DFT through the register template, for unit-testing dft<R,DIR>
extern "C" int emu_dft(int r, int dir, float* v /* 2*r floats in/out, natural order */) {
    cf a[32], b[32];
    for (int i = 0; i < r; ++i) a[i] = mk(v[2 * i], v[2 * i + 1]);
#define D(R)                                                   \
    if (r == R) {                                              \
        if (dir < 0) dft<R, -1>(a); else dft<R, +1>(a);        \
        for (int k = 0; k < R; ++k) b[k] = a[brev<R>(k)];      \
    }
    D(2) D(4) D(8) D(16) D(32)
#undef D
    for (int i = 0; i < r; ++i) { v[2 * i] = b[i].x; v[2 * i + 1] = b[i].y; }
    return 0;
}

// ---- 16-points-per-thread variant (fft_core16.cuh) ------------------------------------
template <class C, class MaskT>
static void run_block16(const float* xa, const float* xb, long long n_in, long long ws, const float* mask_natural,
                        bool real_mask, float* z) {
    std::vector<cf> tw1 = build16_tw1<C>(), tw2 = build16_tw2<C>();
    std::vector<float> cs, cx;
    build16_coef<C>(mask_natural, real_mask, cs, cx);
    std::vector<cf> tile(C::TILE), regs((size_t)C::T * 16), out((size_t)C::T * 16);
    for (int t = 0; t < C::T; ++t) load_window16<C>(&regs[t * 16], t, xa, xb, ws, n_in);
    for (int t = 0; t < C::T; ++t) fwd16_stage1<C>(&regs[t * 16], t, tw1.data(), tile.data());
    for (int t = 0; t < C::T; ++t) fwd16_stage2<C>(&regs[t * 16], t, tw2.data(), tile.data());
    for (int t = 0; t < C::T; ++t) mid16_load_dft<C>(&regs[t * 16], t, tile.data());
    for (int t = 0; t < C::T; ++t)          // the pair exchange is a warp shuffle (lane ^ 16) on the device
        for (int j = 0; j < 16; ++j) {
            const int bj = (j & 1) << 3 | (j & 2) << 1 | (j & 4) >> 1 | (j & 8) >> 3;
            const cf mine = regs[t * 16 + bj];
            const bool shared = C::N3 == 32 && sizeof(MaskT) == sizeof(float);   // as in mid16_stage3
            const int ci = shared ? C::s3_row(t) : t, CS = shared ? 256 : C::T;
            const MaskT s = ((const MaskT*)cs.data())[j * CS + ci];
            if (C::N3 == 32) {
                const cf other = regs[(t ^ 16) * 16 + bj];
                cf x = ((const cf*)cx.data())[j * CS + ci];
                if (shared && C::s3_half(t)) x.y = -x.y;
                out[t * 16 + j] = mid16_combine<MaskT>(mine, other, s, x);
            } else {
                out[t * 16 + j] = mask_mul(mine, s);
            }
        }
    for (int t = 0; t < C::T; ++t) mid16_idft_store<C>(&out[t * 16], t, tile.data());
    for (int t = 0; t < C::T; ++t) inv16_stage2<C>(&regs[t * 16], t, tw2.data(), tile.data());
    for (int t = 0; t < C::T; ++t) inv16_stage1<C>(&regs[t * 16], t, tw1.data(), tile.data());
    for (int t = 0; t < C::T; ++t)
        for (int n1 = 0; n1 < 16; ++n1) {
            const int n = n1 * C::M1 + t;
            const cf v = regs[t * 16 + brev<16>(n1)];
            z[2 * n] = v.x;
            z[2 * n + 1] = v.y;
        }
}

extern "C" int emu_fir16_block(int n, const float* xa, const float* xb, long long n_in, long long ws,
                               const float* mask_natural, int real_mask, float* z) {
    if (n == 8192) {
        if (real_mask) run_block16<Fir16Cfg<32>, float>(xa, xb, n_in, ws, mask_natural, true, z);
        else run_block16<Fir16Cfg<32>, cf>(xa, xb, n_in, ws, mask_natural, false, z);
        return 0;
    }
    if (n == 4096) {
        if (real_mask) run_block16<Fir16Cfg<16>, float>(xa, xb, n_in, ws, mask_natural, true, z);
        else run_block16<Fir16Cfg<16>, cf>(xa, xb, n_in, ws, mask_natural, false, z);
        return 0;
    }
    return -1;
}
