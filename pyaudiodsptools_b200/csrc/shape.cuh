// shape.cuh — the reference's pointwise wave-shapers, usable as a store epilogue of
// the FIR kernels and by a standalone pointwise kernel.
//   kind 1: CreateSaturator.apply   pyAudioDspTools/EffectSaturator.py:41-48
//   kind 2: CreateSoftClipper.apply pyAudioDspTools/EffectSoftClipper.py:37-44
// float32 arithmetic in the reference's operation order with individually rounded operations (no FMA
// contraction), so the saturator is bit-identical to numpy; the soft clipper differs only by powf's last ulps.
#pragma once
#include "fft_core.cuh"

namespace adt {

struct ShapeParams {
    int kind;      // 0 none, 1 saturator, 2 soft clipper
    int mode;      // saturator exponent: 1 'hard', 2 'soft'
    float p0;      // saturator: s = 10^(threshold_dB/20)        soft clipper: drive + 1
    float p1;      // saturator: 1 - s (rounded from double)
    float p2;      // saturator: (s + 1) / 2
    float p3;      // saturator: makeup gain 10^(dB/20)
};

// FAST = false: IEEE divisions, bit-identical to numpy (standalone kernel).
// FAST = true : the two divisions become a reciprocal multiply and __fdividef (<= 2 ulp on the result);
//               used by the fused FIR store epilogue, where two IEEE divisions per sample would cost more
//               than the whole filter (measured 2.95 ms vs 1.07 ms).
template <bool FAST = false>
ADT_HD float shape_apply(const ShapeParams& sp, float x) {
#if defined(__CUDA_ARCH__)
    if (sp.kind == 1) {
        const bool neg = x < 0.0f;
        float a = fabsf(x);
        if (a > sp.p0) {
            const float t = __fsub_rn(a, sp.p0);
            float u = FAST ? t * __frcp_rn(sp.p1) : __fdiv_rn(t, sp.p1);
            if (sp.mode == 2) u = __fmul_rn(u, u);
            const float q = __fadd_rn(1.0f, u);
            a = __fadd_rn(sp.p0, FAST ? __fdividef(t, q) : __fdiv_rn(t, q));
        }
        if (a > 1.0f) a = sp.p2;
        return __fmul_rn(sp.p3, neg ? -a : a);
    }
    if (sp.kind == 2) {
        const bool neg = x < 0.0f;
        const float a = fminf(fabsf(x), 1.0f);
        const float b = fabsf(__fsub_rn(a, 1.0f));
        const float y = __fadd_rn(FAST ? -__powf(b, sp.p0) : -powf(b, sp.p0), 1.0f);
        return neg ? -y : y;
    }
#endif
    return x;
}

}  // namespace adt
