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

ADT_HD float shape_apply(const ShapeParams& sp, float x) {
#if defined(__CUDA_ARCH__)
    if (sp.kind == 1) {
        const bool neg = x < 0.0f;
        float a = fabsf(x);
        if (a > sp.p0) {
            const float t = __fsub_rn(a, sp.p0);
            float u = __fdiv_rn(t, sp.p1);
            if (sp.mode == 2) u = __fmul_rn(u, u);
            a = __fadd_rn(sp.p0, __fdiv_rn(t, __fadd_rn(1.0f, u)));
        }
        if (a > 1.0f) a = sp.p2;
        return __fmul_rn(sp.p3, neg ? -a : a);
    }
    if (sp.kind == 2) {
        const bool neg = x < 0.0f;
        const float a = fminf(fabsf(x), 1.0f);
        const float y = __fadd_rn(-powf(fabsf(__fsub_rn(a, 1.0f)), sp.p0), 1.0f);
        return neg ? -y : y;
    }
#endif
    return x;
}

}  // namespace adt
