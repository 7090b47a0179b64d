// fir_k8192.cu — kernels of the N = 8192 (256 threads, 2 CTAs/SM; the headline kernel) transform (its own translation unit: sizes compile in parallel).
#define ADT_FIR_VARIANT_IMPL
#define ADT_FIR_PINGPONG_IMPL
#include "fir_pingpong.cuh"
#include "fir_variants.cuh"

// resident CTAs per SM the kernels are compiled for (__launch_bounds__ -> register budget)
#ifndef ADT_CTAS_8192
#define ADT_CTAS_8192 2
#endif

namespace adt {
const FirVariant* fir_variant_p32_8192() {
    static const FirVariant v = make_variant32<FirCfg<16, 16>, ADT_CTAS_8192, false, true>("p32");
    return &v;
}
}  // namespace adt
