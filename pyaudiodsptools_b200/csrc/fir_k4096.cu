// fir_k4096.cu — kernels of the N = 4096 (128 threads, 4 CTAs/SM) transform (its own translation unit: sizes compile in parallel).
#define ADT_FIR_VARIANT_IMPL
#include "fir_variants.cuh"

// resident CTAs per SM the kernels are compiled for (__launch_bounds__ -> register budget)
#ifndef ADT_CTAS_4096
#define ADT_CTAS_4096 4
#endif

namespace adt {
const FirVariant* fir_variant_p32_4096() {
    static const FirVariant v = make_variant32<FirCfg<16, 8>, ADT_CTAS_4096, false, true>("p32");
    return &v;
}
}  // namespace adt
