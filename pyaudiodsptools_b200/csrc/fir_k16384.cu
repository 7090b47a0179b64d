// fir_k16384.cu — kernels of the N = 16384 (512 threads, 1 CTA/SM) transform (its own translation unit: sizes compile in parallel).
#define ADT_FIR_VARIANT_IMPL
#include "fir_variants.cuh"

namespace adt {
const FirVariant* fir_variant_p32_16384() {
    static const FirVariant v = make_variant32<FirCfg<16, 32>, 1, true>("p32");   // persistent loop: +1.5 % here
    return &v;
}
}  // namespace adt
