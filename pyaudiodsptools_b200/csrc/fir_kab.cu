// fir_kab.cu — kernels of the A/B family "p16" (16 points per thread; make AB=1) transform (its own translation unit: sizes compile in parallel).
#define ADT_FIR_VARIANT_IMPL
#include "fir_variants.cuh"

namespace adt {
#if ADT_AB_VARIANTS
const FirVariant* fir_variant_p16_4096() {
    static const FirVariant v = make_variant16<Fir16Cfg<16>, 4>("p16");
    return &v;
}
const FirVariant* fir_variant_p16_8192() {
    static const FirVariant v = make_variant16<Fir16Cfg<32>, 2>("p16");
    return &v;
}
#else
const FirVariant* fir_variant_p16_4096() { return nullptr; }
const FirVariant* fir_variant_p16_8192() { return nullptr; }
#endif
}  // namespace adt
