// fir_k32768.cu — cluster transforms (fir_cluster.cuh): N = 32768 on a 4-CTA cluster, N = 16384 on a 2-CTA cluster.
#define ADT_FIR_VARIANT_IMPL
#define ADT_FIR_CLUSTER_IMPL
#include "fir_cluster.cuh"
#include "fir_variants.cuh"

namespace adt {
const FirVariant* fir_variant_c4_32768() {
    static const FirVariant v = make_variant_cluster<FirClusterCfg<32, 32, 4>, 2>("c4");
    return &v;
}
const FirVariant* fir_variant_c2_16384() {
    static const FirVariant v = make_variant_cluster<FirClusterCfg<16, 32, 2>, 2>("c2");
    return &v;
}
const FirVariant* fir_variant_b2_16384() {
    static const FirVariant v = make_variant_cluster2b<FirClusterCfg<16, 32, 2>, 2>("b2");
    return &v;
}
}  // namespace adt
