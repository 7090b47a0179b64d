// fir_variants.cuh — the table entry for one transform size / kernel family, and the templates that fill
// it.  Each size is instantiated in its own translation unit (fir_k*.cu) so the sizes compile in parallel;
// adt_api.cu only sees the accessor functions at the bottom.
//
// A/B families that are not on the default path (16 points per thread "p16", the split launch, the
// persistent loop for sizes where it loses) are compiled only with -DADT_AB_VARIANTS=1 (make AB=1).
#pragma once
#include <vector>

#include "fir_kernel.cuh"
#include "fir_tables.h"

#ifndef ADT_AB_VARIANTS
#define ADT_AB_VARIANTS 0
#endif

namespace adt {

typedef void (*fir_kernel_fn)(const FirKernelArgs, const FirExtra);

struct HostTables {
    std::vector<cf> tw1, tw2;
    std::vector<float> coef_s, coef_x;  // kernel-order mask (and, 16-point variant, the cross coefficients)
};

struct FirVariant {
    const char* name;
    int n, threads;
    int cluster;   // CTAs per thread-block cluster cooperating on one transform (1 = ordinary launch)
    size_t smem;
    fir_kernel_fn cplx, real;
    fir_kernel_fn cplx_i16, real_i16;  // 16-bit PCM in/out (IoI16)
    fir_kernel_fn persist_cplx, persist_real;  // persistent dynamic-queue variant (p32, float32 I/O) or null
    fir_kernel_fn shaped_cplx, shaped_real;    // float32 I/O with the wave-shaper store epilogue
    fir_kernel_fn split_int_cplx, split_int_real, split_edge_cplx, split_edge_real;  // split launch (p32) or null
    fir_kernel_fn tma_cplx, tma_real;          // TMA-fed window load (A/B, ADT_FIR_TMA=1) or null
    fir_kernel_fn vt2_cplx, vt2_real;          // two virtual threads per thread, T/2 threads per CTA (A/B, ADT_FIR_VT=2) or null
    fir_kernel_fn pp_cplx, pp_real;            // ping-pong schedule (fir_pingpong.cuh): 2 groups per CTA, persistent, or null
    fir_kernel_fn smask_real;                  // real mask staged in shared memory by the bulk-copy engine (A/B, ADT_FIR_SMASK=1) or null
    fir_kernel_fn accum_cplx, accum_real;      // y += result: tap segments 1.. of a partitioned (long) filter, or null
    void (*build)(const float* mask, bool real_only, HostTables& out);
};

#ifdef ADT_FIR_VARIANT_IMPL
// 32 points per thread (fft_core.cuh).  PERSIST: also build the persistent dynamic-queue kernel.
template <class C, int MIN_CTAS, bool PERSIST, bool TMA = false>
FirVariant make_variant32(const char* name) {
    FirVariant v;
    v.name = name;
    v.n = C::N;
    v.threads = C::T;
    v.cluster = 1;
    v.smem = (size_t)C::TILE * sizeof(cf);
    v.cplx = fir_block_kernel<C, cf, MIN_CTAS>;
    v.real = fir_block_kernel<C, float, MIN_CTAS>;
    v.cplx_i16 = fir_block_kernel<C, cf, MIN_CTAS, IoI16>;
    v.real_i16 = fir_block_kernel<C, float, MIN_CTAS, IoI16>;
    v.persist_cplx = v.persist_real = nullptr;
    if constexpr (PERSIST || ADT_AB_VARIANTS) {
        v.persist_cplx = fir_persist_kernel<C, cf, MIN_CTAS>;
        v.persist_real = fir_persist_kernel<C, float, MIN_CTAS>;
    }
    v.shaped_cplx = fir_block_kernel<C, cf, MIN_CTAS, IoF32, true>;
    v.shaped_real = fir_block_kernel<C, float, MIN_CTAS, IoF32, true>;
    v.split_int_cplx = v.split_int_real = v.split_edge_cplx = v.split_edge_real = nullptr;
    v.accum_cplx = fir_block_kernel<C, cf, MIN_CTAS, IoF32, false, true>;
    v.accum_real = fir_block_kernel<C, float, MIN_CTAS, IoF32, false, true>;
    v.tma_cplx = v.tma_real = nullptr;
    v.vt2_cplx = v.vt2_real = nullptr;
    v.pp_cplx = v.pp_real = nullptr;
    v.smask_real = nullptr;
    if constexpr (TMA) v.smask_real = fir_smask_kernel<C, MIN_CTAS>;
#ifdef ADT_FIR_PINGPONG_IMPL
    if constexpr (TMA && C::T == 256) {
        v.pp_cplx = fir_pingpong_kernel<C, cf>;
        v.pp_real = fir_pingpong_kernel<C, float>;
    }
#endif
    if constexpr (TMA) {
        v.tma_cplx = fir_tma_kernel<C, cf, MIN_CTAS>;
        v.tma_real = fir_tma_kernel<C, float, MIN_CTAS>;
        v.vt2_cplx = fir_vt2_kernel<C, cf, MIN_CTAS + 1>;
        v.vt2_real = fir_vt2_kernel<C, float, MIN_CTAS + 1>;
    }
#if ADT_AB_VARIANTS
    v.split_int_cplx = fir_split_kernel<C, cf, MIN_CTAS, true>;
    v.split_int_real = fir_split_kernel<C, float, MIN_CTAS, true>;
    v.split_edge_cplx = fir_split_kernel<C, cf, MIN_CTAS, false>;
    v.split_edge_real = fir_split_kernel<C, float, MIN_CTAS, false>;
#endif
    v.build = [](const float* mask, bool real_only, HostTables& out) {
        out.tw1 = build_tw1<C>();
        out.tw2 = build_tw2<C>();
        out.coef_s = permute_mask<C>(mask, real_only);
        out.coef_x.clear();
    };
    return v;
}
// 16 points per thread (fft_core16.cuh)
template <class C, int MIN_CTAS>
FirVariant make_variant16(const char* name) {
    FirVariant v;
    v.name = name;
    v.n = C::N;
    v.threads = C::T;
    v.cluster = 1;
    v.smem = (size_t)C::TILE * sizeof(cf);
    v.cplx = fir16_block_kernel<C, cf, MIN_CTAS>;
    v.real = fir16_block_kernel<C, float, MIN_CTAS>;
    v.cplx_i16 = fir16_block_kernel<C, cf, MIN_CTAS, IoI16>;
    v.real_i16 = fir16_block_kernel<C, float, MIN_CTAS, IoI16>;
    v.persist_cplx = v.persist_real = nullptr;
    v.shaped_cplx = fir16_block_kernel<C, cf, MIN_CTAS, IoF32, true>;
    v.shaped_real = fir16_block_kernel<C, float, MIN_CTAS, IoF32, true>;
    v.split_int_cplx = v.split_int_real = v.split_edge_cplx = v.split_edge_real = nullptr;
    v.tma_cplx = v.tma_real = nullptr;
    v.vt2_cplx = v.vt2_real = nullptr;
    v.pp_cplx = v.pp_real = nullptr;
    v.smask_real = nullptr;
    v.accum_cplx = v.accum_real = nullptr;
    v.build = [](const float* mask, bool real_only, HostTables& out) {
        out.tw1 = build16_tw1<C>();
        out.tw2 = build16_tw2<C>();
        build16_coef<C>(mask, real_only, out.coef_s, out.coef_x);
    };
    return v;
}
#ifdef ADT_FIR_CLUSTER_IMPL
// thread-block-cluster transform (fir_cluster.cuh): float32 I/O, store or accumulate
template <class C, int MIN_CTAS>
FirVariant make_variant_cluster(const char* name) {
    FirVariant v{};
    v.name = name;
    v.n = C::N;
    v.threads = C::T;
    v.cluster = C::CS;
    v.smem = (size_t)C::TILE * sizeof(cf);
    v.cplx = fir_cluster_kernel<C, cf, MIN_CTAS, false>;
    v.real = fir_cluster_kernel<C, float, MIN_CTAS, false>;
    v.accum_cplx = fir_cluster_kernel<C, cf, MIN_CTAS, true>;
    v.accum_real = fir_cluster_kernel<C, float, MIN_CTAS, true>;
    v.build = [](const float* mask, bool real_only, HostTables& out) {
        out.tw1 = build_tw1<C>();
        out.tw2 = build_tw2<C>();
        out.coef_s = permute_mask_cluster<C>(mask, real_only);
        out.coef_x.clear();
    };
    return v;
}
// 2-CTA cluster with the bulk (TMA-engine) DSMEM exchange: tile + staging buffer per CTA
template <class C, int MIN_CTAS>
FirVariant make_variant_cluster2b(const char* name) {
    FirVariant v = make_variant_cluster<C, MIN_CTAS>(name);
    v.smem = (size_t)C::TILE * sizeof(cf) + Cluster2Layout<C>::STAGE_BYTES;
    v.cplx = fir_cluster2b_kernel<C, cf, MIN_CTAS, false>;
    v.real = fir_cluster2b_kernel<C, float, MIN_CTAS, false>;
    v.accum_cplx = fir_cluster2b_kernel<C, cf, MIN_CTAS, true>;
    v.accum_real = fir_cluster2b_kernel<C, float, MIN_CTAS, true>;
    return v;
}
#endif
#endif  // ADT_FIR_VARIANT_IMPL

// one accessor per translation unit (null when the family is not built)
const FirVariant* fir_variant_p32_4096();
const FirVariant* fir_variant_p32_8192();
const FirVariant* fir_variant_p32_16384();
const FirVariant* fir_variant_p16_4096();
const FirVariant* fir_variant_p16_8192();
const FirVariant* fir_variant_c4_32768();   // 4-CTA cluster, N = 32768
const FirVariant* fir_variant_c2_16384();   // 2-CTA cluster, N = 16384 (A/B against the one-CTA kernel)
const FirVariant* fir_variant_b2_16384();   // 2-CTA cluster, N = 16384, bulk DSMEM exchange through a staging buffer

}  // namespace adt
