// adt_api.cu — C ABI (include/adt_b200.h): context, memory, events and the
// FFT FIR engine built on fir_kernel.cuh.  The biquad and the NCCL channel
// sharding live in adt_biquad.cu / adt_comm.cu.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/adt_b200.h"
#include "adt_internal.h"
#define CK ADT_CK
#include "fir_variants.cuh"

#ifndef ADT_FIR_PP_DEFAULT
#define ADT_FIR_PP_DEFAULT 0   /* ping-pong schedule (fir_pingpong.cuh) as the default for sizes that have it */
#endif

using namespace adt;

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
int adt_set_error(adt_ctx* ctx, int status, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->last_error = buf;
    }
    return status;
}

int adt_cuda_fail(adt_ctx* ctx, cudaError_t e, const char* what) {
    return adt_set_error(ctx, ADT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

extern "C" const char* adt_version(void) {
    return ADT_AB_VARIANTS ? "adt_b200 0.2 (sm_100a, ab_variants=1)" : "adt_b200 0.2 (sm_100a, ab_variants=0)";
}

extern "C" const char* adt_status_string(int s) {
    switch (s) {
        case ADT_OK: return "ok";
        case ADT_ERR_INVALID: return "invalid argument";
        case ADT_ERR_CUDA: return "CUDA error";
        case ADT_ERR_NO_DEVICE: return "no CUDA device";
        case ADT_ERR_UNSUPPORTED: return "unsupported geometry";
        case ADT_ERR_NCCL: return "NCCL error";
        case ADT_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

extern "C" int adt_device_count(int* count) {
    if (!count) return ADT_ERR_INVALID;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return ADT_OK;
}

// ---------------------------------------------------------------------------
// kernel table: one entry per supported transform size
// ---------------------------------------------------------------------------
namespace {

const FirVariant* const* all_variants(int* count) {
    static const FirVariant* table[] = {
        fir_variant_p32_4096(),   // N = 4096,  128 threads, 4 CTAs/SM
        fir_variant_p32_8192(),   // N = 8192,  256 threads, 2 CTAs/SM  (headline kernel)
        fir_variant_p32_16384(),  // N = 16384, 512 threads, 1 CTA/SM
        fir_variant_c4_32768(),   // N = 32768, 4-CTA cluster x 256 threads, DSMEM exchange
        fir_variant_c2_16384(),   // N = 16384, 2-CTA cluster (ADT_FIR_KERNEL=c2)
        fir_variant_b2_16384(),   // N = 16384, 2-CTA cluster, bulk DSMEM exchange (ADT_FIR_KERNEL=b2)
        fir_variant_p16_4096(),   // A/B family, null unless built with AB=1
        fir_variant_p16_8192(),
    };
    *count = (int)(sizeof table / sizeof table[0]);
    return table;
}

const FirVariant* find_variant(int n, const char* name = nullptr) {
    int cnt = 0;
    const FirVariant* const* t = all_variants(&cnt);
    if (name && *name)
        for (int i = 0; i < cnt; ++i)
            if (t[i] && t[i]->n == n && !strcmp(t[i]->name, name)) return t[i];
    for (int i = 0; i < cnt; ++i)
        if (t[i] && t[i]->n == n) return t[i];
    return nullptr;
}

}  // namespace

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int adt_ctx_create(int device, adt_ctx** out) {
    if (!out) return ADT_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return ADT_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return ADT_ERR_INVALID;
    adt_ctx* ctx = new (std::nothrow) adt_ctx();
    if (!ctx) return ADT_ERR_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; e == cudaSuccess && i < ADT_COPY_STREAMS; ++i) {
        e = cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->copy_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fence, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete ctx;
        return ADT_ERR_CUDA;
    }
    int n_var = 0;
    const FirVariant* const* vars = all_variants(&n_var);
    for (int vi = 0; vi < n_var; ++vi) {
        const FirVariant* v = vars[vi];
        if (!v) continue;
        for (fir_kernel_fn f : {v->cplx, v->real, v->cplx_i16, v->real_i16, v->persist_cplx, v->persist_real,
                                v->shaped_cplx, v->shaped_real, v->split_int_cplx, v->split_int_real,
                                v->split_edge_cplx, v->split_edge_real, v->tma_cplx, v->tma_real, v->vt2_cplx, v->vt2_real, v->accum_cplx,
                                v->accum_real}) {
            if (!f) continue;
            e = cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->smem);
            if (e == cudaSuccess && getenv("ADT_FIR_CARVEOUT"))  // tuning knob: % of the 228 KB given to shared memory
                e = cudaFuncSetAttribute((const void*)f, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         atoi(getenv("ADT_FIR_CARVEOUT")));
            if (e != cudaSuccess) {
                delete ctx;
                return ADT_ERR_CUDA;
            }
        }
        if (v->smask_real &&   // tile + real mask table
            cudaFuncSetAttribute((const void*)v->smask_real, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(v->smem + (size_t)v->n * sizeof(float))) != cudaSuccess) {
            delete ctx;
            return ADT_ERR_CUDA;
        }
        for (fir_kernel_fn f : {v->pp_cplx, v->pp_real}) {   // two groups, two tiles per CTA
            if (!f) continue;
            if (cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * v->smem)) != cudaSuccess) {
                delete ctx;
                return ADT_ERR_CUDA;
            }
        }
    }
    *out = ctx;
    return ADT_OK;
}

extern "C" int adt_ctx_destroy(adt_ctx* ctx) {
    if (!ctx) return ADT_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < ADT_COPY_STREAMS; ++i) {
        if (ctx->copy_done[i]) cudaEventDestroy(ctx->copy_done[i]);
        if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
    }
    if (ctx->fence) cudaEventDestroy(ctx->fence);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ADT_OK;
}

extern "C" const char* adt_last_error(adt_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int adt_ctx_sync(adt_ctx* ctx) {
    if (!ctx) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}

extern "C" int adt_ctx_launch_count(adt_ctx* ctx, uint64_t* count) {
    if (!ctx || !count) return ADT_ERR_INVALID;
    *count = ctx->launches;
    return ADT_OK;
}

extern "C" int adt_ctx_device_name(adt_ctx* ctx, char* buf, size_t len) {
    if (!ctx || !buf || !len) return ADT_ERR_INVALID;
    cudaDeviceProp p;
    CK(ctx, cudaGetDeviceProperties(&p, ctx->device));
    snprintf(buf, len, "%s (sm_%d%d, %d SMs)", p.name, p.major, p.minor, p.multiProcessorCount);
    return ADT_OK;
}

extern "C" int adt_ctx_pci_bus_id(adt_ctx* ctx, char* buf, size_t len) {
    if (!ctx || !buf || len < 13) return ADT_ERR_INVALID;
    CK(ctx, cudaDeviceGetPCIBusId(buf, (int)len, ctx->device));
    return ADT_OK;
}

// ---------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------
extern "C" int adt_malloc(adt_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return ADT_OK;
}
extern "C" int adt_free(adt_ctx* ctx, void* dptr) {
    if (!ctx) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaFree(dptr));
    return ADT_OK;
}
extern "C" int adt_malloc_host(adt_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return ADT_OK;
}
extern "C" int adt_free_host(adt_ctx* ctx, void* hptr) {
    if (!ctx) return ADT_ERR_INVALID;
    CK(ctx, cudaFreeHost(hptr));
    return ADT_OK;
}
extern "C" int adt_memset(adt_ctx* ctx, void* dptr, int value, size_t bytes) {
    if (!ctx || (!dptr && bytes)) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return ADT_OK;
}
extern "C" int adt_memcpy_h2d(adt_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || ((!dst || !src) && bytes)) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
extern "C" int adt_memcpy_d2h(adt_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || ((!dst || !src) && bytes)) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
extern "C" int adt_memcpy_d2d(adt_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || ((!dst || !src) && bytes)) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return ADT_OK;
}

// Concurrent H2D and D2H of the given sizes on two copy streams, returns when both are done: the transfer
// ceiling of this box for a *_host call (bench.py reports it next to the end-to-end number).
extern "C" int adt_copy_roundtrip_host(adt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes_h2d, void* dst_host,
                                       const void* src_dev, size_t bytes_d2h) {
    if (!ctx || ((!dst_dev || !src_host) && bytes_h2d) || ((!dst_host || !src_dev) && bytes_d2h)) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (bytes_h2d) CK(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes_h2d, cudaMemcpyHostToDevice, ctx->copy_stream[0]));
    if (bytes_d2h) CK(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes_d2h, cudaMemcpyDeviceToHost, ctx->copy_stream[1]));
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream[0]));
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream[1]));
    return ADT_OK;
}

// ---------------------------------------------------------------------------
// events
// ---------------------------------------------------------------------------
struct adt_event {
    adt_ctx* ctx;
    cudaEvent_t ev;
};

extern "C" int adt_event_create(adt_ctx* ctx, adt_event** out) {
    if (!ctx || !out) return ADT_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    adt_event* e = new (std::nothrow) adt_event{ctx, nullptr};
    if (!e) return ADT_ERR_NOMEM;
    cudaError_t err = cudaEventCreate(&e->ev);
    if (err != cudaSuccess) {
        delete e;
        return adt_cuda_fail(ctx, err, "cudaEventCreate");
    }
    *out = e;
    return ADT_OK;
}
extern "C" int adt_event_destroy(adt_event* ev) {
    if (!ev) return ADT_ERR_INVALID;
    cudaEventDestroy(ev->ev);
    delete ev;
    return ADT_OK;
}
extern "C" int adt_event_record(adt_event* ev) {
    if (!ev) return ADT_ERR_INVALID;
    CK(ev->ctx, cudaSetDevice(ev->ctx->device));
    CK(ev->ctx, cudaEventRecord(ev->ev, ev->ctx->stream));
    return ADT_OK;
}
extern "C" int adt_event_elapsed_ms(adt_event* a, adt_event* b, float* ms) {
    if (!a || !b || !ms) return ADT_ERR_INVALID;
    CK(b->ctx, cudaEventSynchronize(b->ev));
    CK(b->ctx, cudaEventElapsedTime(ms, a->ev, b->ev));
    return ADT_OK;
}

// ---------------------------------------------------------------------------
// FIR engine
// ---------------------------------------------------------------------------
// One tap segment of a filter: its own transform size, block geometry, mask and twiddle tables.  Ordinary
// filters have exactly one; filters too long for one transform are partitioned in time:
// y = sum_s (h_s * x) delayed by s*Ls, segment 0 stores and segments 1.. accumulate into the same output.
struct FirSeg {
    adt_fir_desc d{};
    const FirVariant* var = nullptr;
    void* d_mask = nullptr;
    cf* d_coef_x = nullptr;
    cf* d_tw1 = nullptr;
    cf* d_tw2 = nullptr;
    int resident_ctas = 0;  // CTAs resident at once (SMs x CTAs per SM), computed on first launch
};

struct adt_fir {
    adt_ctx* ctx = nullptr;
    adt_fir_desc d{};          // segment 0 (chunk / n_channels are common to all segments)
    std::vector<FirSeg> seg;
    int hist_back = 0;         // max `back` over the segments: samples of history the streaming state keeps
    // streaming state: two [n_channels][hist_pitch] buffers (history ++ newest chunk)
    FirShape shape{0, 0, 0.f, 0.f, 0.f, 0.f};  // store epilogue (adt_fir_set_epilogue)
    unsigned int* d_counter = nullptr;  // work queue heads of the persistent variant, one per stream slot
    float* d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    int64_t hist_pitch = 0;
    float* d_io = nullptr;  // [n_channels][chunk] staging for apply_host
    // scratch for process_host (per copy stream)
    float* d_in[ADT_COPY_STREAMS] = {};
    float* d_out[ADT_COPY_STREAMS] = {};
    size_t in_cap[ADT_COPY_STREAMS] = {};
    size_t out_cap[ADT_COPY_STREAMS] = {};
};

__global__ void fir_set_counter(unsigned int* c, unsigned int v) { *c = v; }


// slot: 0 = the context stream, 1 + i = copy stream i (launches on different streams may overlap, so
// per-launch device state — the persistent variant's work counter — exists once per slot)
static int fir_launch_seg(adt_fir* f, FirSeg& sg, bool accum, cudaStream_t s, const void* x, int64_t in_pitch,
                          int64_t n_in, int64_t in_shift, void* y, int64_t out_pitch, int64_t n_out, int32_t n_rows,
                          bool i16, int slot) {
    adt_ctx* ctx = f->ctx;
    if (n_rows <= 0 || n_out <= 0) return ADT_OK;
    const int64_t blocks = (n_out + sg.d.hop - 1) / sg.d.hop;
    const int64_t pairs = (n_rows + 1) / 2;
    if (blocks > 0x7fffffffLL)
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "too many blocks per row: %lld", (long long)blocks);
    FirKernelArgs a;
    a.x = x;
    a.y = y;
    a.mask = sg.d_mask;
    a.coef_x = sg.d_coef_x;
    a.tw1 = sg.d_tw1;
    a.tw2 = sg.d_tw2;
    a.n_rows = n_rows;
    a.blocks_per_row = (int)blocks;
    a.n_items = blocks * pairs;
    a.g.hop = sg.d.hop;
    a.g.n0 = sg.d.n0;
    a.g.back = sg.d.back;
    a.g.in_shift = in_shift;
    a.g.n_in = n_in;
    a.g.n_out = n_out;
    a.g.in_pitch = in_pitch;
    a.g.out_pitch = out_pitch;
    FirExtra ex;
    ex.work_counter = nullptr;
    ex.shape = f->shape;
    const bool shaped = f->shape.kind != 0;
    if (shaped && i16)
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "the wave-shaper epilogue is built for float32 I/O only");
    if (f->seg.size() > 1 && (shaped || i16))
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "partitioned (long) filters support plain float32 I/O only");
    if (accum && !sg.var->accum_real)
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "kernel family %s has no accumulate variant", sg.var->name);
    fir_kernel_fn k = shaped ? (sg.d.mask_is_real ? sg.var->shaped_real : sg.var->shaped_cplx)
                      : i16  ? (sg.d.mask_is_real ? sg.var->real_i16 : sg.var->cplx_i16)
                             : (sg.d.mask_is_real ? sg.var->real : sg.var->cplx);
    if (a.n_items > 0x7fffffffLL)
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "too many work items: %lld", (long long)a.n_items);
    if (accum) k = sg.d.mask_is_real ? sg.var->accum_real : sg.var->accum_cplx;
    if (!k)
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "kernel family %s (N = %d) has no %s variant", sg.var->name, sg.var->n,
                             shaped ? "store-epilogue" : i16 ? "int16" : "such");
    static const int tma_mode = getenv("ADT_FIR_TMA") ? atoi(getenv("ADT_FIR_TMA")) : 0;   // A/B: TMA-fed window load
    if (tma_mode && !shaped && !i16 && !accum && sg.var->tma_real) k = sg.d.mask_is_real ? sg.var->tma_real : sg.var->tma_cplx;
    static const int vt_mode = getenv("ADT_FIR_VT") ? atoi(getenv("ADT_FIR_VT")) : 0;   // A/B: two virtual threads per thread
    int threads = sg.var->threads;
    if (vt_mode == 2 && !shaped && !i16 && !accum && sg.var->vt2_real) {
        k = sg.d.mask_is_real ? sg.var->vt2_real : sg.var->vt2_cplx;
        threads /= 2;
    }
    if (sg.resident_ctas == 0) {  // CTAs in flight at once = how far ahead the L2 prefetch looks
        int per_sm = 0, sms = 0;
        CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, threads, sg.var->smem));
        CK(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        sg.resident_ctas = per_sm > 0 ? per_sm * sms : sms;
    }
    // L2 prefetch distance in units of one wave of resident CTAs (measured: flat between 0.25 and 1; 0 = off costs 9 %)
    const double pf = getenv("ADT_FIR_PREFETCH") ? atof(getenv("ADT_FIR_PREFETCH")) : 0.5;
    a.prefetch_ahead = (int)(pf * sg.resident_ctas);
    unsigned grid = (unsigned)a.n_items;
    // persistent dynamic-queue variant: only worth it when there are several waves of items
    // measured: +1.5 % for the 1-CTA/SM N = 16384 kernel, -5 % for N = 8192 -> default on for 16384 only
    const int persist_mode = getenv("ADT_FIR_PERSIST") ? atoi(getenv("ADT_FIR_PERSIST")) : (sg.d.fft_size == 16384);
    fir_kernel_fn kp = sg.d.mask_is_real ? sg.var->persist_real : sg.var->persist_cplx;
    bool persistent = false;
    if (persist_mode && !i16 && !shaped && !accum && kp && (persist_mode == 2 || a.n_items >= 4LL * sg.resident_ctas)) {   // 2 = force (tests)
        if (!f->d_counter) CK(ctx, cudaMalloc((void**)&f->d_counter, (ADT_COPY_STREAMS + 1) * sizeof(unsigned int)));
        grid = (unsigned)sg.resident_ctas;
        fir_set_counter<<<1, 1, 0, s>>>(f->d_counter + slot, grid);   // first unclaimed item = grid size
        ctx->launches++;
        ex.work_counter = f->d_counter + slot;
        k = kp;
        persistent = true;
    }
    ex.blk_count = (int)blocks;
    ex.blk_offset = ex.blk_skip_len = 0;
    ex.blk_skip_from = (int)blocks;
    // split launch: interior blocks by a kernel without any bounds code, edge blocks by the generic one
    static const int split_mode = getenv("ADT_FIR_SPLIT") ? atoi(getenv("ADT_FIR_SPLIT")) : 0;
    fir_kernel_fn k_int = sg.d.mask_is_real ? sg.var->split_int_real : sg.var->split_int_cplx;
    fir_kernel_fn k_edge = sg.d.mask_is_real ? sg.var->split_edge_real : sg.var->split_edge_cplx;
    if (split_mode && !i16 && !shaped && !accum && !persistent && k_int) {
        const int64_t N = sg.d.fft_size, hop = sg.d.hop, back = sg.d.back;
        int64_t b_lo = (back - in_shift + hop - 1) / hop;            // first block with ws >= 0
        if (b_lo < 0) b_lo = 0;
        int64_t b_hi = (n_in - N + back - in_shift) / hop + 1;        // one past the last block with ws + N <= n_in
        if (n_in - N + back - in_shift < 0) b_hi = 0;
        if (b_hi > n_out / hop) b_hi = n_out / hop;                   // ... and a full hop inside the output
        if (b_hi > blocks) b_hi = blocks;
        if (b_hi - b_lo >= 8 && (b_hi - b_lo) * pairs >= sg.resident_ctas) {
            FirExtra ei = ex, ee = ex;
            ei.blk_count = (int)(b_hi - b_lo); ei.blk_offset = (int)b_lo; ei.blk_skip_from = ei.blk_count; ei.blk_skip_len = 0;
            ee.blk_count = (int)(blocks - (b_hi - b_lo)); ee.blk_offset = 0; ee.blk_skip_from = (int)b_lo; ee.blk_skip_len = (int)(b_hi - b_lo);
            k_int<<<(unsigned)(ei.blk_count * pairs), sg.var->threads, sg.var->smem, s>>>(a, ei);
            CK(ctx, cudaGetLastError());
            ctx->launches++;
            if (ee.blk_count > 0) {
                k_edge<<<(unsigned)(ee.blk_count * pairs), sg.var->threads, sg.var->smem, s>>>(a, ee);
                CK(ctx, cudaGetLastError());
                ctx->launches++;
            }
            return ADT_OK;
        }
    }
    // ping-pong schedule: one persistent CTA of two groups per SM; the work counter starts at the first unclaimed item
    static const int pp_mode = getenv("ADT_FIR_PP") ? atoi(getenv("ADT_FIR_PP")) : ADT_FIR_PP_DEFAULT;
    fir_kernel_fn kpp = sg.d.mask_is_real ? sg.var->pp_real : sg.var->pp_cplx;
    if (pp_mode && kpp && !i16 && !shaped && !accum && !persistent && sg.var->cluster == 1 && vt_mode != 2 && !tma_mode) {
        int sms = 0;
        CK(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        long long ctas = (a.n_items + 1) / 2;
        if (ctas > sms) ctas = sms;
        if (!f->d_counter) CK(ctx, cudaMalloc((void**)&f->d_counter, (ADT_COPY_STREAMS + 1) * sizeof(unsigned int)));
        fir_set_counter<<<1, 1, 0, s>>>(f->d_counter + slot, (unsigned)(2 * ctas));
        ctx->launches++;
        ex.work_counter = f->d_counter + slot;
        kpp<<<(unsigned)ctas, 2 * sg.var->threads, 2 * sg.var->smem, s>>>(a, ex);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
        return ADT_OK;
    }
    static const int smask_mode = getenv("ADT_FIR_SMASK") ? atoi(getenv("ADT_FIR_SMASK")) : 0;   // A/B: mask table in shared memory
    if (smask_mode && sg.var->smask_real && sg.d.mask_is_real && !i16 && !shaped && !accum && !persistent && vt_mode != 2 && !tma_mode) {
        sg.var->smask_real<<<grid, sg.var->threads, sg.var->smem + (size_t)sg.var->n * sizeof(float), s>>>(a, ex);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
        return ADT_OK;
    }
    if (sg.var->cluster > 1) {   // one thread-block cluster per work item
        if (a.n_items * sg.var->cluster > 0x7fffffffLL)
            return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "too many work items: %lld", (long long)a.n_items);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(a.n_items * sg.var->cluster));
        cfg.blockDim = dim3((unsigned)sg.var->threads);
        cfg.dynamicSmemBytes = sg.var->smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)sg.var->cluster;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CK(ctx, cudaLaunchKernelEx(&cfg, k, a, ex));
        ctx->launches++;
        return ADT_OK;
    }
    k<<<grid, threads, sg.var->smem, s>>>(a, ex);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    return ADT_OK;
}

// every tap segment in turn on the same stream: segment 0 stores, the others accumulate
static int fir_launch(adt_fir* f, cudaStream_t s, const void* x, int64_t in_pitch, int64_t n_in, int64_t in_shift,
                      void* y, int64_t out_pitch, int64_t n_out, int32_t n_rows, bool i16 = false, int slot = 0) {
    for (size_t i = 0; i < f->seg.size(); ++i) {
        int rc = fir_launch_seg(f, f->seg[i], i > 0, s, x, in_pitch, n_in, in_shift, y, out_pitch, n_out, n_rows, i16, slot);
        if (rc) return rc;
    }
    return ADT_OK;
}

extern "C" int adt_fir_destroy(adt_fir* f) {
    if (!f) return ADT_ERR_INVALID;
    cudaSetDevice(f->ctx->device);
    cudaDeviceSynchronize();
    for (FirSeg& sg : f->seg) {
        cudaFree(sg.d_mask);
        cudaFree(sg.d_coef_x);
        cudaFree(sg.d_tw1);
        cudaFree(sg.d_tw2);
    }
    cudaFree(f->d_counter);
    cudaFree(f->d_hist[0]);
    cudaFree(f->d_hist[1]);
    cudaFree(f->d_io);
    for (int i = 0; i < ADT_COPY_STREAMS; ++i) {
        cudaFree(f->d_in[i]);
        cudaFree(f->d_out[i]);
    }
    delete f;
    return ADT_OK;
}

static int fir_build_segment(adt_ctx* ctx, const adt_fir_desc* desc, const float* mask, FirSeg& sg) {
    // Default kernel family (measured): "p32" (32 points/thread) for every size; "p16"
    // (16 points/thread, 32 warps/SM) is an A/B family of the AB=1 build.  ADT_FIR_KERNEL overrides.
    const char* want = getenv("ADT_FIR_KERNEL");
    if (!want || !*want) want = "p32";
    const FirVariant* var = find_variant(desc->fft_size, want);
    if (!var)
        return adt_set_error(ctx, ADT_ERR_UNSUPPORTED, "fft_size %d is not one of the built transform sizes", desc->fft_size);
    if (desc->hop < 1 || desc->n0 < 0 || desc->back < 0 || (int64_t)desc->n0 + desc->hop > desc->fft_size)
        return adt_set_error(ctx, ADT_ERR_INVALID, "bad block geometry: hop=%d n0=%d back=%d N=%d", desc->hop, desc->n0,
                             desc->back, desc->fft_size);
    sg.d = *desc;
    sg.var = var;
    HostTables ht;
    var->build(mask, desc->mask_is_real != 0, ht);
    cudaError_t e = cudaMalloc(&sg.d_mask, ht.coef_s.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sg.d_tw1, ht.tw1.size() * sizeof(cf));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sg.d_tw2, ht.tw2.size() * sizeof(cf));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sg.d_coef_x, ht.coef_x.size() * sizeof(float) + 8);
    if (e == cudaSuccess)
        e = cudaMemcpy(sg.d_mask, ht.coef_s.data(), ht.coef_s.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !ht.coef_x.empty())
        e = cudaMemcpy(sg.d_coef_x, ht.coef_x.data(), ht.coef_x.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(sg.d_tw1, ht.tw1.data(), ht.tw1.size() * sizeof(cf), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(sg.d_tw2, ht.tw2.data(), ht.tw2.size() * sizeof(cf), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return adt_cuda_fail(ctx, e, "adt_fir_create (tables)");
    return ADT_OK;
}

extern "C" int adt_fir_create_segmented(adt_ctx* ctx, int32_t n_segments, const adt_fir_desc* descs,
                                        const float* const* masks, adt_fir** out) {
    if (!ctx || !descs || !masks || !out || n_segments < 1 || n_segments > 64) return ADT_ERR_INVALID;
    *out = nullptr;
    const adt_fir_desc* desc = &descs[0];
    if (desc->chunk < 0 || desc->n_channels < 0 || ((desc->chunk > 0) != (desc->n_channels > 0)))
        return adt_set_error(ctx, ADT_ERR_INVALID, "chunk and n_channels must both be > 0 or both be 0");
    for (int i = 0; i < n_segments; ++i)
        if (!masks[i] || descs[i].chunk != desc->chunk || descs[i].n_channels != desc->n_channels)
            return adt_set_error(ctx, ADT_ERR_INVALID, "segment %d: null mask or chunk / n_channels differ from segment 0", i);
    CK(ctx, cudaSetDevice(ctx->device));
    adt_fir* f = new (std::nothrow) adt_fir();
    if (!f) return ADT_ERR_NOMEM;
    f->ctx = ctx;
    f->d = *desc;
    f->seg.resize((size_t)n_segments);
    for (int i = 0; i < n_segments; ++i) {
        int rc = fir_build_segment(ctx, &descs[i], masks[i], f->seg[(size_t)i]);
        if (rc) {
            adt_fir_destroy(f);
            return rc;
        }
        if (descs[i].back > f->hist_back) f->hist_back = descs[i].back;
    }
    cudaError_t e = cudaSuccess;
    if (desc->n_channels > 0) {
        f->hist_pitch = ((int64_t)f->hist_back + desc->chunk + 31) / 32 * 32;
        const size_t hb = (size_t)desc->n_channels * f->hist_pitch * sizeof(float);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaMalloc((void**)&f->d_hist[i], hb);
            if (e == cudaSuccess) e = cudaMemset(f->d_hist[i], 0, hb);
        }
        if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_io, (size_t)desc->n_channels * desc->chunk * sizeof(float));
    }
    if (e != cudaSuccess) {
        adt_fir_destroy(f);
        return adt_cuda_fail(ctx, e, "adt_fir_create");
    }
    *out = f;
    return ADT_OK;
}

extern "C" int adt_fir_create(adt_ctx* ctx, const adt_fir_desc* desc, const float* mask, adt_fir** out) {
    if (!ctx || !desc || !mask || !out) return ADT_ERR_INVALID;
    return adt_fir_create_segmented(ctx, 1, desc, &mask, out);
}

// Attach (kind 1 saturator / 2 soft clipper) or detach (kind 0) a wave-shaper applied to every output
// sample in the kernel's store phase.  params = {p0, p1, p2, p3, mode} as for adt_shape_apply_*.
extern "C" int adt_fir_set_epilogue(adt_fir* f, int kind, const float* params) {
    if (!f) return ADT_ERR_INVALID;
    if (kind == 0) {
        f->shape = FirShape{0, 0, 0.f, 0.f, 0.f, 0.f};
        return ADT_OK;
    }
    if (kind < 1 || kind > 2 || !params) return adt_set_error(f->ctx, ADT_ERR_INVALID, "bad epilogue kind %d", kind);
    const int mode = kind == 1 ? (int)params[4] : 0;
    if (kind == 1 && mode != 1 && mode != 2) return adt_set_error(f->ctx, ADT_ERR_INVALID, "saturator mode must be 1 or 2");
    f->shape = FirShape{kind, mode, params[0], params[1], params[2], params[3]};
    return ADT_OK;
}

extern "C" int adt_fir_reset(adt_fir* f) {
    if (!f) return ADT_ERR_INVALID;
    adt_ctx* ctx = f->ctx;
    if (!f->d_hist[0]) return ADT_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t hb = (size_t)f->d.n_channels * f->hist_pitch * sizeof(float);
    CK(ctx, cudaMemsetAsync(f->d_hist[0], 0, hb, ctx->stream));
    CK(ctx, cudaMemsetAsync(f->d_hist[1], 0, hb, ctx->stream));
    return ADT_OK;
}

static int check_buffers(adt_fir* f, const void* x, int64_t in_pitch, int64_t n_in, const void* y, int64_t out_pitch,
                         int64_t n_out, int32_t n_rows) {
    if (!f) return ADT_ERR_INVALID;
    if (n_rows < 0 || n_in < 0 || n_out < 0 || in_pitch < n_in || out_pitch < n_out || ((!x || !y) && n_rows > 0))
        return adt_set_error(f->ctx, ADT_ERR_INVALID, "bad buffer shape: rows=%d n_in=%lld pitch=%lld n_out=%lld pitch=%lld",
                             n_rows, (long long)n_in, (long long)in_pitch, (long long)n_out, (long long)out_pitch);
    return ADT_OK;
}

static int fir_process_dev_impl(adt_fir* f, const void* x, int64_t in_pitch, int64_t n_in, void* y, int64_t out_pitch,
                                int64_t n_out, int32_t n_rows, bool i16) {
    int rc = check_buffers(f, x, in_pitch, n_in, y, out_pitch, n_out, n_rows);
    if (rc) return rc;
    CK(f->ctx, cudaSetDevice(f->ctx->device));
    return fir_launch(f, f->ctx->stream, x, in_pitch, n_in, 0, y, out_pitch, n_out, n_rows, i16);
}

extern "C" int adt_fir_process_dev(adt_fir* f, const float* x, int64_t in_pitch, int64_t n_in, float* y,
                                   int64_t out_pitch, int64_t n_out, int32_t n_rows) {
    return fir_process_dev_impl(f, x, in_pitch, n_in, y, out_pitch, n_out, n_rows, false);
}
extern "C" int adt_fir_process_dev_i16(adt_fir* f, const int16_t* x, int64_t in_pitch, int64_t n_in, int16_t* y,
                                       int64_t out_pitch, int64_t n_out, int32_t n_rows) {
    return fir_process_dev_impl(f, x, in_pitch, n_in, y, out_pitch, n_out, n_rows, true);
}

static int ensure_cap(adt_ctx* ctx, float** p, size_t* cap, size_t need) {
    if (*cap >= need) return ADT_OK;
    if (*p) CK(ctx, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CK(ctx, cudaMalloc((void**)p, need));
    *cap = need;
    return ADT_OK;
}

// H2D -> kernel -> D2H pipelined over row groups on the copy streams; es = bytes per sample.
static int fir_process_host_impl(adt_fir* f, const void* xv, int64_t in_pitch, int64_t n_in, void* yv,
                                 int64_t out_pitch, int64_t n_out, int32_t n_rows, bool i16) {
    int rc = check_buffers(f, xv, in_pitch, n_in, yv, out_pitch, n_out, n_rows);
    if (rc) return rc;
    adt_ctx* ctx = f->ctx;
    if (n_rows == 0 || n_out == 0) return ADT_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t es = i16 ? sizeof(int16_t) : sizeof(float);
    const char* x = static_cast<const char*>(xv);
    char* y = static_cast<char*>(yv);
    // row groups of ~48 MB, an even number of rows (channel pairs stay together); pitches keep 128-byte rows
    const int64_t din_pitch = (n_in + 63) / 64 * 64, dout_pitch = (n_out + 63) / 64 * 64;
    static const int64_t group_mb = getenv("ADT_FIR_GROUP_MB") ? atoi(getenv("ADT_FIR_GROUP_MB")) : 48;  // tuning knob
    int64_t g_rows = (group_mb << 20) / (int64_t)((din_pitch > dout_pitch ? din_pitch : dout_pitch) * es);
    g_rows = g_rows < 2 ? 2 : (g_rows & ~1LL);
    if (g_rows > n_rows) g_rows = (n_rows + 1) & ~1LL;
    // everything queued on the context stream so far must be done before the copy streams start
    CK(ctx, cudaEventRecord(ctx->fence, ctx->stream));
    int gi = 0;
    for (int64_t r0 = 0; r0 < n_rows; r0 += g_rows, ++gi) {
        const int si = gi % ADT_COPY_STREAMS;
        cudaStream_t s = ctx->copy_stream[si];
        const int32_t rows = (int32_t)((n_rows - r0) < g_rows ? (n_rows - r0) : g_rows);
        if (gi < ADT_COPY_STREAMS) CK(ctx, cudaStreamWaitEvent(s, ctx->fence, 0));
        if (f->in_cap[si] < (size_t)g_rows * din_pitch * es || f->out_cap[si] < (size_t)g_rows * dout_pitch * es) {
            CK(ctx, cudaStreamSynchronize(s));
            rc = ensure_cap(ctx, &f->d_in[si], &f->in_cap[si], (size_t)g_rows * din_pitch * es);
            if (rc) return rc;
            rc = ensure_cap(ctx, &f->d_out[si], &f->out_cap[si], (size_t)g_rows * dout_pitch * es);
            if (rc) return rc;
        }
        if (n_in > 0)
            CK(ctx, cudaMemcpy2DAsync(f->d_in[si], din_pitch * es, x + (size_t)r0 * in_pitch * es, in_pitch * es,
                                      n_in * es, rows, cudaMemcpyHostToDevice, s));
        rc = fir_launch(f, s, f->d_in[si], din_pitch, n_in, 0, f->d_out[si], dout_pitch, n_out, rows, i16, 1 + si);
        if (rc) return rc;
        CK(ctx, cudaMemcpy2DAsync(y + (size_t)r0 * out_pitch * es, out_pitch * es, f->d_out[si], dout_pitch * es,
                                  n_out * es, rows, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < ADT_COPY_STREAMS; ++i) CK(ctx, cudaStreamSynchronize(ctx->copy_stream[i]));
    return ADT_OK;
}

extern "C" int adt_fir_process_host(adt_fir* f, const float* x, int64_t in_pitch, int64_t n_in, float* y,
                                    int64_t out_pitch, int64_t n_out, int32_t n_rows) {
    return fir_process_host_impl(f, x, in_pitch, n_in, y, out_pitch, n_out, n_rows, false);
}
extern "C" int adt_fir_process_host_i16(adt_fir* f, const int16_t* x, int64_t in_pitch, int64_t n_in, int16_t* y,
                                        int64_t out_pitch, int64_t n_out, int32_t n_rows) {
    return fir_process_host_impl(f, x, in_pitch, n_in, y, out_pitch, n_out, n_rows, true);
}

// One streaming step on device buffers: hist[cur] = [last `back` samples | new chunk].
static int fir_apply_common(adt_fir* f, const float* in, cudaMemcpyKind in_kind, float* out_dev) {
    adt_ctx* ctx = f->ctx;
    const int C = f->d.chunk, rows = f->d.n_channels, back = f->hist_back;
    float* cur = f->d_hist[f->cur];
    float* nxt = f->d_hist[f->cur ^ 1];
    cudaStream_t s = ctx->stream;
    CK(ctx, cudaMemcpy2DAsync(cur + back, f->hist_pitch * sizeof(float), in, (size_t)C * sizeof(float),
                              (size_t)C * sizeof(float), rows, in_kind, s));
    // the chunk starts at buffer offset `back`; a segment's window for block b starts `its own back` before b*hop
    int rc = fir_launch(f, s, cur, f->hist_pitch, (int64_t)back + C, back, out_dev, C, C, rows);
    if (rc) return rc;
    if (back > 0)
        CK(ctx, cudaMemcpy2DAsync(nxt, f->hist_pitch * sizeof(float), cur + C, f->hist_pitch * sizeof(float),
                                  (size_t)back * sizeof(float), rows, cudaMemcpyDeviceToDevice, s));
    f->cur ^= 1;
    return ADT_OK;
}

extern "C" int adt_fir_apply_dev(adt_fir* f, const float* in_dev, float* out_dev) {
    if (!f || !in_dev || !out_dev) return ADT_ERR_INVALID;
    if (!f->d_hist[0]) return adt_set_error(f->ctx, ADT_ERR_INVALID, "fir was created without streaming state");
    CK(f->ctx, cudaSetDevice(f->ctx->device));
    return fir_apply_common(f, in_dev, cudaMemcpyDeviceToDevice, out_dev);
}

extern "C" int adt_fir_apply_host(adt_fir* f, const float* in_host, float* out_host) {
    if (!f || !in_host || !out_host) return ADT_ERR_INVALID;
    if (!f->d_hist[0]) return adt_set_error(f->ctx, ADT_ERR_INVALID, "fir was created without streaming state");
    adt_ctx* ctx = f->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = fir_apply_common(f, in_host, cudaMemcpyHostToDevice, f->d_io);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(out_host, f->d_io, (size_t)f->d.n_channels * f->d.chunk * sizeof(float),
                            cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
