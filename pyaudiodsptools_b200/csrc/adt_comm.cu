// adt_comm.cu — channel sharding across the GPUs of one box.
//
// The reference has no distributed anything; its unit of independence is "one
// device object per mono channel" (Example2.py:13-22).  Sharding by channel
// therefore needs no data-path collective: the only exchanges are a scatter
// of input rows from a root rank and a gather of output rows (grouped
// ncclSend/ncclRecv over NVLink/NVSwitch) plus a broadcast for small
// parameter blobs.  libnccl is dlopen()ed on first use so that the library
// loads (and every non-comm entry point works) on machines without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "adt_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define LOAD(field, sym)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
    if (!api.field) return;
        LOAD(GetUniqueId, "ncclGetUniqueId")
        LOAD(CommInitRank, "ncclCommInitRank")
        LOAD(CommDestroy, "ncclCommDestroy")
        LOAD(Send, "ncclSend")
        LOAD(Recv, "ncclRecv")
        LOAD(GroupStart, "ncclGroupStart")
        LOAD(GroupEnd, "ncclGroupEnd")
        LOAD(Broadcast, "ncclBroadcast")
        LOAD(AllReduce, "ncclAllReduce")
        LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
        api.ok = true;
    });
    return api;
}

}  // namespace

struct adt_comm {
    adt_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int* d_flag = nullptr;
};

#define NK(ctx, call)                                                                                  \
    do {                                                                                               \
        ncclResult_t r__ = (call);                                                                     \
        if (r__ != ncclSuccess) return adt_set_error(ctx, ADT_ERR_NCCL, "%s: %s", #call, nccl().GetErrorString(r__)); \
    } while (0)

static_assert(sizeof(ncclUniqueId) == ADT_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");

extern "C" int adt_comm_unique_id(unsigned char id[ADT_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) return ADT_ERR_INVALID;
    if (!nccl().ok) return ADT_ERR_NCCL;
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != ncclSuccess) return ADT_ERR_NCCL;
    memcpy(id, &u, sizeof u);
    return ADT_OK;
}

extern "C" int adt_comm_create(adt_ctx* ctx, const unsigned char id[ADT_NCCL_UNIQUE_ID_BYTES], int32_t rank,
                               int32_t world, adt_comm** out) {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return ADT_ERR_INVALID;
    *out = nullptr;
    if (!nccl().ok) return adt_set_error(ctx, ADT_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    adt_comm* c = new (std::nothrow) adt_comm();
    if (!c) return ADT_ERR_NOMEM;
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return adt_set_error(ctx, ADT_ERR_NCCL, "ncclCommInitRank: %s", nccl().GetErrorString(r));
    }
    cudaError_t e = cudaMalloc((void**)&c->d_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->d_flag, 0, sizeof(int));
    if (e != cudaSuccess) {
        nccl().CommDestroy(c->comm);
        delete c;
        return adt_cuda_fail(ctx, e, "adt_comm_create");
    }
    *out = c;
    return ADT_OK;
}

extern "C" int adt_comm_destroy(adt_comm* c) {
    if (!c) return ADT_ERR_INVALID;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    cudaFree(c->d_flag);
    delete c;
    return ADT_OK;
}

// Grouped point-to-point exchange between the root and every other rank.  `counts[r]` rows of `pitch` floats
// belong to rank r, stored back to back in the root's matrix (offset = prefix sum).  to_root = false: scatter
// (root sends), true: gather (root receives).  An error inside the group still closes the group before
// returning, so a failed call never leaves NCCL waiting for ncclGroupEnd.
static int comm_exchange_rows(adt_comm* c, const float* full_src, float* full_dst, const float* shard_src,
                              float* shard_dst, const int64_t* counts, int64_t pitch, int32_t root, bool to_root) {
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    int64_t my_off = 0, total = 0;
    for (int r = 0; r < c->world; ++r) {
        if (counts[r] < 0) return adt_set_error(ctx, ADT_ERR_INVALID, "negative row count for rank %d", r);
        if (r < c->rank) my_off += counts[r];
        total += counts[r];
    }
    if (total == 0 || pitch == 0) return ADT_OK;
    const size_t mine = (size_t)counts[c->rank] * pitch;
    ncclResult_t first = ncclSuccess;
    const char* what = "";
    auto note = [&](ncclResult_t r, const char* w) {
        if (r != ncclSuccess && first == ncclSuccess) {
            first = r;
            what = w;
        }
    };
    note(nccl().GroupStart(), "ncclGroupStart");
    if (first == ncclSuccess) {
        if (c->rank == root) {
            int64_t off = 0;
            for (int r = 0; r < c->world; ++r) {
                const size_t n = (size_t)counts[r] * pitch;
                if (r != root && n) {
                    if (to_root)
                        note(nccl().Recv(full_dst + (size_t)off * pitch, n, ncclFloat, r, c->comm, ctx->stream), "ncclRecv");
                    else
                        note(nccl().Send(full_src + (size_t)off * pitch, n, ncclFloat, r, c->comm, ctx->stream), "ncclSend");
                }
                off += counts[r];
            }
        } else if (mine) {
            if (to_root)
                note(nccl().Send(shard_src, mine, ncclFloat, root, c->comm, ctx->stream), "ncclSend");
            else
                note(nccl().Recv(shard_dst, mine, ncclFloat, root, c->comm, ctx->stream), "ncclRecv");
        }
        note(nccl().GroupEnd(), "ncclGroupEnd");   // always reached once the group was opened
    }
    if (first != ncclSuccess) return adt_set_error(ctx, ADT_ERR_NCCL, "%s: %s", what, nccl().GetErrorString(first));
    if (c->rank == root && mine) {   // the root's own rows: a local copy on the same stream
        if (to_root)
            ADT_CK(ctx, cudaMemcpyAsync(full_dst + (size_t)my_off * pitch, shard_src, mine * sizeof(float),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
        else
            ADT_CK(ctx, cudaMemcpyAsync(shard_dst, full_src + (size_t)my_off * pitch, mine * sizeof(float),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return ADT_OK;
}

extern "C" int adt_comm_scatterv_rows(adt_comm* c, const float* full, float* shard, const int64_t* row_counts,
                                      int64_t pitch, int32_t root) {
    if (!c || !row_counts || pitch < 0 || root < 0 || root >= c->world) return ADT_ERR_INVALID;
    if ((c->rank == root && !full) || (!shard && row_counts[c->rank] > 0)) return ADT_ERR_INVALID;
    return comm_exchange_rows(c, full, nullptr, nullptr, shard, row_counts, pitch, root, false);
}

extern "C" int adt_comm_gatherv_rows(adt_comm* c, const float* shard, float* full, const int64_t* row_counts,
                                     int64_t pitch, int32_t root) {
    if (!c || !row_counts || pitch < 0 || root < 0 || root >= c->world) return ADT_ERR_INVALID;
    if ((c->rank == root && !full) || (!shard && row_counts[c->rank] > 0)) return ADT_ERR_INVALID;
    return comm_exchange_rows(c, nullptr, full, shard, nullptr, row_counts, pitch, root, true);
}

// uniform shards: the same exchange with rows_per_rank everywhere
extern "C" int adt_comm_scatter_rows(adt_comm* c, const float* full, float* shard, int64_t rows_per_rank, int64_t pitch,
                                     int32_t root) {
    if (!c || rows_per_rank < 0) return ADT_ERR_INVALID;
    std::vector<int64_t> counts((size_t)c->world, rows_per_rank);
    return adt_comm_scatterv_rows(c, full, shard, counts.data(), pitch, root);
}

extern "C" int adt_comm_gather_rows(adt_comm* c, const float* shard, float* full, int64_t rows_per_rank, int64_t pitch,
                                    int32_t root) {
    if (!c || rows_per_rank < 0) return ADT_ERR_INVALID;
    std::vector<int64_t> counts((size_t)c->world, rows_per_rank);
    return adt_comm_gatherv_rows(c, shard, full, counts.data(), pitch, root);
}

extern "C" int adt_comm_broadcast(adt_comm* c, void* buf, size_t bytes, int32_t root) {
    if (!c || (!buf && bytes) || root < 0 || root >= c->world) return ADT_ERR_INVALID;
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    if (bytes == 0) return ADT_OK;
    NK(ctx, nccl().Broadcast(buf, buf, bytes, ncclUint8, root, c->comm, ctx->stream));
    return ADT_OK;
}

extern "C" int adt_comm_barrier(adt_comm* c) {
    if (!c) return ADT_ERR_INVALID;
    adt_ctx* ctx = c->ctx;
    ADT_CK(ctx, cudaSetDevice(ctx->device));
    NK(ctx, nccl().AllReduce(c->d_flag, c->d_flag, 1, ncclInt32, ncclSum, c->comm, ctx->stream));
    ADT_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ADT_OK;
}
