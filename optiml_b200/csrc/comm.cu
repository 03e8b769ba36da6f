// K4: the per-iteration exchange of matvec shards.  One ncclAllGather (in place) on the context's
// stream.  libnccl is resolved at run time with dlopen so that single-GPU use has no NCCL
// dependency and the process shares whichever libnccl.so.2 PyTorch already loaded.
#include "common.cuh"
#include <dlfcn.h>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;

#ifdef SVMB200_HOST_EMULATION
// tests/cuda_emu: ranks are threads of one process and the six NCCL entry points are in-process stand-ins
extern "C" {
ncclResult_t emu_ncclGetUniqueId(ncclUniqueId*);
ncclResult_t emu_ncclCommInitRank(ncclComm_t*, int, ncclUniqueId, int);
ncclResult_t emu_ncclCommDestroy(ncclComm_t);
ncclResult_t emu_ncclAllGather(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
const char* emu_ncclGetErrorString(ncclResult_t);
ncclResult_t emu_ncclGetVersion(int*);
}
static int nccl_load() {
    static int bound = 0;
    g_nccl.lib = &bound;
    g_nccl.GetUniqueId = emu_ncclGetUniqueId;
    g_nccl.CommInitRank = emu_ncclCommInitRank;
    g_nccl.CommDestroy = emu_ncclCommDestroy;
    g_nccl.AllGather = emu_ncclAllGather;
    g_nccl.GetErrorString = emu_ncclGetErrorString;
    g_nccl.GetVersion = emu_ncclGetVersion;
    return SVMB200_OK;
}
#else
static int nccl_load() {
    if (g_nccl.lib) return SVMB200_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        svmb200_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return SVMB200_ERR_NCCL;
    }
#define NCCL_SYM(field, name)                                               \
    *(void**)(&g_nccl.field) = dlsym(h, name);                              \
    if (!g_nccl.field) {                                                    \
        svmb200_set_error("libnccl is missing symbol %s", name);           \
        return SVMB200_ERR_NCCL;                                            \
    }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllGather, "ncclAllGather")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
    NCCL_SYM(GetVersion, "ncclGetVersion")
#undef NCCL_SYM
    g_nccl.lib = h;
    return SVMB200_OK;
}
#endif

#define SVM_NCCL(call)                                                                        \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != 0) {                                                                       \
            svmb200_set_error("%s failed: %s", #call, g_nccl.GetErrorString(r__));           \
            return SVMB200_ERR_NCCL;                                                          \
        }                                                                                     \
    } while (0)

extern "C" int svmb200_comm_unique_id(void* id128) {
    SVM_CHECK_ARG(id128 != nullptr, "null id buffer");
    SVM_TRY(nccl_load());
    ncclUniqueId id;
    SVM_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return SVMB200_OK;
}

extern "C" int svmb200_comm_init(svmb200_ctx* ctx, const void* id128, int rank, int nranks) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(id128 != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / size");
    SVM_CHECK_ARG(nranks <= SVM_MAX_RANKS, "more ranks than SVM_MAX_RANKS (16)");
    if (ctx->nccl_comm) {
        svmb200_set_error("communicator already initialised");
        return SVMB200_ERR_STATE;
    }
    if (nranks == 1) {
        ctx->rank = 0;
        ctx->nranks = 1;
        return SVMB200_OK;
    }
    SVM_TRY(nccl_load());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    SVM_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return SVMB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer-memory arena for the fused matvec + exchange (K2 writes into every peer, K3 waits on flags).
extern "C" int svmb200_comm_p2p_export(svmb200_ctx* ctx, size_t arena_bytes, void* handle64) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(handle64 != nullptr && arena_bytes >= (1u << 20), "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    if (ctx->arena) {
        svmb200_set_error("arena already exported");
        return SVMB200_ERR_STATE;
    }
    SVM_CUDA(cudaMalloc(&ctx->arena, arena_bytes));
    SVM_CUDA(cudaMemset(ctx->arena, 0, arena_bytes));
    ctx->arena_bytes = arena_bytes;
    cudaIpcMemHandle_t h;
    SVM_CUDA(cudaIpcGetMemHandle(&h, ctx->arena));
    memcpy(handle64, &h, sizeof(h));
    return SVMB200_OK;
}

extern "C" int svmb200_comm_p2p_attach(svmb200_ctx* ctx, const void* handles, int nranks) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(handles != nullptr && nranks == ctx->nranks && nranks <= SVM_MAX_RANKS, "bad argument");
    if (!ctx->arena) {
        svmb200_set_error("call svmb200_comm_p2p_export first");
        return SVMB200_ERR_STATE;
    }
    for (int r = 0; r < nranks; ++r) {
        if (r == ctx->rank) {
            ctx->peer_arena[r] = ctx->arena;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            svmb200_set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            for (int q = 0; q < r; ++q)
                if (q != ctx->rank && ctx->peer_arena[q]) {
                    cudaIpcCloseMemHandle(ctx->peer_arena[q]);
                    ctx->peer_arena[q] = nullptr;
                }
            return SVMB200_ERR_CUDA;
        }
        ctx->peer_arena[r] = static_cast<unsigned char*>(p);
    }
    ctx->p2p_enabled = true;
    ctx->xseq = 0;
    return SVMB200_OK;
}

// One process, N GPUs: the reference's API is a single Python process calling SVC.fit (ml/svm/_base.py:631-636).  The
// contexts of `ctxs` (one per device, all owned by the calling thread) become ranks 0..n-1 of a group whose exchange
// arenas are mapped by plain peer access -- no torchrun, no torch.distributed, no NCCL, no IPC.
extern "C" int svmb200_comm_local_group(svmb200_ctx** ctxs, int n, size_t arena_bytes) {
    SVM_CHECK_ARG(ctxs != nullptr && n >= 1 && n <= SVM_MAX_RANKS && arena_bytes >= (1u << 20), "bad argument");
    for (int i = 0; i < n; ++i) {
        SVM_CHECK_ARG(ctxs[i] != nullptr, "null context");
        SVM_CHECK_ARG(ctxs[i]->nranks == 1 && ctxs[i]->nccl_comm == nullptr && ctxs[i]->arena == nullptr,
                      "context already belongs to a communicator");
#ifndef SVMB200_HOST_EMULATION
        for (int j = 0; j < i; ++j) SVM_CHECK_ARG(ctxs[j]->device != ctxs[i]->device, "one context per device");
#endif
    }
    if (n == 1) return SVMB200_OK;
    for (int i = 0; i < n; ++i) {
        SVM_CUDA(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; ++j) {
            if (j == i || ctxs[j]->device == ctxs[i]->device) continue;
            int can = 0;
            SVM_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device));
            if (!can) {
                svmb200_set_error("device %d cannot access device %d: no peer path for the fused exchange", ctxs[i]->device,
                                  ctxs[j]->device);
                return SVMB200_ERR_CUDA;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                svmb200_set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", ctxs[i]->device, ctxs[j]->device,
                                  cudaGetErrorString(e));
                return SVMB200_ERR_CUDA;
            }
            cudaGetLastError();  // clear cudaErrorPeerAccessAlreadyEnabled
        }
        SVM_CUDA(cudaMalloc(&ctxs[i]->arena, arena_bytes));
        SVM_CUDA(cudaMemset(ctxs[i]->arena, 0, arena_bytes));
        ctxs[i]->arena_bytes = arena_bytes;
    }
    for (int i = 0; i < n; ++i) {
        for (int r = 0; r < n; ++r) ctxs[i]->peer_arena[r] = ctxs[r]->arena;
        ctxs[i]->rank = i;
        ctxs[i]->nranks = n;
        ctxs[i]->p2p_enabled = true;
        ctxs[i]->local_group = true;
        ctxs[i]->xseq = 0;
    }
    return SVMB200_OK;
}

extern "C" int svmb200_comm_p2p_disable(svmb200_ctx* ctx) {
    SVM_CHECK_ARG(ctx != nullptr, "null argument");
    ctx->p2p_enabled = false;  // keep the mappings; the solver falls back to ncclAllGather
    return SVMB200_OK;
}

extern "C" int svmb200_comm_p2p_enabled(svmb200_ctx* ctx, int* enabled) {
    SVM_CHECK_ARG(ctx != nullptr && enabled != nullptr, "null argument");
    *enabled = ctx->p2p_enabled ? 1 : 0;
    return SVMB200_OK;
}

// Tear-down of the peer mappings.  Peers may still be executing matvec launches that STORE into this rank's arena when
// a rank leaves early (an exception on one rank of a sharded fit): freeing the arena then turns their stores into
// illegal-address faults (sticky CUDA error, Xid) instead of the clean exchange time-out.  A cross-rank barrier cannot
// be relied on from an error path (the peer may be the one that died), so a multi-rank context leaves its arena
// allocated and its peer mappings open until the process exits -- 64 MB, reclaimed by the driver with the process.
static void p2p_release(svmb200_ctx* ctx) {
    if (ctx->local_group) {
        // one host thread owns every rank: the caller destroys the group's contexts together, after their streams idle
        for (int r = 0; r < SVM_MAX_RANKS; ++r) ctx->peer_arena[r] = nullptr;
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
        ctx->p2p_enabled = false;
        ctx->local_group = false;
        return;
    }
    const bool peers_may_still_store = ctx->nranks > 1;
    for (int r = 0; r < SVM_MAX_RANKS; ++r) {
        if (!peers_may_still_store && ctx->peer_arena[r] && ctx->peer_arena[r] != ctx->arena)
            cudaIpcCloseMemHandle(ctx->peer_arena[r]);
        ctx->peer_arena[r] = nullptr;
    }
    if (ctx->arena && !peers_may_still_store) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    ctx->p2p_enabled = false;
}

extern "C" int svmb200_comm_destroy(svmb200_ctx* ctx) {
    if (!ctx) return SVMB200_OK;
    if (ctx->arena) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        p2p_release(ctx);
    }
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->nranks = 1;
    return SVMB200_OK;
}

int svm_comm_allgather(svmb200_ctx* ctx, double* dbuf, int64_t count_per_rank) {
    if (ctx->nranks <= 1) return SVMB200_OK;
    if (ctx->local_group) {
        svmb200_set_error("a single-process group has no collective: this operation needs the fused exchange or a per-shard call");
        return SVMB200_ERR_STATE;
    }
    if (!ctx->nccl_comm) {
        svmb200_set_error("multi-rank context without communicator");
        return SVMB200_ERR_STATE;
    }
    SVM_NCCL(g_nccl.AllGather(dbuf + (size_t)ctx->rank * count_per_rank, dbuf, (size_t)count_per_rank, ncclFloat64,
                              (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SVMB200_OK;
}
