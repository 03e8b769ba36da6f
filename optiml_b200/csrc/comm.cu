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

extern "C" int svmb200_comm_destroy(svmb200_ctx* ctx) {
    if (!ctx) return SVMB200_OK;
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
    if (!ctx->nccl_comm) {
        svmb200_set_error("multi-rank context without communicator");
        return SVMB200_ERR_STATE;
    }
    SVM_NCCL(g_nccl.AllGather(dbuf + (size_t)ctx->rank * count_per_rank, dbuf, (size_t)count_per_rank, ncclFloat64,
                              (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SVMB200_OK;
}
