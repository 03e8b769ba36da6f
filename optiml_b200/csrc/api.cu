// Context, memory and host-pointer entry points of the svmb200 C ABI.
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>

static thread_local char g_err[1024] = "";

void svmb200_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* svmb200_last_error(void) { return g_err; }
#ifndef SVMB200_HOST_EMULATION
extern "C" const char* svmb200_version(void) { return "svmb200 0.1 (sm_100a)"; }
#else
// tests/cuda_emu: the host mirror refuses to load a library that identifies itself like this (_native.load_library)
extern "C" const char* svmb200_version(void) { return "svmb200 0.1 (HOST EMULATION, test infrastructure only)"; }
#endif

extern "C" int svmb200_device_count(int* count) {
    SVM_CHECK_ARG(count != nullptr, "null argument");
    *count = 0;
    SVM_CUDA(cudaGetDeviceCount(count));
    return SVMB200_OK;
}

extern "C" int svmb200_ctx_create(int device, svmb200_ctx** out) {
    SVM_CHECK_ARG(out != nullptr, "null argument");
    *out = nullptr;
    int count = 0;
    SVM_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) {
        svmb200_set_error("no CUDA device available (this library has no CPU fallback)");
        return SVMB200_ERR_CUDA;
    }
    SVM_CHECK_ARG(device >= 0 && device < count, "device index out of range");
    SVM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SVM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        svmb200_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                          prop.minor);
        return SVMB200_ERR_CUDA;
    }
    svmb200_ctx* ctx = new svmb200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        svmb200_set_error("cannot create stream/events: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return SVMB200_ERR_CUDA;
    }
    if (const char* ev = getenv("SVMB200_SYMMETRIC")) ctx->symmetric = atoi(ev) != 0;
    *out = ctx;
    return SVMB200_OK;
}

// Opt-in: solvers created on this context from now on read only the upper triangle of their matrix (half the HBM bytes per
// iteration; K2s, k2_symv.cuh).  The caller asserts symmetry -- every Hessian of the SVM dual is symmetric
// (optiml/ml/svm/_base.py:554, 1098-1099).  Applies to solvers that hold the whole matrix on one rank; sharded solvers and
// lockstep batches keep the full pass.
extern "C" int svmb200_ctx_set_symmetric(svmb200_ctx* ctx, int on) {
    SVM_CHECK_ARG(ctx != nullptr, "null context");
    ctx->symmetric = on != 0;
    return SVMB200_OK;
}

extern "C" int svmb200_ctx_get_symmetric(svmb200_ctx* ctx, int* on) {
    SVM_CHECK_ARG(ctx != nullptr && on != nullptr, "null argument");
    *on = ctx->symmetric ? 1 : 0;
    return SVMB200_OK;
}

int svm_scratch_reserve(svmb200_ctx* ctx, void** buf, size_t* have, size_t need) {
    if (*have >= need) return SVMB200_OK;
    SVM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    SVM_CUDA(cudaMalloc(buf, need));
    *have = need;
    return SVMB200_OK;
}

extern "C" int svmb200_ctx_destroy(svmb200_ctx* ctx) {
    if (!ctx) return SVMB200_OK;
    cudaSetDevice(ctx->device);
    svmb200_comm_destroy(ctx);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        svm_release_matvec_scratch(ctx);
        svm_release_solver_cache(ctx);
        if (ctx->norm_buf) cudaFree(ctx->norm_buf);
        if (ctx->mp_buf) cudaFree(ctx->mp_buf);
        if (ctx->batch_buf) cudaFree(ctx->batch_buf);
        if (ctx->idx_buf) cudaFree(ctx->idx_buf);
        if (ctx->gbar) cudaFree(ctx->gbar);
        if (ctx->persist_buf) cudaFree(ctx->persist_buf);
        svm_release_variance_cache(ctx);
        cudaStreamDestroy(ctx->stream);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    delete ctx;
    return SVMB200_OK;
}

extern "C" int svmb200_ctx_info(svmb200_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes,
                                size_t* total_bytes) {
    SVM_TRY(svm_use(ctx));
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    size_t f = 0, t = 0;
    SVM_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SVMB200_OK;
}

extern "C" int svmb200_malloc(svmb200_ctx* ctx, size_t bytes, void** dptr) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(dptr != nullptr, "null argument");
    *dptr = nullptr;
    SVM_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return SVMB200_OK;
}
extern "C" int svmb200_free(svmb200_ctx* ctx, void* dptr) {
    SVM_TRY(svm_use(ctx));
    if (dptr) {
        SVM_CUDA(cudaStreamSynchronize(ctx->stream));
        SVM_CUDA(cudaFree(dptr));
    }
    return SVMB200_OK;
}
extern "C" int svmb200_memset(svmb200_ctx* ctx, void* dptr, int value, size_t bytes) {
    SVM_TRY(svm_use(ctx));
    SVM_CUDA(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return SVMB200_OK;
}
extern "C" int svmb200_h2d(svmb200_ctx* ctx, void* dptr, const void* host, size_t bytes) {
    SVM_TRY(svm_use(ctx));
    SVM_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SVM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVMB200_OK;
}
extern "C" int svmb200_d2h(svmb200_ctx* ctx, void* host, const void* dptr, size_t bytes) {
    SVM_TRY(svm_use(ctx));
    SVM_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SVM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVMB200_OK;
}
extern "C" int svmb200_d2d(svmb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    SVM_TRY(svm_use(ctx));
    SVM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SVMB200_OK;
}
// copy between two contexts of this process (replicating X over the ranks of a single-process group): waits for the
// source context's stream on the host, then enqueues the copy on the destination context's stream
extern "C" int svmb200_copy_peer(svmb200_ctx* dst_ctx, void* dst, svmb200_ctx* src_ctx, const void* src, size_t bytes) {
    SVM_CHECK_ARG(dst_ctx != nullptr && src_ctx != nullptr && dst != nullptr && src != nullptr, "null argument");
    SVM_TRY(svm_use(src_ctx));
    SVM_CUDA(cudaStreamSynchronize(src_ctx->stream));
    SVM_TRY(svm_use(dst_ctx));
#ifndef SVMB200_HOST_EMULATION
    if (dst_ctx->device != src_ctx->device) {
        SVM_CUDA(cudaMemcpyPeerAsync(dst, dst_ctx->device, src, src_ctx->device, bytes, dst_ctx->stream));
        return SVMB200_OK;
    }
#endif
    SVM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, dst_ctx->stream));
    return SVMB200_OK;
}
extern "C" int svmb200_sync(svmb200_ctx* ctx) {
    SVM_TRY(svm_use(ctx));
    SVM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVMB200_OK;
}
extern "C" int svmb200_host_alloc_pinned(size_t bytes, void** host) {
    SVM_CHECK_ARG(host != nullptr, "null argument");
    SVM_CUDA(cudaMallocHost(host, bytes ? bytes : 16));
    return SVMB200_OK;
}
extern "C" int svmb200_host_free_pinned(void* host) {
    if (host) SVM_CUDA(cudaFreeHost(host));
    return SVMB200_OK;
}
extern "C" int svmb200_timer_start(svmb200_ctx* ctx) {
    SVM_TRY(svm_use(ctx));
    SVM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return SVMB200_OK;
}
extern "C" int svmb200_timer_stop_ms(svmb200_ctx* ctx, float* ms) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(ms != nullptr, "null argument");
    SVM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    SVM_CUDA(cudaEventSynchronize(ctx->ev1));
    SVM_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return SVMB200_OK;
}
extern "C" int svmb200_launch_count(svmb200_ctx* ctx, uint64_t* launches) {
    SVM_CHECK_ARG(ctx != nullptr && launches != nullptr, "null argument");
    *launches = ctx->launches;
    return SVMB200_OK;
}

// ---------------------------------------------------------------------------------------------
// small RAII helper for the host-pointer entry points
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        SVM_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return SVMB200_OK;
    }
    double* d() { return static_cast<double*>(p); }
};

// upload an n x d host matrix into a device matrix with an even leading dimension (TMA: 16-byte
// row stride); pad columns are zero
int upload_padded(svmb200_ctx* ctx, const double* host, int64_t n, int64_t d, DevBuf& buf, int64_t* ld) {
    const int64_t l = d + (d & 1);
    SVM_TRY(buf.alloc((size_t)n * l * sizeof(double)));
    if (l != d) SVM_CUDA(cudaMemsetAsync(buf.p, 0, (size_t)n * l * sizeof(double), ctx->stream));
    SVM_CUDA(cudaMemcpy2DAsync(buf.p, (size_t)l * sizeof(double), host, (size_t)d * sizeof(double),
                               (size_t)d * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    *ld = l;
    return SVMB200_OK;
}
}  // namespace

extern "C" int svmb200_kernel_matrix_host(svmb200_ctx* ctx, const double* x_host, int64_t nx, const double* y_host,
                                          int64_t ny, int64_t d, int kernel, double gamma, double coef0, double degree,
                                          double* out_host) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(x_host && out_host && nx > 0 && d > 0, "bad argument");
    const bool same = (y_host == nullptr) || (y_host == x_host && ny == nx);
    if (same) ny = nx;
    SVM_CHECK_ARG(ny > 0, "bad argument");
    DevBuf dx, dy;
    int64_t ldx = 0, ldy = 0;
    SVM_TRY(upload_padded(ctx, x_host, nx, d, dx, &ldx));
    if (!same) SVM_TRY(upload_padded(ctx, y_host, ny, d, dy, &ldy));
    const int64_t ldo = svmb200_padded_ld(ny);
    // row blocks of at most ~1 GiB so that huge Gram matrices never need a full device copy
    int64_t rows_per_block = std::max<int64_t>(128, ((1ll << 30) / (ldo * 8)) / 128 * 128);
    rows_per_block = std::min(rows_per_block, round_up64(nx, 128));
    DevBuf dout;
    SVM_TRY(dout.alloc((size_t)rows_per_block * ldo * sizeof(double)));
    for (int64_t r0 = 0; r0 < nx; r0 += rows_per_block) {
        const int64_t nr = std::min(rows_per_block, nx - r0);
        SVM_TRY(svmb200_gram(ctx, dx.d(), nx, ldx, same ? dx.d() : dy.d(), ny, same ? ldx : ldy, d, same ? 1 : 0, kernel,
                             gamma, coef0, degree, nullptr, nullptr, 0.0, r0, nr, dout.d(), ldo));
        SVM_CUDA(cudaMemcpy2DAsync(out_host + r0 * ny, (size_t)ny * sizeof(double), dout.p, (size_t)ldo * sizeof(double),
                                   (size_t)ny * sizeof(double), (size_t)nr, cudaMemcpyDeviceToHost, ctx->stream));
        SVM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SVMB200_OK;
}

// out[j] = sum_i dual_coef[i] k(SV_i, X_j) + b with the support vectors already in HBM (dsv: nsv x d, leading dimension
// ldsv).  k(SV_i, X_j) == k(X_j, SV_i) term by term for every kernel here, so the block K(X_chunk, SV) (rows = test
// points) is built with K1 and contracted with dual_coef by the streaming matvec K2; blocks of <= 2 GiB.
static int decision_core(svmb200_ctx* ctx, const double* dsv, int64_t nsv, int64_t ldsv, const double* dual_coef_host,
                         const double* x_host, int64_t m, int64_t d, int kernel, double gamma, double coef0, double degree,
                         double intercept, double* out_host) {
    DevBuf dx, dcoef, dblock, dres;
    int64_t ldx = 0;
    SVM_TRY(upload_padded(ctx, x_host, m, d, dx, &ldx));
    const int64_t ldo = svmb200_padded_ld(nsv);
    SVM_TRY(dcoef.alloc((size_t)ldo * sizeof(double)));
    SVM_CUDA(cudaMemsetAsync(dcoef.p, 0, (size_t)ldo * sizeof(double), ctx->stream));
    SVM_CUDA(cudaMemcpyAsync(dcoef.p, dual_coef_host, (size_t)nsv * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int64_t rows_per_block = std::max<int64_t>(128, ((1ll << 31) / (ldo * 8)) / 128 * 128);
    rows_per_block = std::min(rows_per_block, round_up64(m, 128));
    SVM_TRY(dblock.alloc((size_t)rows_per_block * ldo * sizeof(double)));
    SVM_TRY(dres.alloc((size_t)m * sizeof(double)));
    for (int64_t r0 = 0; r0 < m; r0 += rows_per_block) {
        const int64_t nr = std::min(rows_per_block, m - r0);
        SVM_TRY(svmb200_gram(ctx, dx.d(), m, ldx, dsv, nsv, ldsv, d, 0, kernel, gamma, coef0, degree, nullptr,
                             nullptr, 0.0, r0, nr, dblock.d(), ldo));
        SVM_TRY(svm_launch_matvec(ctx, dblock.d(), nr, ldo, dcoef.d(), dres.d() + r0, nullptr));
    }
    SVM_CUDA(cudaMemcpyAsync(out_host, dres.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SVM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t j = 0; j < m; ++j) out_host[j] += intercept;
    return SVMB200_OK;
}

extern "C" int svmb200_decision(svmb200_ctx* ctx, const double* sv_host, int64_t nsv, const double* dual_coef_host,
                                const double* x_host, int64_t m, int64_t d, int kernel, double gamma, double coef0,
                                double degree, double intercept, double* out_host) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(sv_host && dual_coef_host && x_host && out_host, "null argument");
    SVM_CHECK_ARG(nsv > 0 && m > 0 && d > 0, "empty operand");
    DevBuf dsv;
    int64_t ldsv = 0;
    SVM_TRY(upload_padded(ctx, sv_host, nsv, d, dsv, &ldsv));
    return decision_core(ctx, dsv.d(), nsv, ldsv, dual_coef_host, x_host, m, d, kernel, gamma, coef0, degree, intercept,
                         out_host);
}

extern "C" int svmb200_decision_device(svmb200_ctx* ctx, const double* dsv, int64_t nsv, int64_t ldsv,
                                       const double* dual_coef_host, const double* x_host, int64_t m, int64_t d, int kernel,
                                       double gamma, double coef0, double degree, double intercept, double* out_host) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(dsv && dual_coef_host && x_host && out_host, "null argument");
    SVM_CHECK_ARG(nsv > 0 && m > 0 && d > 0 && ldsv >= d && ldsv % 2 == 0, "bad operand");
    return decision_core(ctx, dsv, nsv, ldsv, dual_coef_host, x_host, m, d, kernel, gamma, coef0, degree, intercept, out_host);
}

extern "C" int svmb200_bcqp_pg_host(svmb200_ctx* ctx, const double* Q_host, const double* q_host, const double* lb_host,
                                    const double* ub_host, const double* x0_host, int64_t n, double eps, int64_t max_iter,
                                    double* x_out, double* g_out, double* f_hist, double* ng_hist, int64_t* iter,
                                    int* status) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(Q_host && q_host && ub_host && n > 1, "bad argument");
    if (ctx->nranks != 1) {
        svmb200_set_error("svmb200_bcqp_pg_host is a single-GPU convenience entry point");
        return SVMB200_ERR_STATE;
    }
    const int64_t ld = svmb200_padded_ld(n);
    DevBuf dq;
    SVM_TRY(dq.alloc((size_t)n * ld * sizeof(double)));
    SVM_CUDA(cudaMemsetAsync(dq.p, 0, (size_t)n * ld * sizeof(double), ctx->stream));
    SVM_CUDA(cudaMemcpy2DAsync(dq.p, (size_t)ld * sizeof(double), Q_host, (size_t)n * sizeof(double),
                               (size_t)n * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    svmb200_pg* pg = nullptr;
    SVM_TRY(svmb200_pg_create(ctx, dq.d(), n, ld, 0, n, SVMB200_HESSIAN_PLAIN, q_host, lb_host, ub_host, x0_host, eps,
                              max_iter, &pg));
    int rc = svmb200_pg_run(pg, -1, iter, status);
    if (rc == SVMB200_OK) rc = svmb200_pg_state(pg, x_out, g_out, nullptr, nullptr);
    if (rc == SVMB200_OK && (f_hist || ng_hist)) rc = svmb200_pg_history(pg, f_hist, ng_hist, nullptr);
    svmb200_pg_destroy(pg);
    return rc;
}
