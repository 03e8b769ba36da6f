// K2 (streaming FP64 matvec), K3 (vector phase) and the projected-gradient driver.
//
// Algorithm restated from optiml/opti/constrained/projected_gradient.py:76-143 (reference, NumPy):
//   loop:  f = x'Qx/2 + q'x ; g = Qx + q ; d = -g projected on the active box faces ; ng = |d|
//          callback ; ng <= eps -> optimal ; iter >= max_iter -> stopped
//          max_t = largest feasible step ; den = d'Qd ; t = den<=1e-16 ? max_t : min(-g'd/den, max_t)
//          x += t d ; iter += 1
// The reference streams Q three times per iteration.  Here Q is streamed ONCE per iteration:
//   w = Q d  (K2),  den = d'w,  g <- g + t w,  f = x'(g+q)/2,  -g'd == d'd  (K3).
// All O(n) work runs in one 8-CTA thread-block cluster (distributed-shared-memory reduction for
// den, global partials for the reductions that are only consumed one kernel later).
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------ K2
// One CTA owns R consecutive rows and walks the whole (padded) row length; every thread issues
// R*U independent 128-bit streaming loads per step (L1 no-allocate: Q is touched once per pass),
// the vector operand u comes through L1/L2.  Per-row reduction order depends only on (NT, U, ld),
// never on the number of GPUs, so row results are bit-identical for any row sharding.
constexpr int MV_R = 8;
constexpr int MV_NT = 256;
constexpr int MV_U = 2;

__device__ __forceinline__ double2 ld_stream_f64x2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

template <int R, int NT, int U>
__global__ void __launch_bounds__(NT) matvec_rows_kernel(const double* __restrict__ Q, long long ld, long long nrows,
                                                         const double* __restrict__ u, double* __restrict__ w,
                                                         const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    const long long row_base = (long long)blockIdx.x * R;
    const int nvec = (int)(ld >> 1);
    const double2* __restrict__ u2 = reinterpret_cast<const double2*>(u);
    const double2* rows[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        long long rr = row_base + r;
        if (rr >= nrows) rr = nrows - 1;  // clamp: read a valid row, result discarded below
        rows[r] = reinterpret_cast<const double2*>(Q + rr * ld);
    }
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;

    int c = threadIdx.x;
    // full steps: all U column groups in range
    for (; c + (U - 1) * NT < nvec; c += U * NT) {
        double2 qv[U][R];
        double2 uv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) qv[j][r] = ld_stream_f64x2(rows[r] + c + j * NT);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) uv[j] = __ldg(u2 + c + j * NT);
#pragma unroll
        for (int j = 0; j < U; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r] = fma(qv[j][r].x, uv[j].x, acc[r]);
                acc[r] = fma(qv[j][r].y, uv[j].y, acc[r]);
            }
        }
    }
    // tail
    for (; c < nvec; c += NT) {
        double2 qv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) qv[r] = ld_stream_f64x2(rows[r] + c);
        double2 uv = __ldg(u2 + c);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r] = fma(qv[r].x, uv.x, acc[r]);
            acc[r] = fma(qv[r].y, uv.y, acc[r]);
        }
    }
    // warp butterfly, then fixed-order sum over warps
    __shared__ double red[NT / 32][R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double v = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r] = v;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) red[wid][r] = acc[r];
    }
    __syncthreads();
    if (threadIdx.x < R) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) v += red[k][threadIdx.x];
        long long rr = row_base + threadIdx.x;
        if (rr < nrows) w[rr] = v;
    }
}

int svm_launch_matvec(svmb200_ctx* ctx, const double* dQ, int64_t nrows, int64_t ld, const double* du, double* dw,
                      const int* d_done) {
    if (nrows <= 0) return SVMB200_OK;
    if (ld % 2 != 0 || ld <= 0) {
        svmb200_set_error("matvec: ld must be a positive multiple of 2");
        return SVMB200_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(dQ) & 15) || (reinterpret_cast<uintptr_t>(du) & 15)) {
        svmb200_set_error("matvec: operands must be 16-byte aligned");
        return SVMB200_ERR_ARG;
    }
    const unsigned grid = (unsigned)((nrows + MV_R - 1) / MV_R);
    matvec_rows_kernel<MV_R, MV_NT, MV_U><<<grid, MV_NT, 0, ctx->stream>>>(dQ, (long long)ld, (long long)nrows, du, dw,
                                                                           d_done);
    ctx->launches++;
    SVM_CUDA(cudaGetLastError());
    return SVMB200_OK;
}

extern "C" int svmb200_matvec(svmb200_ctx* ctx, const double* dQ, int64_t nrows, int64_t ld, const double* du,
                              double* dw) {
    SVM_TRY(svm_use(ctx));
    return svm_launch_matvec(ctx, dQ, nrows, ld, du, dw, nullptr);
}

// ------------------------------------------------------------------------------------------ K3
constexpr int VP_CL = 8;      // CTAs per cluster (portable maximum)
constexpr int VP_NT = 1024;   // threads per CTA

struct PGDeviceState {
    long long iter;  // iterations completed (== index of the state whose f/ng are stored below)
    int done;
    int status;
    double f, ng, s, maxt, t, den;
};

struct VecArgs {
    double *x, *g, *d, *u;
    const double *q, *lb, *ub, *w;
    double *part_s, *part_f, *part_mt;  // VP_CL entries each
    double *hist_f, *hist_ng;
    long long hist_cap;
    PGDeviceState* st;
    long long n;         // matrix dimension
    int svr;             // 0: nvars = n ; 1: nvars = 2n, Q = [[M,-M],[-M,M]]
    double eps;
    long long max_iter;
};

enum { VP_INIT = 0, VP_STEP = 1, VP_FINALISE = 2 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// element-wise pieces, written with explicit round-to-nearest intrinsics so that nvcc cannot
// contract them into FMAs: NumPy rounds t*d and x + (t*d) separately (projected_gradient.py:129).
__device__ __forceinline__ double axpy_rn(double a, double x, double y) { return __dadd_rn(y, __dmul_rn(a, x)); }

struct TailAcc {
    double s, f, mt;
};

__device__ __forceinline__ double project_dir(double g, double x, double lb, double ub) {
    // projected_gradient.py:83-87
    double d = -g;
    if ((__dsub_rn(ub, x) <= 1e-12) && (d > 0.0)) d = 0.0;
    if ((__dsub_rn(x, lb) <= 1e-12) && (d < 0.0)) d = 0.0;
    return d;
}

__device__ __forceinline__ void tail_accumulate(TailAcc& a, double d, double x, double g, double q, double lb, double ub) {
    a.s = __dadd_rn(a.s, __dmul_rn(d, d));
    a.f = __dadd_rn(a.f, __dmul_rn(x, __dadd_rn(g, q)));
    // projected_gradient.py:111-114 (correctly rounded IEEE division, exact min)
    if (d > 0.0) a.mt = fmin(a.mt, __ddiv_rn(__dsub_rn(ub, x), d));
    else if (d < 0.0) a.mt = fmin(a.mt, __ddiv_rn(__dsub_rn(lb, x), d));
}

template <int MODE>
__global__ void __cluster_dims__(VP_CL, 1, 1) __launch_bounds__(VP_NT, 1) pg_vector_kernel(VecArgs a, long long k) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ double sm_a[VP_NT / 32], sm_b[VP_NT / 32], sm_c[VP_NT / 32];
    __shared__ double sm_den;  // this CTA's partial of d'w, read by the peers through DSMEM
    __shared__ double sm_peer[VP_CL];

    PGDeviceState* st = a.st;
    if (st->done) return;  // set by an earlier kernel: uniform over the cluster

    const long long n = a.n;
    const long long chunk = (n + VP_CL - 1) / VP_CL;
    const long long j0 = crank * chunk;
    const long long j1 = (j0 + chunk < n) ? (j0 + chunk) : n;

    double t = 0.0;
    if (MODE != VP_INIT) {
        // ---- finalise the state reached by the previous kernel: reductions in fixed CTA order
        double s = 0.0, f2 = 0.0, mt = INFINITY;
#pragma unroll
        for (int r = 0; r < VP_CL; ++r) {
            s = __dadd_rn(s, a.part_s[r]);
            f2 = __dadd_rn(f2, a.part_f[r]);
            mt = fmin(mt, a.part_mt[r]);
        }
        const double f = 0.5 * f2;
        const double ng = sqrt(s);
        if (crank == 0 && tid == 0) {
            if (k < a.hist_cap) {
                a.hist_f[k] = f;
                a.hist_ng[k] = ng;
            }
            st->f = f;
            st->ng = ng;
            st->s = s;
            st->maxt = mt;
            st->iter = k;
        }
        int stop = 0;
        if (ng <= a.eps) stop = SVMB200_STATUS_OPTIMAL;          // projected_gradient.py:100-102
        else if (k >= a.max_iter) stop = SVMB200_STATUS_STOPPED;  // projected_gradient.py:104-106
        if (stop) {
            if (crank == 0 && tid == 0) {
                st->status = stop;
                __threadfence();
                st->done = 1;
            }
            return;
        }
        if (MODE == VP_FINALISE) return;

        // ---- den = d'Qd = d'w  (projected_gradient.py:121)
        double dp = 0.0;
        for (long long j = j0 + tid; j < j1; j += VP_NT) {
            const double wj = a.w[j];
            if (a.svr) {
                // w_full = [w ; -w], d = [d1 ; d2]  =>  d'w_full = sum (d1_j - d2_j) w_j = sum u_j w_j
                dp = __dadd_rn(dp, __dmul_rn(a.u[j], wj));
            } else {
                dp = __dadd_rn(dp, __dmul_rn(a.d[j], wj));
            }
        }
        dp = warp_sum(dp);
        if (lane == 0) sm_a[wid] = dp;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int i = 0; i < VP_NT / 32; ++i) v = __dadd_rn(v, sm_a[i]);
            sm_den = v;
        }
        cluster.sync();
        if (tid < VP_CL) sm_peer[tid] = *cluster.map_shared_rank(&sm_den, tid);
        __syncthreads();
        double den = 0.0;
#pragma unroll
        for (int r = 0; r < VP_CL; ++r) den = __dadd_rn(den, sm_peer[r]);
        // projected_gradient.py:123-127 ; -g'd equals d'd term by term, so the numerator is s
        t = (den <= 1e-16) ? mt : fmin(__ddiv_rn(s, den), mt);
        if (crank == 0 && tid == 0) {
            st->t = t;
            st->den = den;
        }
    }

    // ---- x += t d ; g += t w ; new direction, partial reductions for the next state
    TailAcc acc;
    acc.s = 0.0;
    acc.f = 0.0;
    acc.mt = INFINITY;
    for (long long j = j0 + tid; j < j1; j += VP_NT) {
        const double wj = a.w[j];
        double x = a.x[j], q = a.q[j], g;
        if (MODE == VP_INIT) {
            g = __dadd_rn(wj, q);  // g = Q x0 + q  (opti/_base.py:291)
        } else {
            x = axpy_rn(t, a.d[j], x);
            g = axpy_rn(t, wj, a.g[j]);
            a.x[j] = x;
        }
        a.g[j] = g;
        const double lb = a.lb[j], ub = a.ub[j];
        const double dn = project_dir(g, x, lb, ub);
        a.d[j] = dn;
        tail_accumulate(acc, dn, x, g, q, lb, ub);
        double uj = dn;
        if (a.svr) {
            const long long i2 = j + n;
            double x2 = a.x[i2], q2 = a.q[i2], g2;
            if (MODE == VP_INIT) {
                g2 = __dadd_rn(-wj, q2);
            } else {
                x2 = axpy_rn(t, a.d[i2], x2);
                g2 = axpy_rn(t, -wj, a.g[i2]);
                a.x[i2] = x2;
            }
            a.g[i2] = g2;
            const double lb2 = a.lb[i2], ub2 = a.ub[i2];
            const double dn2 = project_dir(g2, x2, lb2, ub2);
            a.d[i2] = dn2;
            tail_accumulate(acc, dn2, x2, g2, q2, lb2, ub2);
            uj = __dsub_rn(dn, dn2);
        }
        a.u[j] = uj;
    }
    acc.s = warp_sum(acc.s);
    acc.f = warp_sum(acc.f);
    acc.mt = warp_min(acc.mt);
    if (lane == 0) {
        sm_a[wid] = acc.s;
        sm_b[wid] = acc.f;
        sm_c[wid] = acc.mt;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0, f = 0.0, mt = INFINITY;
        for (int i = 0; i < VP_NT / 32; ++i) {
            s = __dadd_rn(s, sm_a[i]);
            f = __dadd_rn(f, sm_b[i]);
            mt = fmin(mt, sm_c[i]);
        }
        a.part_s[crank] = s;
        a.part_f[crank] = f;
        a.part_mt[crank] = mt;
        if (crank == 0 && MODE != VP_INIT) st->iter = k + 1;
    }
    if (MODE == VP_STEP) cluster.sync();  // keep sm_den alive until every peer has read it
}

// ------------------------------------------------------------------------------------------ driver
struct svmb200_pg {
    svmb200_ctx* ctx = nullptr;
    const double* dQ = nullptr;
    int64_t n = 0, ld = 0, row0 = 0, nrows = 0, nvars = 0, rows_per_rank = 0;
    int svr = 0;
    double eps = 1e-6;
    int64_t max_iter = 1000;
    int64_t hist_cap = 0;
    // device buffers
    double *x = nullptr, *g = nullptr, *d = nullptr, *u = nullptr, *w = nullptr;
    double *q = nullptr, *lb = nullptr, *ub = nullptr;
    double *part = nullptr, *hist_f = nullptr, *hist_ng = nullptr;
    PGDeviceState* st = nullptr;
    PGDeviceState* st_host = nullptr;  // pinned
    // host-side cursor
    int64_t k_next = 0;     // next iteration whose STEP kernel has not been enqueued
    bool finished = false;  // device reported done
    // stats of the last run
    float last_ms = 0.f, last_mv_ms = 0.f;
    int64_t last_passes = 0;
    bool profile = false;
    std::vector<cudaEvent_t> mv_ev;
};

static VecArgs make_vec_args(svmb200_pg* pg) {
    VecArgs a;
    a.x = pg->x;
    a.g = pg->g;
    a.d = pg->d;
    a.u = pg->u;
    a.q = pg->q;
    a.lb = pg->lb;
    a.ub = pg->ub;
    a.w = pg->w;
    a.part_s = pg->part;
    a.part_f = pg->part + VP_CL;
    a.part_mt = pg->part + 2 * VP_CL;
    a.hist_f = pg->hist_f;
    a.hist_ng = pg->hist_ng;
    a.hist_cap = pg->hist_cap;
    a.st = pg->st;
    a.n = pg->n;
    a.svr = pg->svr;
    a.eps = pg->eps;
    a.max_iter = pg->max_iter;
    return a;
}

template <int MODE>
static int launch_vec(svmb200_pg* pg, long long k) {
    VecArgs a = make_vec_args(pg);
    pg_vector_kernel<MODE><<<VP_CL, VP_NT, 0, pg->ctx->stream>>>(a, k);
    pg->ctx->launches++;
    SVM_CUDA(cudaGetLastError());
    return SVMB200_OK;
}

static int pg_product(svmb200_pg* pg, bool timed) {
    // w[row0 : row0+nrows] = Q_shard u, then all ranks exchange their shards (K4)
    svmb200_ctx* ctx = pg->ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed && pg->profile) {
        SVM_CUDA(cudaEventCreate(&e0));
        SVM_CUDA(cudaEventCreate(&e1));
        SVM_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    SVM_TRY(svm_launch_matvec(ctx, pg->dQ, pg->nrows, pg->ld, pg->u, pg->w + pg->row0, &pg->st->done));
    if (e1) {
        SVM_CUDA(cudaEventRecord(e1, ctx->stream));
        pg->mv_ev.push_back(e0);
        pg->mv_ev.push_back(e1);
    }
    if (ctx->nranks > 1) SVM_TRY(svm_comm_allgather(ctx, pg->w, pg->rows_per_rank));
    pg->last_passes++;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_destroy(svmb200_pg* pg) {
    if (!pg) return SVMB200_OK;
    if (pg->ctx) cudaSetDevice(pg->ctx->device);
    for (cudaEvent_t e : pg->mv_ev) cudaEventDestroy(e);
    double* bufs[] = {pg->x, pg->g, pg->d, pg->u, pg->w, pg->q, pg->lb, pg->ub, pg->part, pg->hist_f, pg->hist_ng};
    for (double* b : bufs)
        if (b) cudaFree(b);
    if (pg->st) cudaFree(pg->st);
    if (pg->st_host) cudaFreeHost(pg->st_host);
    delete pg;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_create(svmb200_ctx* ctx, const double* dQ, int64_t n, int64_t ld, int64_t row0, int64_t nrows,
                                 int hessian, const double* q_host, const double* lb_host, const double* ub_host,
                                 const double* x0_host, double eps, int64_t max_iter, svmb200_pg** out) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(out != nullptr, "out is null");
    *out = nullptr;
    SVM_CHECK_ARG(dQ != nullptr && q_host != nullptr && ub_host != nullptr, "null input");
    SVM_CHECK_ARG(n > 1, "Q is too small");  // opti/_base.py:249-250
    SVM_CHECK_ARG(ld >= n && ld % 2 == 0, "ld must be >= n and even");
    SVM_CHECK_ARG(hessian == SVMB200_HESSIAN_PLAIN || hessian == SVMB200_HESSIAN_SVR, "bad hessian layout");
    SVM_CHECK_ARG(max_iter > 0, "max_iter must be > 0");  // opti/_base.py:73-74
    const int P = ctx->nranks;
    const int64_t rpr = (n + P - 1) / P;
    if (P > 1) {
        const int64_t exp_row0 = (int64_t)ctx->rank * rpr;
        int64_t exp_rows = n - exp_row0;
        if (exp_rows > rpr) exp_rows = rpr;
        if (exp_rows < 0) exp_rows = 0;
        SVM_CHECK_ARG(row0 == exp_row0 && nrows == exp_rows, "row shard does not match ceil(n/nranks) partition");
    } else {
        SVM_CHECK_ARG(row0 == 0 && nrows == n, "single-rank solve needs the whole matrix");
    }
    svmb200_pg* pg = new svmb200_pg();
    pg->ctx = ctx;
    pg->dQ = dQ;
    pg->n = n;
    pg->ld = ld;
    pg->row0 = row0;
    pg->nrows = nrows;
    pg->svr = hessian == SVMB200_HESSIAN_SVR;
    pg->nvars = pg->svr ? 2 * n : n;
    pg->rows_per_rank = rpr;
    pg->eps = eps;
    pg->max_iter = max_iter;
    pg->hist_cap = max_iter + 1 < (1ll << 24) ? max_iter + 1 : (1ll << 24);
    const size_t nv = (size_t)pg->nvars * sizeof(double);
    int rc = SVMB200_OK;
    auto fail = [&](int code) {
        svmb200_pg_destroy(pg);
        return code;
    };
#define PG_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            svmb200_set_error("%s:%d %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return fail(SVMB200_ERR_CUDA);                                                             \
        }                                                                                              \
    } while (0)
    PG_CUDA(cudaMalloc(&pg->x, nv));
    PG_CUDA(cudaMalloc(&pg->g, nv));
    PG_CUDA(cudaMalloc(&pg->d, nv));
    PG_CUDA(cudaMalloc(&pg->q, nv));
    PG_CUDA(cudaMalloc(&pg->lb, nv));
    PG_CUDA(cudaMalloc(&pg->ub, nv));
    PG_CUDA(cudaMalloc(&pg->u, (size_t)ld * sizeof(double)));
    PG_CUDA(cudaMalloc(&pg->w, (size_t)(rpr * P) * sizeof(double)));
    PG_CUDA(cudaMalloc(&pg->part, 3 * VP_CL * sizeof(double)));
    PG_CUDA(cudaMalloc(&pg->hist_f, (size_t)pg->hist_cap * sizeof(double)));
    PG_CUDA(cudaMalloc(&pg->hist_ng, (size_t)pg->hist_cap * sizeof(double)));
    PG_CUDA(cudaMalloc(&pg->st, sizeof(PGDeviceState)));
    PG_CUDA(cudaMallocHost(&pg->st_host, sizeof(PGDeviceState)));
    cudaStream_t s = ctx->stream;
    PG_CUDA(cudaMemsetAsync(pg->st, 0, sizeof(PGDeviceState), s));
    PG_CUDA(cudaMemsetAsync(pg->u, 0, (size_t)ld * sizeof(double), s));
    PG_CUDA(cudaMemsetAsync(pg->w, 0, (size_t)(rpr * P) * sizeof(double), s));
    PG_CUDA(cudaMemsetAsync(pg->d, 0, nv, s));
    PG_CUDA(cudaMemsetAsync(pg->g, 0, nv, s));
    // bounds / start point: opti/constrained/_base.py:61-65 (lb = 0, x0 = (lb+ub)/2)
    std::vector<double> lbv((size_t)pg->nvars, 0.0), x0v((size_t)pg->nvars), u0((size_t)n);
    if (lb_host) memcpy(lbv.data(), lb_host, nv);
    for (int64_t i = 0; i < pg->nvars; ++i) x0v[i] = x0_host ? x0_host[i] : (lbv[i] + ub_host[i]) / 2;
    for (int64_t j = 0; j < n; ++j) u0[j] = pg->svr ? x0v[j] - x0v[j + n] : x0v[j];
    PG_CUDA(cudaMemcpyAsync(pg->q, q_host, nv, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyAsync(pg->ub, ub_host, nv, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyAsync(pg->lb, lbv.data(), nv, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyAsync(pg->x, x0v.data(), nv, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyAsync(pg->u, u0.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaStreamSynchronize(s));  // staging vectors go out of scope
#undef PG_CUDA
    // g0 = Q x0 + q, first direction and its partial reductions
    pg->last_passes = 0;
    rc = pg_product(pg, false);
    if (rc == SVMB200_OK) rc = launch_vec<VP_INIT>(pg, 0);
    if (rc != SVMB200_OK) return fail(rc);
    pg->k_next = 0;
    *out = pg;
    return SVMB200_OK;
}

static int pg_poll(svmb200_pg* pg) {
    SVM_CUDA(cudaMemcpyAsync(pg->st_host, pg->st, sizeof(PGDeviceState), cudaMemcpyDeviceToHost, pg->ctx->stream));
    SVM_CUDA(cudaStreamSynchronize(pg->ctx->stream));
    if (pg->st_host->done) pg->finished = true;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_run(svmb200_pg* pg, int64_t max_new, int64_t* iter, int* status) {
    SVM_CHECK_ARG(pg != nullptr, "null solver");
    svmb200_ctx* ctx = pg->ctx;
    SVM_TRY(svm_use(ctx));
    for (cudaEvent_t e : pg->mv_ev) cudaEventDestroy(e);
    pg->mv_ev.clear();
    pg->last_passes = 0;
    pg->last_ms = pg->last_mv_ms = 0.f;
    SVM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (!pg->finished) {
        const bool to_end = max_new < 0;
        int64_t budget = to_end ? (pg->max_iter - pg->k_next) : max_new;
        if (budget > pg->max_iter - pg->k_next) budget = pg->max_iter - pg->k_next;
        // enqueue in batches; the device-side done flag turns the remainder of a batch into no-ops
        const int64_t BATCH = 64;
        while (budget > 0 && !pg->finished) {
            const int64_t nb = budget < BATCH ? budget : BATCH;
            for (int64_t i = 0; i < nb; ++i) {
                SVM_TRY(pg_product(pg, true));
                SVM_TRY(launch_vec<VP_STEP>(pg, pg->k_next));
                pg->k_next++;
            }
            budget -= nb;
            SVM_TRY(pg_poll(pg));
            if (pg->finished) pg->k_next = pg->st_host->iter;
        }
        if (!pg->finished) {
            // make the state at callback point k_next visible (f, |d|, stopping tests)
            SVM_TRY(launch_vec<VP_FINALISE>(pg, pg->k_next));
            SVM_TRY(pg_poll(pg));
        }
    }
    SVM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    SVM_CUDA(cudaEventSynchronize(ctx->ev1));
    SVM_CUDA(cudaEventElapsedTime(&pg->last_ms, ctx->ev0, ctx->ev1));
    for (size_t i = 0; i + 1 < pg->mv_ev.size(); i += 2) {
        float ms = 0.f;
        SVM_CUDA(cudaEventElapsedTime(&ms, pg->mv_ev[i], pg->mv_ev[i + 1]));
        pg->last_mv_ms += ms;
    }
    if (iter) *iter = pg->st_host->iter;
    if (status) *status = pg->st_host->done ? pg->st_host->status : SVMB200_STATUS_UNKNOWN;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_state(svmb200_pg* pg, double* x_host, double* g_host, double* f, double* ng) {
    SVM_CHECK_ARG(pg != nullptr, "null solver");
    SVM_TRY(svm_use(pg->ctx));
    cudaStream_t s = pg->ctx->stream;
    const size_t nv = (size_t)pg->nvars * sizeof(double);
    if (x_host) SVM_CUDA(cudaMemcpyAsync(x_host, pg->x, nv, cudaMemcpyDeviceToHost, s));
    if (g_host) SVM_CUDA(cudaMemcpyAsync(g_host, pg->g, nv, cudaMemcpyDeviceToHost, s));
    SVM_CUDA(cudaMemcpyAsync(pg->st_host, pg->st, sizeof(PGDeviceState), cudaMemcpyDeviceToHost, s));
    SVM_CUDA(cudaStreamSynchronize(s));
    if (f) *f = pg->st_host->f;
    if (ng) *ng = pg->st_host->ng;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_history(svmb200_pg* pg, double* f_hist_host, double* ng_hist_host, int64_t* count) {
    SVM_CHECK_ARG(pg != nullptr, "null solver");
    SVM_TRY(svm_use(pg->ctx));
    cudaStream_t s = pg->ctx->stream;
    SVM_CUDA(cudaMemcpyAsync(pg->st_host, pg->st, sizeof(PGDeviceState), cudaMemcpyDeviceToHost, s));
    SVM_CUDA(cudaStreamSynchronize(s));
    int64_t cnt = pg->st_host->iter + 1;
    if (cnt > pg->hist_cap) cnt = pg->hist_cap;
    if (f_hist_host) SVM_CUDA(cudaMemcpyAsync(f_hist_host, pg->hist_f, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (ng_hist_host) SVM_CUDA(cudaMemcpyAsync(ng_hist_host, pg->hist_ng, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    SVM_CUDA(cudaStreamSynchronize(s));
    if (count) *count = cnt;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_stats(svmb200_pg* pg, float* ms, int64_t* passes, float* matvec_ms) {
    SVM_CHECK_ARG(pg != nullptr, "null solver");
    if (ms) *ms = pg->last_ms;
    if (passes) *passes = pg->last_passes;
    if (matvec_ms) *matvec_ms = pg->last_mv_ms;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_set_profile(svmb200_pg* pg, int on) {
    SVM_CHECK_ARG(pg != nullptr, "null solver");
    pg->profile = on != 0;
    return SVMB200_OK;
}

extern "C" int svmb200_pg_device_x(svmb200_pg* pg, double** dx) {
    SVM_CHECK_ARG(pg != nullptr && dx != nullptr, "null argument");
    *dx = pg->x;
    return SVMB200_OK;
}

// ------------------------------------------------------------------------------------------ K5
extern "C" int svmb200_masked_product(svmb200_ctx* ctx, const double* dQ, int64_t n, int64_t ld, int64_t row0,
                                      int64_t nrows, const double* beta_host, double* v_host) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(dQ && beta_host && v_host, "null argument");
    SVM_CHECK_ARG(ld >= n && ld % 2 == 0, "ld must be >= n and even");
    const int P = ctx->nranks;
    const int64_t rpr = (n + P - 1) / P;
    double *du = nullptr, *dw = nullptr;
    SVM_CUDA(cudaMalloc(&du, (size_t)ld * sizeof(double)));
    if (cudaMalloc(&dw, (size_t)(rpr * P) * sizeof(double)) != cudaSuccess) {
        cudaFree(du);
        svmb200_set_error("masked_product: out of device memory");
        return SVMB200_ERR_CUDA;
    }
    int rc = SVMB200_OK;
    cudaStream_t s = ctx->stream;
    cudaMemsetAsync(du, 0, (size_t)ld * sizeof(double), s);
    cudaMemsetAsync(dw, 0, (size_t)(rpr * P) * sizeof(double), s);
    cudaMemcpyAsync(du, beta_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s);
    rc = svm_launch_matvec(ctx, dQ, nrows, ld, du, dw + row0, nullptr);
    if (rc == SVMB200_OK && P > 1) rc = svm_comm_allgather(ctx, dw, rpr);
    if (rc == SVMB200_OK) {
        cudaMemcpyAsync(v_host, dw, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            svmb200_set_error("masked_product: %s", cudaGetErrorString(e));
            rc = SVMB200_ERR_CUDA;
        }
    }
    cudaFree(du);
    cudaFree(dw);
    return rc;
}
