// K3: the O(n) vector phases of the solvers -- device code (included by pg.cu only).
//   pg_vector_body   projected gradient        (optiml/opti/constrained/projected_gradient.py:76-143)
//   fw_vector_body   Frank-Wolfe               (optiml/opti/constrained/frank_wolfe.py:88-165)
//   al_vector_body   augmented Lagrangian + stochastic update rules (constrained/_base.py:224-410, stochastic/*.py)
// and their kernel entry points: one problem per launch, or a batch of problems in lockstep (blockIdx.y).
#pragma once
#include "k2_matvec.cuh"
#include "al_math.cuh"

// ------------------------------------------------------------------------------------------ K3
// Grid-wide vector phase, no inter-CTA synchronisation: a CTA first reduces (redundantly, in a fixed
// order) the per-row-block shares of u'w written by K2 and the per-CTA partials of |d|^2, x'(g+q) and
// max_t written by the previous K3 launch, takes the step on its slice, and leaves its own partials
// for the next launch.
constexpr int VP_NT = 256;
constexpr int VP_MAXC = 128;   // at most this many CTAs (fixed: the reduction shape must not follow the GPU)
constexpr int VP_ELEMS = 512;  // target elements per CTA

struct PGDeviceState {
    long long iter;  // index of the state whose f/ng are stored below
    int done;
    int status;
    double f, ng, s, maxt, t, den;
    double best_lb;  // Frank-Wolfe: best lower bound so far
    // augmented Lagrangian: launch index that raised `done` (a CTA of that same launch that starts late must not
    // leave early: it still owes its slice of the gradient) and the multiplier of the equality row
    long long done_k;
    double mu;
};

struct VecArgs {
    double *x, *g, *d, *u;
    const double *q, *lb, *ub;
    const double* gathered;  // per rank: [rpr results w | rpr/MV_GROUP shares of u'w]
    long long rpr, stride;
    double* part;            // 2 x 3 x VP_MAXC : |d|^2, x'(g+q), max_t per CTA, double-buffered by state parity
    double *hist_f, *hist_ng;
    long long hist_cap;
    PGDeviceState* st;
    long long n;   // matrix dimension
    int svr;       // 0: nvars = n ; 1: nvars = 2n, Q = [[M,-M],[-M,M]]
    int nctas;
    double eps;
    long long max_iter;
    double fw_t;  // Frank-Wolfe stabilisation parameter t in [0, 1)
    // fused exchange: `gathered_ll` (tagged 16-byte entries, see ll_store) replaces `gathered`
    const ulonglong2* gathered_ll;
    unsigned tag;
    int* fault;
    long long ll_pstride;  // batched solves: entries between the two parity copies of this problem's gathered buffer
    // label signs (nvars entries of +-1, or null): the resident matrix is M and the problem is posed on
    // Q = (s s') o M.  Q u = s o (M (s o u)) is exact for s = +-1, so the vector kernels sign w on the way in and
    // u on the way out and K2 never sees the signs -- several such problems can share one pass over M (one-vs-rest)
    const double* sgn;
};

// (L1-bypassing load: inside the persistent kernel of k_persistent.cuh the products come from other CTAs of the SAME
// launch, and a line cached two iterations ago would be stale; between separate launches it makes no difference)
__device__ __forceinline__ double gathered_at(const VecArgs& a, size_t idx) {
    return a.gathered_ll != nullptr ? ll_load(a.gathered_ll + idx, a.tag, a.fault) : __ldcg(a.gathered + idx);
}

// w = Q u for Q = (s s') o M:  the vector handed to K2 is s o u and the product that comes back is signed again
// (multiplications by +-1 are exact; label signs never combine with the SVR block layout)
__device__ __forceinline__ double apply_sign(const VecArgs& a, long long j, double v) {
    return a.sgn != nullptr ? __dmul_rn(a.sgn[j], v) : v;
}

enum { VP_INIT = 0, VP_STEP = 1, VP_FINALISE = 2 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct Quad {
    double a, b, c, m;  // three sums and one min
};

// all threads receive the block-wide result; fixed tree: lanes (butterfly), then warps in order
__device__ __forceinline__ Quad block_reduce(Quad v, double (*sm)[4]) {
    v.a = warp_sum(v.a);
    v.b = warp_sum(v.b);
    v.c = warp_sum(v.c);
    v.m = warp_min(v.m);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();  // protect sm from the previous use
    if (lane == 0) {
        sm[wid][0] = v.a;
        sm[wid][1] = v.b;
        sm[wid][2] = v.c;
        sm[wid][3] = v.m;
    }
    __syncthreads();
    Quad r;
    r.a = r.b = r.c = 0.0;
    r.m = INFINITY;
#pragma unroll
    for (int i = 0; i < VP_NT / 32; ++i) {
        r.a = __dadd_rn(r.a, sm[i][0]);
        r.b = __dadd_rn(r.b, sm[i][1]);
        r.c = __dadd_rn(r.c, sm[i][2]);
        r.m = fmin(r.m, sm[i][3]);
    }
    return r;
}

// element-wise pieces, written with explicit round-to-nearest intrinsics so that nvcc cannot
// contract them into FMAs: NumPy rounds t*d and x + (t*d) separately (projected_gradient.py:129).
__device__ __forceinline__ double axpy_rn(double a, double x, double y) { return __dadd_rn(y, __dmul_rn(a, x)); }

__device__ __forceinline__ double project_dir(double g, double x, double lb, double ub) {
    // projected_gradient.py:83-87
    double d = -g;
    if ((__dsub_rn(ub, x) <= 1e-12) && (d > 0.0)) d = 0.0;
    if ((__dsub_rn(x, lb) <= 1e-12) && (d < 0.0)) d = 0.0;
    return d;
}

__device__ __forceinline__ void tail_accumulate(Quad& a, double d, double x, double g, double q, double lb, double ub) {
    a.a = __dadd_rn(a.a, __dmul_rn(d, d));
    a.b = __dadd_rn(a.b, __dmul_rn(x, __dadd_rn(g, q)));
    // projected_gradient.py:111-114 (correctly rounded IEEE division, exact min): (ub - x) / d where d > 0, (lb - x) / d
    // where d < 0, nothing where d == 0.  ONE branch-free division per variable -- the divide is a ~30-instruction dependent
    // sequence, and with a branch per sign the compiler emitted two of them per variable and could not interleave the
    // variables of a thread; same operands, same quotient, same bits.
    const double num = __dsub_rn(d > 0.0 ? ub : lb, x);
    const double quot = __ddiv_rn(num, d != 0.0 ? d : 1.0);
    a.m = (d != 0.0) ? fmin(a.m, quot) : a.m;
}

// One variable (and, for the SVR block layout, its twin j + n) of the vector phase.  The operands do not depend on the
// step length, so a thread loads those of its first PG_PREFETCH variables BEFORE the reductions that produce t: the
// kernel is a chain of dependent latencies (loads -> block reduce -> divide -> loads -> update -> block reduce), and this
// takes one memory round trip (~1 us with the vectors evicted from L2 by the matrix stream) off it.  Same arithmetic,
// same order of accumulation per thread.
constexpr int PG_PREFETCH = 2;   // covers the whole slice of a CTA when n <= 65 536 (512 variables per CTA)

struct PGElem {
    double w, x, q, lb, ub, d, g;
    double x2, q2, lb2, ub2, d2, g2;
};

// the operands that the PREVIOUS vector launch (or the host, before INIT) left behind: everything but w
template <int MODE>
__device__ __forceinline__ PGElem pg_load_elem_own(const VecArgs& a, long long j, long long n) {
    PGElem e;
    e.w = 0.0;
    // x, d, g change every iteration and are read here possibly BEFORE griddepcontrol.wait: L2 loads (no L1 line of an
    // earlier launch can be served); q and the bounds are constants of the solve
    e.x = __ldcg(a.x + j);
    e.q = a.q[j];
    e.lb = a.lb[j];
    e.ub = a.ub[j];
    e.d = e.g = 0.0;
    if (MODE != VP_INIT) {
        e.d = __ldcg(a.d + j);
        e.g = __ldcg(a.g + j);
    }
    e.x2 = e.q2 = e.lb2 = e.ub2 = e.d2 = e.g2 = 0.0;
    if (a.svr) {
        const long long i2 = j + n;
        e.x2 = __ldcg(a.x + i2);
        e.q2 = a.q[i2];
        e.lb2 = a.lb[i2];
        e.ub2 = a.ub[i2];
        if (MODE != VP_INIT) {
            e.d2 = __ldcg(a.d + i2);
            e.g2 = __ldcg(a.g + i2);
        }
    }
    return e;
}

// w_j of the product this launch consumes (written by the matvec launch in front of it / by the peers)
__device__ __forceinline__ double pg_load_w(const VecArgs& a, long long j, unsigned rpr) {
    const unsigned rk = (unsigned)j / rpr;
    return gathered_at(a, (size_t)rk * a.stride + ((unsigned)j - rk * rpr));
}

template <int MODE>
__device__ __forceinline__ PGElem pg_load_elem(const VecArgs& a, long long j, long long n, unsigned rpr) {
    PGElem e = pg_load_elem_own<MODE>(a, j, n);
    e.w = pg_load_w(a, j, rpr);
    return e;
}

template <int MODE>
__device__ __forceinline__ void pg_step_elem(const VecArgs& a, long long j, long long n, double t, const PGElem& e, Quad& acc) {
    const double wj = apply_sign(a, j, e.w);
    double x = e.x, g;
    if (MODE == VP_INIT) {
        g = __dadd_rn(wj, e.q);  // g = Q x0 + q  (opti/_base.py:291)
    } else {
        x = axpy_rn(t, e.d, x);
        g = axpy_rn(t, wj, e.g);
        a.x[j] = x;
    }
    a.g[j] = g;
    const double dn = project_dir(g, x, e.lb, e.ub);
    a.d[j] = dn;
    tail_accumulate(acc, dn, x, g, e.q, e.lb, e.ub);
    double uj = dn;
    if (a.svr) {
        const long long i2 = j + n;
        double x2 = e.x2, g2;
        if (MODE == VP_INIT) {
            g2 = __dadd_rn(-wj, e.q2);
        } else {
            x2 = axpy_rn(t, e.d2, x2);
            g2 = axpy_rn(t, -wj, e.g2);
            a.x[i2] = x2;
        }
        a.g[i2] = g2;
        const double dn2 = project_dir(g2, x2, e.lb2, e.ub2);
        a.d[i2] = dn2;
        tail_accumulate(acc, dn2, x2, g2, e.q2, e.lb2, e.ub2);
        uj = __dsub_rn(dn, dn2);
    }
    a.u[j] = apply_sign(a, j, uj);
}

template <int MODE>
__device__ __forceinline__ void pg_vector_body(const VecArgs& a, const long long k) {
    __shared__ double sm[VP_NT / 32][4];
    PGDeviceState* st = a.st;
    // `done` is raised inside the stop branch below, which every CTA of that launch takes anyway;
    // a CTA that starts late and already sees the flag returns here instead -- same outcome.
    if (*reinterpret_cast<volatile int*>(&st->done)) return;
    if (a.fault != nullptr && *reinterpret_cast<volatile int*>(a.fault)) return;  // broken exchange: see ll_load
    const int tid = threadIdx.x;
    const long long n = a.n;
    const long long chunk = (n + a.nctas - 1) / a.nctas;
    const long long j0 = (long long)blockIdx.x * chunk;
    const long long j1 = (j0 + chunk < n) ? (j0 + chunk) : n;
    const unsigned rpr = (unsigned)a.rpr, gpr = rpr / MV_GROUP;  // rows / groups per rank (n < 2^31)
    // per-CTA partials are double-buffered by state parity: this launch reads the sums of state k and leaves those of
    // state k+1 (INIT: of state 0) in the other half, so a CTA that starts late never reads a half-updated set
    const double* part_r = a.part + (size_t)(k & 1) * 3 * VP_MAXC;
    double* part_w = a.part + (size_t)((MODE == VP_INIT ? k : k + 1) & 1) * 3 * VP_MAXC;

    // Operands first.  Everything but w was left behind by the PREVIOUS vector launch, which had completed before the
    // matvec launch in front of this one even started -- so those loads are issued BEFORE griddepcontrol.wait: this grid is
    // scheduled while the matvec grid drains (programmatic dependent launch) and its operands are in flight by the time
    // the product is complete.  w and the shares of u'w are read after the wait.
    PGElem pre[PG_PREFETCH];
    if (MODE != VP_FINALISE) {
#pragma unroll
        for (int e = 0; e < PG_PREFETCH; ++e) {
            const long long j = j0 + tid + (long long)e * VP_NT;
            if (j < j1) pre[e] = pg_load_elem_own<MODE>(a, j, n);
        }
    }
    pdl_wait();               // the product this launch consumes must be complete and visible
    pdl_launch_dependents();  // the next product may be scheduled behind us; it waits for our completion itself
    if (MODE != VP_FINALISE) {
#pragma unroll
        for (int e = 0; e < PG_PREFETCH; ++e) {
            const long long j = j0 + tid + (long long)e * VP_NT;
            if (j < j1) pre[e].w = pg_load_w(a, j, rpr);
        }
    }
    double t = 0.0;
    if (MODE != VP_INIT) {
        // ---- reductions over the whole problem, identical in every CTA
        Quad r;
        r.a = r.b = r.c = 0.0;
        r.m = INFINITY;
        if (tid < a.nctas) {  // written by other CTAs: L1-bypassing loads (see gathered_at)
            r.a = __ldcg(part_r + tid);
            r.b = __ldcg(part_r + VP_MAXC + tid);
            r.m = __ldcg(part_r + 2 * VP_MAXC + tid);
        }
        if (MODE == VP_STEP) {
            // u'w: one share per 64-row group, thread-strided in global group order
            const unsigned ngrp = (unsigned)((n + MV_GROUP - 1) / MV_GROUP);
            for (unsigned b0 = tid; b0 < ngrp; b0 += 4 * VP_NT) {
                double v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned b = b0 + e * VP_NT;
                    const unsigned rk = b / gpr;
                    v[e] = b < ngrp ? gathered_at(a, (size_t)rk * a.stride + rpr + (b - rk * gpr)) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) r.c = __dadd_rn(r.c, v[e]);
            }
        }
        r = block_reduce(r, sm);
        const double s = r.a, f = 0.5 * r.b, mt = r.m, den = r.c;
        const double ng = sqrt(s);
        if (blockIdx.x == 0 && tid == 0) {
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
            if (blockIdx.x == 0 && tid == 0) {
                st->status = stop;
                __threadfence();
                st->done = 1;
            }
            return;
        }
        if (MODE == VP_FINALISE) return;
        // projected_gradient.py:121-127 ; -g'd equals d'd term by term, so the numerator is s
        t = (den <= 1e-16) ? mt : fmin(__ddiv_rn(s, den), mt);
        if (blockIdx.x == 0 && tid == 0) {
            st->t = t;
            st->den = den;
        }
    }

    // ---- x += t d ; g += t w ; new direction, partial reductions for the next state
    Quad acc;
    acc.a = acc.b = acc.c = 0.0;
    acc.m = INFINITY;
#pragma unroll
    for (int e = 0; e < PG_PREFETCH; ++e) {
        const long long j = j0 + tid + (long long)e * VP_NT;
        if (j < j1) pg_step_elem<MODE>(a, j, n, t, pre[e], acc);
    }
    for (long long j = j0 + tid + (long long)PG_PREFETCH * VP_NT; j < j1; j += VP_NT) {  // chunks > 512 (n > 65 536)
        const PGElem el = pg_load_elem<MODE>(a, j, n, rpr);
        pg_step_elem<MODE>(a, j, n, t, el, acc);
    }
    acc = block_reduce(acc, sm);
    if (tid == 0) {
        part_w[blockIdx.x] = acc.a;
        part_w[VP_MAXC + blockIdx.x] = acc.b;
        part_w[2 * VP_MAXC + blockIdx.x] = acc.m;
    }
}

// ------------------------------------------------------------------------------------------ K3' Frank-Wolfe
// Vector phase of the reference's FrankWolfe.minimize (optiml/opti/constrained/frank_wolfe.py:88-165), same
// structure as pg_vector_kernel: y = ub where g < 0 else lb; lower bound f + g'(y-x); relative gap against the
// best bound; d = y - x (y clipped to x +- t(ub-lb) when stabilised); a = den<=1e-16 ? 1 : min(-g'd/den, 1).
// Partials per CTA: x'(g+q), g'(y-x) with the UNclipped y, g'd.  History slot 2 holds the gap.
__device__ __forceinline__ void fw_direction(double g, double x, double lb, double ub, double t, double& d, double& gy) {
    const double y = (g < 0.0) ? ub : lb;                       // frank_wolfe.py:100
    gy = __dmul_rn(g, __dsub_rn(y, x));                          // term of g'(y - x), frank_wolfe.py:104
    double yc = y;
    if (t > 0.0) {                                               // frank_wolfe.py:128-130
        const double radius = __dmul_rn(t, __dsub_rn(ub, lb));
        yc = fmin(fmax(y, __dsub_rn(x, radius)), __dadd_rn(x, radius));
    }
    d = __dsub_rn(yc, x);                                        // frank_wolfe.py:135
}

template <int MODE>
__device__ __forceinline__ void fw_vector_body(const VecArgs& a, const long long k) {
    __shared__ double sm[VP_NT / 32][4];
    PGDeviceState* st = a.st;
    if (*reinterpret_cast<volatile int*>(&st->done)) return;
    if (a.fault != nullptr && *reinterpret_cast<volatile int*>(a.fault)) return;  // broken exchange: see ll_load
    const int tid = threadIdx.x;
    const long long n = a.n;
    const long long chunk = (n + a.nctas - 1) / a.nctas;
    const long long j0 = (long long)blockIdx.x * chunk;
    const long long j1 = (j0 + chunk < n) ? (j0 + chunk) : n;
    const unsigned rpr = (unsigned)a.rpr, gpr = rpr / MV_GROUP;
    const double* part_r = a.part + (size_t)(k & 1) * 3 * VP_MAXC;  // double-buffered by state parity (see pg_vector_kernel)
    double* part_w = a.part + (size_t)((MODE == VP_INIT ? k : k + 1) & 1) * 3 * VP_MAXC;

    double step = 0.0;
    if (MODE != VP_INIT) {
        Quad r;
        r.a = r.b = r.c = 0.0;
        r.m = 0.0;  // used as a fourth SUM here (g'd), not a min
        double gd_part = 0.0;
        if (tid < a.nctas) {
            r.a = part_r[tid];                 // x'(g+q)
            r.b = part_r[VP_MAXC + tid];       // g'(y-x)
            gd_part = part_r[2 * VP_MAXC + tid];  // g'd
        }
        if (MODE == VP_STEP) {
            const unsigned ngrp = (unsigned)((n + MV_GROUP - 1) / MV_GROUP);
            for (unsigned b0 = tid; b0 < ngrp; b0 += 4 * VP_NT) {
                double v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned b = b0 + e * VP_NT;
                    const unsigned rk = b / gpr;
                    v[e] = b < ngrp ? gathered_at(a, (size_t)rk * a.stride + rpr + (b - rk * gpr)) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) r.c = __dadd_rn(r.c, v[e]);
            }
        }
        // reduce g'd through the sum slot of a second Quad (block_reduce's fourth slot is a min)
        Quad r2;
        r2.a = gd_part;
        r2.b = r2.c = 0.0;
        r2.m = INFINITY;
        r = block_reduce(r, sm);
        r2 = block_reduce(r2, sm);
        const double f = 0.5 * r.a, gy = r.b, den = r.c, gd = r2.a;
        const double lbv = __dadd_rn(f, gy);
        const double prev_best = (k == 0) ? -INFINITY : st->best_lb;  // written by the previous launch
        const double best = lbv > prev_best ? lbv : prev_best;        // frank_wolfe.py:105-106
        const double gap = __ddiv_rn(__dsub_rn(f, best), fmax(fabs(f), 1.0));  // frank_wolfe.py:109
        if (blockIdx.x == 0 && tid == 0) {
            if (k < a.hist_cap) {
                a.hist_f[k] = f;
                a.hist_ng[k] = gap;
            }
            st->f = f;
            st->ng = gap;
            st->s = best;
            st->best_lb = best;  // idempotent (max): a CTA that starts late and reads the new value computes the same
            st->iter = k;
        }
        int stop = 0;
        if (gap <= a.eps) stop = SVMB200_STATUS_OPTIMAL;          // frank_wolfe.py:120-122
        else if (k >= a.max_iter) stop = SVMB200_STATUS_STOPPED;  // frank_wolfe.py:124-126
        if (stop) {
            if (blockIdx.x == 0 && tid == 0) {
                st->status = stop;
                __threadfence();
                st->done = 1;
            }
            return;
        }
        if (MODE == VP_FINALISE) return;  // best_lb is committed by the STEP launch of the same k
        step = (den <= 1e-16) ? 1.0 : fmin(__ddiv_rn(-gd, den), 1.0);  // frank_wolfe.py:145-149
        if (blockIdx.x == 0 && tid == 0) {
            st->t = step;
            st->den = den;
        }
    }
    Quad acc;
    acc.a = acc.b = acc.c = 0.0;
    acc.m = INFINITY;
    for (long long j = j0 + tid; j < j1; j += VP_NT) {
        const unsigned rk = (unsigned)j / rpr;
        const double wj = apply_sign(a, j, gathered_at(a, (size_t)rk * a.stride + ((unsigned)j - rk * rpr)));
        double x = a.x[j], q = a.q[j], g;
        if (MODE == VP_INIT) {
            g = __dadd_rn(wj, q);
        } else {
            x = axpy_rn(step, a.d[j], x);
            g = axpy_rn(step, wj, a.g[j]);
            a.x[j] = x;
        }
        a.g[j] = g;
        double dn, gyj;
        fw_direction(g, x, a.lb[j], a.ub[j], a.fw_t, dn, gyj);
        a.d[j] = dn;
        acc.a = __dadd_rn(acc.a, __dmul_rn(x, __dadd_rn(g, q)));
        acc.b = __dadd_rn(acc.b, gyj);
        acc.c = __dadd_rn(acc.c, __dmul_rn(g, dn));
        double uj = dn;
        if (a.svr) {
            const long long i2 = j + n;
            double x2 = a.x[i2], q2 = a.q[i2], g2;
            if (MODE == VP_INIT) {
                g2 = __dadd_rn(-wj, q2);
            } else {
                x2 = axpy_rn(step, a.d[i2], x2);
                g2 = axpy_rn(step, -wj, a.g[i2]);
                a.x[i2] = x2;
            }
            a.g[i2] = g2;
            double dn2, gy2;
            fw_direction(g2, x2, a.lb[i2], a.ub[i2], a.fw_t, dn2, gy2);
            a.d[i2] = dn2;
            acc.a = __dadd_rn(acc.a, __dmul_rn(x2, __dadd_rn(g2, q2)));
            acc.b = __dadd_rn(acc.b, gy2);
            acc.c = __dadd_rn(acc.c, __dmul_rn(g2, dn2));
            uj = __dsub_rn(dn, dn2);
        }
        a.u[j] = apply_sign(a, j, uj);
    }
    acc = block_reduce(acc, sm);
    if (tid == 0) {
        part_w[blockIdx.x] = acc.a;
        part_w[VP_MAXC + blockIdx.x] = acc.b;
        part_w[2 * VP_MAXC + blockIdx.x] = acc.c;
    }
}

// ------------------------------------------------------------------------------------------ K3'' augmented Lagrangian
// Vector phase of the reference's full-batch stochastic optimisers (stochastic/adagrad.py:84-125 and siblings) on
// AugmentedLagrangianQuadratic (constrained/_base.py:224-410), SURVEY.md 8f-3.  One launch per iteration after the
// streaming pass w = Q xe:  every CTA reduces the sums the previous launch left (ALSums) and the shares of xe'w,
// updates the multiplier of the equality row, applies the optimality test of the PREVIOUS iteration
// (opti/_base.py:129-149), evaluates L(xe) and the epoch limit, then takes the step on its slice: gradient, update
// rule, box multipliers at the new point, next evaluation point, and its terms of the next ALSums.  The arithmetic
// lives in al_math.cuh (shared with the host emulation of the CPU tests).  The second history array holds the
// primal cost x'Qx/2 + q'x (what ml/svm/_base.py:289-291 stores for a Lagrangian dual).
struct ALArgs {
    ALParams p;
    double *lam_lb, *lam_ub, *s1, *s2, *s3, *step, *xpre;
    const double* A;                      // equality row (nvars), null without an equality constraint
    const double *lr, *mom, *bc1, *bc2;   // per-iteration scalars, epochs + 1 entries each
    double* part;                         // 2 x AL_NSUMS x VP_MAXC per-CTA sums, double-buffered by state parity
    double* mu;                           // 2 entries, by state parity
};

template <int NV>
__device__ __forceinline__ void block_reduce_n(double (&v)[NV], double (*sm)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();  // protect sm from the previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sm[wid][i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < VP_NT / 32; ++w) r = __dadd_rn(r, sm[w][i]);
        v[i] = r;
    }
}

__device__ __forceinline__ void al_sums_to_array(const ALSums& s, double (&v)[AL_NSUMS + 1]) {
    v[0] = s.ax_pre; v[1] = s.ax_eval; v[2] = s.qx; v[3] = s.dx2; v[4] = s.dlam2; v[5] = s.c2; v[6] = s.cc2; v[7] = s.lamc;
}
__device__ __forceinline__ void al_array_to_sums(const double (&v)[AL_NSUMS + 1], ALSums& s) {
    s.ax_pre = v[0]; s.ax_eval = v[1]; s.qx = v[2]; s.dx2 = v[3]; s.dlam2 = v[4]; s.c2 = v[5]; s.cc2 = v[6]; s.lamc = v[7];
}

__device__ __forceinline__ ALElem al_load(const VecArgs& a, const ALArgs& al, long long j) {
    ALElem e;
    e.x = a.x[j];
    e.lam_lb = al.lam_lb[j];
    e.lam_ub = al.lam_ub[j];
    e.s1 = al.s1[j];
    e.s2 = al.s2[j];
    e.s3 = al.s3[j];
    e.step = al.step[j];
    return e;
}
__device__ __forceinline__ void al_store(const VecArgs& a, const ALArgs& al, long long j, const ALElem& e, double xpre) {
    a.x[j] = e.x;
    al.lam_lb[j] = e.lam_lb;
    al.lam_ub[j] = e.lam_ub;
    al.s1[j] = e.s1;
    al.s2[j] = e.s2;
    al.s3[j] = e.s3;
    al.step[j] = e.step;
    al.xpre[j] = xpre;
}

// INIT is launched with k = -1 (it prepares the sums of state 0)
template <int MODE>
__device__ __forceinline__ void al_vector_body(const VecArgs& a, const ALArgs& al, const long long k) {
    __shared__ double sm[VP_NT / 32][AL_NSUMS + 1];
    PGDeviceState* st = a.st;
    if (*reinterpret_cast<volatile int*>(&st->done)) {
        __threadfence();
        if (*reinterpret_cast<volatile long long*>(&st->done_k) != k) return;
    }
    if (a.fault != nullptr && *reinterpret_cast<volatile int*>(a.fault)) return;  // broken exchange: see ll_load
    const int tid = threadIdx.x;
    const long long n = a.n;
    const long long chunk = (n + a.nctas - 1) / a.nctas;
    const long long j0 = (long long)blockIdx.x * chunk;
    const long long j1 = (j0 + chunk < n) ? (j0 + chunk) : n;
    const unsigned rpr = (unsigned)a.rpr, gpr = rpr / MV_GROUP;
    const ALParams& p = al.p;
    double* part_w = al.part + (size_t)((k + 1) & 1) * AL_NSUMS * VP_MAXC;  // sums of state k+1
    double v[AL_NSUMS + 1];

    if (MODE == VP_INIT) {
        ALSums acc = {};
        for (long long j = j0 + tid; j < j1; j += VP_NT) {
            const double x = a.x[j];
            al_init_sums(x, a.q[j], al.A ? al.A[j] : 0.0, a.lb[j], a.ub[j], acc);
            double uj = x;
            if (a.svr) {
                const long long i2 = j + n;
                const double x2 = a.x[i2];
                al_init_sums(x2, a.q[i2], al.A ? al.A[i2] : 0.0, a.lb[i2], a.ub[i2], acc);
                uj = __dsub_rn(x, x2);
            }
            a.u[j] = apply_sign(a, j, uj);
        }
        al_sums_to_array(acc, v);
        v[AL_NSUMS] = 0.0;
        block_reduce_n<AL_NSUMS + 1>(v, sm);
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < AL_NSUMS; ++i) part_w[i * VP_MAXC + blockIdx.x] = v[i];
        }
        return;
    }

    // ---- sums over the whole problem, identical in every CTA
    const double* part_r = al.part + (size_t)(k & 1) * AL_NSUMS * VP_MAXC;
#pragma unroll
    for (int i = 0; i < AL_NSUMS; ++i) v[i] = tid < a.nctas ? part_r[i * VP_MAXC + tid] : 0.0;
    {
        // xe'w: one share per 64-row group, thread-strided in global group order
        double xw = 0.0;
        const unsigned ngrp = (unsigned)((n + MV_GROUP - 1) / MV_GROUP);
        for (unsigned b0 = tid; b0 < ngrp; b0 += 4 * VP_NT) {
            double sh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned b = b0 + e * VP_NT;
                const unsigned rk = b / gpr;
                sh[e] = b < ngrp ? gathered_at(a, (size_t)rk * a.stride + rpr + (b - rk * gpr)) : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) xw = __dadd_rn(xw, sh[e]);
        }
        v[AL_NSUMS] = xw;
    }
    block_reduce_n<AL_NSUMS + 1>(v, sm);
    ALSums S;
    al_array_to_sums(v, S);
    const double mu_prev = k >= 1 ? al.mu[(k - 1) & 1] : 0.0;
    double mu = 0.0, c_eq = 0.0, f = 0.0, pf = 0.0;
    const int rc = al_scalar_phase(p, k, a.max_iter, S, v[AL_NSUMS], mu_prev, mu, c_eq, f, pf);
    const bool lead = blockIdx.x == 0 && tid == 0;
    if (rc == AL_OPTIMAL) {
        // the previous iteration ended the run: x, the multipliers and iter = k - 1 stay, f and g are those of
        // state k - 1 (the reference breaks out before re-evaluating them)
        if (lead) {
            al.mu[k & 1] = mu;
            st->mu = mu;
            st->status = SVMB200_STATUS_OPTIMAL;
            st->done_k = k;
            __threadfence();
            st->done = 1;
        }
        return;
    }
    if (lead) {
        if (k < a.hist_cap) {
            a.hist_f[k] = f;
            a.hist_ng[k] = pf;
        }
        st->f = f;
        st->ng = pf;
        st->iter = k;
        st->mu = mu;
        al.mu[k & 1] = mu;
    }
    ALScalars sc;
    sc.mu = mu;
    sc.ax = S.ax_eval;
    sc.act_eq = c_eq != 0.0;
    sc.lr = al.lr[k];
    sc.mom = al.mom[k];
    sc.mom_next = al.mom[k + 1];
    sc.bc1 = al.bc1[k];
    sc.bc2 = al.bc2[k];
    const bool step = (rc == AL_CONTINUE) && (MODE == VP_STEP);

    ALSums acc = {};
    for (long long j = j0 + tid; j < j1; j += VP_NT) {
        const unsigned rk = (unsigned)j / rpr;
        const double wj = apply_sign(a, j, gathered_at(a, (size_t)rk * a.stride + ((unsigned)j - rk * rpr)));
        const double Aj = al.A ? al.A[j] : 0.0, qj = a.q[j], lbj = a.lb[j], ubj = a.ub[j];
        ALElem e = al_load(a, al, j);
        const double g = al_gradient(p, sc, wj, qj, Aj, lbj, ubj, e);
        a.g[j] = g;
        double uj = e.x;
        if (step) {
            double xpre;
            al_step(p, sc, g, qj, Aj, lbj, ubj, e, xpre, acc);
            al_store(a, al, j, e, xpre);
            uj = e.x;
        }
        if (a.svr) {
            const long long i2 = j + n;
            const double A2 = al.A ? al.A[i2] : 0.0, q2 = a.q[i2], lb2 = a.lb[i2], ub2 = a.ub[i2];
            ALElem e2 = al_load(a, al, i2);
            const double g2 = al_gradient(p, sc, -wj, q2, A2, lb2, ub2, e2);
            a.g[i2] = g2;
            if (step) {
                double xpre2;
                al_step(p, sc, g2, q2, A2, lb2, ub2, e2, xpre2, acc);
                al_store(a, al, i2, e2, xpre2);
            }
            uj = __dsub_rn(uj, e2.x);
        }
        if (step) a.u[j] = apply_sign(a, j, uj);
    }
    if (!step) {
        if (rc == AL_STOPPED && lead) {
            st->status = SVMB200_STATUS_STOPPED;
            st->done_k = k;
            __threadfence();
            st->done = 1;
        }
        return;
    }
    al_sums_to_array(acc, v);
    v[AL_NSUMS] = 0.0;
    block_reduce_n<AL_NSUMS + 1>(v, sm);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < AL_NSUMS; ++i) part_w[i * VP_MAXC + blockIdx.x] = v[i];
    }
}

// ------------------------------------------------------------------------------------------ kernel entry points
// One problem per launch (argument block by value), or a batch of problems that share the resident matrix and
// advance in lockstep (one-vs-rest / multi-target): blockIdx.y selects the problem, the argument
// blocks live in device memory.  A problem that has finished ignores the launches that follow (its `done` flag).
template <int MODE>
__global__ void __launch_bounds__(VP_NT) pg_vector_kernel(const VecArgs a, const long long k) {
    pg_vector_body<MODE>(a, k);   // griddepcontrol.wait / launch_dependents sit inside, behind the operand prefetch
}
template <int MODE>
__global__ void __launch_bounds__(VP_NT) fw_vector_kernel(const VecArgs a, const long long k) {
    pdl_wait();               // the product this launch consumes (and the previous vector launch) must be complete
    pdl_launch_dependents();  // the next product may be scheduled behind us; it waits for our completion itself
    fw_vector_body<MODE>(a, k);
}
template <int MODE>
__global__ void __launch_bounds__(VP_NT) al_vector_kernel(const VecArgs a, const ALArgs al, const long long k) {
    pdl_wait();
    pdl_launch_dependents();
    al_vector_body<MODE>(a, al, k);
}
// The argument blocks are constant over a run except for the fused exchange: the tag and the parity of the gathered
// buffer change with every product and come as launch parameters.
__device__ __forceinline__ VecArgs batch_args(const VecArgs* __restrict__ args, unsigned tag, int parity) {
    VecArgs a = args[blockIdx.y];
    if (a.gathered_ll != nullptr) {
        a.gathered_ll += (size_t)parity * a.ll_pstride;
        a.tag = tag;
    }
    return a;
}
template <int MODE>
__global__ void __launch_bounds__(VP_NT) pg_vector_batch_kernel(const VecArgs* __restrict__ args, const long long k,
                                                                const unsigned tag, const int parity) {
    const VecArgs a = batch_args(args, tag, parity);   // argument blocks: written by a copy long before this launch
    pg_vector_body<MODE>(a, k);
}
template <int MODE>
__global__ void __launch_bounds__(VP_NT) fw_vector_batch_kernel(const VecArgs* __restrict__ args, const long long k,
                                                                const unsigned tag, const int parity) {
    pdl_wait();
    pdl_launch_dependents();
    const VecArgs a = batch_args(args, tag, parity);
    fw_vector_body<MODE>(a, k);
}
template <int MODE>
__global__ void __launch_bounds__(VP_NT) al_vector_batch_kernel(const VecArgs* __restrict__ args,
                                                                const ALArgs* __restrict__ als, const long long k,
                                                                const unsigned tag, const int parity) {
    pdl_wait();
    pdl_launch_dependents();
    const VecArgs a = batch_args(args, tag, parity);
    const ALArgs al = als[blockIdx.y];
    al_vector_body<MODE>(a, al, k);
}
