// Host emulation of the augmented-Lagrangian vector phase  --  TEST INFRASTRUCTURE ONLY.
//
// Runs the launch sequence of csrc/pg.cu (INIT, then per iteration one product w = Q xe and one STEP launch of
// al_vector_kernel) on the CPU, CTA by CTA, with the SAME arithmetic header the CUDA kernel includes
// (optiml_b200/csrc/al_math.cuh, compiled here with g++ -ffp-contract=off).  The CPU tests compare it with the NumPy
// oracle (oracle/al_oracle.py), so the per-variable arithmetic, the order of the phases and the double-buffering by
// state parity are checked without a GPU.  It is not part of the product and nothing under optiml_b200/ loads it.
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libal_emulate.so oracle/al_emulate.cpp
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../optiml_b200/csrc/al_math.cuh"

namespace {
constexpr int VP_MAXC = 128;
constexpr int VP_ELEMS = 512;

struct Problem {
    int64_t n;      // matrix dimension
    int svr;        // nvars = 2n, Q = [[M,-M],[-M,M]]
    const double *M, *q, *lb, *ub, *A;
    ALParams p;
    const double *lr, *mom;
    int64_t epochs;
    int nctas;
};

struct State {
    std::vector<double> x, g, u, w, lam_lb, lam_ub, s1, s2, s3, step, xpre;
    std::vector<double> part;  // 2 x AL_NSUMS x VP_MAXC
    double mu[2] = {0, 0};
    int64_t iter = 0;
    int done = 0, status = 0;
    double f = 0, pf = 0, mu_last = 0;
};

double* sums_field(ALSums& s, int i) {
    double* f[AL_NSUMS] = {&s.ax_pre, &s.ax_eval, &s.qx, &s.dx2, &s.dlam2, &s.c2, &s.cc2, &s.lamc};
    return f[i];
}

void product(const Problem& P, State& S) {  // K2: w = M u, plus the total of u'w in w[n]
    double xw = 0.0;
    for (int64_t i = 0; i < P.n; ++i) {
        double acc = 0.0;
        for (int64_t j = 0; j < P.n; ++j) acc += P.M[i * P.n + j] * S.u[j];
        S.w[i] = acc;
        xw += S.u[i] * acc;
    }
    S.w[P.n] = xw;
}

// one launch; mode 0 = INIT (k = -1), 1 = STEP, 2 = FINALISE
void launch(const Problem& P, State& S, int mode, int64_t k, double* hist_f, double* hist_pf) {
    if (S.done) return;
    const int64_t n = P.n;
    const int64_t chunk = (n + P.nctas - 1) / P.nctas;
    double* part_w = S.part.data() + (size_t)((k + 1) & 1) * AL_NSUMS * VP_MAXC;
    if (mode == 0) {
        for (int c = 0; c < P.nctas; ++c) {
            const int64_t j0 = c * chunk, j1 = (j0 + chunk < n) ? j0 + chunk : n;
            ALSums acc = {};
            for (int64_t j = j0; j < j1; ++j) {
                al_init_sums(S.x[j], P.q[j], P.A ? P.A[j] : 0.0, P.lb[j], P.ub[j], acc);
                double uj = S.x[j];
                if (P.svr) {
                    const int64_t i2 = j + n;
                    al_init_sums(S.x[i2], P.q[i2], P.A ? P.A[i2] : 0.0, P.lb[i2], P.ub[i2], acc);
                    uj = S.x[j] - S.x[i2];
                }
                S.u[j] = uj;
            }
            for (int i = 0; i < AL_NSUMS; ++i) part_w[i * VP_MAXC + c] = *sums_field(acc, i);
        }
        return;
    }
    const double* part_r = S.part.data() + (size_t)(k & 1) * AL_NSUMS * VP_MAXC;
    ALSums T = {};
    for (int i = 0; i < AL_NSUMS; ++i) {
        double t = 0.0;
        for (int c = 0; c < P.nctas; ++c) t += part_r[i * VP_MAXC + c];
        *sums_field(T, i) = t;
    }
    const double mu_prev = k >= 1 ? S.mu[(k - 1) & 1] : 0.0;
    double mu = 0.0, c_eq = 0.0, f = 0.0, pf = 0.0;
    const int rc = al_scalar_phase(P.p, k, P.epochs, T, S.w[n], mu_prev, mu, c_eq, f, pf);
    if (rc == AL_OPTIMAL) {
        S.mu[k & 1] = mu;
        S.mu_last = mu;
        S.status = AL_OPTIMAL;
        S.done = 1;
        return;
    }
    hist_f[k] = f;
    hist_pf[k] = pf;
    S.f = f;
    S.pf = pf;
    S.iter = k;
    S.mu_last = mu;
    S.mu[k & 1] = mu;
    ALScalars sc;
    sc.mu = mu;
    sc.ax = T.ax_eval;
    sc.act_eq = c_eq != 0.0;
    sc.lr = P.lr[k < P.epochs ? k : P.epochs - 1];
    sc.mom = P.mom ? P.mom[k] : 0.0;
    sc.mom_next = P.mom ? P.mom[k + 1] : 0.0;
    sc.bc1 = 1.0 - pow(P.p.beta1, (double)(k + 1));
    sc.bc2 = 1.0 - pow(P.p.beta2, (double)(k + 1));
    const bool step = rc == AL_CONTINUE && mode == 1;
    auto load = [&](int64_t j) {
        ALElem e;
        e.x = S.x[j];
        e.lam_lb = S.lam_lb[j];
        e.lam_ub = S.lam_ub[j];
        e.s1 = S.s1[j];
        e.s2 = S.s2[j];
        e.s3 = S.s3[j];
        e.step = S.step[j];
        return e;
    };
    auto store = [&](int64_t j, const ALElem& e, double xpre) {
        S.x[j] = e.x;
        S.lam_lb[j] = e.lam_lb;
        S.lam_ub[j] = e.lam_ub;
        S.s1[j] = e.s1;
        S.s2[j] = e.s2;
        S.s3[j] = e.s3;
        S.step[j] = e.step;
        S.xpre[j] = xpre;
    };
    for (int c = 0; c < P.nctas; ++c) {
        const int64_t j0 = c * chunk, j1 = (j0 + chunk < n) ? j0 + chunk : n;
        ALSums acc = {};
        for (int64_t j = j0; j < j1; ++j) {
            const double wj = S.w[j];
            ALElem e = load(j);
            const double g = al_gradient(P.p, sc, wj, P.q[j], P.A ? P.A[j] : 0.0, P.lb[j], P.ub[j], e);
            S.g[j] = g;
            double uj = e.x;
            if (step) {
                double xpre;
                al_step(P.p, sc, g, P.q[j], P.A ? P.A[j] : 0.0, P.lb[j], P.ub[j], e, xpre, acc);
                store(j, e, xpre);
                uj = e.x;
            }
            if (P.svr) {
                const int64_t i2 = j + n;
                ALElem e2 = load(i2);
                const double g2 = al_gradient(P.p, sc, -wj, P.q[i2], P.A ? P.A[i2] : 0.0, P.lb[i2], P.ub[i2], e2);
                S.g[i2] = g2;
                if (step) {
                    double xpre2;
                    al_step(P.p, sc, g2, P.q[i2], P.A ? P.A[i2] : 0.0, P.lb[i2], P.ub[i2], e2, xpre2, acc);
                    store(i2, e2, xpre2);
                }
                uj = uj - e2.x;
            }
            if (step) S.u[j] = uj;
        }
        if (step)
            for (int i = 0; i < AL_NSUMS; ++i) part_w[i * VP_MAXC + c] = *sums_field(acc, i);
    }
    if (!step && rc == AL_STOPPED) {
        S.status = AL_STOPPED;
        S.done = 1;
    }
}
}  // namespace

// M: n x n row-major (the resident matrix); vectors have nvars = n (svr = 0) or 2n entries.  `finalise_every` > 0
// inserts a FINALISE launch (with its own product) before every that-many-th STEP, like the step-wise host loop.
extern "C" int al_emulate(int64_t n, int svr, const double* M, const double* q, const double* lb, const double* ub,
                          const double* x0, const double* a, double b, double rho, int rule, int momentum_type,
                          const double* lr, const double* mom, double decay, double beta1, double beta2, double offset,
                          double tol, int64_t epochs, int finalise_every, double* x_out, double* g_out, double* lam_lb_out,
                          double* lam_ub_out, double* mu_out, double* f_hist, double* pf_hist, int64_t* iter_out,
                          int* status_out) {
    Problem P;
    P.n = n;
    P.svr = svr;
    P.M = M;
    P.q = q;
    P.lb = lb;
    P.ub = ub;
    P.A = a;
    P.p.rule = rule;
    P.p.momentum_type = momentum_type;
    P.p.has_eq = a != nullptr;
    P.p.b = b;
    P.p.rho = rho;
    P.p.offset = offset;
    P.p.tol = tol;
    P.p.decay = decay;
    P.p.om_decay = 1.0 - decay;
    P.p.beta1 = beta1;
    P.p.om_beta1 = 1.0 - beta1;
    P.p.beta2 = beta2;
    P.p.om_beta2 = 1.0 - beta2;
    P.lr = lr;
    P.mom = momentum_type == AL_MOM_NONE ? nullptr : mom;
    P.epochs = epochs;
    P.nctas = (int)((n + VP_ELEMS - 1) / VP_ELEMS);
    if (P.nctas > VP_MAXC) P.nctas = VP_MAXC;
    if (P.nctas < 1) P.nctas = 1;
    const int64_t nv = svr ? 2 * n : n;
    State S;
    S.x.assign(x0, x0 + nv);
    for (auto* v : {&S.g, &S.lam_lb, &S.lam_ub, &S.s1, &S.s2, &S.s3, &S.step, &S.xpre}) v->assign(nv, 0.0);
    if (rule == AL_RMSPROP) S.s1.assign(nv, 1.0);
    S.u.assign(n, 0.0);
    S.w.assign(n + 1, 0.0);
    S.part.assign(2 * AL_NSUMS * VP_MAXC, 0.0);
    launch(P, S, 0, -1, f_hist, pf_hist);
    for (int64_t k = 0; k < epochs && !S.done; ++k) {
        if (finalise_every > 0 && k % finalise_every == 0) {
            product(P, S);
            launch(P, S, 2, k, f_hist, pf_hist);
            if (S.done) break;
        }
        product(P, S);
        launch(P, S, 1, k, f_hist, pf_hist);
    }
    const bool pre_jump = momentum_type == AL_MOM_NESTEROV && S.done && S.status == AL_OPTIMAL;
    memcpy(x_out, pre_jump ? S.xpre.data() : S.x.data(), nv * sizeof(double));
    memcpy(g_out, S.g.data(), nv * sizeof(double));
    memcpy(lam_lb_out, S.lam_lb.data(), nv * sizeof(double));
    memcpy(lam_ub_out, S.lam_ub.data(), nv * sizeof(double));
    *mu_out = S.mu_last;
    *iter_out = S.iter;
    *status_out = S.done ? S.status : 0;
    return 0;
}
