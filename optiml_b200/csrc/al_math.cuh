// Arithmetic of the augmented-Lagrangian vector phase (K3''): included by the CUDA kernel (pg.cu).  The header also
// compiles as plain C++ (g++ -ffp-contract=off) so that the CPU tests can replay the launch sequence on the host.
//
// Restates, for one variable at a time and with the structure of AG = [A; -I; I] made explicit:
//   AugmentedLagrangianQuadratic.function_jacobian   optiml/opti/constrained/_base.py:395-407
//   Optimizer.check_lagrangian_dual_optimality       optiml/opti/_base.py:129-149
//   the update rules of optiml/opti/unconstrained/stochastic/{adagrad,gradient_descent,rmsprop,adadelta,adam,
//   amsgrad,adamax}.py (full batch; momentum none / polyak / nesterov)
// Every product, sum, quotient and square root is a separately rounded IEEE operation, as in NumPy: on the device
// the _rn intrinsics keep nvcc from contracting them into FMAs.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define AL_HD __host__ __device__ __forceinline__
#else
#define AL_HD inline
#endif

#if defined(__CUDA_ARCH__)
AL_HD double al_add(double a, double b) { return __dadd_rn(a, b); }
AL_HD double al_sub(double a, double b) { return __dsub_rn(a, b); }
AL_HD double al_mul(double a, double b) { return __dmul_rn(a, b); }
AL_HD double al_div(double a, double b) { return __ddiv_rn(a, b); }
AL_HD double al_sqrt(double a) { return __dsqrt_rn(a); }
#else
AL_HD double al_add(double a, double b) { return a + b; }
AL_HD double al_sub(double a, double b) { return a - b; }
AL_HD double al_mul(double a, double b) { return a * b; }
AL_HD double al_div(double a, double b) { return a / b; }
AL_HD double al_sqrt(double a) { return sqrt(a); }
#endif
AL_HD double al_max(double a, double b) { return a > b ? a : b; }

// update rules / momentum types (values are part of the C ABI, include/svmb200.h)
enum { AL_ADAGRAD = 0, AL_SGD = 1, AL_RMSPROP = 2, AL_ADADELTA = 3, AL_ADAM = 4, AL_AMSGRAD = 5, AL_ADAMAX = 6 };
enum { AL_MOM_NONE = 0, AL_MOM_POLYAK = 1, AL_MOM_NESTEROV = 2 };
enum { AL_CONTINUE = 0, AL_OPTIMAL = 1, AL_STOPPED = 2 };  // == SVMB200_STATUS_*

struct ALParams {
    int rule, momentum_type, has_eq;
    double b, rho, offset, tol;
    double decay, om_decay;  // om_* = 1 - *, rounded once on the host like Python's `1. - self.decay`
    double beta1, om_beta1, beta2, om_beta2;
};

// sums over all variables that one launch leaves for the next one
struct ALSums {
    double ax_pre;   // A . x        at the point reached by the step (multiplier update, optimality test)
    double ax_eval;  // A . xe       at the next evaluation point (= x unless Nesterov: x + momentum * step)
    double qx;       // q . xe
    double dx2;      // |x - past_x|^2
    double dlam2;    // |lambda - past lambda|^2   (box multipliers only)
    double c2;       // |box constraints(x)|^2, unclipped
    double cc2;      // |max(box constraints(xe), 0)|^2
    double lamc;     // lambda . box constraints(xe)
};
constexpr int AL_NSUMS = 8;

// scalars of one state, identical in every thread
struct ALScalars {
    double mu;       // multiplier of the equality row
    double ax;       // A . xe
    int act_eq;      // equality residual != 0
    double lr, mom, mom_next, bc1, bc2;  // this iteration's step size / momentum, next iteration's momentum, 1 - beta^t
};

// per-variable state
struct ALElem {
    double x;                // evaluation point xe (the reference's self.x at the callback point)
    double lam_lb, lam_ub;   // multipliers of  -x <= -lb  and  x <= ub
    double s1, s2, s3;       // rule state: gms | moments | running max (amsgrad)
    double step;             // previous step (momentum, adadelta)
};

// Scalar phase of state k: multiplier of the equality row, optimality test of the previous iteration
// (opti/_base.py:143-147), value of the augmented Lagrangian (constrained/_base.py:398-400), epoch limit
// (adagrad.py:101-103).  Returns AL_OPTIMAL / AL_STOPPED / AL_CONTINUE.
AL_HD int al_scalar_phase(const ALParams& p, long long k, long long epochs, const ALSums& S, double xw, double mu_prev,
                          double& mu, double& c_eq, double& f, double& pf) {
    mu = mu_prev;
    if (k >= 1) {
        double c_pre = 0.0, dmu = 0.0;
        if (p.has_eq) {
            c_pre = al_sub(S.ax_pre, p.b);
            mu = al_add(mu_prev, al_mul(p.rho, c_pre));
            dmu = al_sub(mu, mu_prev);
        }
        const double ndual = al_sqrt(al_add(al_mul(dmu, dmu), S.dlam2));
        const double ndx = al_sqrt(S.dx2);
        const double nc = al_sqrt(al_add(al_mul(c_pre, c_pre), S.c2));
        if (al_add(ndual, ndx) <= p.tol || nc <= p.tol) return AL_OPTIMAL;
    }
    c_eq = p.has_eq ? al_sub(S.ax_eval, p.b) : 0.0;
    pf = al_add(al_mul(0.5, xw), S.qx);                                   // opti/_base.py:282
    const double nrm = al_sqrt(al_add(al_mul(c_eq, c_eq), S.cc2));          // np.linalg.norm(clipped_constraints)
    const double pen = al_mul(al_mul(0.5, p.rho), al_mul(nrm, nrm));        // 0.5 * rho * norm ** 2
    f = al_add(al_add(pf, al_add(al_mul(mu, c_eq), S.lamc)), pen);
    if (k + 1 >= epochs) return AL_STOPPED;
    return AL_CONTINUE;
}

// gradient of the augmented Lagrangian in one variable (constrained/_base.py:402-406); w = (Q xe)_j
AL_HD double al_gradient(const ALParams& p, const ALScalars& s, double w, double q, double A, double lb, double ub,
                         const ALElem& e) {
    const double c_lb = al_sub(-e.x, -lb), c_ub = al_sub(e.x, ub);          // AG x - bh, rows -I and I
    const double x_lb = c_lb > 0.0 ? e.x : 0.0, x_ub = c_ub > 0.0 ? e.x : 0.0;
    const double b_lb = c_lb > 0.0 ? lb : 0.0, b_ub = c_ub > 0.0 ? ub : 0.0;
    double t2 = al_add(-e.lam_lb, e.lam_ub);                                 // dual_x @ AG
    double t3 = al_mul(p.rho, al_add(x_lb, x_ub));                           // rho AG[idx]' AG[idx] x
    double t4 = al_mul(p.rho, al_add(b_lb, b_ub));                           // rho bh[idx] @ AG[idx]
    if (p.has_eq) {
        t2 = al_add(al_mul(s.mu, A), t2);
        if (s.act_eq) {
            t3 = al_add(al_mul(al_mul(p.rho, A), s.ax), t3);
            t4 = al_add(al_mul(al_mul(p.rho, p.b), A), t4);
        }
    }
    return al_sub(al_add(al_add(al_add(w, q), t2), t3), t4);
}

// One optimiser step in one variable, the multiplier update at the new point and this variable's terms of ALSums.
// On return e holds the state at the NEXT evaluation point and x_pre the point before the Nesterov jump.
AL_HD void al_step(const ALParams& p, const ALScalars& s, double g, double q, double A, double lb, double ub, ALElem& e,
                   double& x_pre, ALSums& acc) {
    const double d = -g, g2 = al_mul(g, g);
    double step2 = 0.0;
    switch (p.rule) {
        case AL_ADAGRAD:   // adagrad.py:108-111
            e.s1 = al_add(e.s1, g2);
            step2 = al_div(al_mul(s.lr, d), al_sqrt(al_add(e.s1, p.offset)));
            break;
        case AL_SGD:       // gradient_descent.py
            step2 = al_mul(s.lr, d);
            break;
        case AL_RMSPROP:   // rmsprop.py
            e.s1 = al_add(al_mul(p.decay, e.s1), al_mul(p.om_decay, g2));
            step2 = al_div(al_mul(s.lr, d), al_sqrt(al_add(e.s1, p.offset)));
            break;
        case AL_ADADELTA:  // adadelta.py: gms, then step = lr d sqrt(sms + off) / sqrt(gms + off)
            e.s1 = al_add(al_mul(p.decay, e.s1), al_mul(p.om_decay, g2));
            step2 = al_mul(al_mul(s.lr, d), al_div(al_sqrt(al_add(e.s2, p.offset)), al_sqrt(al_add(e.s1, p.offset))));
            break;
        case AL_ADAM:      // adam.py
            e.s1 = al_add(al_mul(p.beta1, e.s1), al_mul(p.om_beta1, d));
            e.s2 = al_add(al_mul(p.beta2, e.s2), al_mul(p.om_beta2, g2));
            step2 = al_div(al_mul(s.lr, al_div(e.s1, s.bc1)), al_add(al_sqrt(al_div(e.s2, s.bc2)), p.offset));
            break;
        case AL_AMSGRAD:   // amsgrad.py
            e.s1 = al_add(al_mul(p.beta1, e.s1), al_mul(p.om_beta1, d));
            e.s2 = al_add(al_mul(p.beta2, e.s2), al_mul(p.om_beta2, g2));
            e.s3 = al_max(e.s2, e.s3);
            step2 = al_div(al_mul(s.lr, e.s1), al_add(al_sqrt(e.s3), p.offset));
            break;
        default:           // AL_ADAMAX, adamax.py
            e.s1 = al_add(al_mul(p.beta1, e.s1), al_mul(p.om_beta1, d));
            e.s2 = al_max(al_mul(p.beta2, e.s2), fabs(g));
            step2 = al_div(al_mul(s.lr, al_div(e.s1, s.bc1)), al_add(e.s2, p.offset));
            break;
    }
    const double xe = e.x;
    double x_new, step;
    if (p.momentum_type == AL_MOM_POLYAK) {
        step = al_add(al_mul(s.mom, e.step), step2);
        x_new = al_add(xe, step);
    } else if (p.momentum_type == AL_MOM_NESTEROV) {
        x_new = al_add(xe, step2);
        step = al_add(al_mul(s.mom, e.step), step2);   // jump taken before this evaluation + correction
    } else {
        step = step2;
        x_new = al_add(xe, step);
    }
    if (p.rule == AL_ADADELTA) e.s2 = al_add(al_mul(p.decay, e.s2), al_mul(p.om_decay, al_mul(step, step)));
    e.step = step;
    x_pre = x_new;
    // multipliers at the new point (opti/_base.py:139-141)
    const double c_lb = al_sub(-x_new, -lb), c_ub = al_sub(x_new, ub);
    const double l_lb = al_max(al_add(e.lam_lb, al_mul(p.rho, c_lb)), 0.0);
    const double l_ub = al_max(al_add(e.lam_ub, al_mul(p.rho, c_ub)), 0.0);
    const double dl_lb = al_sub(l_lb, e.lam_lb), dl_ub = al_sub(l_ub, e.lam_ub), dx = al_sub(x_new, xe);
    acc.dlam2 = al_add(acc.dlam2, al_add(al_mul(dl_lb, dl_lb), al_mul(dl_ub, dl_ub)));
    acc.c2 = al_add(acc.c2, al_add(al_mul(c_lb, c_lb), al_mul(c_ub, c_ub)));
    acc.dx2 = al_add(acc.dx2, al_mul(dx, dx));
    acc.ax_pre = al_add(acc.ax_pre, al_mul(A, x_new));
    e.lam_lb = l_lb;
    e.lam_ub = l_ub;
    // next evaluation point
    double xn = x_new, e_lb = c_lb, e_ub = c_ub;
    if (p.momentum_type == AL_MOM_NESTEROV) {
        xn = al_add(x_new, al_mul(s.mom_next, step));
        e_lb = al_sub(-xn, -lb);
        e_ub = al_sub(xn, ub);
    }
    e.x = xn;
    const double cc_lb = al_max(e_lb, 0.0), cc_ub = al_max(e_ub, 0.0);
    acc.cc2 = al_add(acc.cc2, al_add(al_mul(cc_lb, cc_lb), al_mul(cc_ub, cc_ub)));
    acc.lamc = al_add(acc.lamc, al_add(al_mul(l_lb, e_lb), al_mul(l_ub, e_ub)));
    acc.ax_eval = al_add(acc.ax_eval, al_mul(A, xn));
    acc.qx = al_add(acc.qx, al_mul(q, xn));
}

// terms of ALSums at the start point (no step, multipliers zero)
AL_HD void al_init_sums(double x, double q, double A, double lb, double ub, ALSums& acc) {
    const double cc_lb = al_max(al_sub(-x, -lb), 0.0), cc_ub = al_max(al_sub(x, ub), 0.0);
    acc.cc2 = al_add(acc.cc2, al_add(al_mul(cc_lb, cc_lb), al_mul(cc_ub, cc_ub)));
    acc.ax_eval = al_add(acc.ax_eval, al_mul(A, x));
    acc.ax_pre = al_add(acc.ax_pre, al_mul(A, x));
    acc.qx = al_add(acc.qx, al_mul(q, x));
}
