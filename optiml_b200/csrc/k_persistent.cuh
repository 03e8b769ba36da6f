// Persistent projected-gradient loop for problems whose matrix fits the SHARED MEMORY of the GPU -- device code
// (included by pg.cu only).
//
// BASELINE config C1 (n = 2 000: a 32 MB matrix) is launch-bound on the two-kernel loop: 14 us per iteration for ~1 us of
// memory traffic ("hard parts").  148 SMs x 227 KB of shared memory hold 33 MB, so here ONE
// cooperative kernel keeps the whole matrix on chip -- CTA b owns rows [b r, (b+1) r) and loads them once -- and runs all
// iterations of the solve with ONE grid barrier per iteration:
//     phase A   w = Q u for the CTA's rows, from shared memory, into the w buffer of this iteration's parity;
//               per-row products u_r w_r for the shares of u'w
//     grid barrier
//     phase B   EVERY CTA runs the whole O(n) vector phase redundantly, with the iterate IN REGISTERS: a thread owns the
//               same <= 8 variables for the whole solve (x, g, d, q, lb, ub: 48 doubles), per iteration it reads their
//               w and writes the new direction to the CTA's private u.  Nothing has to travel back before the next
//               phase A, so there is no second barrier (a first version ran phase B on the nctas CTAs of K3's grid and
//               paid two barriers and K3's chain of memory latencies: 13.6 us per iteration, no better than two
//               launches -- profiles/r2_s4_persistent_kernel_v1_ncu.txt).  The w / product buffers are double-buffered
//               by iteration parity: a CTA that is already in phase A of k + 1 writes the other buffer.
// Scope: projected gradient, plain layout without label-sign views, n <= 2016 on 148 SMs (14 rows per CTA in 221 KB;
// at most four virtual K3 CTAs of <= 512 variables) -- BASELINE config C1 and the folds / classes of small data sets; everything else takes the two-kernel loop.
// Bit-identical to the K2 + K3 loop: a row sum is the thread-strided fma chain, warp butterfly and in-order warp sum of
// matvec_seg_kernel (one column segment: ld <= 3072); the share tree is K2's group combine; phase B evaluates, for
// every virtual CTA c of K3's grid, exactly the per-thread accumulations and reduction trees of pg_vector_body (same
// element helpers, same block-reduce shape), and the cross-CTA partials that K3 passes through memory stay in shared
// memory.  CTA 0 works on the solver's own arrays and records state and history; the stopping test is evaluated
// identically by every CTA, so all leave together.
#pragma once
#include "k3_vector.cuh"

constexpr int PK_NT = 256;
constexpr int PK_RMAX = 14;   // rows per CTA (accumulators per thread): ceil(2016 / 148)
constexpr int PK_UMAX = 6;    // 128-bit operand slots per thread: ld <= 2 * PK_NT * PK_UMAX = 3072 columns
constexpr int PK_SROUNDS = (2 * PK_NT * PK_UMAX / MV_GROUP + PK_NT / 64 - 1) / (PK_NT / 64);  // share rounds: 4 groups each

constexpr int PK_VMAX = 4;                                 // virtual K3 CTAs (512 variables each)
constexpr int PK_SLOTS = 2 * PK_VMAX;                      // variables per thread: two per virtual CTA
constexpr int PK_NGRP_MAX = 2 * PK_NT * PK_UMAX / MV_GROUP;  // 48 share groups

struct PersistArgs {
    const double* Q;        // n x ld, all rows on this GPU
    long long ld, n;
    int rows_per_cta;
    double* wbuf;           // 2 x ld : w = Q u, double-buffered by iteration parity
    double* prod;           // 2 x ld : per-row products u_r * w_r, likewise
    double* priv;           // (grid - 1) private copies of u (ld each) for CTAs 1..grid-1
    unsigned* gbar;         // grid barrier: [0] arrivals, [32] generation (separate 128-byte lines), zero between launches
    VecArgs v;              // the solver's own arrays and state (CTA 0 works on them)
    long long k0;           // first iteration of this launch
    long long niter;        // iterations to run unless the stopping test fires first
};

#ifndef SVMB200_HOST_EMULATION
__device__ __forceinline__ unsigned pk_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pk_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned pk_arrive(unsigned* p) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p) : "memory");
    return old;
}
__device__ __forceinline__ void pk_backoff(unsigned long long) {}
__device__ __forceinline__ double2 pk_ld_cg_f64x2(const double2* p) {
    double2 r;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
#else
__device__ __forceinline__ unsigned pk_ld_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void pk_st_release(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned pk_arrive(unsigned* p) { return __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
__device__ __forceinline__ void pk_backoff(unsigned long long spins) { (void)emu::spin_wait(spins); }
__device__ __forceinline__ double2 pk_ld_cg_f64x2(const double2* p) { return *p; }
#endif

// all CTAs of the (cooperatively launched, hence co-resident) grid; sense by generation, the last arriver re-arms.
// The arrival is an acq_rel RMW at gpu scope (cumulative over the CTA's writes through the bar.sync before it), the
// waiters poll with relaxed loads and acquire once.
__device__ __forceinline__ void pk_grid_barrier(unsigned* bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* count = bar;
        unsigned* gen = bar + 32;
        const unsigned my_gen = *reinterpret_cast<volatile unsigned*>(gen);
        if (pk_arrive(count) == nblocks - 1) {
            *reinterpret_cast<volatile unsigned*>(count) = 0u;
            pk_st_release(gen, my_gen + 1u);
        } else {
            for (unsigned long long spins = 0; *reinterpret_cast<volatile unsigned*>(gen) == my_gen; ++spins) pk_backoff(spins);
            (void)pk_ld_acquire(gen);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PK_NT, 1) pg_persistent_kernel(const PersistArgs a) {
#ifndef SVMB200_HOST_EMULATION
    extern __shared__ __align__(16) unsigned char pk_smem_raw[];
    double* qs = reinterpret_cast<double*>(pk_smem_raw);
#else
    double* qs = reinterpret_cast<double*>(emu::dynamic_smem());
#endif
    __shared__ double red[PK_NT / 32][PK_RMAX];
    __shared__ double share_red[PK_SROUNDS][PK_NT / 32];
    __shared__ double shares[PK_NGRP_MAX];
    __shared__ double part[2][3][PK_VMAX];            // per-virtual-CTA partials, double-buffered by state parity (as K3's)
    __shared__ double smq[PK_VMAX][PK_NT / 32][4];
    __shared__ double sm[VP_NT / 32][4];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
    long long r1 = r0 + a.rows_per_cta;
    if (r1 > a.n) r1 = a.n;
    const int myrows = r1 > r0 ? (int)(r1 - r0) : 0;
    const int nvec = (int)(a.ld >> 1);
    const long long n = a.n;
    const int nctas = a.v.nctas;
    const long long chunk = (n + nctas - 1) / nctas;
    const unsigned ngrp = (unsigned)((n + MV_GROUP - 1) / MV_GROUP);
    PGDeviceState* st = a.v.st;
    const bool lead = blockIdx.x == 0;

    // ---- the CTA's rows, once: global -> shared (the matrix is read-only for the whole solve)
    {
        const double2* src = reinterpret_cast<const double2*>(a.Q + r0 * a.ld);
        double2* dst = reinterpret_cast<double2*>(qs);
        const long long total = (long long)myrows * nvec;
        for (long long i = tid; i < total; i += PK_NT) dst[i] = ld_stream_f64x2(src + i);
    }
    // ---- the iterate in registers: slot s = 2 c + e is variable c * chunk + tid + 256 e of virtual K3 CTA c (the
    // element a thread of K3's CTA c visits in its e-th loop trip); u in a private copy (CTA 0: the solver's own)
    double* u_priv = lead ? a.v.u : a.priv + (size_t)(blockIdx.x - 1) * (size_t)a.ld;
    long long slot_j[PK_SLOTS];
    bool slot_on[PK_SLOTS];
    double xs[PK_SLOTS], gs[PK_SLOTS], ds[PK_SLOTS], qv_[PK_SLOTS], lbs[PK_SLOTS], ubs[PK_SLOTS];
#pragma unroll
    for (int sl = 0; sl < PK_SLOTS; ++sl) {
        const int c = sl >> 1, e = sl & 1;
        const long long j0 = (long long)c * chunk;
        const long long j1 = (j0 + chunk < n) ? (j0 + chunk) : n;
        slot_j[sl] = j0 + tid + (long long)e * VP_NT;
        slot_on[sl] = c < nctas && slot_j[sl] < j1;
        xs[sl] = gs[sl] = ds[sl] = qv_[sl] = lbs[sl] = ubs[sl] = 0.0;
        if (slot_on[sl]) {
            const long long j = slot_j[sl];
            xs[sl] = a.v.x[j];
            gs[sl] = a.v.g[j];
            ds[sl] = a.v.d[j];
            qv_[sl] = a.v.q[j];
            lbs[sl] = a.v.lb[j];
            ubs[sl] = a.v.ub[j];
        }
    }
    if (!lead)
        for (long long i = tid; i < a.ld; i += PK_NT) u_priv[i] = a.v.u[i];
    if (tid < 2 * 3 * PK_VMAX) {
        const int par = tid / (3 * PK_VMAX), q3 = (tid / PK_VMAX) % 3, c = tid % PK_VMAX;
        part[par][q3][c] = c < nctas ? a.v.part[(size_t)par * 3 * VP_MAXC + (size_t)q3 * VP_MAXC + c] : 0.0;
    }
    __syncthreads();

    long long k = a.k0;
    for (long long it = 0; it < a.niter; ++it, ++k) {
        double* wk = a.wbuf + (size_t)(k & 1) * a.ld;
        double* pk = a.prod + (size_t)(k & 1) * a.ld;
        // ================= phase A: w = Q u for my rows (matvec_seg_kernel's arithmetic, one column segment) =================
        // The GPU issues a warp's instructions in order: independent chains only overlap if they are interleaved in the
        // PROGRAM.  Per-row / per-slot branches keep the compiler from doing that (v3 of this kernel spent 2.4 k cycles of
        // an iteration in sixteen serialised butterflies), so the loops below are branch-free: rows beyond the CTA's own
        // re-read its last row and are discarded, columns beyond the matrix multiply by a zero operand.
        double u_row = 0.0;
        if (myrows > 0) {
            double2 uv[PK_UMAX];
            const double2* u2 = reinterpret_cast<const double2*>(u_priv);
            const int mcount = (nvec + PK_NT - 1) / PK_NT;   // operand slots in use (4 at n = 2000)
#pragma unroll
            for (int m = 0; m < PK_UMAX; ++m) {
                const int c = tid + m * PK_NT;
                uv[m] = (m < mcount && c < nvec) ? u2[c] : double2{0.0, 0.0};
            }
            if (tid < myrows) u_row = u_priv[r0 + tid];   // for the row's term of u'w below: in flight with the rest
            const double2* q2 = reinterpret_cast<const double2*>(qs);
            int roff[PK_RMAX];
#pragma unroll
            for (int r = 0; r < PK_RMAX; ++r) roff[r] = (r < myrows ? r : myrows - 1) * nvec;
            double acc[PK_RMAX];
#pragma unroll
            for (int r = 0; r < PK_RMAX; ++r) acc[r] = 0.0;
#pragma unroll
            for (int m = 0; m < PK_UMAX; ++m) {
                if (m < mcount) {  // uniform
                    const int c = tid + m * PK_NT;
                    const int cc = c < nvec ? c : 0;   // (uv[m] is zero there)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double2 qv[PK_RMAX / 2];
#pragma unroll
                        for (int r = 0; r < PK_RMAX / 2; ++r) qv[r] = q2[roff[h * (PK_RMAX / 2) + r] + cc];
#pragma unroll
                        for (int r = 0; r < PK_RMAX / 2; ++r) {
                            double& ar = acc[h * (PK_RMAX / 2) + r];
                            ar = fma(qv[r].x, uv[m].x, ar);
                            ar = fma(qv[r].y, uv[m].y, ar);
                        }
                    }
                }
            }
            // warp butterfly, level by level over all rows, then fixed-order sum over warps
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double other[PK_RMAX];
#pragma unroll
                for (int r = 0; r < PK_RMAX; ++r) other[r] = __shfl_xor_sync(0xffffffffu, acc[r], o);
#pragma unroll
                for (int r = 0; r < PK_RMAX; ++r) acc[r] += other[r];
            }
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < PK_RMAX; ++r) red[wid][r] = acc[r];
            }
        }
        __syncthreads();
        if (tid < myrows) {
            double p8 = 0.0;
#pragma unroll
            for (int kk = 0; kk < PK_NT / 32; ++kk) p8 += red[kk][tid];
            double w = 0.0;
            w += p8;                                     // the segment combine of K2 with its single segment
            const long long rr = r0 + tid;
            wk[rr] = w;
            pk[rr] = __dmul_rn(u_row, w);                // term of this row in its group's share of u'w
        }
        pk_grid_barrier(a.gbar, gridDim.x);

        // ================= phase B (every CTA, on its own copy of the iterate) =================
        double wv[PK_SLOTS];
#pragma unroll
        for (int sl = 0; sl < PK_SLOTS; ++sl) wv[sl] = slot_on[sl] ? __ldcg(wk + slot_j[sl]) : 0.0;  // in flight under the shares
        // ---- shares of u'w: warps 2p, 2p+1 take groups p, p + 4, ...; butterfly inside each warp, then warp 2p + warp 2p+1
        {
            double dv[PK_SROUNDS];
#pragma unroll
            for (int r = 0; r < PK_SROUNDS; ++r) {
                const unsigned grp = (unsigned)r * (PK_NT / 64) + (unsigned)(wid >> 1);
                const long long rr = (long long)grp * MV_GROUP + (wid & 1) * 32 + lane;
                dv[r] = (grp < ngrp && rr < n) ? __ldcg(pk + rr) : 0.0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {   // level by level over all rounds (see phase A)
                double other[PK_SROUNDS];
#pragma unroll
                for (int r = 0; r < PK_SROUNDS; ++r) other[r] = __shfl_xor_sync(0xffffffffu, dv[r], o);
#pragma unroll
                for (int r = 0; r < PK_SROUNDS; ++r) dv[r] = __dadd_rn(dv[r], other[r]);
            }
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < PK_SROUNDS; ++r) share_red[r][wid] = dv[r];
            }
            __syncthreads();
            if ((unsigned)tid < ngrp) {
                const int r = tid / (PK_NT / 64), p = tid % (PK_NT / 64);
                shares[tid] = __dadd_rn(share_red[r][2 * p], share_red[r][2 * p + 1]);
            }
            __syncthreads();
        }
        // ---- K3's reductions over the whole problem (pg_vector_body, MODE == VP_STEP)
        const double (*pr)[PK_VMAX] = part[k & 1];
        double (*pw)[PK_VMAX] = part[(k + 1) & 1];
        Quad r;
        r.a = r.b = r.c = 0.0;
        r.m = INFINITY;
        if (tid < nctas) {
            r.a = pr[0][tid];
            r.b = pr[1][tid];
            r.m = pr[2][tid];
        }
        {
            // u'w: thread-strided over the groups, four per step (ngrp <= 48 < VP_NT: one step, one live term)
            double v4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned bq = (unsigned)tid + (unsigned)e * VP_NT;
                v4[e] = ((unsigned)tid < ngrp && bq < ngrp) ? shares[bq] : 0.0;
            }
            if ((unsigned)tid < ngrp) {
#pragma unroll
                for (int e = 0; e < 4; ++e) r.c = __dadd_rn(r.c, v4[e]);
            }
        }
        r = block_reduce(r, sm);
        const double s = r.a, f = 0.5 * r.b, mt = r.m, den = r.c;
        const double ng = sqrt(s);
        if (lead && tid == 0) {
            if (k < a.v.hist_cap) {
                a.v.hist_f[k] = f;
                a.v.hist_ng[k] = ng;
            }
            st->f = f;
            st->ng = ng;
            st->s = s;
            st->maxt = mt;
            st->iter = k;
        }
        int stop = 0;
        if (ng <= a.v.eps) stop = SVMB200_STATUS_OPTIMAL;
        else if (k >= a.v.max_iter) stop = SVMB200_STATUS_STOPPED;
        if (stop) {
            if (lead && tid == 0) {
                st->status = stop;
                __threadfence();
                st->done = 1;
            }
            break;
        }
        const double t = (den <= 1e-16) ? mt : fmin(__ddiv_rn(s, den), mt);
        if (lead && tid == 0) {
            st->t = t;
            st->den = den;
        }
        // ---- update, new direction and the partials of the next state (pg_step_elem, plain layout, no sign view), the
        // two variables of virtual CTA c accumulated in K3's loop order, then the warp halves of its block reduce
        Quad accq[PK_VMAX];
#pragma unroll
        for (int c = 0; c < PK_VMAX; ++c) {
            accq[c].a = accq[c].b = accq[c].c = 0.0;
            accq[c].m = INFINITY;
        }
        double dnew[PK_SLOTS];
#pragma unroll
        for (int sl = 0; sl < PK_SLOTS; ++sl) {   // branch-free: idle slots hold zeros, add +0.0 to the sums and skip the min
            const double x = axpy_rn(t, ds[sl], xs[sl]);
            const double g = axpy_rn(t, wv[sl], gs[sl]);
            const double dn = project_dir(g, x, lbs[sl], ubs[sl]);
            const double mprev = accq[sl >> 1].m;
            tail_accumulate(accq[sl >> 1], dn, x, g, qv_[sl], lbs[sl], ubs[sl]);
            accq[sl >> 1].m = slot_on[sl] ? accq[sl >> 1].m : mprev;
            xs[sl] = x;
            gs[sl] = g;
            ds[sl] = dn;
            dnew[sl] = dn;
        }
#pragma unroll
        for (int sl = 0; sl < PK_SLOTS; ++sl)
            if (slot_on[sl]) u_priv[slot_j[sl]] = dnew[sl];
        // the warp halves of K3's block reduce, all virtual CTAs level by level
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < PK_VMAX; ++c) {
                const double oa = __shfl_xor_sync(0xffffffffu, accq[c].a, o), ob = __shfl_xor_sync(0xffffffffu, accq[c].b, o);
                const double om = __shfl_xor_sync(0xffffffffu, accq[c].m, o);
                accq[c].a = __dadd_rn(accq[c].a, oa);
                accq[c].b = __dadd_rn(accq[c].b, ob);
                accq[c].m = fmin(accq[c].m, om);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < PK_VMAX; ++c) {
                smq[c][wid][0] = accq[c].a;
                smq[c][wid][1] = accq[c].b;
                smq[c][wid][3] = accq[c].m;
            }
        }
        __syncthreads();
        if (tid < nctas) {   // the second half of block_reduce, for virtual CTA tid
            double ra = 0.0, rb = 0.0, rm = INFINITY;
#pragma unroll
            for (int i = 0; i < VP_NT / 32; ++i) {
                ra = __dadd_rn(ra, smq[tid][i][0]);
                rb = __dadd_rn(rb, smq[tid][i][1]);
                rm = fmin(rm, smq[tid][i][3]);
            }
            pw[0][tid] = ra;
            pw[1][tid] = rb;
            pw[2][tid] = rm;
        }
        __syncthreads();   // u, the partials and (CTA 0) the iterate are in place for the next phase A / reductions
    }
    // ---- CTA 0 hands the iterate back; the solver's own copy of K3's partials (the FINALISE launch and a later launch
    // read them)
    if (lead) {
#pragma unroll
        for (int sl = 0; sl < PK_SLOTS; ++sl)
            if (slot_on[sl]) {
                a.v.x[slot_j[sl]] = xs[sl];
                a.v.g[slot_j[sl]] = gs[sl];
                a.v.d[slot_j[sl]] = ds[sl];
            }
    }
    __syncthreads();
    if (lead && tid < 2 * 3 * PK_VMAX) {
        const int par = tid / (3 * PK_VMAX), q3 = (tid / PK_VMAX) % 3, c = tid % PK_VMAX;
        if (c < nctas) a.v.part[(size_t)par * 3 * VP_MAXC + (size_t)q3 * VP_MAXC + c] = part[par][q3][c];
    }
}
