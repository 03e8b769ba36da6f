// Persistent projected-gradient loop for problems whose matrix fits the SHARED MEMORY of the GPU -- device code
// (included by pg.cu only).
//
// BASELINE config C1 (n = 2 000: a 32 MB matrix) is launch-bound on the two-kernel loop: 14 us per iteration for ~1 us of
// memory traffic ("hard parts").  148 SMs x 227 KB of shared memory hold 33 MB, so here ONE
// cooperative kernel keeps the whole matrix on chip -- CTA b owns rows [b r, (b+1) r) and loads them once -- and runs all
// iterations of the solve:
//     phase A (every CTA)       w = Q u for the CTA's rows, from shared memory; per-row products u_r w_r for the shares
//     grid barrier
//     phase B (CTA v < nctas)   the 64-row shares of u'w, then the vector phase of K3 (pg_vector_body) as virtual CTA v
//     grid barrier              stop when the stopping test has fired
// Bit-identical to the K2 + K3 loop: a row sum is the thread-strided fma chain, warp butterfly and in-order warp sum of
// matvec_seg_kernel (one column segment: ld <= 3072), the share tree is K2's group combine, phase B is K3's own code.
// What crosses CTAs inside the kernel (u, w, products, per-CTA partials, the done flag) is read with L1-bypassing loads.
#pragma once
#include "k3_vector.cuh"

constexpr int PK_NT = 256;
constexpr int PK_RMAX = 16;   // rows per CTA (accumulators per thread)
constexpr int PK_UMAX = 6;    // 128-bit operand slots per thread: ld <= 2 * PK_NT * PK_UMAX = 3072 columns
constexpr int PK_SROUNDS = (2 * PK_NT * PK_UMAX / MV_GROUP + PK_NT / 64 - 1) / (PK_NT / 64);  // share rounds: 4 groups each

struct PersistArgs {
    const double* Q;        // n x ld, all rows on this GPU
    long long ld, n;
    int rows_per_cta;
    double* prod;           // n per-row products u_r * w_r (scratch)
    unsigned* gbar;         // grid barrier: [0] arrivals, [32] generation (separate 128-byte lines), zero between launches
    VecArgs v;              // vector phase; v.gathered is the plain [w | shares] buffer of a single rank
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

// all CTAs of the (cooperatively launched, hence co-resident) grid; sense by generation, the last arriver re-arms
__device__ __forceinline__ void pk_grid_barrier(unsigned* bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* count = bar;
        unsigned* gen = bar + 32;
        __threadfence();
        const unsigned my_gen = pk_ld_acquire(gen);
        if (pk_arrive(count) == nblocks - 1) {
            *reinterpret_cast<volatile unsigned*>(count) = 0u;
            pk_st_release(gen, my_gen + 1u);
        } else {
            for (unsigned long long spins = 0; pk_ld_acquire(gen) == my_gen; ++spins) pk_backoff(spins);
        }
        __threadfence();
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
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
    long long r1 = r0 + a.rows_per_cta;
    if (r1 > a.n) r1 = a.n;
    const int myrows = r1 > r0 ? (int)(r1 - r0) : 0;
    const int nvec = (int)(a.ld >> 1);
    PGDeviceState* st = a.v.st;

    // ---- the CTA's rows, once: global -> shared (the matrix is read-only for the whole solve)
    {
        const double2* src = reinterpret_cast<const double2*>(a.Q + r0 * a.ld);
        double2* dst = reinterpret_cast<double2*>(qs);
        const long long total = (long long)myrows * nvec;
        for (long long i = tid; i < total; i += PK_NT) dst[i] = ld_stream_f64x2(src + i);
    }
    __syncthreads();

    const unsigned ngrp = (unsigned)((a.n + MV_GROUP - 1) / MV_GROUP);
    for (long long it = 0; it < a.niter; ++it) {
        const long long k = a.k0 + it;
        // ================= phase A: w = Q u for my rows (matvec_seg_kernel's arithmetic, one column segment) =================
        double u_row = 0.0;
        if (myrows > 0) {
            double2 uv[PK_UMAX];
            const double2* u2 = reinterpret_cast<const double2*>(a.v.u);
#pragma unroll
            for (int m = 0; m < PK_UMAX; ++m) {
                const int c = tid + m * PK_NT;
                uv[m] = c < nvec ? pk_ld_cg_f64x2(u2 + c) : double2{0.0, 0.0};
            }
            if (tid < myrows) u_row = __ldcg(a.v.u + r0 + tid);  // for the row's term of u'w below: in flight with the rest
            double acc[PK_RMAX];
#pragma unroll
            for (int r = 0; r < PK_RMAX; ++r) acc[r] = 0.0;
#pragma unroll
            for (int r = 0; r < PK_RMAX; ++r) {
                if (r < myrows) {
                    const double2* row = reinterpret_cast<const double2*>(qs + (size_t)r * a.ld);
#pragma unroll
                    for (int m = 0; m < PK_UMAX; ++m) {
                        const int c = tid + m * PK_NT;
                        if (c < nvec) {
                            const double2 qv = row[c];
                            acc[r] = fma(qv.x, uv[m].x, acc[r]);
                            acc[r] = fma(qv.y, uv[m].y, acc[r]);
                        }
                    }
                }
            }
            // warp butterfly, then fixed-order sum over warps
#pragma unroll
            for (int r = 0; r < PK_RMAX; ++r) {
                if (r < myrows) {  // uniform over the CTA
                    double v = acc[r];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) red[wid][r] = v;
                }
            }
        }
        __syncthreads();
        if (tid < myrows) {
            double part = 0.0;
#pragma unroll
            for (int kk = 0; kk < PK_NT / 32; ++kk) part += red[kk][tid];
            double v = 0.0;
            v += part;                                   // the segment combine of K2 with its single segment
            const long long rr = r0 + tid;
            const_cast<double*>(a.v.gathered)[rr] = v;
            a.prod[rr] = __dmul_rn(u_row, v);            // term of this row in its group's share of u'w
        }
        pk_grid_barrier(a.gbar, gridDim.x);

        // ================= phase B: shares of u'w (K2's group combine) + K3's vector phase on the first nctas CTAs =================
        if ((int)blockIdx.x < a.v.nctas) {
            // warps 2p, 2p+1 take groups p, p + 4, ...: butterfly inside each warp, then warp 2p + warp 2p+1 (K2's tree);
            // all loads first, so the rounds cost one memory round trip together
            double dv[PK_SROUNDS];
#pragma unroll
            for (int r = 0; r < PK_SROUNDS; ++r) {
                const unsigned grp = (unsigned)r * (PK_NT / 64) + (unsigned)(wid >> 1);
                const long long rr = (long long)grp * MV_GROUP + (wid & 1) * 32 + lane;
                dv[r] = (grp < ngrp && rr < a.n) ? __ldcg(a.prod + rr) : 0.0;
            }
#pragma unroll
            for (int r = 0; r < PK_SROUNDS; ++r) {
                if ((unsigned)r * (PK_NT / 64) < ngrp) {  // uniform
                    double v = dv[r];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (lane == 0) share_red[r][wid] = v;
                }
            }
            __syncthreads();
            if ((unsigned)tid < ngrp) {
                const int r = tid / (PK_NT / 64), p = tid % (PK_NT / 64);
                const_cast<double*>(a.v.gathered)[a.v.rpr + tid] = __dadd_rn(share_red[r][2 * p], share_red[r][2 * p + 1]);
            }
            __syncthreads();
            __threadfence_block();
            pg_vector_body<VP_STEP>(a.v, k);
        }
        pk_grid_barrier(a.gbar, gridDim.x);
        if (*reinterpret_cast<volatile int*>(&st->done)) break;
    }
}
