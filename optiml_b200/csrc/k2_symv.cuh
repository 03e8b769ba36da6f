// K2s: the product w = Q u for a SYMMETRIC resident matrix from ONE pass over its upper triangle -- device code (included by
// pg.cu only).  Opt-in (svmb200_pg_set_symmetric / SVMB200_SYMMETRIC=1): the Hessian of the SVM dual is (s s') o K or the
// block matrix of K + 1, K = kernel(X, X) is symmetric (optiml/ml/svm/kernels.py:49-51, 91-95, 125-129), and the
// reference's `Q.dot(d)` (optiml/opti/constrained/projected_gradient.py:113) reads all of it.  The full pass (K2,
// k2_matvec.cuh) is bound by 8 n^2 bytes of HBM traffic per iteration; this pass streams 4 n^2 + O(n BH) bytes and uses
// every element twice: once for the row it stands in and once, transposed, for the row of its column index.
//
//   rows are cut into bands of BH rows; band I owns
//     * its diagonal block  [I BH, (I+1) BH) x [I BH, (I+1) BH)        -- row sums only, all BH x BH elements
//     * panels of BW columns from column (I+1) BH to the last column  -- row sums AND column sums
//   a work item = one diagonal block or one panel.  Thread t of the 256 owns columns 2t, 2t+1 of every 512-column chunk
//   (128-bit streaming loads, a warp reads 512 contiguous bytes of a row): its column sums stay in registers over the
//   BH rows of the band; its row sums are kept for TR rows at a time and reduced across the CTA (transposing butterfly:
//   31 shuffles for TR values instead of 5 TR) while the loads of the next TR rows are already in flight.
//
//   rowpart[seg][r]   row sums of band(r)'s item `seg` (0 = diagonal block, s >= 1 = panel s - 1)
//   colpart[I][c]     column sums over the rows of band I, for c >= (I+1) BH
//   w[r] = sum_seg rowpart[seg][r] + sum_{I < band(r)} colpart[I][r]     (symv_combine_kernel, fixed order)
//
// Which CTA computes an item never changes a bit, and the combine order depends on n only: the product is reproducible
// run to run.  It is NOT bit-identical to K2's (other summation order), which is why the mode is opt-in: the iterate of a
// well-conditioned solve (C1, C3, C4) stays within 1e-8 of the reference's, a chaotic one (C2) does not care
// which rounding it follows.
#pragma once
#include "k2_matvec.cuh"

constexpr int SY_NT = 256;
constexpr int SY_CHUNK = 2 * SY_NT;  // columns one pass of the CTA's threads covers

#ifndef SVMB200_SYMV_TR
#define SVMB200_SYMV_TR 16     // rows whose sums a thread holds at a time
#endif
#ifndef SVMB200_SYMV_NRB
#define SVMB200_SYMV_NRB 8     // sub-blocks of TR rows per band
#endif
#ifndef SVMB200_SYMV_NCH
#define SVMB200_SYMV_NCH 4     // 512-column chunks per panel
#endif
#ifndef SVMB200_SYMV_LB
#define SVMB200_SYMV_LB 8      // 128-bit loads per batch; two batches are in flight
#endif
#ifndef SVMB200_SYMV_MINB
#define SVMB200_SYMV_MINB 2    // CTAs per SM the register budget is cut for
#endif

template <int TR_, int NRB_, int NCH_, int LB_, int MINB_>
struct SymvShape {
    static constexpr int TR = TR_, NRB = NRB_, NCH = NCH_, LB = LB_, MINB = MINB_;
    static constexpr int BH = TR_ * NRB_;          // rows per band
    static constexpr int BW = SY_CHUNK * NCH_;     // columns per panel
    static constexpr int NBATCH = NCH_ * TR_ / LB_;  // load batches per sub-block
    static_assert((TR_ & (TR_ - 1)) == 0 && TR_ >= 2 && TR_ <= 32, "TR: a power of two up to 32");
    static_assert(TR_ % LB_ == 0 && NBATCH % 2 == 0, "batches must tile a sub-block, an even number of them");
    static_assert(BH % MV_GROUP == 0 || MV_GROUP % BH == 0, "bands and 64-row groups must nest");
    static_assert(BH % 2 == 0, "panels start on 16-byte boundaries");
};
using SymvDefault = SymvShape<SVMB200_SYMV_TR, SVMB200_SYMV_NRB, SVMB200_SYMV_NCH, SVMB200_SYMV_LB, SVMB200_SYMV_MINB>;

// panels of band `band`: columns (band+1) BH ... n - 1 in steps of BW (columns >= n meet u = 0: nothing to add)
__host__ __device__ inline long long symv_npanels(long long n, long long band, int BH, int BW) {
    const long long first = (band + 1) * BH;
    return first >= n ? 0 : (n - first + BW - 1) / BW;
}

// work list of an n x n pass: full panels first, then the narrow last panels, the diagonal blocks at the end -- the SMs
// that run out of large items fill the tail of the grid with small ones (the order never changes a bit of the result)
template <class S>
inline void symv_build_items(long long n, long long ld, std::vector<int2>& items) {
    const long long nbands = (n + S::BH - 1) / S::BH;
    items.clear();
    for (int pass = 0; pass < 2; ++pass) {
        for (long long I = 0; I < nbands; ++I) {
            const long long np = symv_npanels(n, I, S::BH, S::BW);
            for (long long p = 0; p < np; ++p) {
                const long long c0 = (I + 1) * S::BH + p * S::BW;
                const bool full = c0 + S::BW <= ld && (I + 1) * S::BH <= n;
                if (full == (pass == 0)) items.push_back(make_int2((int)I, (int)(p + 1)));
            }
        }
    }
    for (long long I = 0; I < nbands; ++I) items.push_back(make_int2((int)I, 0));
}

struct SymvArgs {
    const double* Q;     // n x ld, symmetric in its leading n x n block
    long long ld, n, n_pad;
    const double* u;     // ld entries, zero beyond n
    double* rowpart;     // [1 + max panels][n_pad]
    double* colpart;     // [bands][ld]
    const int2* items;   // {band, seg}: the work list, large items first
    const int* done;
};

// sums v[i] over the 32 lanes for all i < TR at once: the butterfly halves the set of rows a lane is responsible for at
// every step instead of carrying all of them to the end.  On return lane l holds the total of row l & (TR - 1).
template <int TR>
__device__ __forceinline__ double symv_warp_rows(double (&v)[TR], const int lane) {
#pragma unroll
    for (int o = 16; o >= TR; o >>= 1) {
#pragma unroll
        for (int i = 0; i < TR; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
#pragma unroll
    for (int o = TR / 2; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const double send = up ? v[i] : v[i + o];
            const double keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

template <class S, bool COLS>
__device__ __forceinline__ void symv_item(const SymvArgs& a, const long long r0, const long long c0, const int width,
                                          const int band, const int seg, const double* ush,
                                          double (*red)[SY_NT / 32][S::TR]) {
    constexpr int TR = S::TR, NRB = S::NRB, NCH = S::NCH, LB = S::LB, NBATCH = S::NBATCH, BPC = TR / LB;
    const int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nch = (width + SY_CHUNK - 1) / SY_CHUNK;  // chunks this item has (a diagonal block or a last panel: fewer)
    // the thread's columns, relative to c0; a column beyond the item reads column 0 of it against u = 0
    int coff[NCH];
    bool cok[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = k * SY_CHUNK + 2 * tid;
        cok[k] = c < width;
        coff[k] = cok[k] ? c : 0;
    }
    double2 colacc[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) colacc[k] = make_double2(0.0, 0.0);
    const double* __restrict__ ucol = a.u + c0;

    // rows of sub-block rb: clamped to the last row of the matrix (a valid read; ush is zero there and the row sum is dropped)
    auto load_batch = [&](const double* qb, const int last, const int b, double2 (&q)[LB]) {
        const int k = b / BPC, rbase = (b % BPC) * LB;
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int r = rbase + j < last ? rbase + j : last;
            q[j] = ld_stream_f64x2(reinterpret_cast<const double2*>(qb + (long long)r * a.ld + coff[k]));
        }
    };

    double2 q[2][LB];
    {
        const long long rows_left = a.n - r0;
        load_batch(a.Q + r0 * a.ld + c0, (int)(rows_left < TR ? rows_left : TR) - 1, 0, q[0]);
    }
#pragma unroll 1
    for (int rb = 0; rb < NRB; ++rb) {
        const long long rbase = r0 + (long long)rb * TR;
        if (rbase >= a.n) break;
        const long long rows_left = a.n - rbase;
        const int last = (int)(rows_left < TR ? rows_left : TR) - 1;
        const double* qb = a.Q + rbase * a.ld + c0;
        const bool more = rb + 1 < NRB && rbase + TR < a.n;
        double rowacc[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) rowacc[r] = 0.0;
#pragma unroll
        for (int b = 0; b < NBATCH; ++b) {
            if (b + 1 < NBATCH) {
                if ((b + 1) / BPC < nch) load_batch(qb, last, b + 1, q[(b + 1) & 1]);
            } else if (more) {  // the first batch of the next sub-block, in flight across the reduction below
                const long long nleft = rows_left - TR;
                load_batch(qb + (long long)TR * a.ld, (int)(nleft < TR ? nleft : TR) - 1, 0, q[0]);
            }
            const int k = b / BPC, rb0 = (b % BPC) * LB;
            if (k >= nch) continue;  // uniform over the CTA
            double2 uc = __ldg(reinterpret_cast<const double2*>(ucol + coff[k]));
            if (!cok[k]) uc = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const double2 v = q[b & 1][j];
                rowacc[rb0 + j] = fma(v.x, uc.x, rowacc[rb0 + j]);
                rowacc[rb0 + j] = fma(v.y, uc.y, rowacc[rb0 + j]);
                if (COLS) {
                    const double ur = ush[rb * TR + rb0 + j];
                    colacc[k].x = fma(v.x, ur, colacc[k].x);
                    colacc[k].y = fma(v.y, ur, colacc[k].y);
                }
            }
        }
        // row sums of the sub-block: lanes, then warps in index order
        const double tot = symv_warp_rows<TR>(rowacc, lane);
        if (lane < TR) red[rb & 1][wid][lane] = tot;
        __syncthreads();
        if (tid < TR && tid <= last) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < SY_NT / 32; ++w) v += red[rb & 1][w][tid];
            a.rowpart[(size_t)seg * a.n_pad + rbase + tid] = v;
        }
    }
    if (COLS) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            if (cok[k]) *reinterpret_cast<double2*>(a.colpart + (size_t)band * a.ld + c0 + coff[k]) = colacc[k];
        }
    }
}

template <class S>
__global__ void __launch_bounds__(SY_NT, S::MINB) symv_tile_kernel(const SymvArgs a) {
    pdl_wait();               // u (and the done flag) come from the vector launch before this one
    pdl_launch_dependents();
    if (a.done != nullptr && *a.done) return;
    constexpr int BH = S::BH, BW = S::BW;
    __shared__ double ush[BH];
    __shared__ double red[2][SY_NT / 32][S::TR];
    const int2 it = a.items[blockIdx.x];
    const long long r0 = (long long)it.x * BH;
    for (int i = (int)threadIdx.x; i < BH; i += SY_NT) ush[i] = r0 + i < a.n ? a.u[r0 + i] : 0.0;
    __syncthreads();
    if (it.y == 0) {
        long long c1 = r0 + BH;
        if (c1 > a.ld) c1 = a.ld;
        symv_item<S, false>(a, r0, r0, (int)(c1 - r0), it.x, 0, ush, red);
    } else {
        const long long c0 = r0 + BH + (long long)(it.y - 1) * BW;
        long long c1 = c0 + BW;
        if (c1 > a.ld) c1 = a.ld;
        symv_item<S, true>(a, r0, c0, (int)(c1 - c0), it.x, it.y, ush, red);
    }
}

// w[r] = row sums of r's band in item order + column sums of the bands above it in band order; one CTA per 64-row group,
// which also leaves the group's share of u'w where K2 leaves it (same tree: butterfly inside each warp, warp 0 + warp 1)
struct SymvCombineArgs {
    const double* rowpart;
    const double* colpart;
    long long ld, n, n_pad;
    int BH, BW;
    const double* u_rows;   // u at the rows, or null
    double* w;
    double* denpart;        // one per 64-row group, or null
    const int* done;
};

__global__ void __launch_bounds__(MV_GROUP) symv_combine_kernel(const SymvCombineArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    if (a.done != nullptr && *a.done) return;
    __shared__ double red[2];
    const long long rr = (long long)blockIdx.x * MV_GROUP + threadIdx.x;
    double dv = 0.0;
    if (rr < a.n) {
        const long long band = rr / a.BH;
        const int nseg = 1 + (int)symv_npanels(a.n, band, a.BH, a.BW);
        double v = 0.0;
        for (int s = 0; s < nseg; ++s) v += __ldcg(a.rowpart + (size_t)s * a.n_pad + rr);
        const double* cp = a.colpart + rr;
        long long I = 0;
        for (; I + 8 <= band; I += 8) {
            double t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg(cp + (size_t)(I + j) * a.ld);
#pragma unroll
            for (int j = 0; j < 8; ++j) v += t[j];
        }
        for (; I < band; ++I) v += __ldcg(cp + (size_t)I * a.ld);
        a.w[rr] = v;
        if (a.u_rows != nullptr) dv = __dmul_rn(a.u_rows[rr], v);
    }
    if (a.denpart != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dv = __dadd_rn(dv, __shfl_xor_sync(0xffffffffu, dv, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dv;
        __syncthreads();
        if (threadIdx.x == 0) a.denpart[blockIdx.x] = __dadd_rn(red[0], red[1]);
    }
}
