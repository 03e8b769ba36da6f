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
#include <algorithm>
#include <utility>

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
#define SVMB200_SYMV_LB 8      // 128-bit copies per batch
#endif
#ifndef SVMB200_SYMV_PLAN_OVERHEAD
#define SVMB200_SYMV_PLAN_OVERHEAD 8192.0   // cost of starting an item, in matrix elements (the planner's model; ~3 us)
#endif
#ifndef SVMB200_SYMV_STAGES
#define SVMB200_SYMV_STAGES 3  // batches in flight per thread (ring depth)
#endif
#ifndef SVMB200_SYMV_MINB
#define SVMB200_SYMV_MINB 2    // CTAs per SM the register budget is cut for
#endif

template <int TR_, int NRB_, int NCH_, int LB_, int MINB_, int STAGES_ = 3>
struct SymvShape {
    static constexpr int TR = TR_, NRB = NRB_, NCH = NCH_, LB = LB_, MINB = MINB_, STAGES = STAGES_;
    static constexpr int RING_BYTES = STAGES_ * LB_ * SY_NT * 16;   // dynamic shared memory: the copy ring
    static constexpr int BH = TR_ * NRB_;          // rows per band
    static constexpr int BW = SY_CHUNK * NCH_;     // columns per panel
    static constexpr int NBATCH = NCH_ * TR_ / LB_;  // load batches per sub-block
    static_assert((TR_ & (TR_ - 1)) == 0 && TR_ >= 2 && TR_ <= 32, "TR: a power of two up to 32");
    static_assert(TR_ % LB_ == 0, "batches must tile a sub-block");
    static_assert(BH % MV_GROUP == 0 || MV_GROUP % BH == 0, "bands and 64-row groups must nest");
    static_assert(BH % 2 == 0, "panels start on 16-byte boundaries");
};
using SymvDefault = SymvShape<SVMB200_SYMV_TR, SVMB200_SYMV_NRB, SVMB200_SYMV_NCH, SVMB200_SYMV_LB, SVMB200_SYMV_MINB, SVMB200_SYMV_STAGES>;

// ------------------------------------------------------------------------------------------ the plan of a pass
// One rank (P = 1) works through the upper triangle of the whole matrix.  With the matrix cut into P row blocks
// (svmb200_shard_rows, rank r holds block r in full), the P x P grid of blocks is symmetric as well, and every pair of
// off-diagonal blocks (p, q), (q, p) needs ONE of them: rank p takes block (p, q) for the (P - 1) / 2 ranks q that follow
// it cyclically, and for an even P the pair at distance P / 2 is halved (the lower rank takes the first h columns of
// its block, the upper rank the rows from h on of its own, h = half the upper rank's rows rounded down to a band) -- every
// rank streams the same share, half of its block row.  What a rank computes from block (p, q):
//     row sums     -> its own rows                         (rowpart, as on one rank)
//     column sums  -> rows of rank q: reduced over its bands and SENT to q   (symv_send_kernel -> q's inbox, tagged entries)
// and its diagonal block is a small one-rank problem.  The owner of a row adds, in a fixed order, the row sums of its
// own items, the column sums of its own bands above the row and the column sums it was sent (ascending sender rank) --
// symv_combine_kernel -- and publishes the finished product exactly as K2 does (own w, or tagged entries into every
// rank's gathered buffer), so the vector kernels do not know which pass fed them.
struct SymvItem {
    int lr0;     // first LOCAL row of the band
    int rows;    // rows of the band (BH; the short bands at the end of a shard: SH; a last band may be ragged)
    int band;    // index of the band (row of colpart, entry of nseg)
    int c0;      // first column
    int width;   // columns (even; a last panel is narrower than BW)
    int seg;     // slot of the item's row sums in rowpart
    int cols;    // 1: column sums too (everything but a diagonal block)
};
struct SymvSend {
    int dest;    // owner of the columns
    int c0;      // first column
    int ncols;   // columns (= rows of dest) covered
    int lr0;     // local row of dest that column c0 is
    int band0, band1;   // local bands whose column sums are added
};
struct SymvRecv {
    int src;       // sending rank
    int lr0, lr1;  // my local rows it sends column sums for
};
struct SymvPlan {
    std::vector<SymvItem> items;   // large items first
    std::vector<int> nseg;         // per local band: items (= row-sum slots) of the band
    std::vector<int> band_lr0;     // per local band: its first local row (+ one entry: nrows)
    std::vector<int> band_of_unit; // band of every `unit` rows (band boundaries are multiples of unit)
    std::vector<SymvSend> sends;
    std::vector<SymvRecv> recvs;   // ascending sender
    int nseg_max = 0, unit = 1;
    long long row0 = 0, nrows = 0, nbands = 0, streamed_elems = 0;
};

inline void symv_block(long long n, long long rpr, int r, long long* row0, long long* nrows) {
    long long r0 = (long long)r * rpr;
    if (r0 > n) r0 = n;
    *row0 = r0;
    *nrows = n - r0 < rpr ? n - r0 : rpr;
}

// does rank p read (part of) block (p, q)?  *c_lo, *c_hi: the columns (relative to block q's first); *b_lo: first local
// band of p that takes part
template <class S>
inline bool symv_takes(long long n, long long rpr, int P, int p, int q, long long* c_lo, long long* c_hi, long long* b_lo) {
    long long row0q, nrq, row0p, nrp;
    symv_block(n, rpr, q, &row0q, &nrq);
    symv_block(n, rpr, p, &row0p, &nrp);
    *c_lo = 0;
    *c_hi = nrq;
    *b_lo = 0;
    if (p == q || nrq <= 0 || nrp <= 0) return false;
    const int d = ((q - p) % P + P) % P;
    if (2 * d < P) return true;
    if (2 * d > P) return false;
    // the halved pair: lo = min(p, q) reads the first h columns of (lo, hi); hi reads its rows from h on of (hi, lo)
    const int hi = p > q ? p : q;
    long long row0h, nrh;
    symv_block(n, rpr, hi, &row0h, &nrh);
    const long long h = nrh / 2 / S::BH * S::BH;
    if (p < q) {
        *c_hi = h;
        return h > 0;
    }
    *b_lo = h / S::BH;
    return true;
}

// list scheduling of item costs (in launch order) on `slots` resident CTAs: the finish time of the pass
inline double symv_makespan(const std::vector<double>& cost, int slots) {
    std::vector<double> heap((size_t)slots, 0.0);   // min-heap of slot finish times
    auto sift = [&](size_t i) {
        for (;;) {
            size_t l = 2 * i + 1, r = l + 1, m = i;
            if (l < heap.size() && heap[l] < heap[m]) m = l;
            if (r < heap.size() && heap[r] < heap[m]) m = r;
            if (m == i) return;
            std::swap(heap[i], heap[m]);
            i = m;
        }
    };
    double end = 0.0;
    for (double c : cost) {
        heap[0] += c;
        if (heap[0] > end) end = heap[0];
        sift(0);
    }
    return end;
}

// `slots`: CTAs of the tile kernel that are resident at a time (2 per SM).  A rank of an 8-GPU solve has only ~2 waves of
// full-size items: the finish time would be set by the last, nearly empty wave.  The plan therefore ends the grid on small
// items: the last `s` bands of the shard are cut into SHORT bands of SH = BH / 4 rows, and the panels at the end of the list
// into halves and quarters (still whole 512-column chunks).  How many of each is chosen by simulating the CTA scheduler
// (items start in list order, largest first, on the first free slot).  Band boundaries up to half the shard stay multiples
// of BH, which is what the halved block pair of an even rank count is cut at (symv_takes).
template <class S>
inline void symv_build_plan(long long n, long long ld, int rank, int P, long long rpr, SymvPlan& plan, int slots = 296) {
    plan = SymvPlan();
    symv_block(n, rpr, rank, &plan.row0, &plan.nrows);
    const long long row0 = plan.row0, nrows = plan.nrows;
    constexpr int SH = S::BH / 4 >= S::TR ? S::BH / 4 : S::TR;   // rows of a short band
    plan.unit = SH;
    // the last block also owns the padding columns [n, ld): they meet u = 0, and the width of a panel stays even
    auto block_end = [&](int r) {
        long long r0, nr;
        symv_block(n, rpr, r, &r0, &nr);
        return r == P - 1 ? ld : r0 + nr;
    };
    const double overhead = SVMB200_SYMV_PLAN_OVERHEAD;   // start-up of an item (first copies, u of the rows, the final barrier) in elements
    auto cost_of = [&](const SymvItem& it) { return (double)it.rows * it.width + overhead; };
    const long long full_bands = nrows / S::BH;
    // A grid of many waves has no tail worth grading (one GPU at C4: 15 waves, < 1 % to gain) -- and simulating hundreds of
    // candidate plans of thousands of items would cost more host time than it can save: the wide search is for short grids.
    long long est_items = 0;
    {
        const long long span = (P == 1 ? n / 2 : n / 2);   // columns a band reads, on average
        est_items = full_bands * ((span + S::BW - 1) / S::BW);
    }
    const bool wide_search = est_items < 8ll * (slots > 0 ? slots : 1);
    std::vector<long long> shorts = {0};   // how many of the last full bands become short ones
    if (wide_search)
        for (long long sb : {1, 2, 4, 8})
            if (SH < S::BH && sb <= full_bands / 2) shorts.push_back(sb);
    double best = -1.0;
    std::vector<int> best_lr0;
    for (long long sb : shorts) {
        // ---- band layout
        std::vector<int> lr0s;
        const long long tall_rows = (full_bands - sb) * S::BH + (sb == 0 ? nrows - full_bands * S::BH : 0);
        for (long long r = 0; r < tall_rows; r += S::BH) lr0s.push_back((int)r);
        for (long long r = tall_rows; r < nrows; r += SH) lr0s.push_back((int)r);
        lr0s.push_back((int)nrows);
        const long long nb = (long long)lr0s.size() - 1;
        std::vector<SymvItem> full, rest, diag;   // seg is assigned at the end, in list order
        for (long long I = 0; I < nb; ++I) {
            const int lr0 = lr0s[(size_t)I], rows = lr0s[(size_t)I + 1] - lr0;
            const long long g0 = row0 + lr0;
            long long dend = g0 + rows;   // the diagonal block: rows x rows, clipped to this rank's block (never active: rows end there)
            if (dend % 2) dend += 1;      // an odd last row count: the pair's second column is padding (u = 0) or the next row's, see below
            if (dend > block_end(rank)) dend = block_end(rank);
            SymvItem d = {lr0, rows, (int)I, (int)g0, (int)(dend - g0), 0, 0};
            diag.push_back(d);
            // column intervals of the band: the rest of its own block and the blocks (or half block) it reads; blocks that
            // follow each other are one interval -- panels do not care whose rows their columns are
            std::vector<std::pair<long long, long long>> iv;
            if (dend < block_end(rank)) iv.push_back({dend, block_end(rank)});
            for (int k = 1; k < P; ++k) {
                const int q = (rank + k) % P;
                long long c_lo, c_hi, b_lo, row0q, nrq;
                if (!symv_takes<S>(n, rpr, P, rank, q, &c_lo, &c_hi, &b_lo) || lr0 < b_lo * S::BH) continue;
                symv_block(n, rpr, q, &row0q, &nrq);
                iv.push_back({row0q + c_lo, c_hi == nrq ? block_end(q) : row0q + c_hi});
            }
            std::sort(iv.begin(), iv.end());
            std::vector<std::pair<long long, long long>> merged;
            for (const auto& v : iv) {
                if (!merged.empty() && merged.back().second == v.first) merged.back().second = v.second;
                else merged.push_back(v);
            }
            for (const auto& v : merged) {
                for (long long c = v.first; c < v.second && c < n; c += S::BW) {   // a panel that starts at or beyond n meets u = 0 only
                    const long long w = v.second - c < S::BW ? v.second - c : S::BW;
                    SymvItem it = {lr0, rows, (int)I, (int)c, (int)w, 0, 1};
                    (w == S::BW && rows == S::BH ? full : rest).push_back(it);
                }
            }
        }
        // ---- grade the tail: keep a whole number of waves of full panels, cut the others into halves / quarters; the
        // full-width panels of the short bands may be cut likewise
        auto build = [&](size_t keep, double f4, int short_parts, std::vector<SymvItem>& out) {
            out.clear();
            const size_t nf = full.size();
            if (keep > nf) keep = nf;
            const bool can2 = S::NCH % 2 == 0, can4 = S::NCH % 4 == 0;
            const size_t n4 = can4 ? (size_t)(f4 * (double)(nf - keep)) : 0;
            auto push_parts = [&](const SymvItem& it, int parts) {
                for (int p = 0; p < parts; ++p) {
                    SymvItem piece = it;
                    piece.width = it.width / parts;
                    piece.c0 = it.c0 + p * piece.width;
                    out.push_back(piece);
                }
            };
            for (size_t i = 0; i < nf; ++i) push_parts(full[i], i < keep ? 1 : (i >= nf - n4 ? 4 : (can2 ? 2 : 1)));
            for (const SymvItem& it : rest) {
                const bool whole_short = it.width == S::BW && it.rows < S::BH;
                push_parts(it, whole_short && ((short_parts == 2 && can2) || (short_parts == 4 && can4)) ? short_parts : 1);
            }
            // largest first (the ragged items sit between the whole panels and their pieces), diagonal blocks at the end
            std::stable_sort(out.begin(), out.end(), [&](const SymvItem& a, const SymvItem& b) { return cost_of(a) > cost_of(b); });
            out.insert(out.end(), diag.begin(), diag.end());
        };
        const size_t nslots = (size_t)(slots > 0 ? slots : 1), waves = full.size() / nslots;
        std::vector<size_t> keeps = {full.size(), waves * nslots};
        if (waves >= 1) keeps.push_back((waves - 1) * nslots);
        if (wide_search)
            for (size_t k = 1; k <= 8; ++k)   // and in steps of a quarter of a wave below the total (up to two waves)
                if (k * (nslots / 4 + 1) < full.size()) keeps.push_back(full.size() - k * (nslots / 4 + 1));
        const double f4s[] = {0.0, 1.0};
        std::vector<SymvItem> cand;
        for (int short_parts : {1, 4}) {
            if (sb == 0 && short_parts > 1) break;
            for (size_t keep : keeps) {
                for (double f4 : f4s) {
                    build(keep, f4, short_parts, cand);
                    std::vector<double> cost;
                    cost.reserve(cand.size());
                    for (const SymvItem& it : cand) cost.push_back(cost_of(it));
                    const double t = symv_makespan(cost, (int)nslots);
#ifdef SYMV_PLAN_DEBUG
                    printf("short %lld x%d keep %zu f4 %.2f items %zu makespan %.0f\n", sb, short_parts, keep, f4, cand.size(), t);
#endif
                    if (best < 0.0 || t < best * (1.0 - 1e-3)) {   // a finer plan must pay for itself
                        best = t;
                        plan.items = cand;
                        best_lr0 = lr0s;
                    }
                    if (keep == full.size()) break;   // nothing to cut
                }
            }
        }
    }
    plan.band_lr0 = best_lr0;
    plan.nbands = (long long)best_lr0.size() - 1;
    plan.nseg.assign((size_t)plan.nbands, 0);
    plan.band_of_unit.assign((size_t)((nrows + SH - 1) / SH), 0);
    for (long long I = 0; I < plan.nbands; ++I)
        for (long long r = best_lr0[(size_t)I]; r < best_lr0[(size_t)I + 1]; r += SH) plan.band_of_unit[(size_t)(r / SH)] = (int)I;
    for (SymvItem& it : plan.items) {
        it.seg = plan.nseg[(size_t)it.band]++;
        plan.streamed_elems += (long long)it.rows * it.width;
    }
    for (int v : plan.nseg) plan.nseg_max = v > plan.nseg_max ? v : plan.nseg_max;
    for (int q = 0; q < P; ++q) {
        long long c_lo, c_hi, b_lo, row0q, nrq;
        symv_block(n, rpr, q, &row0q, &nrq);
        // (the first b_lo bands of a shard are BH rows tall -- short bands only come after its middle -- so b_lo is a band index)
        if (symv_takes<S>(n, rpr, P, rank, q, &c_lo, &c_hi, &b_lo) && b_lo * S::BH < nrows) {
            SymvSend sd = {q, (int)(row0q + c_lo), (int)(c_hi - c_lo), (int)c_lo, (int)b_lo, (int)plan.nbands};
            plan.sends.push_back(sd);
        }
        if (symv_takes<S>(n, rpr, P, q, rank, &c_lo, &c_hi, &b_lo) && b_lo * S::BH < nrq) {
            SymvRecv rv = {q, (int)c_lo, (int)c_hi};
            plan.recvs.push_back(rv);
        }
    }
}

struct SymvArgs {
    const double* Q;     // this rank's rows x ld; the whole matrix is symmetric
    long long ld, nrows, row0, n_pad;
    const double* u;     // ld entries, zero beyond n
    double* rowpart;     // [max items per band][n_pad]
    double* colpart;     // [local bands][ld]
    const SymvItem* items;
    const int* done;
    const int* fault;    // the context's sticky exchange-fault flag, or null
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

// ---- the matrix stream: cp.async (16 bytes, L2 only) into a ring of thread-private shared-memory slots.  Every thread
// keeps STAGES batches of LB copies in flight without holding a register for them -- the in-flight bytes per SM
// (CTAs x 256 threads x STAGES x LB x 16 B, ~190 KB) are what a 7 TB/s stream needs at ~1.5 us of loaded latency; with the
// data held in registers the same kernel stalled at 6.2 TB/s for every tile shape (profiles/r2_sy2_symv_sweep.log).
// A slot is written and read by ONE thread, so there is no block-wide barrier anywhere in the stream: completion is the
// thread's own cp.async.wait_group, and a slot is refilled only after the FMAs that consumed it have been issued.
#ifndef SVMB200_HOST_EMULATION
__device__ __forceinline__ void sy_cp_async16(double2* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void sy_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void sy_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
#else
__device__ __forceinline__ void sy_cp_async16(double2* smem_dst, const double* gsrc) { *smem_dst = *reinterpret_cast<const double2*>(gsrc); }
__device__ __forceinline__ void sy_cp_async_commit() {}
template <int PENDING>
__device__ __forceinline__ void sy_cp_async_wait() {}
#endif

template <class S, bool COLS>
__device__ __forceinline__ void symv_item(const SymvArgs& a, const long long r0, const int rows, const long long c0,
                                          const int width, const int band, const int seg, const double* ush,
                                          double (*wsum)[S::BH], double2* ring) {
    constexpr int TR = S::TR, NCH = S::NCH, LB = S::LB, NBATCH = S::NBATCH, BPC = TR / LB, ST = S::STAGES;
    static_assert(ST >= 2 && ST <= NBATCH, "ring depth: between 2 batches and one sub-block");
    const int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nch = (width + SY_CHUNK - 1) / SY_CHUNK;  // chunks this item has (a diagonal block or a last panel: fewer)
    // r0: first LOCAL row of the band, rows: its height (<= BH)
    const int nrb = (rows + TR - 1) / TR;               // sub-blocks this item has (the last band: fewer)
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
    // u at the thread's columns: the same for every sub-block of the item, so it is fetched ONCE, up front -- as a load in
    // front of every batch it was the top stall of the first ring version (an exposed L2 round trip per 8 rows)
    double2 ucs[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        ucs[k] = make_double2(0.0, 0.0);
        if (cok[k]) ucs[k] = __ldg(reinterpret_cast<const double2*>(a.u + c0 + coff[k]));
    }
    const double* __restrict__ qitem = a.Q + r0 * a.ld + c0;
    double2* const myring = ring + tid;   // ring[slot][j][thread]

    // batch b of sub-block rb -> ring slot; rows beyond the matrix re-read its last row (a valid address; ush is zero
    // there and the row sum is dropped).  Always commits, so that the group count per step is constant.
    auto issue = [&](const int rb, const int b, const int slot) {
        const int k = b / BPC, rbase = (b % BPC) * LB;
        if (rb < nrb && k < nch) {
            const int last = (rows - rb * TR < TR ? rows - rb * TR : TR) - 1;
            const double* qb = qitem + (long long)(rb * TR) * a.ld + coff[k];
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int r = rbase + j < last ? rbase + j : last;
                sy_cp_async16(myring + (slot * LB + j) * SY_NT, qb + (long long)r * a.ld);
            }
        }
        sy_cp_async_commit();
    };

#pragma unroll
    for (int b = 0; b < ST; ++b) issue(0, b, b);
    int slot = 0;
#pragma unroll 1
    for (int rb = 0; rb < nrb; ++rb) {
        double rowacc[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) rowacc[r] = 0.0;
#pragma unroll
        for (int b = 0; b < NBATCH; ++b) {
            sy_cp_async_wait<ST - 1>();   // the oldest batch in flight -- the one consumed now -- has landed
            const int k = b / BPC, rb0 = (b % BPC) * LB;
            if (k < nch) {  // uniform over the CTA
                const double2 uc = ucs[k];
#pragma unroll
                for (int j = 0; j < LB; ++j) {
                    const double2 v = myring[(slot * LB + j) * SY_NT];
                    rowacc[rb0 + j] = fma(v.x, uc.x, rowacc[rb0 + j]);
                    rowacc[rb0 + j] = fma(v.y, uc.y, rowacc[rb0 + j]);
                    if (COLS) {
                        const double ur = ush[rb * TR + rb0 + j];
                        colacc[k].x = fma(v.x, ur, colacc[k].x);
                        colacc[k].y = fma(v.y, ur, colacc[k].y);
                    }
                }
            }
            // refill the slot just consumed with the batch ST steps ahead (it may belong to the next sub-block)
            if (b + ST < NBATCH) issue(rb, b + ST, slot);
            else issue(rb + 1, b + ST - NBATCH, slot);
            slot = slot + 1 == ST ? 0 : slot + 1;
        }
        // row sums of the sub-block over the warp's lanes; the warps meet once per item, not once per sub-block, so a
        // warp that reduces never holds up the stream of the others
        const double tot = symv_warp_rows<TR>(rowacc, lane);
        if (lane < TR) wsum[wid][rb * TR + lane] = tot;
    }
    sy_cp_async_wait<0>();
    __syncthreads();
    for (int i = tid; i < rows; i += SY_NT) {   // warps in index order
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < SY_NT / 32; ++w) v += wsum[w][i];
        a.rowpart[(size_t)seg * a.n_pad + r0 + i] = v;
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
    if (a.fault != nullptr && *a.fault) return;  // the exchange is broken: nothing downstream will be used
    constexpr int BH = S::BH;
#ifndef SVMB200_HOST_EMULATION
    extern __shared__ __align__(16) unsigned char sy_smem_raw[];
    double2* ring = reinterpret_cast<double2*>(sy_smem_raw);
#else
    double2* ring = reinterpret_cast<double2*>(emu::dynamic_smem());
#endif
    __shared__ double ush[BH];
    __shared__ double wsum[SY_NT / 32][BH];   // row sums per warp
    const SymvItem it = a.items[blockIdx.x];
    const long long r0 = it.lr0;
    for (int i = (int)threadIdx.x; i < BH; i += SY_NT) ush[i] = i < it.rows ? a.u[a.row0 + r0 + i] : 0.0;
    __syncthreads();
    if (it.cols) symv_item<S, true>(a, r0, it.rows, it.c0, it.width, it.band, it.seg, ush, wsum, ring);
    else symv_item<S, false>(a, r0, it.rows, it.c0, it.width, it.band, it.seg, ush, wsum, ring);
}

// ---- column sums that belong to other ranks' rows: added over this rank's bands (band order) and stored into the
// owner's inbox as tagged entries (k2_matvec.cuh, ll_store).  One thread per column.
struct SymvSendArgs {
    const double* colpart;
    long long ld;
    int nsend;
    SymvSend sends[SVM_MAX_RANKS];
    ulonglong2* inbox[SVM_MAX_RANKS];   // per send: this rank's slot in the destination's inbox (entry of its local row 0)
    unsigned tag;
    const int* done;
    const int* fault;
};

__global__ void __launch_bounds__(256) symv_send_kernel(const SymvSendArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    if (a.done != nullptr && *a.done) return;
    if (a.fault != nullptr && *a.fault) return;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < a.nsend; ++s) {
        const SymvSend sd = a.sends[s];
        if (idx >= sd.ncols) {
            idx -= sd.ncols;
            continue;
        }
        const double* cp = a.colpart + sd.c0 + idx;
        double v = 0.0;
        int I = sd.band0;
        for (; I + 16 <= sd.band1; I += 16) {
            double t[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) t[j] = __ldcg(cp + (size_t)(I + j) * a.ld);
#pragma unroll
            for (int j = 0; j < 16; ++j) v += t[j];
        }
        for (; I < sd.band1; ++I) v += __ldcg(cp + (size_t)I * a.ld);
        ll_store(a.inbox[s] + sd.lr0 + idx, v, a.tag);
        return;
    }
}

// w[r] = row sums of r's band in item order + column sums of the bands above it; one CTA per 64-row group, which also
// leaves the group's share of u'w where K2 leaves it (same tree: butterfly inside each warp, warp 0 + warp 1).  The
// column sums of a row (up to n / BH of them, each a separate cache line) are split over SY_CPARTS threads in contiguous
// ranges -- a single thread would spend n / BH / 8 memory round trips on them -- and joined in a fixed tree.
constexpr int SY_CPARTS = 8;

struct SymvCombineArgs {
    const double* rowpart;
    const double* colpart;
    long long ld, nrows, row0, n_pad;
    int unit;               // rows per entry of band_of_unit
    const int* band_of_unit;   // band of every `unit` local rows
    const int* nseg;        // per local band: row-sum slots in use
    const double* u_rows;   // u at this rank's rows, or null
    double* w;              // nrows results (no exchange)
    double* denpart;        // one per 64-row group, or null
    const int* done;
    // column sums sent by other ranks (sharded pass): inbox + (src * rpr + local row), tagged entries
    int nrecv;
    SymvRecv recvs[SVM_MAX_RANKS];
    const ulonglong2* inbox;
    long long rpr;
    unsigned tag;
    int* fault;
    // fused exchange of the finished product (nranks_x > 0), as in MatvecArgs
    int nranks_x;
    long long share_off;    // slot of the first u'w share relative to the slot of row 0
    ulonglong2* peer_w[SVM_MAX_RANKS];
};

__global__ void __launch_bounds__(MV_GROUP * SY_CPARTS) symv_combine_kernel(const SymvCombineArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    if (a.done != nullptr && *a.done) return;
    if (a.fault != nullptr && *a.fault) return;
    __shared__ double part[SY_CPARTS][MV_GROUP];
    __shared__ double red[2];
    const int t = (int)threadIdx.x % MV_GROUP, p = (int)threadIdx.x / MV_GROUP;   // warps share a part
    // the last groups have the most column sums to add: they go first, the short ones fill the tail of the grid
    const unsigned group = gridDim.x - 1u - blockIdx.x;
    const long long rr = (long long)group * MV_GROUP + t;   // local row
    double v = 0.0;
    if (rr < a.nrows) {
        const long long band = a.band_of_unit[rr / a.unit];
        const long long per = (band + SY_CPARTS - 1) / SY_CPARTS;
        long long I = p * per, I1 = I + per;
        if (I1 > band) I1 = band;
        const double* cp = a.colpart + a.row0 + rr;
        for (; I + 16 <= I1; I += 16) {
            double tt[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) tt[j] = __ldcg(cp + (size_t)(I + j) * a.ld);
#pragma unroll
            for (int j = 0; j < 16; ++j) v += tt[j];
        }
        for (; I < I1; ++I) v += __ldcg(cp + (size_t)I * a.ld);
        if (p == 0) {   // the row sums of the band's own items go in front of part 0
            const int nseg = a.nseg[band];
            double rs = 0.0;
            for (int s = 0; s < nseg; ++s) rs += __ldcg(a.rowpart + (size_t)s * a.n_pad + rr);
            v = rs + v;
        }
    }
    part[p][t] = v;
    __syncthreads();
    double dv = 0.0;
    if (p == 0 && rr < a.nrows) {
        static_assert(SY_CPARTS == 8, "the join below is written for eight parts");
        double w = ((part[0][t] + part[1][t]) + (part[2][t] + part[3][t])) +
                   ((part[4][t] + part[5][t]) + (part[6][t] + part[7][t]));
        for (int s = 0; s < a.nrecv; ++s) {   // ascending sender
            const SymvRecv rv = a.recvs[s];
            if (rr >= rv.lr0 && rr < rv.lr1) w += ll_load(a.inbox + (size_t)rv.src * a.rpr + rr, a.tag, a.fault);
        }
        if (a.nranks_x > 0) {
            for (int q = 0; q < a.nranks_x; ++q) ll_store(a.peer_w[q] + rr, w, a.tag);   // NVLink stores (one is local)
        } else {
            a.w[rr] = w;
        }
        if (a.u_rows != nullptr) dv = __dmul_rn(a.u_rows[rr], w);
    }
    if (a.denpart != nullptr) {
        if (p == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dv = __dadd_rn(dv, __shfl_xor_sync(0xffffffffu, dv, o));
            if ((t & 31) == 0) red[t >> 5] = dv;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double tot = __dadd_rn(red[0], red[1]);
            if (a.nranks_x > 0) {
                for (int q = 0; q < a.nranks_x; ++q) ll_store(a.peer_w[q] + a.share_off + group, tot, a.tag);
            } else {
                a.denpart[group] = tot;
            }
        }
    }
}
