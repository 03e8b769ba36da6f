// Host-side helpers of the fit path that are O(n d) on the host in the reference and sit on the critical path of a
// sharded fit (every rank repeats them): the variance behind gamma='scale' and the gather of the support vectors.
//
// svmb200_host_variance reproduces NumPy's X.var() BIT FOR BIT (optiml/ml/svm/kernels.py:93, 127 evaluate
// 1 / (d * X.var()) and gamma feeds every Gram entry, so a differently rounded variance would move the whole fit):
//   numpy/_core/_methods.py _var:   mean = add.reduce(x) / N ;  var = add.reduce((x - mean) * (x - mean)) / N
//   numpy/_core/src/umath/loops_utils.h.src  pairwise sum: blocks of <= 128 elements summed with 8 interleaved
//   accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), plus a scalar tail; larger ranges split at
//   n/2 rounded down to a multiple of 8.
// NumPy walks that tree on one thread and materialises two n x d temporaries; here the second pass is fused and the
// top levels of the (deterministic) tree are spread over a few threads -- same additions, same order, same bits
// (tests/test_host_logic.py checks equality with np.var on ragged shapes).
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr int64_t PW_BLOCK = 128;

template <bool SQUARED_DEV>
inline double leaf_value(const double* a, int64_t i, double mean) {
    if (!SQUARED_DEV) return a[i];
    const double t = a[i] - mean;
    return t * t;
}

template <bool SQUARED_DEV>
double pairwise_sum(const double* a, int64_t n, double mean) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += leaf_value<SQUARED_DEV>(a, i, mean);
        return res;
    }
    if (n <= PW_BLOCK) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = leaf_value<SQUARED_DEV>(a, j, mean);
        int64_t i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += leaf_value<SQUARED_DEV>(a, i + j, mean);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += leaf_value<SQUARED_DEV>(a, i, mean);
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum<SQUARED_DEV>(a, n2, mean) + pairwise_sum<SQUARED_DEV>(a + n2, n - n2, mean);
}

struct Range {
    int64_t off, len;
};

// the subtrees `depth` levels below the root, left to right
void split(int64_t off, int64_t n, int depth, std::vector<Range>& out) {
    if (depth == 0 || n <= PW_BLOCK) {
        out.push_back({off, n});
        return;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    split(off, n2, depth - 1, out);
    split(off + n2, n - n2, depth - 1, out);
}

// recombine the subtree sums exactly as the recursion would
double combine(int64_t n, int depth, const double* sums, size_t& next) {
    if (depth == 0 || n <= PW_BLOCK) return sums[next++];
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    const double l = combine(n2, depth - 1, sums, next);
    const double r = combine(n - n2, depth - 1, sums, next);
    return l + r;
}

template <bool SQUARED_DEV>
double tree_sum(const double* a, int64_t n, double mean, int threads) {
    int depth = 0;
    while ((1 << depth) < threads) ++depth;
    if (threads <= 1 || n < (int64_t)1 << 16) return pairwise_sum<SQUARED_DEV>(a, n, mean);
    std::vector<Range> parts;
    split(0, n, depth, parts);
    std::vector<double> sums(parts.size());
    std::vector<std::thread> pool;
    const size_t nt = (size_t)threads < parts.size() ? (size_t)threads : parts.size();
    for (size_t t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            for (size_t p = t; p < parts.size(); p += nt) sums[p] = pairwise_sum<SQUARED_DEV>(a + parts[p].off, parts[p].len, mean);
        });
    for (auto& th : pool) th.join();
    size_t next = 0;
    return combine(n, depth, sums.data(), next);
}

}  // namespace

extern "C" int svmb200_host_variance(const double* x_host, int64_t count, int threads, double* var) {
    SVM_CHECK_ARG(x_host != nullptr && var != nullptr && count > 0, "bad argument");
    if (threads < 1) threads = 1;
    if (threads > 16) threads = 16;
    const double mean = tree_sum<false>(x_host, count, 0.0, threads) / (double)count;
    *var = tree_sum<true>(x_host, count, mean, threads) / (double)count;
    return SVMB200_OK;
}

// out[i] = x[idx[i]] for rows of d doubles (ml/svm/_base.py:869, 1425: support_vectors_ = X[sv])
extern "C" int svmb200_host_gather_rows(const double* x_host, int64_t d, const int64_t* idx, int64_t nidx, double* out,
                                        int threads) {
    SVM_CHECK_ARG(x_host != nullptr && idx != nullptr && out != nullptr && d > 0 && nidx >= 0, "bad argument");
    if (threads < 1) threads = 1;
    if (threads > 16) threads = 16;
    auto work = [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) memcpy(out + i * d, x_host + idx[i] * d, (size_t)d * sizeof(double));
    };
    if (threads == 1 || nidx * d < (int64_t)1 << 16) {
        work(0, nidx);
        return SVMB200_OK;
    }
    std::vector<std::thread> pool;
    const int64_t chunk = (nidx + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t i0 = t * chunk, i1 = i0 + chunk < nidx ? i0 + chunk : nidx;
        if (i0 < i1) pool.emplace_back(work, i0, i1);
    }
    for (auto& th : pool) th.join();
    return SVMB200_OK;
}
