// Device-side replacements of the O(n d) host steps of a fit: every rank of a sharded fit repeats them, and a rank of an
// 8-GPU job owns two host cores -- on the host they were a third of the 8-GPU fit outside the solver loop.
//
//   svmb200_device_variance   X.var() behind gamma='scale' (optiml/ml/svm/kernels.py:93, 127) BIT FOR BIT, from the copy of
//                             X that is in HBM anyway, plus the all-finite test of sklearn's input validation
//                             (check_pairwise_arrays, kernels.py:50, 92, 126) in the same pass
//   svmb200_gather_rows       support_vectors_ = X[sv] (ml/svm/_base.py:869, 1425) gathered in HBM, where the decision
//                             function needs them; the host copy is made on first access
//
// NumPy's reduction order (numpy/_core/src/umath/loops_utils.h.src, pairwise sum; numpy/_core/_methods.py, _var):
//   n < 8: sequential;  n <= 128: eight interleaved accumulators r[j] += a[i + j], combined as
//   ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail sequentially;  larger: split at n/2 rounded down to a
//   multiple of 8 and add the two halves.  mean = sum / N;  var = sum((x - mean) * (x - mean)) / N.
// The recursion's leaves (<= 128 consecutive elements) are independent: eight lanes of a warp own the eight accumulators
// of one leaf (coalesced 64-byte loads), three xor-shuffles are exactly NumPy's combine tree, lane 0 adds the tail.  The
// leaf sums go to the host (0.5 MB at n d = 6.4e6), which walks the upper levels of the same recursion.  Same additions
// in the same order as csrc/hostmath.cu and np.var -- tests compare all three for equality on ragged shapes.
#include "common.cuh"
#include <math.h>

namespace {

constexpr long long PW_BLOCK = 128;
constexpr int LEAF_LANES = 8;
constexpr int LEAF_NT = 256;

struct LeafArgs {
    const double* X;
    long long cols, ld, count;       // logical element i lives at X[(i / cols) * ld + i % cols]
    const long long* leaf_off;       // nleaves + 1 offsets into the flattened array
    int nleaves;
    double mean;
    double* sums;                    // one per leaf
    int* nonfinite;                  // raised when an element is NaN or +-inf (first pass only)
};

__device__ __forceinline__ double flat_at(const LeafArgs& a, long long i) {
    if (a.ld == a.cols) return a.X[i];
    const long long r = i / a.cols;
    return a.X[r * a.ld + (i - r * a.cols)];
}

template <bool SQUARED_DEV>
__device__ __forceinline__ double leaf_value(const LeafArgs& a, long long i, int& bad) {
    const double v = flat_at(a, i);
    if (!SQUARED_DEV) {
        if (!isfinite(v)) bad = 1;
        return v;
    }
    const double t = __dsub_rn(v, a.mean);
    return __dmul_rn(t, t);
}

template <bool SQUARED_DEV>
__global__ void __launch_bounds__(LEAF_NT) pairwise_leaf_kernel(const LeafArgs a) {
    const int leaf = (int)(blockIdx.x * (LEAF_NT / LEAF_LANES) + threadIdx.x / LEAF_LANES);
    const int j = (int)(threadIdx.x % LEAF_LANES);
    const bool live = leaf < a.nleaves;
    long long off = 0, len = 0;
    if (live) {
        off = a.leaf_off[leaf];
        len = a.leaf_off[leaf + 1] - off;
    }
    int bad = 0;
    double r = 0.0;
    const long long body = len < 8 ? 0 : len - (len % 8);
    if (body > 0) {
        r = leaf_value<SQUARED_DEV>(a, off + j, bad);
        for (long long i = 8; i < body; i += 8) r = __dadd_rn(r, leaf_value<SQUARED_DEV>(a, off + i + j, bad));
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)): IEEE addition is commutative, so every lane of the octet ends with the same bits
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (live && j == 0) {
        double res = body > 0 ? r : 0.0;
        for (long long i = body; i < len; ++i) res = __dadd_rn(res, leaf_value<SQUARED_DEV>(a, off + i, bad));
        a.sums[leaf] = res;
    }
    if (!SQUARED_DEV && bad) *a.nonfinite = 1;  // benign race: every writer stores the same value
}

void leaf_offsets(long long off, long long n, std::vector<long long>& out) {
    if (n <= PW_BLOCK) {
        out.push_back(off);
        return;
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    leaf_offsets(off, n2, out);
    leaf_offsets(off + n2, n - n2, out);
}

double combine(long long n, const double* sums, size_t& next) {
    if (n <= PW_BLOCK) return sums[next++];
    long long n2 = n / 2;
    n2 -= n2 % 8;
    const double l = combine(n2, sums, next);
    const double r = combine(n - n2, sums, next);
    return l + r;
}

struct VarianceCache {
    long long count = -1;      // the flattened length the leaf table was built for
    int nleaves = 0;
    long long* d_off = nullptr;
    double* d_sums = nullptr;
    int* d_flag = nullptr;
    double* h_sums = nullptr;  // pinned
    int* h_flag = nullptr;     // pinned
    size_t cap_leaves = 0;
};

__global__ void gather_rows_kernel(const double* __restrict__ X, long long ld_src, long long d, const long long* __restrict__ idx,
                                   long long nidx, double* __restrict__ out, long long ld_out) {
    // one warp per output row, pad columns written as zero (the row stride stays 16-byte aligned for TMA)
    const long long row = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= nidx) return;
    const double* src = X + idx[row] * ld_src;
    double* dst = out + row * ld_out;
    for (long long c = threadIdx.x % 32; c < ld_out; c += 32) dst[c] = c < d ? src[c] : 0.0;
}

}  // namespace

void svm_release_variance_cache(svmb200_ctx* ctx) {
    VarianceCache* c = static_cast<VarianceCache*>(ctx->var_cache);
    if (!c) return;
    if (c->d_off) cudaFree(c->d_off);
    if (c->d_sums) cudaFree(c->d_sums);
    if (c->d_flag) cudaFree(c->d_flag);
    if (c->h_sums) cudaFreeHost(c->h_sums);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    delete c;
    ctx->var_cache = nullptr;
}

// var = X.var() over the rows x cols logical elements of a device matrix with leading dimension ld (NumPy's bits);
// *nonfinite = 1 when some element is NaN or infinite (var is then whatever NumPy would have produced from it).
// `want_variance == 0` runs the finite test only.
extern "C" int svmb200_device_variance(svmb200_ctx* ctx, const double* dX, int64_t rows, int64_t cols, int64_t ld,
                                       int want_variance, double* var, int* nonfinite) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(dX != nullptr && rows > 0 && cols > 0 && ld >= cols, "bad argument");
    SVM_CHECK_ARG(var != nullptr || !want_variance, "var is null");
    const long long count = (long long)rows * cols;
    if (!ctx->var_cache) ctx->var_cache = new VarianceCache();
    VarianceCache& c = *static_cast<VarianceCache*>(ctx->var_cache);
    cudaStream_t s = ctx->stream;
    if (c.count != count) {
        std::vector<long long> off;
        off.reserve((size_t)(count / 64 + 2));
        leaf_offsets(0, count, off);
        off.push_back(count);
        const size_t nl = off.size() - 1;
        SVM_CHECK_ARG(nl < (1u << 30), "matrix too large");
        if (nl > c.cap_leaves) {
            SVM_CUDA(cudaStreamSynchronize(s));
            if (c.d_off) cudaFree(c.d_off);
            if (c.d_sums) cudaFree(c.d_sums);
            if (c.h_sums) cudaFreeHost(c.h_sums);
            c.d_off = nullptr;
            c.d_sums = nullptr;
            c.h_sums = nullptr;
            c.cap_leaves = 0;
            c.count = -1;
            SVM_CUDA(cudaMalloc(&c.d_off, (nl + 1) * sizeof(long long)));
            SVM_CUDA(cudaMalloc(&c.d_sums, nl * sizeof(double)));
            SVM_CUDA(cudaMallocHost(&c.h_sums, nl * sizeof(double)));
            c.cap_leaves = nl;
        }
        if (!c.d_flag) SVM_CUDA(cudaMalloc(&c.d_flag, sizeof(int)));
        if (!c.h_flag) SVM_CUDA(cudaMallocHost(&c.h_flag, sizeof(int)));
        SVM_CUDA(cudaMemcpyAsync(c.d_off, off.data(), (nl + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
        SVM_CUDA(cudaStreamSynchronize(s));  // the staging vector goes out of scope
        c.nleaves = (int)nl;
        c.count = count;
    }
    LeafArgs a;
    a.X = dX;
    a.cols = cols;
    a.ld = ld;
    a.count = count;
    a.leaf_off = c.d_off;
    a.nleaves = c.nleaves;
    a.mean = 0.0;
    a.sums = c.d_sums;
    a.nonfinite = c.d_flag;
    const unsigned grid = (unsigned)((c.nleaves + LEAF_NT / LEAF_LANES - 1) / (LEAF_NT / LEAF_LANES));
    SVM_CUDA(cudaMemsetAsync(c.d_flag, 0, sizeof(int), s));
    pairwise_leaf_kernel<false><<<grid, LEAF_NT, 0, s>>>(a);
    ctx->launches++;
    SVM_CUDA(cudaGetLastError());
    SVM_CUDA(cudaMemcpyAsync(c.h_flag, c.d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (want_variance) SVM_CUDA(cudaMemcpyAsync(c.h_sums, c.d_sums, (size_t)c.nleaves * sizeof(double), cudaMemcpyDeviceToHost, s));
    SVM_CUDA(cudaStreamSynchronize(s));
    if (nonfinite) *nonfinite = *c.h_flag;
    if (!want_variance) return SVMB200_OK;
    size_t next = 0;
    const double mean = combine(count, c.h_sums, next) / (double)count;
    a.mean = mean;
    pairwise_leaf_kernel<true><<<grid, LEAF_NT, 0, s>>>(a);
    ctx->launches++;
    SVM_CUDA(cudaGetLastError());
    SVM_CUDA(cudaMemcpyAsync(c.h_sums, c.d_sums, (size_t)c.nleaves * sizeof(double), cudaMemcpyDeviceToHost, s));
    SVM_CUDA(cudaStreamSynchronize(s));
    next = 0;
    *var = combine(count, c.h_sums, next) / (double)count;
    return SVMB200_OK;
}

// dOut[i][:] = dX[idx[i]][:] for i < nidx (rows of d doubles; leading dimensions ld_src / ld_out, pad columns zero).
// idx_host: int64 row indices on the host (the support set, ml/svm/_base.py:867-869).  Asynchronous on the context's stream.
extern "C" int svmb200_gather_rows(svmb200_ctx* ctx, const double* dX, int64_t nrows_src, int64_t ld_src, int64_t d,
                                   const int64_t* idx_host, int64_t nidx, double* dOut, int64_t ld_out) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(dX != nullptr && dOut != nullptr && d > 0 && ld_src >= d && ld_out >= d && nidx >= 0, "bad argument");
    if (nidx == 0) return SVMB200_OK;
    SVM_CHECK_ARG(idx_host != nullptr, "null index");
    for (int64_t i = 0; i < nidx; ++i) SVM_CHECK_ARG(idx_host[i] >= 0 && idx_host[i] < nrows_src, "row index out of range");
    SVM_TRY(svm_scratch_reserve(ctx, &ctx->idx_buf, &ctx->idx_bytes, (size_t)nidx * sizeof(long long)));
    cudaStream_t s = ctx->stream;
    SVM_CUDA(cudaMemcpyAsync(ctx->idx_buf, idx_host, (size_t)nidx * sizeof(long long), cudaMemcpyHostToDevice, s));
    SVM_CUDA(cudaStreamSynchronize(s));  // pageable source: the caller may reuse idx_host right away
    const int rows_per_cta = 8;
    gather_rows_kernel<<<(unsigned)((nidx + rows_per_cta - 1) / rows_per_cta), rows_per_cta * 32, 0, s>>>(
        dX, ld_src, d, static_cast<const long long*>(ctx->idx_buf), nidx, dOut, ld_out);
    ctx->launches++;
    SVM_CUDA(cudaGetLastError());
    return SVMB200_OK;
}
