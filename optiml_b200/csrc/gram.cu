// K1: Gram / Hessian build.  out = s_a s_b^T o (k(A, B) + bias), FP64.
//
// Restates  optiml/ml/svm/kernels.py:49-51 (linear), :91-95 (poly), :125-129 (gaussian through
// sklearn euclidean_distances: D = -2AB' + |a|^2 + |b|^2, max(D,0), diag = 0 when B is A) fused with
// optiml/ml/svm/_base.py:554,628 (Q = K o yy' + yy') / :1098-1099,1178 (M = K + 1).
//
// Structure (sm_100a): persistent CTAs, one 128x128 output tile at a time.
//   warp 8      : TMA producer -- cp.async.bulk.tensor 2D boxes of 128 rows x 16 doubles (128 B,
//                 SWIZZLE_128B) for A and B into a 4-stage shared-memory ring, mbarrier full/empty.
//   warps 0..7  : FP64 tensor-core contraction, mma.sync.m8n8k4 (DMMA; tcgen05 has no f64 kind),
//                 64x32 accumulator block per warp in registers, fused exp/pow/sign/bias epilogue,
//                 128-bit stores.  The producer keeps prefetching the next tile during the epilogue.
// Shared-memory reads are bank-conflict free: with the 128B swizzle the 16-byte chunk c of row r
// lives at chunk c^(r&7); the four k-slots of an m8n8k4 fragment are mapped to chunks {s, s+4}
// (s = k-step), so the 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs.
#include "common.cuh"
#include <cudaTypedefs.h>
#include <math.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int MMA_WARPS = 8;
constexpr int GRAM_THREADS = (MMA_WARPS + 1) * 32;
constexpr int TILE_BYTES = BM * BK * 8;            // 16 KB per operand per stage
constexpr int STAGE_BYTES = 2 * TILE_BYTES;        // 32 KB
constexpr int GRAM_SMEM = STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8;

struct GramArgs {
    const double* norm_a;  // |a_i|^2 per row of A (gaussian) or null
    const double* norm_b;
    const double* sign_a;  // +-1 per row or null
    const double* sign_b;
    double* out;
    long long ldo;
    long long row0, nrows;  // rows of A handled by this launch
    long long nb;           // valid columns (rows of B)
    int kchunks;
    int tiles_m, tiles_n;
    int same;
    double gamma, coef0, degree, bias;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// The transcendental bodies are kept out of line: the epilogue is unrolled 64x per thread (static
// accumulator indices) and 64 inlined copies of exp()/pow() would not fit the instruction cache.
__device__ __noinline__ double exp_outofline(double x) { return exp(x); }
__device__ __noinline__ double pow_outofline(double x, double y) { return pow(x, y); }

template <int KERNEL>
__device__ __forceinline__ double kernel_epilogue(double dot, double na, double nb, bool diag, const GramArgs& p) {
    if (KERNEL == SVMB200_KERNEL_LINEAR) return dot;
    if (KERNEL == SVMB200_KERNEL_POLY) {
        // (gamma * <a,b> + coef0) ** degree with separately rounded product and sum (NumPy semantics)
        return pow_outofline(__dadd_rn(__dmul_rn(p.gamma, dot), p.coef0), p.degree);
    }
    // gaussian: D = (-2<a,b> + |a|^2) + |b|^2 ; clamp ; exact zero on the diagonal of a self-Gram
    double dist = __dadd_rn(__dadd_rn(-2.0 * dot, na), nb);
    dist = fmax(dist, 0.0);
    if (diag) dist = 0.0;
    return exp_outofline(__dmul_rn(-p.gamma, dist));
}

template <int KERNEL>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GramArgs p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment required by the 128B swizzle pattern
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, MMA_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int ntiles = p.tiles_m * p.tiles_n;

    if (warp == MMA_WARPS) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
                const int arow = (int)p.row0 + tm * BM;
                const int brow = tn * BN;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const uint32_t dst = smem_base + stage * STAGE_BYTES;
                    tma_load_2d(dst, &map_a, fb, kc * BK, arow);
                    tma_load_2d(dst + TILE_BYTES, &map_b, fb, kc * BK, brow);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        return;
    }

    // ===================== DMMA consumers =====================
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps -> 64 x 32 per warp
    const int g = lane >> 2, t = lane & 3;
    // byte offset of this lane's k-slot inside a 128-byte row, before the per-step XOR
    const int hi = (t >> 1) * 4, sub = (t & 1) * 8;
    int stage = 0;
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(full0 + 8 * stage, phase);
            const unsigned char* sa = smem + stage * STAGE_BYTES + (wm * 64 + g) * 128;
            const unsigned char* sb = smem + stage * STAGE_BYTES + TILE_BYTES + (wn * 32 + g) * 128;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int off = (((s + hi) ^ g) << 4) + sub;
                double a[8], b[4];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) a[mi] = *reinterpret_cast<const double*>(sa + mi * 8 * 128 + off);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double*>(sb + ni * 8 * 128 + off);
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * stage);
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }

        // ---- fused epilogue: kernel function, bias, label signs, 128-bit stores
        const long long col_base = (long long)tn * BN + wn * 32 + 2 * t;
        double nbv[4][2], sbv[4][2];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long c = col_base + ni * 8 + e;
                const bool ok = c < p.nb;
                nbv[ni][e] = (KERNEL == SVMB200_KERNEL_GAUSSIAN && ok) ? p.norm_b[c] : 0.0;
                sbv[ni][e] = (p.sign_b != nullptr && ok) ? p.sign_b[c] : 1.0;
            }
        }
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            const long long i = p.row0 + (long long)tm * BM + wm * 64 + mi * 8 + g;  // row of A
            if (i >= p.row0 + p.nrows) continue;
            const double na = (KERNEL == SVMB200_KERNEL_GAUSSIAN) ? p.norm_a[i] : 0.0;
            const double sa_i = (p.sign_a != nullptr) ? p.sign_a[i] : 1.0;
            double* orow = p.out + (i - p.row0) * p.ldo;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const long long c = col_base + ni * 8;
                if (c >= p.ldo) continue;
                double2 v;
                double* ve = reinterpret_cast<double*>(&v);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double val = 0.0;
                    if (c + e < p.nb) {
                        val = kernel_epilogue<KERNEL>(acc[mi][ni][e], na, nbv[ni][e], p.same && (i == c + e), p);
                        if (p.bias != 0.0) val = __dadd_rn(val, p.bias);
                        val = __dmul_rn(val, __dmul_rn(sa_i, sbv[ni][e]));
                    }
                    ve[e] = val;
                }
                *reinterpret_cast<double2*>(orow + c) = v;
            }
        }
    }
}

// |x_i|^2 per row: one warp per row, fixed lane-strided order + butterfly
__global__ void row_sqnorm_kernel(const double* __restrict__ X, long long n, long long ld, long long d,
                                  double* __restrict__ out) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const double* x = X + row * ld;
    double s = 0.0;
    for (long long k = lane; k < d; k += 32) s = fma(x[k], x[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

int make_tensor_map(svmb200_ctx* ctx, CUtensorMap* map, const double* base, int64_t rows, int64_t d, int64_t ld) {
    if (!ctx->encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SVM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
            svmb200_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return SVMB200_ERR_CUDA;
        }
        ctx->encode_tiled = fn;
    }
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ctx->encode_tiled);
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        svmb200_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%lld ld=%lld)", (int)r,
                          (long long)rows, (long long)d, (long long)ld);
        return SVMB200_ERR_CUDA;
    }
    return SVMB200_OK;
}

template <int KERNEL>
int launch_gram(svmb200_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const GramArgs& args) {
    static bool configured = false;
    if (!configured) {
        SVM_CUDA(cudaFuncSetAttribute(gram_kernel<KERNEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM));
        configured = true;
    }
    const int ntiles = args.tiles_m * args.tiles_n;
    const int grid = ntiles < ctx->sm_count ? ntiles : ctx->sm_count;
    gram_kernel<KERNEL><<<grid, GRAM_THREADS, GRAM_SMEM, ctx->stream>>>(ma, mb, args);
    ctx->launches++;
    SVM_CUDA(cudaGetLastError());
    return SVMB200_OK;
}

}  // namespace

extern "C" int64_t svmb200_padded_ld(int64_t ncols) { return round_up64(ncols < 1 ? 1 : ncols, 16); }

extern "C" int svmb200_gram(svmb200_ctx* ctx, const double* dA, int64_t na, int64_t lda, const double* dB, int64_t nb,
                            int64_t ldb, int64_t d, int same, int kernel, double gamma, double coef0, double degree,
                            const double* dsign_a, const double* dsign_b, double bias, int64_t row0, int64_t nrows,
                            double* dout, int64_t ldo) {
    SVM_TRY(svm_use(ctx));
    SVM_CHECK_ARG(dA && dB && dout, "null matrix");
    SVM_CHECK_ARG(na > 0 && nb > 0 && d > 0, "empty operand");
    SVM_CHECK_ARG(lda >= d && ldb >= d && lda % 2 == 0 && ldb % 2 == 0, "lda/ldb must be even and >= d (16-byte row stride for TMA)");
    SVM_CHECK_ARG((reinterpret_cast<uintptr_t>(dA) & 15) == 0 && (reinterpret_cast<uintptr_t>(dB) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(dout) & 15) == 0, "matrices must be 16-byte aligned");
    SVM_CHECK_ARG(ldo >= nb && ldo % 2 == 0, "ldo must be even and >= nb");
    SVM_CHECK_ARG(row0 >= 0 && nrows >= 0 && row0 + nrows <= na, "row range outside A");
    SVM_CHECK_ARG(kernel >= SVMB200_KERNEL_LINEAR && kernel <= SVMB200_KERNEL_GAUSSIAN, "unknown kernel id");
    SVM_CHECK_ARG(na < (1ll << 31) && nb < (1ll << 31) && d < (1ll << 31), "dimension too large");
    if (nrows == 0) return SVMB200_OK;

    double *norm_a = nullptr, *norm_b = nullptr;
    if (kernel == SVMB200_KERNEL_GAUSSIAN) {
        SVM_CUDA(cudaMalloc(&norm_a, (size_t)na * sizeof(double)));
        const int wpb = 8;
        row_sqnorm_kernel<<<(unsigned)((na + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(dA, na, lda, d, norm_a);
        ctx->launches++;
        if (dB == dA && nb == na) {
            norm_b = norm_a;
        } else {
            if (cudaMalloc(&norm_b, (size_t)nb * sizeof(double)) != cudaSuccess) {
                cudaFree(norm_a);
                svmb200_set_error("gram: out of device memory");
                return SVMB200_ERR_CUDA;
            }
            row_sqnorm_kernel<<<(unsigned)((nb + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(dB, nb, ldb, d, norm_b);
            ctx->launches++;
        }
    }
    int rc = SVMB200_OK;
    CUtensorMap ma, mb;
    rc = make_tensor_map(ctx, &ma, dA, na, d, lda);
    if (rc == SVMB200_OK) rc = make_tensor_map(ctx, &mb, dB, nb, d, ldb);
    if (rc == SVMB200_OK) {
        GramArgs a;
        a.norm_a = norm_a;
        a.norm_b = norm_b;
        a.sign_a = dsign_a;
        a.sign_b = dsign_b;
        a.out = dout;
        a.ldo = ldo;
        a.row0 = row0;
        a.nrows = nrows;
        a.nb = nb;
        a.kchunks = (int)((d + BK - 1) / BK);
        a.tiles_m = (int)((nrows + BM - 1) / BM);
        a.tiles_n = (int)((ldo + BN - 1) / BN);
        a.same = same;
        a.gamma = gamma;
        a.coef0 = coef0;
        a.degree = degree;
        a.bias = bias;
        if (kernel == SVMB200_KERNEL_LINEAR) rc = launch_gram<SVMB200_KERNEL_LINEAR>(ctx, ma, mb, a);
        else if (kernel == SVMB200_KERNEL_POLY) rc = launch_gram<SVMB200_KERNEL_POLY>(ctx, ma, mb, a);
        else rc = launch_gram<SVMB200_KERNEL_GAUSSIAN>(ctx, ma, mb, a);
    }
    if (norm_a || norm_b) {
        // norms are consumed by the kernel just enqueued
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess && rc == SVMB200_OK) {
            svmb200_set_error("gram kernel failed: %s", cudaGetErrorString(e));
            rc = SVMB200_ERR_CUDA;
        }
        if (norm_b && norm_b != norm_a) cudaFree(norm_b);
        if (norm_a) cudaFree(norm_a);
    }
    return rc;
}
