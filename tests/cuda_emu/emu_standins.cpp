// tests/cuda_emu -- stand-ins for the parts of the library that cannot be emulated  (TEST INFRASTRUCTURE ONLY).
//
// gram.cu (TMA + DMMA tensor-core kernel) is NOT compiled into the emulated library.  svmb200_gram below is a plain
// host loop with the same contract (include/svmb200.h) so that whole fits can run through the emulated solver
// kernels; it says nothing about K1 itself, which is checked on the GPU.  (comm.cu IS compiled: CUDA IPC and the six
// NCCL entry points it binds have in-process stand-ins in emu_runtime.cpp, ranks being threads.)
#include "common.cuh"

extern "C" int64_t svmb200_padded_ld(int64_t ncols) { return round_up64(ncols < 1 ? 1 : ncols, 16); }

extern "C" int svmb200_gram(svmb200_ctx* ctx, const double* dA, int64_t na, int64_t lda, const double* dB, int64_t nb,
                            int64_t ldb, int64_t d, int same, int kernel, double gamma, double coef0, double degree,
                            const double* dsign_a, const double* dsign_b, double bias, int64_t row0, int64_t nrows,
                            double* dout, int64_t ldo) {
    SVM_TRY(svm_use(ctx));
    // the argument contract of csrc/gram.cu, message for message
    SVM_CHECK_ARG(dA && dB && dout, "null matrix");
    SVM_CHECK_ARG(na > 0 && nb > 0 && d > 0, "empty operand");
    SVM_CHECK_ARG(lda >= d && ldb >= d && lda % 2 == 0 && ldb % 2 == 0, "lda/ldb must be even and >= d (16-byte row stride for TMA)");
    SVM_CHECK_ARG((reinterpret_cast<uintptr_t>(dA) & 15) == 0 && (reinterpret_cast<uintptr_t>(dB) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(dout) & 15) == 0, "matrices must be 16-byte aligned");
    SVM_CHECK_ARG(ldo >= nb && ldo % 2 == 0, "ldo must be even and >= nb");
    SVM_CHECK_ARG(row0 >= 0 && nrows >= 0 && row0 + nrows <= na, "row range outside A");
    SVM_CHECK_ARG(kernel >= SVMB200_KERNEL_LINEAR && kernel <= SVMB200_KERNEL_LAPLACIAN, "unknown kernel id");
    SVM_CHECK_ARG((kernel != SVMB200_KERNEL_GAUSSIAN && kernel != SVMB200_KERNEL_LAPLACIAN) || gamma >= 0.0,
                  "gamma must be >= 0 for the gaussian / laplacian kernels");
    for (int64_t i = row0; i < row0 + nrows; ++i) {
        const double* a = dA + i * lda;
        double* out = dout + (i - row0) * ldo;
        double na2 = 0.0;
        for (int64_t k = 0; k < d; ++k) na2 += a[k] * a[k];
        for (int64_t j = 0; j < nb; ++j) {
            const double* b = dB + j * ldb;
            double dot = 0.0, nb2 = 0.0, l1 = 0.0;
            for (int64_t k = 0; k < d; ++k) {
                dot += a[k] * b[k];
                nb2 += b[k] * b[k];
                l1 += fabs(a[k] - b[k]);
            }
            double kv = 0.0;
            switch (kernel) {
                case SVMB200_KERNEL_LINEAR: kv = dot; break;
                case SVMB200_KERNEL_POLY: kv = pow(gamma * dot + coef0, degree); break;
                case SVMB200_KERNEL_GAUSSIAN: {
                    double dist = na2 + nb2 - 2.0 * dot;
                    if (dist < 0.0 || (same && i == j)) dist = 0.0;
                    kv = exp(-gamma * dist);
                    break;
                }
                case SVMB200_KERNEL_SIGMOID: kv = tanh(gamma * dot + coef0); break;
                case SVMB200_KERNEL_LAPLACIAN: kv = exp(-gamma * l1); break;
            }
            const double si = dsign_a ? dsign_a[i] : 1.0, sj = dsign_b ? dsign_b[j] : 1.0;
            out[j] = si * sj * (kv + bias);
        }
        for (int64_t j = nb; j < ldo; ++j) out[j] = 0.0;
    }
    ctx->launches++;
    return SVMB200_OK;
}
