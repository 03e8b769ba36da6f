// FP64 pipe probes on B200: vector DFMA rate vs DMMA rate, alone and mixed on the same SM.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
// mixed: even warps DMMA, odd warps DFMA
__global__ void mixed_kernel(double* out, int iters, double a, double b) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp >= 4) {
        iters *= 13;  // same duration as the DMMA warps when alone
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[i];
    } else {
        double c[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = threadIdx.x;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    }
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        float ms = timeit([&]() { dfma_kernel<8><<<148, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DFMA ILP8  148 CTAs x %4d thr: %8.3f ms  %7.2f TFLOP/s\n", threads, ms, 2.0 * 148 * threads * 8.0 * iters / ms / 1e9);
    }
    {
        float ms = timeit([&]() { dfma_kernel<1><<<148, 256>>>(out, iters, 1.0000001, 1e-9); });
        printf("DFMA ILP1  148 CTAs x  256 thr: %8.3f ms  %7.2f TFLOP/s  (dependent chain: %.1f clk/DFMA at 1.9 GHz)\n", ms, 2.0 * 148 * 256 * 1.0 * iters / ms / 1e9, ms * 1e-3 * 1.9e9 / iters);
    }
    for (int threads : {128, 256, 512}) {
        float ms = timeit([&]() { dmma_kernel<16><<<148, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
        printf("DMMA 16acc 148 CTAs x %4d thr: %8.3f ms  %7.2f TFLOP/s\n", threads, ms, 512.0 * 148 * (threads / 32) * 16.0 * (iters / 4) / ms / 1e9);
    }
    {
        float ms = timeit([&]() { mixed_kernel<<<148, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
        double dm = 512.0 * 148 * 4 * 16.0 * (iters / 4), df = 2.0 * 148 * 128 * 8.0 * (iters / 4) * 13;
        printf("MIXED 4 DMMA warps + 4 DFMA warps: %8.3f ms  DMMA %.2f TF + DFMA %.2f TF (if both ran the whole time)\n", ms, dm / ms / 1e9, df / ms / 1e9);
        float a = timeit([&]() { dmma_kernel<16><<<148, 128>>>(out, iters / 4, 1.0000001, 1e-9); });
        float b = timeit([&]() { dfma_kernel<8><<<148, 128>>>(out, iters / 4 * 13, 1.0000001, 1e-9); });
        printf("   alone: 4 DMMA warps %.3f ms, 4 DFMA warps %.3f ms\n", a, b);
    }
    return 0;
}
