// tests/cuda_emu -- host stand-in for <cuda_runtime.h>  (TEST INFRASTRUCTURE ONLY, never part of the product).
//
// Lets the CPU test-suite compile optiml_b200/csrc/*.cu with g++ and run the REAL kernel source thread by thread: every
// CUDA thread of a block is a fiber, __syncthreads(), the warp shuffles and the m8n8k4 tensor-core product are fiber
// barriers / warp collectives, mbarriers, tiled (TMA) copies and named barriers are modelled (emu_runtime.cpp), blocks
// run one after the other, "device" memory is host memory with canaries.  What this checks: indexing, barrier
// placement, producer / consumer protocols, launch sequences, double-buffering, arithmetic order -- not timing, not
// memory ordering between CTAs (blocks are serial), not PTX.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

// ---- language extensions
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __grid_constant__
#define EMU_NOINLINE __attribute__((noinline))   // build.py rewrites __noinline__ (libstdc++ uses that token itself)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct alignas(8) int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
struct alignas(16) ulonglong2 { unsigned long long x, y; };

namespace emu {
extern thread_local uint3 threadIdx_, blockIdx_;
extern thread_local dim3 blockDim_, gridDim_;
void syncthreads();
double shfl_xor(double v, int lane_mask);
bool spin_wait(unsigned long long spins);  // called from a wait loop on peer memory: yield; true once 20 s have passed
// K1 (csrc/gram.cu): dynamic shared memory, mbarriers, tiled (TMA) copies, the FP64 tensor-core product, named barriers
unsigned char* dynamic_smem();
uint32_t smem_offset(const void* p);                            // stand-in for the 32-bit shared-window address
void mbar_init(uint32_t bar, uint32_t count);
void mbar_arrive(uint32_t bar, uint32_t expect_tx_bytes);       // arrive (+ expect_tx)
void mbar_wait(uint32_t bar, uint32_t parity);                  // until the phase of that parity has completed
void named_barrier(int id, int nthreads, bool wait);            // bar.sync / bar.arrive
void syncwarp();
void dmma_m8n8k4(double& c0, double& c1, double a, double b);   // D = A(8x4, row) * B(4x8, col) + C, one warp
}  // namespace emu
struct CUtensorMap;
namespace emu {
void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1);  // box copy, zero fill, swizzle, complete_tx
}  // namespace emu
#define threadIdx (emu::threadIdx_)
#define blockIdx (emu::blockIdx_)
#define blockDim (emu::blockDim_)
#define gridDim (emu::gridDim_)

// ---- device intrinsics used by the kernels
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() {}
static inline void __syncwarp() { emu::syncwarp(); }
static inline int __double2loint(double v) { long long r; memcpy(&r, &v, 8); return (int)(unsigned)(r & 0xffffffffll); }
static inline int __double2hiint(double v) { long long r; memcpy(&r, &v, 8); return (int)(unsigned)((unsigned long long)r >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    const unsigned long long r = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double v; memcpy(&v, &r, 8); return v;
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) { return emu::shfl_xor(v, lane_mask); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline unsigned atomicInc(unsigned* addr, unsigned val) {
    const unsigned old = *addr;
    *addr = old >= val ? 0u : old + 1u;
    return old;
}

// ---- runtime API subset
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp {
    int major, minor, multiProcessorCount;
};

cudaError_t cudaGetDeviceCount(int* count);
cudaError_t cudaSetDevice(int device);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* prop, int device);
cudaError_t cudaMemGetInfo(size_t* free_bytes, size_t* total_bytes);
cudaError_t cudaMalloc(void** p, size_t bytes);
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMallocHost((void**)p, bytes); }
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaMemset(void* p, int value, size_t bytes);
// CUDA IPC between "ranks" that are threads of this process: the handle carries the pointer
struct cudaIpcMemHandle_t {
    char reserved[64];
};
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
enum { cudaErrorPeerAccessAlreadyEnabled = 704 };
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);

enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaEnableDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* res);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- kernel launches: `k<<<grid, block, smem, stream>>>(args)` is rewritten by tests/cuda_emu/build.py into
// emu::launch(grid, block, smem, [=]() { k(args); })
#include <functional>
namespace emu {
void launch(dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()>& thread_body);
// cooperative launch: every block is alive at the same time (one host thread per block) so that a grid barrier built on
// global atomics can complete; grids of a few dozen blocks only
void launch_cooperative(dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()>& thread_body);
}
