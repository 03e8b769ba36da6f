// Exploration harness for K2 (streaming FP64 matvec): sweeps load width, rows per CTA, unroll, column
// segmentation and a TMA-bulk (cp.async.bulk + mbarrier) variant on a 20 GB and a 2.5 GB matrix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mv_sweep mv_sweep.cu
#include <cuda_runtime.h>
#include <string.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ double2 ld128(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
struct d4 { double x, y, z, w; };
__device__ __forceinline__ d4 ld256(const void* p) {
    d4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

// ---------------------------------------------------------------- variant A: LDG, (row block) x (column segment)
// work item = R rows x SEGV vectors; items enumerated segment-major inside a row block.
__device__ __forceinline__ double2 ld128ef(const double2* p, uint64_t pol) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
template <int R, int NT, int U, int W /*16 or 32 bytes*/, int ORDER = 0, int EV = 0, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) mv_ldg(const double* __restrict__ Q, long long ld, long long nrows,
                                             const double* __restrict__ u, double* __restrict__ wpart,
                                             int nseg, int seg_elems, unsigned* __restrict__ cnt, double* __restrict__ w) {
    constexpr int EPV = W / 8;  // elements per vector
    const long long item = blockIdx.x;
    const long long nrb_ = (nrows + R - 1) / R;
    const long long rb = ORDER == 0 ? item / nseg : item % nrb_;
    const int seg = (int)(ORDER == 0 ? item % nseg : item / nrb_);
    const long long row_base = rb * R;
    uint64_t pol = 0;
    if (EV) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const long long c0 = (long long)seg * seg_elems;
    long long c1 = c0 + seg_elems;
    if (c1 > ld) c1 = ld;
    const int nvec = (int)((c1 - c0) / EPV);
    const char* rows[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        long long rr = row_base + r;
        if (rr >= nrows) rr = nrows - 1;
        rows[r] = reinterpret_cast<const char*>(Q + rr * ld + c0);
    }
    const char* ub = reinterpret_cast<const char*>(u + c0);
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    int c = threadIdx.x;
    for (; c + (U - 1) * NT < nvec; c += U * NT) {
        if (W == 16) {
            double2 qv[U][R], uv[U];
#pragma unroll
            for (int j = 0; j < U; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) qv[j][r] = EV ? ld128ef(reinterpret_cast<const double2*>(rows[r]) + c + j * NT, pol) : ld128(reinterpret_cast<const double2*>(rows[r]) + c + j * NT);
#pragma unroll
            for (int j = 0; j < U; ++j) uv[j] = __ldg(reinterpret_cast<const double2*>(ub) + c + j * NT);
#pragma unroll
            for (int j = 0; j < U; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) { acc[r] = fma(qv[j][r].x, uv[j].x, acc[r]); acc[r] = fma(qv[j][r].y, uv[j].y, acc[r]); }
        } else {
            d4 qv[U][R]; d4 uv[U];
#pragma unroll
            for (int j = 0; j < U; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) qv[j][r] = ld256(rows[r] + (size_t)(c + j * NT) * 32);
#pragma unroll
            for (int j = 0; j < U; ++j) { const double2* p = reinterpret_cast<const double2*>(ub + (size_t)(c + j * NT) * 32); double2 a = __ldg(p), b = __ldg(p + 1); uv[j].x = a.x; uv[j].y = a.y; uv[j].z = b.x; uv[j].w = b.y; }
#pragma unroll
            for (int j = 0; j < U; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r] = fma(qv[j][r].x, uv[j].x, acc[r]); acc[r] = fma(qv[j][r].y, uv[j].y, acc[r]);
                    acc[r] = fma(qv[j][r].z, uv[j].z, acc[r]); acc[r] = fma(qv[j][r].w, uv[j].w, acc[r]);
                }
        }
    }
    for (; c < nvec; c += NT) {
        if (W == 16) {
            double2 uv = __ldg(reinterpret_cast<const double2*>(ub) + c);
#pragma unroll
            for (int r = 0; r < R; ++r) { double2 q = ld128(reinterpret_cast<const double2*>(rows[r]) + c); acc[r] = fma(q.x, uv.x, acc[r]); acc[r] = fma(q.y, uv.y, acc[r]); }
        } else {
            const double2* p = reinterpret_cast<const double2*>(ub + (size_t)c * 32); double2 a = __ldg(p), b = __ldg(p + 1);
#pragma unroll
            for (int r = 0; r < R; ++r) { d4 q = ld256(rows[r] + (size_t)c * 32); acc[r] = fma(q.x, a.x, acc[r]); acc[r] = fma(q.y, a.y, acc[r]); acc[r] = fma(q.z, b.x, acc[r]); acc[r] = fma(q.w, b.y, acc[r]); }
        }
    }
    __shared__ double red[NT / 32][R];
    __shared__ unsigned last_flag;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double v = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r] = v;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) red[wid][r] = acc[r];
    }
    __syncthreads();
    if (threadIdx.x < R) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) v += red[k][threadIdx.x];
        const long long rr = row_base + threadIdx.x;
        if (nseg == 1) { if (rr < nrows) w[rr] = v; }
        else wpart[(size_t)seg * nrows + (rr < nrows ? rr : nrows - 1)] = v;  // clamp rows share a slot: harmless
    }
    if (nseg > 1) {
        // last-arriver combine in fixed segment order
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last_flag = (atomicInc(&cnt[rb], nseg - 1) == (unsigned)(nseg - 1));
        __syncthreads();
        if (last_flag && threadIdx.x < R) {
            __threadfence();
            const long long rr = row_base + threadIdx.x;
            if (rr < nrows) {
                double v = 0.0;
                for (int s = 0; s < nseg; ++s) v += __ldcg(&wpart[(size_t)s * nrows + rr]);
                w[rr] = v;
            }
        }
    }
}

// ---------------------------------------------------------------- variant B: TMA bulk (1D) ring
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra WL;\nDN:\n}\n" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// persistent: each CTA walks items (row block x segment) with stride gridDim.x; stage = R rows x CH doubles
template <int R, int CH, int STAGES, int NCW /*consumer warps*/>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) mv_tma(const double* __restrict__ Q, long long ld, long long nrows,
                                                            const double* __restrict__ u, double* __restrict__ w,
                                                            int nseg, int seg_elems, double* __restrict__ wpart, unsigned* __restrict__ cnt) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int STAGE_BYTES = R * CH * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t full0 = s32(bars), empty0 = s32(bars + STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mb_init(full0 + 8 * s, 1); mb_init(empty0 + 8 * s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long nrb = (nrows + R - 1) / R;
    const long long nitems = nrb * nseg;
    if (warp == NCW) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
                const long long rb = item / nseg; const int seg = (int)(item % nseg);
                const long long c0 = (long long)seg * seg_elems; long long c1 = c0 + seg_elems; if (c1 > ld) c1 = ld;
                for (long long c = c0; c < c1; c += CH) {
                    const int len = (int)((c1 - c) < CH ? (c1 - c) : CH);
                    mb_wait(empty0 + 8 * stage, phase ^ 1);
                    mb_expect(full0 + 8 * stage, (uint32_t)(R * len * 8));
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        long long rr = rb * R + r; if (rr >= nrows) rr = nrows - 1;
                        bulk_g2s(s32(smem + stage * STAGE_BYTES + r * CH * 8), Q + rr * ld + c, (uint32_t)(len * 8), full0 + 8 * stage);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }
    // consumers: NCW warps; thread t handles vectors t, t+NCW*32, ... of each row chunk
    constexpr int NTC = NCW * 32;
    __shared__ double red[NCW][R];
    int stage = 0; uint32_t phase = 0;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const long long rb = item / nseg; const int seg = (int)(item % nseg);
        const long long c0 = (long long)seg * seg_elems; long long c1 = c0 + seg_elems; if (c1 > ld) c1 = ld;
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        for (long long c = c0; c < c1; c += CH) {
            const int len = (int)((c1 - c) < CH ? (c1 - c) : CH);
            const int nvec = len >> 1;
            mb_wait(full0 + 8 * stage, phase);
            const double2* sbase = reinterpret_cast<const double2*>(smem + stage * STAGE_BYTES);
            const double2* u2 = reinterpret_cast<const double2*>(u + c);
            for (int v = threadIdx.x; v < nvec; v += NTC) {
                const double2 uv = __ldg(u2 + v);
#pragma unroll
                for (int r = 0; r < R; ++r) { const double2 q = sbase[r * (CH / 2) + v]; acc[r] = fma(q.x, uv.x, acc[r]); acc[r] = fma(q.y, uv.y, acc[r]); }
            }
            __syncwarp();
            if (lane == 0) mb_arrive(empty0 + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double v = acc[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[r] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NTC));
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) red[warp][r] = acc[r];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NTC));
        if (threadIdx.x < R) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < NCW; ++k) v += red[k][threadIdx.x];
            const long long rr = rb * R + threadIdx.x;
            if (rr < nrows) { if (nseg == 1) w[rr] = v; else wpart[(size_t)seg * nrows + rr] = v; }
        }
        // (segment combine omitted in this exploration variant when nseg > 1: timing only)
    }
}

// ---------------------------------------------------------------- harness
struct Result { const char* name; double gbs; float ms; };
static std::vector<Result> results;

static int g_sustain = 0;
template <typename F>
void timeit(const char* name, double bytes, F launch, int reps = 10) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    if (g_sustain > 0) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < g_sustain; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-60s SUSTAINED x%d: %8.4f ms/launch  %7.1f GB/s\n", name, g_sustain, ms / g_sustain, bytes / (ms / g_sustain) / 1e6);
        fflush(stdout);
        return;
    }
    float best = 1e30f, tot = 0;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms); tot += ms;
    }
    CK(cudaGetLastError());
    printf("%-44s avg %8.4f ms  best %8.4f ms   %7.1f GB/s avg  %7.1f GB/s best\n", name, tot / reps, best, bytes / (tot / reps) / 1e6, bytes / best / 1e6);
    fflush(stdout);
}

template <int R, int NT, int U, int W, int ORDER = 0, int EV = 0, int MINB = 1>
void run_ldg(const char* tag, const double* Q, long long ld, long long nrows, const double* u, double* w, double* wpart, unsigned* cnt, int seg_elems) {
    int nseg = seg_elems <= 0 ? 1 : (int)((ld + seg_elems - 1) / seg_elems);
    if (seg_elems <= 0) seg_elems = (int)ld;
    long long nrb = (nrows + R - 1) / R;
    char name[128]; snprintf(name, sizeof(name), "%s ldg R%d NT%d U%d W%d seg%d(x%d) o%d ev%d mb%d", tag, R, NT, U, W, seg_elems, nseg, ORDER, EV, MINB);
    timeit(name, 8.0 * nrows * ld, [&]() { mv_ldg<R, NT, U, W, ORDER, EV, MINB><<<(unsigned)(nrb * nseg), NT>>>(Q, ld, nrows, u, wpart, nseg, seg_elems, cnt, w); });
}

template <int R, int CH, int STAGES, int NCW>
void run_tma(const char* tag, const double* Q, long long ld, long long nrows, const double* u, double* w, double* wpart, unsigned* cnt, int seg_elems, int ctas_per_sm) {
    int nseg = seg_elems <= 0 ? 1 : (int)((ld + seg_elems - 1) / seg_elems);
    if (seg_elems <= 0) seg_elems = (int)ld;
    const int smem = STAGES * R * CH * 8 + 2 * STAGES * 8;
    CK(cudaFuncSetAttribute(mv_tma<R, CH, STAGES, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    char name[128]; snprintf(name, sizeof(name), "%s tma R%d CH%d ST%d CW%d seg%d grid%dx148 (%d KB)", tag, R, CH, STAGES, NCW, seg_elems, ctas_per_sm, smem / 1024);
    timeit(name, 8.0 * nrows * ld, [&]() { mv_tma<R, CH, STAGES, NCW><<<148 * ctas_per_sm, (NCW + 1) * 32, smem>>>(Q, ld, nrows, u, w, nseg, seg_elems, wpart, cnt); });
}

__global__ void fill(double* p, size_t n, double v) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; size_t st = (size_t)gridDim.x * blockDim.x; for (; i < n; i += st) p[i] = v + (double)(i % 1000) * 1e-3; }
__global__ void copyk(const double2* __restrict__ a, double2* __restrict__ b, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; size_t st = (size_t)gridDim.x * blockDim.x; for (; i < n; i += st) b[i] = a[i]; }

int main(int argc, char** argv) {
    const long long n = 50000, ld = 50000;
    double *Q, *u, *w, *wpart; unsigned* cnt;
    CK(cudaMalloc(&Q, (size_t)n * ld * 8)); CK(cudaMalloc(&u, ld * 8)); CK(cudaMalloc(&w, n * 8)); CK(cudaMalloc(&wpart, (size_t)64 * n * 8)); CK(cudaMalloc(&cnt, n * 4));
    CK(cudaMemset(cnt, 0, n * 4));
    fill<<<148 * 8, 256>>>(Q, (size_t)n * ld, 0.5); fill<<<148, 256>>>(u, ld, 0.25); CK(cudaDeviceSynchronize());
    // reference points: device copy bandwidth (read+write) on 8 GB
    { double* b; CK(cudaMalloc(&b, (size_t)4e9)); timeit("cudaMemcpy D2D 4 GB (r+w bytes)", 8e9, [&]() { cudaMemcpyAsync(b, Q, (size_t)4e9, cudaMemcpyDeviceToDevice); });
      timeit("copy kernel 4 GB (r+w bytes)", 8e9, [&]() { copyk<<<148 * 16, 512>>>((const double2*)Q, (double2*)b, (size_t)4e9 / 16); }); CK(cudaFree(b)); }
    if (argc > 1 && strcmp(argv[1], "shard") == 0) {
        // round 2: the pass over a 1/8 row shard (6272 x 50000 = 2.5 GB) is the least efficient piece of the 8-GPU loop
        // (6.6-6.8 TB/s against 7.0-7.2 on the full matrix): smaller work items for a shorter ramp and tail?
        const long long nrows = 6272;
        const char* tag = "[shard 6272]";
        g_sustain = 2000;
        run_ldg<4, 256, 4, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<2, 256, 8, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<2, 256, 4, 16, 0, 0, 4>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<2, 256, 4, 16, 0, 0, 6>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<4, 256, 4, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 4096);
        run_ldg<2, 256, 8, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 4096);
        run_ldg<4, 128, 4, 16, 0, 0, 6>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<2, 128, 8, 16, 0, 0, 6>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<1, 256, 16, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<4, 256, 4, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        const long long full = 50000;
        g_sustain = 300;
        run_ldg<4, 256, 4, 16, 0, 0, 3>("[full 50000]", Q, ld, full, u, w, wpart, cnt, 8192);
        run_ldg<2, 256, 8, 16, 0, 0, 3>("[full 50000]", Q, ld, full, u, w, wpart, cnt, 8192);
        run_ldg<2, 256, 4, 16, 0, 0, 4>("[full 50000]", Q, ld, full, u, w, wpart, cnt, 8192);
        return 0;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const long long nrows = pass == 0 ? 50000 : 6250;
        const char* tag = pass == 0 ? "[20GB]" : "[2.5GB]";
        g_sustain = 0;
        run_ldg<4, 256, 4, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        g_sustain = pass == 0 ? 400 : 2000;
        run_ldg<4, 256, 4, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<4, 256, 4, 16, 0, 1, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<4, 128, 4, 32, 0, 0, 4>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<4, 128, 8, 16, 0, 0, 4>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<8, 256, 2, 16, 0, 0, 2>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_ldg<4, 256, 4, 16, 0, 0, 2>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
        run_tma<4, 1024, 3, 4>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192, 2);
        run_tma<8, 512, 3, 4>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192, 2);
        run_tma<4, 1024, 3, 2>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192, 2);
        run_ldg<4, 256, 4, 16, 0, 0, 3>(tag, Q, ld, nrows, u, w, wpart, cnt, 8192);
    }
    return 0;
}
