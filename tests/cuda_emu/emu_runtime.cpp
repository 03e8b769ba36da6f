// tests/cuda_emu -- fiber runtime behind the host stand-in for the CUDA runtime  (TEST INFRASTRUCTURE ONLY).
//
// A launch runs its blocks one after the other; inside a block every CUDA thread is a fiber with its own stack.
// A fiber runs until it reaches __syncthreads() or a warp shuffle (or returns); when no fiber of the block can run,
// the scheduler releases the barriers that are complete:
//   * __syncthreads(): all threads of the block that have not returned are waiting at it;
//   * __shfl_xor_sync(full mask): all 32 lanes of the warp are waiting at it (a lane that has returned, or that waits
//     at __syncthreads() instead, is an error -- on the GPU that is undefined behaviour with a full mask).
// Anything else is a deadlock and fails the launch (sticky error, reported by cudaGetLastError / the next sync).
// The order in which runnable fibers are resumed -- and the order in which the blocks of a launch run -- is selectable
// (forward, reverse, seeded shuffle): code that is correctly synchronised and whose reductions have a fixed shape
// gives the same bits under every order; a missing barrier, or a result that depends on which CTA arrives last at a
// ticket, usually does not.
// "Device" memory is host memory, filled with 0xFF on allocation (NaNs: nothing may rely on zero-initialisation) and
// fenced by canaries that cudaFree checks.
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <sched.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace emu {

thread_local uint3 threadIdx_ = {0, 0, 0}, blockIdx_ = {0, 0, 0};
thread_local dim3 blockDim_(1, 1, 1), gridDim_(1, 1, 1);

// ---------------------------------------------------------------------------------------------- context switch
#if defined(__x86_64__)
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#else
#error "tests/cuda_emu needs x86-64 (hand-written fiber switch)"
#endif

enum State { READY, AT_SYNC, AT_SHFL, AT_MMA, DONE };

struct Fiber {
    void* sp = nullptr;
    State state = DONE;
    double shfl_val = 0.0;
    int shfl_mask = 0;
    double mma[4] = {0, 0, 0, 0};  // a, b, c0, c1 of a pending m8n8k4 product
    bool progressed = false;       // made progress since it was last resumed (a polling loop that only yields has not)
};

// mbarrier of the emulated block: arrivals still expected in the current phase, transaction bytes still in flight
struct MBar {
    uint32_t init_count = 0, pending = 0, phase = 0;
    long long tx = 0;
};

constexpr size_t STACK_BYTES = 64 * 1024;

struct BlockRunner {
    std::vector<Fiber> fibers;
    char* stacks = nullptr;
    std::vector<char*>* owner = nullptr;
    ~BlockRunner() {
        if (stacks && owner) owner->push_back(stacks);
    }
    void* sched_sp = nullptr;
    int current = -1;
    const std::function<void()>* body = nullptr;
    // K1 support: dynamic shared memory (1024-byte aligned), mbarriers by shared offset, named barriers
    unsigned char* smem = nullptr;
    size_t smem_bytes = 0;
    std::map<uint32_t, MBar> mbars;
    int named_count[16] = {};
    unsigned named_gen[16] = {};
};

static thread_local BlockRunner* g_run = nullptr;
static thread_local int g_sticky_error = 0;
static thread_local std::string g_sticky_text;
static int g_order = 0;  // 0 forward, 1 reverse, 2 shuffle
static unsigned g_seed = 1;

static void fail(const std::string& what) {
    if (!g_sticky_error) {
        g_sticky_error = cudaErrorLaunchFailure;
        g_sticky_text = what;
        fprintf(stderr, "[cuda_emu] %s\n", what.c_str());
    }
}

static void yield_to_scheduler() {
    BlockRunner* r = g_run;
    emu_switch(&r->fibers[r->current].sp, r->sched_sp);
}

extern "C" void emu_fiber_main() {
    BlockRunner* r = g_run;
    (*r->body)();
    r->fibers[r->current].state = DONE;
    yield_to_scheduler();
    abort();  // a finished fiber is never resumed
}

void syncthreads() {
    g_run->fibers[g_run->current].state = AT_SYNC;
    yield_to_scheduler();
}

double shfl_xor(double v, int lane_mask) {
    Fiber& f = g_run->fibers[g_run->current];
    f.shfl_val = v;
    f.shfl_mask = lane_mask;
    f.state = AT_SHFL;
    yield_to_scheduler();
    return g_run->fibers[g_run->current].shfl_val;
}

// a polling loop inside a kernel (mbarrier wait, named barrier): let the other threads of the block run
static void yield_polling() {
    Fiber& f = g_run->fibers[g_run->current];
    f.progressed = false;
    yield_to_scheduler();   // state stays READY: resumed in the next round
}

unsigned char* dynamic_smem() { return g_run->smem; }

// The dynamic shared memory of a block starts 16 bytes past a 1024-byte boundary -- in the host address AND in the
// emulated shared-window offsets -- so a kernel that needs an aligned tile has to round up itself, whether it does so
// on the generic address or on the window offset.
constexpr uint32_t SMEM_WINDOW_SKEW = 16;

uint32_t smem_offset(const void* p) {
    const unsigned char* q = static_cast<const unsigned char*>(p);
    if (q < g_run->smem || q >= g_run->smem + g_run->smem_bytes) {
        fail("shared-window address of a pointer outside the dynamic shared memory");
        return 0;
    }
    return (uint32_t)(q - g_run->smem) + SMEM_WINDOW_SKEW;
}
static unsigned char* smem_ptr(uint32_t off) { return g_run->smem + (off - SMEM_WINDOW_SKEW); }

static void mbar_complete_if_done(MBar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1u;
        b.pending = b.init_count;
    }
}
void mbar_init(uint32_t bar, uint32_t count) {
    MBar& b = g_run->mbars[bar];
    b.init_count = b.pending = count;
    b.phase = 0;
    b.tx = 0;
}
void mbar_arrive(uint32_t bar, uint32_t expect_tx_bytes) {
    auto it = g_run->mbars.find(bar);
    if (it == g_run->mbars.end() || it->second.pending == 0) {
        fail("arrival at an mbarrier that was not initialised or expects no more arrivals in this phase");
        return;
    }
    it->second.tx += expect_tx_bytes;
    it->second.pending -= 1;
    mbar_complete_if_done(it->second);
}
void mbar_wait(uint32_t bar, uint32_t parity) {
    for (;;) {
        auto it = g_run->mbars.find(bar);
        if (it == g_run->mbars.end()) {
            fail("wait on an mbarrier that was not initialised");
            return;
        }
        if (it->second.phase != (parity & 1u)) return;  // the phase of that parity has completed
        if (g_sticky_error) return;
        yield_polling();
    }
}

void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    // box of map->box[1] rows x map->box[0] elements starting at (inner c0, outer c1); out-of-bounds elements are
    // zero; rows are 128 bytes apart in shared memory and, with the 128-byte swizzle, the 16-byte chunk j of a row
    // whose shared address has bits [7:9] = r lands at chunk j ^ r
    const size_t row_bytes = (size_t)map->box[0] * map->elem_bytes;
    unsigned char* base = smem_ptr(dst);
    if (map->swizzle != CU_TENSOR_MAP_SWIZZLE_128B || row_bytes != 128 || (reinterpret_cast<uintptr_t>(base) & 1023u) != 0) {
        fail("emulated TMA: only 128-byte rows with the 128-byte swizzle into 1024-byte aligned tiles are modelled");
        return;
    }
    if (base + (size_t)map->box[1] * row_bytes > g_run->smem + g_run->smem_bytes) {
        fail("emulated TMA: destination tile exceeds the dynamic shared memory");
        return;
    }
    for (uint32_t r = 0; r < map->box[1]; ++r) {
        const long long gr = (long long)c1 + r;
        for (uint32_t e = 0; e < map->box[0]; ++e) {
            const long long gc = (long long)c0 + e;
            double v = 0.0;
            if (gr >= 0 && gr < (long long)map->dim[1] && gc >= 0 && gc < (long long)map->dim[0])
                memcpy(&v, map->base + (size_t)gr * map->row_stride + (size_t)gc * map->elem_bytes, sizeof(v));
            const uint32_t byte = e * map->elem_bytes, chunk = byte >> 4, within = byte & 15u;
            memcpy(base + (size_t)r * 128 + (((chunk ^ (r & 7u)) << 4) | within), &v, sizeof(v));
        }
    }
    auto it = g_run->mbars.find(bar);
    if (it == g_run->mbars.end()) {
        fail("emulated TMA: completion mbarrier was not initialised");
        return;
    }
    it->second.tx -= (long long)map->box[1] * (long long)row_bytes;
    mbar_complete_if_done(it->second);
}

void named_barrier(int id, int nthreads, bool wait) {
    BlockRunner* r = g_run;
    if (id < 0 || id >= 16) {
        fail("named barrier id out of range");
        return;
    }
    const unsigned gen = r->named_gen[id];
    if (++r->named_count[id] >= nthreads) {
        r->named_count[id] = 0;
        ++r->named_gen[id];
        return;
    }
    if (!wait) return;  // bar.arrive
    while (r->named_gen[id] == gen && !g_sticky_error) yield_polling();
}

void syncwarp() { (void)shfl_xor(0.0, 0); }

void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    Fiber& f = g_run->fibers[g_run->current];
    f.mma[0] = a;
    f.mma[1] = b;
    f.mma[2] = c0;
    f.mma[3] = c1;
    f.state = AT_MMA;
    yield_to_scheduler();
    const Fiber& g = g_run->fibers[g_run->current];
    c0 = g.mma[2];
    c1 = g.mma[3];
}

bool spin_wait(unsigned long long spins) {
    static thread_local std::chrono::steady_clock::time_point t0;
    if (spins == 0) t0 = std::chrono::steady_clock::now();
    if ((spins & 63) == 63) {
        sched_yield();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) return true;
    }
    return false;
}

static void prepare_fiber(BlockRunner& r, int t) {
    char* top = r.stacks + (size_t)(t + 1) * STACK_BYTES;
    top = reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(top) & ~uintptr_t(15));
    void** sp = reinterpret_cast<void**>(top);
    *--sp = nullptr;                                    // keeps (rsp + 8) 16-byte aligned at entry
    *--sp = reinterpret_cast<void*>(&emu_fiber_main);   // `ret` of the first switch lands here
    for (int i = 0; i < 6; ++i) *--sp = nullptr;        // rbp rbx r12 r13 r14 r15
    r.fibers[t].sp = sp;
    r.fibers[t].state = READY;
}

static bool run_block(BlockRunner& r, int nt, std::mt19937& rng) {
    for (int t = 0; t < nt; ++t) prepare_fiber(r, t);
    r.mbars.clear();
    for (int i = 0; i < 16; ++i) r.named_count[i] = 0, r.named_gen[i] = 0;
    std::vector<int> order(nt);
    for (int t = 0; t < nt; ++t) order[t] = t;
    int live = nt;
    unsigned long long idle_rounds = 0;
    while (live > 0) {
        if (g_order == 1) std::reverse(order.begin(), order.end());
        else if (g_order == 2) std::shuffle(order.begin(), order.end(), rng);
        bool progress = false;
        for (int t : order) {
            if (r.fibers[t].state != READY) continue;
            r.current = t;
            threadIdx_.x = (unsigned)t;
            r.fibers[t].progressed = true;   // a polling loop that only yields clears it again
            emu_switch(&r.sched_sp, r.fibers[t].sp);
            if (r.fibers[t].state == DONE) --live;
            progress = progress || r.fibers[t].progressed;
        }
        if (live == 0) break;
        // release the barriers that are complete
        int at_sync = 0, ready = 0;
        for (int t = 0; t < nt; ++t) {
            at_sync += r.fibers[t].state == AT_SYNC;
            ready += r.fibers[t].state == READY;
        }
        bool released = false;
        if (at_sync == live) {
            for (int t = 0; t < nt; ++t)
                if (r.fibers[t].state == AT_SYNC) r.fibers[t].state = READY;
            released = true;
        }
        for (int w0 = 0; w0 < nt && !released; w0 += 32) {
            const int w1 = std::min(w0 + 32, nt);
            int at_shfl = 0, at_mma = 0;
            for (int t = w0; t < w1; ++t) {
                at_shfl += r.fibers[t].state == AT_SHFL;
                at_mma += r.fibers[t].state == AT_MMA;
            }
            if (at_shfl == w1 - w0) {
                double vals[32];
                for (int t = w0; t < w1; ++t) vals[t - w0] = r.fibers[t].shfl_val;
                for (int t = w0; t < w1; ++t) {
                    const int src = (t - w0) ^ r.fibers[t].shfl_mask;
                    r.fibers[t].shfl_val = (src >= 0 && src < w1 - w0) ? vals[src] : vals[t - w0];
                    r.fibers[t].state = READY;
                }
                released = true;
            } else if (at_mma == 32) {
                // mma.sync.m8n8k4 f64: lane l holds A[l/4][l%4], B[l%4][l/4] and C[l/4][2(l%4) + {0,1}]
                double A[8][4], B[4][8];
                for (int l = 0; l < 32; ++l) {
                    A[l >> 2][l & 3] = r.fibers[w0 + l].mma[0];
                    B[l & 3][l >> 2] = r.fibers[w0 + l].mma[1];
                }
                for (int l = 0; l < 32; ++l) {
                    const int i = l >> 2;
                    for (int e = 0; e < 2; ++e) {
                        const int j = 2 * (l & 3) + e;
                        double c = r.fibers[w0 + l].mma[2 + e];
                        for (int k = 0; k < 4; ++k) c = fma(A[i][k], B[k][j], c);
                        r.fibers[w0 + l].mma[2 + e] = c;
                    }
                    r.fibers[w0 + l].state = READY;
                }
                released = true;
            } else if ((at_shfl > 0 || at_mma > 0) && ready == 0 && at_sync + at_shfl + at_mma == live) {
                // everybody is blocked and this warp can never complete its collective
                fail("warp collective (shuffle / mma with a full mask) reached by only part of the warp (block " +
                     std::to_string(blockIdx_.x) + "," + std::to_string(blockIdx_.y) + ", warp " + std::to_string(w0 / 32) + ")");
                return false;
            }
        }
        if (released || progress) {
            idle_rounds = 0;
            continue;
        }
        if (ready == 0) {
            fail("deadlock: __syncthreads() not reached by all live threads of block " + std::to_string(blockIdx_.x) + "," +
                 std::to_string(blockIdx_.y));
            return false;
        }
        if (++idle_rounds > 100000) {  // only polling loops are running and nothing they poll ever changes
            fail("deadlock: threads of block " + std::to_string(blockIdx_.x) + "," + std::to_string(blockIdx_.y) +
                 " poll an mbarrier / named barrier that nobody completes");
            return false;
        }
    }
    return true;
}

static thread_local uint64_t g_launches = 0, g_blocks = 0;

void launch(dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()>& thread_body) {
    if (g_sticky_error) return;
    if (block.y != 1 || block.z != 1 || grid.z != 1 || block.x == 0 || block.x > 1024) {
        fail("unsupported launch shape");
        return;
    }
    const int nt = (int)block.x;
    BlockRunner r;
    r.fibers.resize(nt);
    {
        // fiber stacks are recycled between launches (zero-filling 16 MB per launch would dominate tiny kernels)
        struct StackPool : std::vector<char*> {
            ~StackPool() {  // a rank thread of the multi-rank tests ends: give its stacks back
                for (char* p : *this) free(p);
            }
        };
        static thread_local StackPool pool;
        r.stacks = pool.empty() ? static_cast<char*>(malloc((size_t)(1024 + 1) * STACK_BYTES)) : pool.back();
        if (!pool.empty()) pool.pop_back();
        if (!r.stacks) {
            fail("out of memory for fiber stacks");
            return;
        }
        r.owner = &pool;
    }
    r.body = &thread_body;
    std::vector<unsigned char> smem_storage;
    if (dynamic_smem_bytes > 0) {
        // handed out at an address that is NOT 1024-byte aligned (see SMEM_WINDOW_SKEW)
        smem_storage.assign(dynamic_smem_bytes + 2048, 0xFF);
        unsigned char* aligned = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_storage.data()) + 1023) & ~uintptr_t(1023));
        r.smem = aligned + SMEM_WINDOW_SKEW;
        r.smem_bytes = dynamic_smem_bytes;
    }
    BlockRunner* outer = g_run;
    g_run = &r;
    std::mt19937 rng(g_seed + (unsigned)g_launches * 7919u);
    blockDim_ = block;
    gridDim_ = grid;
    ++g_launches;
    // blocks run one after the other -- in index order, or (order 2) in a seeded shuffle: which CTA arrives last at a
    // ticket must not change a bit of the result
    std::vector<uint64_t> blocks((size_t)grid.x * grid.y);
    for (size_t i = 0; i < blocks.size(); ++i) blocks[i] = i;
    if (g_order == 2) std::shuffle(blocks.begin(), blocks.end(), rng);
    else if (g_order == 1) std::reverse(blocks.begin(), blocks.end());
    for (uint64_t id : blocks) {
        if (g_sticky_error) break;
        blockIdx_ = {(unsigned)(id % grid.x), (unsigned)(id / grid.x), 0};
        ++g_blocks;
        if (!run_block(r, nt, rng)) break;
    }
    g_run = outer;
}

// Cooperative launch: one HOST THREAD per block, all alive together -- the fiber runtime (g_run, threadIdx_, the
// "__shared__" statics, the stack pool) is thread_local, so every block gets its own.  A grid barrier built on global
// atomics then completes exactly as on the GPU: the polling thread of a block spins on the host (spin_wait yields the
// CPU) until the other blocks arrive.  An error in any block becomes the launching thread's sticky error.
void launch_cooperative(dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()>& thread_body) {
    if (g_sticky_error) return;
    const size_t nblocks = (size_t)grid.x * grid.y;
    if (block.y != 1 || block.z != 1 || grid.z != 1 || block.x == 0 || block.x > 1024 || nblocks == 0 || nblocks > 160) {
        fail("unsupported cooperative launch shape (the emulation runs one host thread per block: <= 160 blocks)");
        return;
    }
    ++g_launches;
    g_blocks += nblocks;
    std::vector<std::string> errors(nblocks);
    std::vector<std::thread> threads;
    const unsigned seed = g_seed + (unsigned)g_launches * 7919u;
    for (size_t id = 0; id < nblocks; ++id) {
        threads.emplace_back([&, id]() {
            const int nt = (int)block.x;
            BlockRunner r;
            r.fibers.resize(nt);
            r.stacks = static_cast<char*>(malloc((size_t)(nt + 1) * STACK_BYTES));
            if (!r.stacks) {
                errors[id] = "out of memory for fiber stacks";
                return;
            }
            r.owner = nullptr;
            r.body = &thread_body;
            std::vector<unsigned char> smem_storage;
            if (dynamic_smem_bytes > 0) {
                smem_storage.assign(dynamic_smem_bytes + 2048, 0xFF);
                unsigned char* aligned = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_storage.data()) + 1023) & ~uintptr_t(1023));
                r.smem = aligned + SMEM_WINDOW_SKEW;
                r.smem_bytes = dynamic_smem_bytes;
            }
            g_run = &r;
            blockDim_ = block;
            gridDim_ = grid;
            blockIdx_ = {(unsigned)(id % grid.x), (unsigned)(id / grid.x), 0};
            std::mt19937 rng(seed + (unsigned)id);
            run_block(r, nt, rng);
            g_run = nullptr;
            if (g_sticky_error) errors[id] = g_sticky_text;
            free(r.stacks);
            r.stacks = nullptr;
        });
    }
    for (auto& t : threads) t.join();
    for (const std::string& e : errors)
        if (!e.empty()) {
            fail("cooperative launch: " + e);
            break;
        }
}

// ---------------------------------------------------------------------------------------------- memory
constexpr size_t GUARD = 256;
constexpr unsigned char CANARY = 0xA5;
struct Alloc {
    size_t bytes;
    bool pinned;
};
static std::mutex g_mem_mutex;
static std::map<void*, Alloc> g_allocs;

static cudaError_t alloc_fenced(void** p, size_t bytes, bool pinned) {
    if (!p) return cudaErrorInvalidValue;
    void* raw = nullptr;
    if (posix_memalign(&raw, 256, bytes + 2 * GUARD) != 0) return cudaErrorMemoryAllocation;
    unsigned char* base = static_cast<unsigned char*>(raw);
    memset(base, CANARY, GUARD);
    memset(base + GUARD, 0xFF, bytes);  // NaN doubles, -1 integers: nothing may rely on fresh memory being zero
    memset(base + GUARD + bytes, CANARY, GUARD);
    *p = base + GUARD;
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    g_allocs[*p] = Alloc{bytes, pinned};
    return cudaSuccess;
}

static cudaError_t free_fenced(void* p) {
    if (!p) return cudaSuccess;
    Alloc a;
    {
        std::lock_guard<std::mutex> lock(g_mem_mutex);
        auto it = g_allocs.find(p);
        if (it == g_allocs.end()) {
            fail("free of a pointer that was not allocated (or double free)");
            return cudaErrorInvalidValue;
        }
        a = it->second;
        g_allocs.erase(it);
    }
    unsigned char* base = static_cast<unsigned char*>(p) - GUARD;
    for (size_t i = 0; i < GUARD; ++i) {
        if (base[i] != CANARY || base[GUARD + a.bytes + i] != CANARY) {
            fail("out-of-bounds write detected around an allocation of " + std::to_string(a.bytes) + " bytes");
            break;
        }
    }
    free(base);
    return cudaSuccess;
}

}  // namespace emu

// ---------------------------------------------------------------------------------------------- runtime API subset
cudaError_t cudaGetDeviceCount(int* count) {
    *count = 1;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* prop, int) {
    prop->major = 10;
    prop->minor = 0;
    prop->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t* free_bytes, size_t* total_bytes) {
    *free_bytes = *total_bytes = (size_t)180 << 30;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t bytes) { return emu::alloc_fenced(p, bytes, false); }
cudaError_t cudaFree(void* p) { return emu::free_fenced(p); }
cudaError_t cudaMallocHost(void** p, size_t bytes) { return emu::alloc_fenced(p, bytes, true); }
cudaError_t cudaFreeHost(void* p) { return emu::free_fenced(p); }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
    memmove(dst, src, bytes);
    return emu::g_sticky_error;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r)
        memmove(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t) {
    memset(p, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int value, size_t bytes) {
    memset(p, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    memset(h, 0, sizeof(*h));
    memcpy(h->reserved, &p, sizeof(p));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    memcpy(p, h.reserved, sizeof(*p));
    return *p ? cudaSuccess : cudaErrorInvalidValue;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = reinterpret_cast<cudaStream_t>(new int(0));
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    delete reinterpret_cast<int*>(s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return emu::g_sticky_error; }
struct emu_event {
    std::chrono::steady_clock::time_point t;
};
cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = new emu_event();
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return emu::g_sticky_error; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
static CUresult emu_encode_tiled(CUtensorMap* map, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* gdim,
                                 const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estride,
                                 CUtensorMapInterleave interleave, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill) {
    // the constraints of the real encoder that this code base can run into
    if (type != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 || rank != 2 || interleave != CU_TENSOR_MAP_INTERLEAVE_NONE) return CUDA_ERROR_INVALID_VALUE;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (gstride[0] & 15) != 0) return CUDA_ERROR_INVALID_VALUE;
    if (box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256 || estride[0] != 1 || estride[1] != 1) return CUDA_ERROR_INVALID_VALUE;
    if (swizzle == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * 8 > 128) return CUDA_ERROR_INVALID_VALUE;
    memset(map, 0, sizeof(*map));
    map->base = static_cast<const unsigned char*>(base);
    map->dim[0] = gdim[0];
    map->dim[1] = gdim[1];
    map->row_stride = gstride[0];
    map->box[0] = box[0];
    map->box[1] = box[1];
    map->swizzle = (uint32_t)swizzle;
    map->elem_bytes = 8;
    return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* res) {
    if (strcmp(symbol, "cuTensorMapEncodeTiled") == 0) {
        *fn = reinterpret_cast<void*>(&emu_encode_tiled);
        *res = cudaDriverEntryPointSuccess;
    } else {
        *fn = nullptr;
        *res = cudaDriverEntryPointSymbolNotFound;
    }
    return cudaSuccess;
}
cudaError_t cudaGetLastError() { return emu::g_sticky_error; }
const char* cudaGetErrorString(cudaError_t e) {
    if (e == cudaSuccess) return "no error";
    return emu::g_sticky_text.empty() ? "emulated CUDA error" : emu::g_sticky_text.c_str();
}

// ---------------------------------------------------------------------------------------------- test hooks
extern "C" {
// 0: threads resumed in index order, 1: reversed after every barrier, 2: shuffled (seeded)
void emu_set_schedule(int order, unsigned seed) {
    emu::g_order = order;
    emu::g_seed = seed;
}
int emu_live_allocations(void) {
    std::lock_guard<std::mutex> lock(emu::g_mem_mutex);
    return (int)emu::g_allocs.size();
}
int emu_sticky_error(void) { return emu::g_sticky_error; }
void emu_clear_error(void) {
    emu::g_sticky_error = 0;
    emu::g_sticky_text.clear();
}
uint64_t emu_launches(void) { return emu::g_launches; }
uint64_t emu_blocks(void) { return emu::g_blocks; }
}

// ---------------------------------------------------------------------------------------------- NCCL stand-in
// Ranks are THREADS of the test process (one svmb200 context each).  A communicator is a slot in a "world" keyed by
// the unique id; ncclCommInitRank blocks until every rank has joined (like NCCL), ncclAllGather is a barrier, a copy
// of every peer's slice out of the peer's own buffer, and a second barrier.  comm.cu binds these instead of
// dlopen("libnccl.so.2") when it is compiled for the emulation.
namespace emu {
struct World {
    int nranks = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t generation = 0;
    std::vector<void*> bufs;
    bool wait(double seconds = 60.0) {  // all ranks arrive, or time out (a rank died): false
        std::unique_lock<std::mutex> lock(m);
        const uint64_t gen = generation;
        if (++arrived == nranks) {
            arrived = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        return cv.wait_for(lock, std::chrono::duration<double>(seconds), [&] { return generation != gen; });
    }
};
struct Comm {
    std::shared_ptr<World> world;
    int rank;
};
static std::mutex g_world_mutex;
static std::map<uint64_t, std::shared_ptr<World>> g_worlds;
static uint64_t g_next_id = 1;
}  // namespace emu

struct ncclUniqueIdEmu {
    char internal[128];
};
extern "C" {
int emu_ncclGetUniqueId(ncclUniqueIdEmu* id) {
    std::lock_guard<std::mutex> lock(emu::g_world_mutex);
    memset(id, 0, sizeof(*id));
    const uint64_t v = emu::g_next_id++;
    memcpy(id->internal, &v, sizeof(v));
    return 0;
}
int emu_ncclCommInitRank(void** comm, int nranks, ncclUniqueIdEmu id, int rank) {
    uint64_t key = 0;
    memcpy(&key, id.internal, sizeof(key));
    std::shared_ptr<emu::World> w;
    {
        std::lock_guard<std::mutex> lock(emu::g_world_mutex);
        auto& slot = emu::g_worlds[key];
        if (!slot) {
            slot = std::make_shared<emu::World>();
            slot->nranks = nranks;
            slot->bufs.assign((size_t)nranks, nullptr);
        }
        w = slot;
    }
    if (w->nranks != nranks || rank < 0 || rank >= nranks) return 4;
    if (!w->wait()) return 6;
    *comm = new emu::Comm{w, rank};
    return 0;
}
int emu_ncclCommDestroy(void* comm) {
    delete static_cast<emu::Comm*>(comm);
    return 0;
}
static std::atomic<uint64_t> g_allgathers{0};
uint64_t emu_allgather_calls(void) { return g_allgathers.load(); }
int emu_ncclAllGather(const void*, void* recv, size_t count, int dtype, void* comm, cudaStream_t) {
    emu::Comm* c = static_cast<emu::Comm*>(comm);
    ++g_allgathers;
    if (dtype != 8) return 4;  // ncclFloat64
    emu::World& w = *c->world;
    w.bufs[(size_t)c->rank] = recv;
    if (!w.wait()) return 6;
    for (int r = 0; r < w.nranks; ++r) {
        if (r == c->rank) continue;
        memcpy(static_cast<double*>(recv) + (size_t)r * count, static_cast<const double*>(w.bufs[(size_t)r]) + (size_t)r * count,
               count * sizeof(double));
    }
    if (!w.wait()) return 6;
    return 0;
}
const char* emu_ncclGetErrorString(int code) { return code == 6 ? "emulated NCCL: a rank did not arrive" : "emulated NCCL error"; }
int emu_ncclGetVersion(int* v) {
    *v = 22809;
    return 0;
}
}
