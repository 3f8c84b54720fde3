// cuda_runtime.h of tests/cuda_emu -- TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH AND NOT A FALLBACK.
//
// The build container has no GPU.  To check the LOGIC of the sm_100a kernels there (indexing, hashing, the delta-scoring
// bookkeeping, the packed scan tables) this header lets g++ compile find_tfbs_b200/csrc/{kernels.cuh,tfbs.cu} unchanged and runs
// every kernel launch on the host: one block at a time, every CUDA thread of the block as a cooperative fiber (ucontext), with
// __syncthreads / __syncwarp / warp shuffles / ballot implemented as fiber barriers.  Races, memory-ordering bugs and anything
// about speed are invisible here; those are what the `-m gpu` tests and bench.py on a real B200 are for.
//
// The result is tests/cuda_emu/libtfbs_emu.so, loaded only by tests/test_kernel_logic_emulated.py.  The product library
// libtfbs_b200.so is built by nvcc from the same sources, contains no host implementation of the kernels and still refuses to
// create a context without a B200 (tests/test_abi.py::test_no_gpu_fails_loudly).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define TFBS_EMULATED 1

// ---- qualifiers ------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ /* libstdc++ spells __attribute__((__noinline__)): must expand to nothing */
#define __shared__ static thread_local  /* blocks run one at a time per host thread; two contexts may run on two threads */
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- the fiber scheduler -----------------------------------------------------------------------------------------
namespace emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 96 * 1024;

struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    bool done;
};

struct Barrier {
    unsigned alive = 0, arrived = 0, generation = 0;
};

// Barrier of the lanes named by a *_sync mask.  Disjoint masks (e.g. four groups of 8 lanes) synchronise independently, like on
// the hardware, and so do different masks that share lanes (a group barrier and a later full-warp one).  Lanes of the mask that
// have left the kernel count as arrived.
struct MaskBarrier {
    unsigned mask = 0, arrived = 0, generation = 0;
};

struct State {
    ucontext_t sched;
    Fiber fibers[kMaxThreads];
    char* stacks = nullptr;
    Barrier block_bar;
    Barrier warp_bar[kMaxThreads / 32];            // alive lanes per warp (bookkeeping for the mask barriers)
    MaskBarrier mask_bar[kMaxThreads / 32][32];
    unsigned alive_mask[kMaxThreads / 32];
    unsigned long long warp_slot[kMaxThreads / 32][32];
    unsigned n_threads = 0;
    int cur = -1;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    void (*entry)(void*) = nullptr;
    void* entry_arg = nullptr;
    alignas(16) unsigned char dyn_smem[256 * 1024];
};

inline State& S() {
    static thread_local State* s = nullptr;  // one emulated device per host thread that launches (contexts are per thread)
    if (!s) {
        s = new State();
        s->stacks = (char*)malloc(kStackBytes * kMaxThreads);
    }
    return *s;
}

inline void yield() {
    State& s = S();
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void barrier_wait(Barrier& b) {
    const unsigned gen = b.generation;
    if (++b.arrived >= b.alive) {
        b.arrived = 0;
        ++b.generation;
        return;
    }
    while (b.generation == gen) yield();
}

// a thread that has left the kernel counts as arrived at every later barrier
inline void barrier_leave(Barrier& b) {
    --b.alive;
    if (b.alive && b.arrived >= b.alive) {
        b.arrived = 0;
        ++b.generation;
    }
}

inline void mask_barrier_wait(unsigned mask) {
    State& s = S();
    const unsigned w = (unsigned)s.cur >> 5, lane = (unsigned)s.cur & 31u;
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "cuda_emu: lane %u calls a *_sync primitive with mask %08x that does not name it\n", lane, mask); abort(); }
    // the barrier in use for this mask, else a free one (different masks of one warp are different barriers, as on the hardware)
    MaskBarrier* b = nullptr;
    for (MaskBarrier& x : s.mask_bar[w])
        if (x.arrived && x.mask == mask) { b = &x; break; }
    if (!b)
        for (MaskBarrier& x : s.mask_bar[w])
            if (!x.arrived) { b = &x; b->mask = mask; break; }
    if (!b) { fprintf(stderr, "cuda_emu: too many *_sync masks in flight in one warp\n"); abort(); }
    const unsigned gen = b->generation;
    if (++b->arrived >= (unsigned)__builtin_popcount(mask & s.alive_mask[w])) {
        b->arrived = 0;
        ++b->generation;
        return;
    }
    while (b->generation == gen) yield();
}

inline void mask_barrier_leave(unsigned w, unsigned lane) {
    State& s = S();
    s.alive_mask[w] &= ~(1u << lane);
    for (MaskBarrier& b : s.mask_bar[w])
        if (b.arrived && b.arrived >= (unsigned)__builtin_popcount(b.mask & s.alive_mask[w])) {
            b.arrived = 0;
            ++b.generation;
        }
}

inline void trampoline() {
    State& s = S();
    s.entry(s.entry_arg);
    Fiber& f = s.fibers[s.cur];
    f.done = true;
    barrier_leave(s.block_bar);
    barrier_leave(s.warp_bar[s.cur / 32]);
    mask_barrier_leave((unsigned)s.cur / 32, (unsigned)s.cur & 31u);
    swapcontext(&f.ctx, &s.sched);
}

// Runs entry(arg) once per thread of every block, blocks one after the other.
inline void run_grid(dim3 grid, dim3 block, void (*entry)(void*), void* arg) {
    State& s = S();
    const unsigned n = block.x * block.y * block.z;
    if (n == 0 || n > (unsigned)kMaxThreads) { fprintf(stderr, "cuda_emu: bad block size %u\n", n); abort(); }
    s.bdim = block;
    s.gdim = grid;
    s.entry = entry;
    s.entry_arg = arg;
    s.n_threads = n;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                s.bid = uint3{bx, by, bz};
                s.block_bar = Barrier{n, 0, 0};
                for (unsigned w = 0; w < (n + 31) / 32; ++w) {
                    const unsigned lanes = std::min(32u, n - 32 * w);
                    s.warp_bar[w] = Barrier{lanes, 0, 0};
                    s.alive_mask[w] = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
                    for (MaskBarrier& mb : s.mask_bar[w]) mb = MaskBarrier{};
                }
                for (unsigned t = 0; t < n; ++t) {
                    Fiber& f = s.fibers[t];
                    f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    f.done = false;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = s.stacks + kStackBytes * t;
                    f.ctx.uc_stack.ss_size = kStackBytes;
                    f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                }
                unsigned remaining = n;
                while (remaining) {
                    for (unsigned t = 0; t < n; ++t) {
                        Fiber& f = s.fibers[t];
                        if (f.done) continue;
                        s.cur = (int)t;
                        swapcontext(&s.sched, &f.ctx);
                        if (f.done) --remaining;
                    }
                }
                s.cur = -1;
            }
}

template <class K>
struct Launcher {
    K kernel;
    dim3 grid, block;
    template <class... Args>
    void operator()(Args... args) {
        auto call = [&]() { kernel(args...); };
        run_grid(grid, block, [](void* p) { (*static_cast<decltype(call)*>(p))(); }, &call);
    }
};
template <class K>
Launcher<K> launcher(K kernel, dim3 grid, dim3 block) { return Launcher<K>{kernel, grid, block}; }

inline unsigned lane_id() { return (unsigned)S().cur & 31u; }
inline unsigned warp_id() { return (unsigned)S().cur >> 5; }

template <class T>
inline T shuffle(unsigned mask, T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    State& s = S();
    const unsigned w = warp_id();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    s.warp_slot[w][lane_id()] = raw;
    mask_barrier_wait(mask);
    raw = s.warp_slot[w][src_lane & 31u];
    mask_barrier_wait(mask);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

}  // namespace emu

#define threadIdx (emu::S().fibers[emu::S().cur].tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

#define TFBS_LAUNCH(kernel, grid, block, smem, stream) emu::launcher(kernel, dim3(grid), dim3(block))
#define TFBS_DYNAMIC_SHARED(name) unsigned char* name = emu::S().dyn_smem

// ---- device intrinsics ---------------------------------------------------------------------------------------------
inline void __syncthreads() { emu::barrier_wait(emu::S().block_bar); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::mask_barrier_wait(mask); }
template <class T>
inline T __shfl_sync(unsigned m, T v, int src) { return emu::shuffle(m, v, (unsigned)src); }
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return emu::shuffle(m, v, emu::lane_id() ^ (unsigned)lane_mask); }
template <class T>
inline T __shfl_up_sync(unsigned m, T v, unsigned delta) {
    const unsigned lane = emu::lane_id();
    T got = emu::shuffle(m, v, lane >= delta ? lane - delta : lane);
    return lane >= delta ? got : v;
}
template <class T>
inline T __shfl_down_sync(unsigned m, T v, unsigned delta) {
    const unsigned lane = emu::lane_id();
    T got = emu::shuffle(m, v, lane + delta < 32 ? lane + delta : lane);
    return lane + delta < 32 ? got : v;
}
inline unsigned __ballot_sync(unsigned m, int pred) {
    emu::State& s = emu::S();
    const unsigned w = emu::warp_id();
    s.warp_slot[w][emu::lane_id()] = pred ? 1ull : 0ull;
    emu::mask_barrier_wait(m);
    unsigned out = 0;
    for (unsigned l = 0; l < 32; ++l)
        if (((m & s.alive_mask[w]) >> l) & 1u)
            if (s.warp_slot[w][l]) out |= 1u << l;
    emu::mask_barrier_wait(m);
    return out;
}
inline unsigned __reduce_or_sync(unsigned m, unsigned v) {
    emu::State& s = emu::S();
    const unsigned w = emu::warp_id();
    s.warp_slot[w][emu::lane_id()] = v;
    emu::mask_barrier_wait(m);
    unsigned out = 0;
    for (unsigned l = 0; l < 32; ++l)
        if (((m & s.alive_mask[w]) >> l) & 1u) out |= (unsigned)s.warp_slot[w][l];
    emu::mask_barrier_wait(m);
    return out;
}
inline unsigned __activemask() { return emu::S().alive_mask[emu::warp_id()]; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }

// one host thread runs all fibers, so plain read-modify-write is atomic
template <class T, class U>
inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <class T, class U>
inline T atomicSub(T* p, U v) { T old = *p; *p = (T)(old - (T)v); return old; }
template <class T, class U>
inline T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <class T, class U>
inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <class T, class U, class V>
inline T atomicCAS(T* p, U cmp, V val) { T old = *p; if (old == (T)cmp) *p = (T)val; return old; }
template <class T, class U>
inline T atomicXor(T* p, U v) { T old = *p; *p = (T)(old ^ (T)v); return old; }
template <class T, class U>
inline T atomicOr(T* p, U v) { T old = *p; *p = (T)(old | (T)v); return old; }
template <class T, class U>
inline T atomicExch(T* p, U v) { T old = *p; *p = (T)v; return old; }

inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ---- runtime API (host memory stands in for device memory) ---------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct emuStream* cudaStream_t;
struct emuEvent { std::chrono::steady_clock::time_point t; };
typedef emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostRegisterDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin, totalGlobalMem;
};

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated allocation failure"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "cuda_emu (host fibers, test only)");
    p->major = 10;
    p->multiProcessorCount = 3;  // persistent kernels launch one CTA per SM: keep the emulated grid small
    p->sharedMemPerBlockOptin = 227 * 1024;
    p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
// device memory comes back filled with garbage, like the real thing: a kernel that relies on zeroed memory fails here
inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = malloc(n ? n : 1);
    if (*p) memset(*p, 0xCD, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) {
    *p = malloc(n ? n : 1);
    if (*p) memset(*p, 0xCD, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)calloc(1, 8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }            // launches run to completion inside the call
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
