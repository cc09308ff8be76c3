// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  Never part of librnabloom_gpu.so, never loaded by the product path.
//
// There is no GPU in the development container, so index arithmetic / data-movement mistakes in a kernel would only show
// up on the (scarce) B200 box.  This shim lets g++ compile the *unchanged* kernel sources of rna-bloom_b200/csrc for the
// host (-DRB_EMU): one fiber (or, for ThreadSanitizer, one OS thread) per CUDA thread of a CTA, CTAs one after the other,
// __syncthreads = a barrier, atomics = __atomic builtins, the handful of CUDA runtime calls the host code makes = malloc/memcpy.  tests/test_emu_parity.py
// then runs the same parity tests the GPU suite runs (against the oracle) through the resulting library.
// It checks the LOGIC of the kernels, not their performance or their memory-model behaviour on the device.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <stdio.h>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// Two execution modes of a CTA:
//   fibers (default)        one OS thread, one ucontext fiber per CUDA thread; a fiber runs until it reaches a barrier (CTA- or warp-
//                           wide) and the scheduler releases a barrier when every live participant has arrived.  ~10x faster than OS
//                           threads for the barrier-heavy kernels, deterministic, and atomics are trivially atomic.
//   -DRB_EMU_THREADS        one OS thread per CUDA thread (real concurrency): what tests/emu/race_check.cpp needs under TSan.
namespace emu {
inline EmuDim3 g_blockDim, g_gridDim;
inline unsigned char* g_dyn_smem = nullptr;
inline unsigned long long g_warp_xchg[64][32];
template <typename K>
struct Launcher;
}  // namespace emu

#ifdef RB_EMU_THREADS
namespace emu {
inline thread_local EmuDim3 t_threadIdx, t_blockIdx;
inline std::barrier<>* g_cta_barrier = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> g_warp_barrier;
inline void cta_sync() { g_cta_barrier->arrive_and_wait(); }
inline void warp_sync() { g_warp_barrier[t_threadIdx.x >> 5]->arrive_and_wait(); }
inline void yield_now() { std::this_thread::yield(); }

template <typename F>
inline void run_grid(unsigned grid, unsigned block, size_t smem, F&& body) {
    if (grid == 0 || block == 0) return;
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    g_blockDim.x = block;
    g_gridDim.x = grid;
    std::barrier<> outer((ptrdiff_t)block);
    std::unique_ptr<std::barrier<>> inner;
    auto worker = [&](unsigned tid) {
        for (unsigned b = 0; b < grid; ++b) {
            if (tid == 0) {
                inner.reset(new std::barrier<>((ptrdiff_t)block));
                g_cta_barrier = inner.get();
                g_warp_barrier.clear();
                for (unsigned w = 0; w * 32 < block; ++w)
                    g_warp_barrier.emplace_back(new std::barrier<>((ptrdiff_t)std::min(32u, block - w * 32)));
            }
            outer.arrive_and_wait();
            t_threadIdx.x = tid;
            t_blockIdx.x = b;
            body();
            inner->arrive_and_drop();   // a thread that has left the kernel no longer takes part in __syncthreads
            outer.arrive_and_wait();
        }
    };
    if (block == 1) { worker(0); return; }
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
}
}  // namespace emu
#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#else
#include <ucontext.h>
#include <functional>
namespace emu {
enum { F_RUN = 0, F_CTA = 1, F_WARP = 2, F_DONE = 3 };
struct Fiber { ucontext_t ctx; EmuDim3 tidx; int state; };
inline EmuDim3 g_blockIdx;
inline Fiber* g_cur = nullptr;
inline ucontext_t g_main;
inline std::function<void()>* g_body = nullptr;
inline std::vector<Fiber> g_fibers;
inline std::vector<unsigned char> g_stacks;
constexpr size_t kFiberStack = 256 * 1024;
inline void fiber_entry() {
    (*g_body)();
    g_cur->state = F_DONE;
    swapcontext(&g_cur->ctx, &g_main);
}
inline void block_on(int why) { g_cur->state = why; swapcontext(&g_cur->ctx, &g_main); }
inline void cta_sync() { block_on(F_CTA); }
inline void warp_sync() { block_on(F_WARP); }
inline void yield_now() {}

template <typename F>
inline void run_grid(unsigned grid, unsigned block, size_t smem, F&& body) {
    if (grid == 0 || block == 0) return;
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    g_blockDim.x = block;
    g_gridDim.x = grid;
    std::function<void()> fn = body;
    g_body = &fn;
    if (g_stacks.size() < (size_t)block * kFiberStack) g_stacks.resize((size_t)block * kFiberStack);
    g_fibers.resize(block);
    for (unsigned b = 0; b < grid; ++b) {
        g_blockIdx.x = b;
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = g_fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * kFiberStack;
            f.ctx.uc_stack.ss_size = kFiberStack;
            f.ctx.uc_link = &g_main;
            makecontext(&f.ctx, fiber_entry, 0);
            f.tidx.x = t;
            f.state = F_RUN;
        }
        for (;;) {
            unsigned live = 0, ran = 0;
            for (unsigned t = 0; t < block; ++t) {   // run every runnable fiber until it blocks or finishes
                Fiber& f = g_fibers[t];
                if (f.state == F_RUN) { g_cur = &f; swapcontext(&g_main, &f.ctx); ++ran; }
                if (f.state != F_DONE) ++live;
            }
            if (!live) break;
            // release barriers whose live participants have all arrived
            unsigned at_cta = 0;
            for (unsigned t = 0; t < block; ++t) at_cta += g_fibers[t].state == F_CTA;
            bool released = false;
            if (at_cta == live) { for (unsigned t = 0; t < block; ++t) if (g_fibers[t].state == F_CTA) g_fibers[t].state = F_RUN; released = true; }
            for (unsigned w = 0; w * 32 < block; ++w) {
                unsigned lw = 0, aw = 0;
                for (unsigned t = w * 32; t < std::min(block, w * 32 + 32); ++t) { lw += g_fibers[t].state != F_DONE; aw += g_fibers[t].state == F_WARP; }
                if (lw && aw == lw) { for (unsigned t = w * 32; t < std::min(block, w * 32 + 32); ++t) if (g_fibers[t].state == F_WARP) g_fibers[t].state = F_RUN; released = true; }
            }
            if (!released && !ran) { fprintf(stderr, "cuda_emu: CTA %u deadlocked at a barrier (divergent __syncthreads / partial-warp collective?)\n", b); abort(); }
        }
    }
    g_body = nullptr;
}
}  // namespace emu
#define threadIdx (emu::g_cur->tidx)
#define blockIdx emu::g_blockIdx
#endif

namespace emu {
template <typename K>
struct Launcher {
    unsigned grid, block; size_t smem; K kern;
    template <typename... A>
    void operator()(A... args) const { run_grid(grid, block, smem, [&]() { kern(args...); }); }
};
template <typename K>
inline Launcher<K> launcher(long long grid, long long block, size_t smem, void*, K kern) { return Launcher<K>{(unsigned)grid, (unsigned)block, smem, kern}; }
}  // namespace emu

#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
#define RB_LAUNCH(grid, block, smem, stream, ...) emu::launcher((grid), (block), (smem), (void*)(stream), __VA_ARGS__)
#define RB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)

static inline void __syncthreads() { emu::cta_sync(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu::yield_now(); }

// ---- warp collectives (full-mask, convergent uses only): exchange buffer + two warp barriers --------------------------------------
template <typename T>
static inline T emu_shfl(T v, int src_lane) {
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    emu::g_warp_xchg[w][l] = bits;
    emu::warp_sync();
    unsigned long long got = emu::g_warp_xchg[w][(unsigned)src_lane & 31];
    emu::warp_sync();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (int)((threadIdx.x & 31) ^ (unsigned)m)); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { const int l = (int)(threadIdx.x & 31); const T o = emu_shfl(v, l >= d ? l - d : l); return l >= d ? o : v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src); }
static inline unsigned __ballot_sync(unsigned, int pred) {   // every lane of the warp must call it (full mask)
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    emu::g_warp_xchg[w][l] = pred ? 1ULL : 0ULL;
    emu::warp_sync();
    unsigned m = 0;
    const unsigned lanes = std::min(32u, blockDim.x - w * 32);
    for (unsigned i = 0; i < lanes; ++i) m |= (unsigned)emu::g_warp_xchg[w][i] << i;
    emu::warp_sync();
    return m;
}
static inline unsigned __match_any_sync(unsigned, unsigned value) {   // every lane of the warp must call it (full mask)
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    emu::g_warp_xchg[w][l] = value;
    emu::warp_sync();
    unsigned m = 0;
    const unsigned lanes = std::min(32u, blockDim.x - w * 32);
    for (unsigned i = 0; i < lanes; ++i) if ((unsigned)emu::g_warp_xchg[w][i] == value) m |= 1u << i;
    emu::warp_sync();
    return m;
}
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline void __syncwarp() {   // every lane of the warp must call it
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    emu::warp_sync();
}

// ---- atomics -----------------------------------------------------------------------------------------------------------
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// ---- loads / stores / integer intrinsics ---------------------------------------------------------------------------------
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }   // ld.global.cg: coherent at L2, may race with atomics by design
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __vcmpne4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xFF) != ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
    return r;
}
template <typename A, typename B> static inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <typename A, typename B> static inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }

// ---- the slice of the runtime API the host code uses (everything is synchronous) ----------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaLimitMaxL2FetchGranularity = 5,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 2; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated allocation failure"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t n) {
    void* q = nullptr;
    if (posix_memalign(&q, 256, (n + 255) & ~(size_t)255) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    memset(q, 0xA5, n);   // device memory is not zeroed: make reliance on that visible
    *p = (T*)q;
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return posix_memalign(p, 256, (n + 255) & ~(size_t)255) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <typename K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* occ, K, int, size_t) { *occ = 1; return cudaSuccess; }
