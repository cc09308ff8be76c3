// microbench_smem_slice.cu -- the north-star alternative to the L2-resident apply: filter slices staged in SHARED memory with
// cp.async.bulk (TMA bulk copy, SASS UBLKCP), probes as shared-memory loads / atomics, bulk write-back.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o microbench_smem_slice scripts/microbench_smem_slice.cu
//   ./microbench_smem_slice [probes_in_millions (default 1512)] [filter_GiB (default 8)]
//
// What is compared (same number of probes, same 8 GiB bit array, probes already sorted by slice in both cases):
//   L2   : slices of 64 MiB, all CTAs consume one slice after the other (in-order work counter), ld.global.cg + atomicOr when clear,
//          one answer byte per probe -- what ks_apply_probes does today
//   SMEM : sub-slices of 64 KiB, a CTA owns a sub-slice at a time: bulk copy global -> shared (3-stage mbarrier pipeline), probes
//          against shared memory (LDS for look-ups, ATOMS.OR for inserts), bulk copy shared -> global when a bit was set
// The SMEM variant needs the probes sorted 1024x finer (2^17 sub-slices instead of 2^7 slices), i.e. a SECOND tile-sort level:
// its cost (measured in the engine: ~5.8 ps per record and level chip-wide) is printed beside the kernel time.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
// records of region r = [r * per, (r + 1) * per): random offsets inside the region's slice of 2^bits_log2 bits
__global__ void k_fill(uint32_t* rec, uint64_t n, int bits_log2, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        rec[i] = (uint32_t)(mix64(i ^ seed) & ((1ULL << bits_log2) - 1));
}

// ---- L2-resident slices (the engine's apply kernel, reduced to its memory behaviour) -----------------------------------------------
__device__ __forceinline__ uint32_t ld_keep(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t or_keep(uint32_t* p, uint32_t b, uint64_t pol) {
    uint32_t v;
    asm volatile("atom.global.or.L2::cache_hint.b32 %0, [%1], %2, %3;" : "=r"(v) : "l"(p), "r"(b), "l"(pol) : "memory");
    return v;
}
// HINT = 1: slice accesses carry an L2 evict_last policy, the answer stream is written with st.cs
template <int SET, int HINT>
__global__ void __launch_bounds__(256) k_apply_l2(const uint32_t* __restrict__ rec, uint64_t per_region, int n_regions, int slice_log2, int chunk,
                                                  uint32_t* __restrict__ words, uint8_t* __restrict__ ans, int* counter) {
    __shared__ int s_c;
    const int chunks_per_region = (int)((per_region + chunk - 1) / chunk);
    const int total = chunks_per_region * n_regions;
    constexpr int U = 8;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_c = atomicAdd(counter, 1);
        __syncthreads();
        const int c = s_c;
        if (c >= total) break;
        const int r = c / chunks_per_region;
        const uint64_t first = (uint64_t)r * per_region + (uint64_t)(c % chunks_per_region) * chunk;
        const uint32_t n = (uint32_t)min((uint64_t)chunk, (uint64_t)(r + 1) * per_region - first);
        uint32_t* w0 = words + ((uint64_t)r << (slice_log2 - 5));
        for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 256 * U) {
            uint32_t li[U], wd[U];
#pragma unroll
            for (int u = 0; u < U; ++u) li[u] = (i0 + u * 256 < n) ? __ldcs(rec + first + i0 + u * 256) : 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) wd[u] = (i0 + u * 256 < n) ? (HINT ? ld_keep(w0 + (li[u] >> 5), pol) : __ldcg(w0 + (li[u] >> 5))) : 0u;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * 256 < n) {
                    const uint32_t bit = 1u << (li[u] & 31);
                    if (SET && !(wd[u] & bit)) wd[u] = HINT ? or_keep(w0 + (li[u] >> 5), bit, pol) : atomicOr(w0 + (li[u] >> 5), bit);
                    if (HINT) __stcs(ans + first + i0 + u * 256, (uint8_t)((wd[u] & bit) ? 0x80 : 0));
                    else ans[first + i0 + u * 256] = (wd[u] & bit) ? 0x80 : 0;
                }
        }
    }
}

// ---- shared-memory staged sub-slices ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kSubBytes = 64 * 1024;   // one staged sub-slice: 2^19 bits
constexpr int kStages = 3;
constexpr int kSmemThreads = 512;
// CTA b handles sub-slices b, b + grid, ... ; stage s holds sub-slice number (it % kStages)
template <int SET>
__global__ void __launch_bounds__(kSmemThreads) k_apply_smem(const uint32_t* __restrict__ rec, uint64_t per_sub, int n_sub, uint32_t* __restrict__ words,
                                                            uint8_t* __restrict__ ans) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[kStages];
    __shared__ int dirty;
    uint32_t* stage[kStages];
#pragma unroll
    for (int s = 0; s < kStages; ++s) stage[s] = reinterpret_cast<uint32_t*>(smem + (size_t)s * kSubBytes);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int my_n = (n_sub - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto issue = [&](int it) {   // thread 0 only
        const int s = it % kStages;
        const int sub = (int)blockIdx.x + it * (int)gridDim.x;
        mbar_expect_tx(&full[s], kSubBytes);
        bulk_g2s(stage[s], reinterpret_cast<const unsigned char*>(words) + (size_t)sub * kSubBytes, kSubBytes, &full[s]);
    };
    if (threadIdx.x == 0) for (int it = 0; it < kStages - 1 && it < my_n; ++it) issue(it);
    for (int it = 0; it < my_n; ++it) {
        const int s = it % kStages;
        const int sub = (int)blockIdx.x + it * (int)gridDim.x;
        if (threadIdx.x == 0) {
            if (SET) bulk_wait_read<0>();   // the write-back that last read the stage about to be refilled is done
            if (it + kStages - 1 < my_n) issue(it + kStages - 1);
            dirty = 0;
        }
        mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
        __syncthreads();
        const uint32_t* w = stage[s];
        const uint64_t first = (uint64_t)sub * per_sub;
        int set_any = 0;
        for (uint32_t i = threadIdx.x; i < per_sub; i += kSmemThreads) {
            const uint32_t li = __ldcs(rec + first + i);
            const uint32_t bit = 1u << (li & 31);
            uint32_t wd = w[li >> 5];
            if (SET && !(wd & bit)) { wd = atomicOr(stage[s] + (li >> 5), bit); set_any = 1; }
            ans[first + i] = (wd & bit) ? 0x80 : 0;
        }
        if (SET) {
            if (set_any) dirty = 1;
            fence_async_smem();   // generic-proxy writes to the stage -> visible to the bulk copy
            __syncthreads();
            if (threadIdx.x == 0 && dirty) { bulk_s2g(reinterpret_cast<unsigned char*>(words) + (size_t)sub * kSubBytes, stage[s], kSubBytes); bulk_commit(); }
        }
        __syncthreads();
    }
    if (SET && threadIdx.x == 0) bulk_wait_read<0>();
}

int main(int argc, char** argv) {
    const uint64_t n_probes = (uint64_t)(argc > 1 ? atof(argv[1]) : 1512.0) * 1000000ULL;
    const uint64_t filter_bytes = (uint64_t)(argc > 2 ? atof(argv[2]) : 8.0) * (1ULL << 30);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t *words, *rec;
    uint8_t* ans;
    int* counter;
    CK(cudaMalloc(&words, filter_bytes));
    CK(cudaMalloc(&rec, n_probes * 4 + 4096));
    CK(cudaMalloc(&ans, n_probes + 4096));
    CK(cudaMalloc(&counter, 64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    printf("%d SMs, %.0f M probes against %.1f GiB of bits\n", sms, n_probes / 1e6, filter_bytes / 1073741824.0);
    // ---- L2 variant: 64 MiB slices
    {
        const int slice_log2 = 29;
        const int n_regions = (int)(filter_bytes >> (slice_log2 - 3));
        const uint64_t per = n_probes / n_regions;
        k_fill<<<sms * 8, 256>>>(rec, per * n_regions, slice_log2, 1);
        for (int hint = 0; hint < 2; ++hint)
        for (int set = 0; set < 2; ++set)
            for (int rep = 0; rep < 2; ++rep) {   // rep 0: bits clear (insert: every probe sets); rep 1: bits set
                if (rep == 0) CK(cudaMemset(words, 0, filter_bytes));
                CK(cudaMemset(counter, 0, 4));
                CK(cudaEventRecord(e0));
                if (set && hint) k_apply_l2<1, 1><<<sms * 6, 256>>>(rec, per, n_regions, slice_log2, 2048, words, ans, counter);
                else if (set) k_apply_l2<1, 0><<<sms * 6, 256>>>(rec, per, n_regions, slice_log2, 2048, words, ans, counter);
                else if (hint) k_apply_l2<0, 1><<<sms * 6, 256>>>(rec, per, n_regions, slice_log2, 2048, words, ans, counter);
                else k_apply_l2<0, 0><<<sms * 6, 256>>>(rec, per, n_regions, slice_log2, 2048, words, ans, counter);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("L2   slices 64 MiB %s %-7s %-10s %8.3f ms  %7.1f G probes/s\n", hint ? "evict_last+st.cs" : "no hints        ", set ? "set" : "lookup",
                       rep ? "(bits set)" : "(clear)", ms, per * n_regions / ms / 1e6);
            }
    }
    // ---- SMEM variant: 64 KiB sub-slices, 3-stage bulk-copy pipeline
    {
        const int n_sub = (int)(filter_bytes / kSubBytes);
        const uint64_t per = n_probes / n_sub;
        k_fill<<<sms * 8, 256>>>(rec, per * n_sub, 19, 2);
        const size_t smem = (size_t)kStages * kSubBytes;
        CK(cudaFuncSetAttribute(k_apply_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_apply_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int set = 0; set < 2; ++set)
            for (int rep = 0; rep < 2; ++rep) {
                if (rep == 0) CK(cudaMemset(words, 0, filter_bytes));
                CK(cudaEventRecord(e0));
                if (set) k_apply_smem<1><<<sms, kSmemThreads, smem>>>(rec, per, n_sub, words, ans);
                else k_apply_smem<0><<<sms, kSmemThreads, smem>>>(rec, per, n_sub, words, ans);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double sort2 = per * n_sub * 5.8e-12 * 1e3;
                printf("SMEM sub-slices 64 KiB %-7s %-10s %8.3f ms  %7.1f G probes/s   (+ second sort level ~%.2f ms -> %7.1f G probes/s)\n", set ? "set" : "lookup",
                       rep ? "(bits set)" : "(clear)", ms, per * n_sub / ms / 1e6, sort2, per * n_sub / (ms + sort2) / 1e6);
            }
        // correctness: after the set passes every probed bit is set -> a lookup pass answers 0x80 everywhere
        CK(cudaMemset(ans, 0, per * n_sub));
        k_apply_smem<0><<<sms, kSmemThreads, smem>>>(rec, per, n_sub, words, ans);
        CK(cudaDeviceSynchronize());
        uint8_t* h = (uint8_t*)malloc(1 << 20);
        CK(cudaMemcpy(h, ans + (per * n_sub) / 2, 1 << 20, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < (1 << 20); ++i) bad += h[i] != 0x80;
        printf("SMEM check: %d of %d sampled answers wrong\n", bad, 1 << 20);
    }
    CK(cudaGetLastError());
    return 0;
}
