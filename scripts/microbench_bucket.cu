// Feasibility numbers for the bucketed pipeline (measurement tool, not part of the library):
//  (1) random 4-byte probes / atomics confined to an L2-sized slice (after a sequential prefetch of the slice)
//  (2) radix scatter of 8-byte records into B buckets with warp-aggregated cursors
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
template <int OP>
__global__ void __launch_bounds__(256) k_slice(uint32_t* a, uint64_t mask_words, uint64_t n_per_thread, uint64_t seed, uint32_t* sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t* p = a + (mix64(seed ^ (t * n_per_thread + i + j)) & mask_words);
            if (OP == 0) v[j] = __ldcg(p);
            else if (OP == 1) { atomicOr(p, 1u << (j & 31)); v[j] = 0; }
            else if (OP == 2) v[j] = atomicOr(p, 1u << (j & 31));
            else v[j] = atomicCAS(p, 0u, 1u);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc ^= v[j];
    }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void k_prefetch(const uint4* a, uint64_t n, uint32_t* sink) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { uint4 v = __ldcg(a + i); acc ^= v.x ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}
// scatter: record i -> bucket (hash & (B-1)), position from a warp-aggregated atomicAdd on the bucket cursor
__global__ void __launch_bounds__(256) k_scatter(const uint64_t* __restrict__ in, uint64_t n, int log_b, uint64_t cap, uint64_t* __restrict__ out, unsigned int* __restrict__ cursor) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = __ldcs(in + i);
        const int b = (int)(mix64(r) >> (64 - log_b));
        const unsigned peers = __match_any_sync(__activemask(), b);
        const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&cursor[b], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        const unsigned int p = base + __popc(peers & ((1u << lane) - 1u));
        if (p < cap) __stcs(out + (uint64_t)b * cap + p, r);
    }
}
__global__ void k_fill(uint64_t* a, uint64_t n) { for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) a[i] = mix64(i * 3 + 1); }
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    uint32_t* sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int mib : {8, 16, 32, 64, 128}) {
        const uint64_t words = (uint64_t)mib << 18;
        uint32_t* a; CK(cudaMalloc(&a, words * 4)); CK(cudaMemset(a, 0, words * 4));
        const char* names[4] = {"ld.cg", "red.or", "atom.or", "atom.cas"};
        for (int op = 0; op < 4; ++op) {
            const int grid = sms * 8; const uint64_t per = 512;
            k_prefetch<<<sms * 4, 256>>>((const uint4*)a, words / 4, sink);
            CK(cudaEventRecord(e0));
            if (op == 0) k_slice<0><<<grid, 256>>>(a, words - 1, per, 7, sink);
            else if (op == 1) k_slice<1><<<grid, 256>>>(a, words - 1, per, 7, sink);
            else if (op == 2) k_slice<2><<<grid, 256>>>(a, words - 1, per, 7, sink);
            else k_slice<3><<<grid, 256>>>(a, words - 1, per, 7, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("slice %4d MiB  %-9s %8.1f G ops/s\n", mib, names[op], (double)grid * 256 * per / ms / 1e6);
        }
        CK(cudaFree(a));
    }
    const uint64_t n = 1ull << 28;  // 268 M records = 2 GiB
    uint64_t *in, *out; unsigned int* cur;
    CK(cudaMalloc(&in, n * 8)); k_fill<<<1184, 256>>>(in, n); CK(cudaDeviceSynchronize());
    for (int lb : {6, 8, 9, 10, 12}) {
        const uint64_t cap = (n >> lb) + (n >> (lb + 4)) + 65536;
        CK(cudaMalloc(&out, (cap << lb) * 8)); CK(cudaMalloc(&cur, 4 << lb));
        // distinct-ish records: reuse `in` filled by index
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(cur, 0, 4 << lb));
            CK(cudaEventRecord(e0));
            k_scatter<<<sms * 16, 256>>>(in, n, lb, cap, out, cur);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        printf("scatter into %5d buckets: %7.2f G records/s  (%.0f GB/s read+write)\n", 1 << lb, n / ms / 1e6, n * 16.0 / ms / 1e6);
        CK(cudaFree(out)); CK(cudaFree(cur));
    }
    return 0;
}
