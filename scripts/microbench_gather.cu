// Random 32 B-sector access over a multi-GiB array: the practical HBM ceiling for Bloom-filter probes on B200.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mb_gather scripts/microbench_gather.cu
// Run under gpurun; prints G probes/s per access flavour.  (Measurement tool only; not part of the library.)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
template <int FLAVOR>
__device__ __forceinline__ uint32_t probe(uint32_t* p) {
    uint32_t v = 0;
    if (FLAVOR == 0) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 1) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 2) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 3) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 4) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 5) asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 6) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 7) v = atomicOr(p, 0u);                       // atom with return, value unchanged
    else if (FLAVOR == 8) { atomicOr(p, 1u); }                       // red (no return), dirties the sector
    else if (FLAVOR == 9) v = atomicOr(p, 1u);                       // atom with return, dirties
    else if (FLAVOR == 10) v = atomicCAS(p, 0u, 1u);
    else if (FLAVOR == 11) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    } else if (FLAVOR == 13) asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAVOR == 14) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    }
    else if (FLAVOR == 12) { asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p)); if (!(v & 2u)) atomicOr(p, 2u); }  // test then set
    return v;
}
template <int FLAVOR, int ILP>
__global__ void __launch_bounds__(256) k_gather(uint32_t* a, uint64_t mask_words, uint64_t n_per_thread, uint64_t seed, uint32_t* sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += ILP) {
        uint32_t v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = probe<FLAVOR>(a + (mix64(seed ^ (t * n_per_thread + i + j)) & mask_words));
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc ^= v[j];
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int FLAVOR>
void run(const char* name, uint32_t* a, uint64_t words, uint32_t* sink, int blocks_per_sm, int sms) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint64_t per = 256;
    const int grid = sms * blocks_per_sm * 8;
    k_gather<FLAVOR, 8><<<grid, 256>>>(a, words - 1, per, 1, sink);  // warm-up
    CK(cudaEventRecord(e0));
    k_gather<FLAVOR, 8><<<grid, 256>>>(a, words - 1, per, 2, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double probes = (double)grid * 256 * per;
    printf("%-34s %8.2f G probes/s  (%6.1f GB/s at 32 B/probe)  %.2f ms\n", name, probes / ms / 1e6, probes * 32 / ms / 1e6, ms);
}
int main(int argc, char** argv) {
    const int gran = argc > 1 ? atoi(argv[1]) : 0;
    const uint64_t gib = argc > 2 ? atoll(argv[2]) : 8;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set fetch granularity %d -> %s\n", gran, cudaGetErrorString(e)); }
    size_t lim = 0; CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity)); printf("cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs, L2 %d MB, array %llu GiB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, (unsigned long long)gib);
    const uint64_t words = gib << 28;
    uint32_t *a, *sink; CK(cudaMalloc(&a, words * 4)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(a, 0, words * 4));
    const int sms = p.multiProcessorCount;
    run<0>("ld.global.cg", a, words, sink, 8, sms);
    run<1>("ld.global.ca", a, words, sink, 8, sms);
    run<2>("ld.global.cv", a, words, sink, 8, sms);
    run<3>("ld.global.cs", a, words, sink, 8, sms);
    run<4>("ld.global.nc.L1::no_allocate", a, words, sink, 8, sms);
    run<5>("ld.global.L2::64B", a, words, sink, 8, sms);
    run<6>("ld.relaxed.gpu", a, words, sink, 8, sms);
    run<11>("ld L2::cache_hint evict_first", a, words, sink, 8, sms);
    run<14>("ld L2::cache_hint evict_last", a, words, sink, 8, sms);
    run<13>("ld.global.L2::128B", a, words, sink, 8, sms);
    run<7>("atom.or 0 (return)", a, words, sink, 8, sms);
    run<8>("red.or 1", a, words, sink, 8, sms);
    run<9>("atom.or 1 (return)", a, words, sink, 8, sms);
    run<10>("atom.cas", a, words, sink, 8, sms);
    CK(cudaMemset(a, 0, words * 4));
    run<12>("ld.cg then red if clear", a, words, sink, 8, sms);
    return 0;
}
