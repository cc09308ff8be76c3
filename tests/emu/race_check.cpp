// race_check.cpp -- TEST INFRASTRUCTURE ONLY: the host emulation of the kernels (cuda_emu.h) under ThreadSanitizer.
// A missing __syncthreads() is invisible to the parity tests of the emulation (OS threads rarely interleave badly) and only
// shows up on the GPU at full size; TSan sees the unordered shared-memory accesses directly.  Built and run by
// tests/test_emu_parity.py::test_kernels_are_race_free_under_tsan:
//   g++ -std=c++20 -O1 -g -fsanitize=thread -DRB_EMU -DRB_EMU_THREADS -I tests/emu -pthread tests/emu/race_check.cpp -o tests/emu/race_check
// (RB_EMU_THREADS: one OS thread per CUDA thread; the default fiber mode of cuda_emu.h is sequential and would hide every race)
#include "../../rna-bloom_b200/csrc/rnabloom_gpu.cu"

#include <cstdio>

#define REQ(call) do { int32_t rc_ = (call); if (rc_) { fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, rb_last_error(ctx)); return 2; } } while (0)

int main(int argc, char** argv) {
    const int n_reads = argc > 1 ? atoi(argv[1]) : 300;
    const int L = 150, stride = 160, k = 25;
    rb_ctx* ctx = nullptr;
    if (rb_ctx_create(0, &ctx)) { fprintf(stderr, "ctx: %s\n", rb_last_error(nullptr)); return 2; }
    rb_graph* g = nullptr;
    REQ(rb_graph_create(ctx, (1 << 26) + 5, (1 << 24) + 3, 64, 3, 3, 1, k, 0, 0, &g));   // with the small test slices: 1026 probe regions
    REQ(rb_graph_set_engine(g, RB_ENGINE_SLICED));
    void* packed = nullptr;
    REQ(rb_dev_alloc(ctx, &packed, (int64_t)n_reads * stride / 4 + 64));
    REQ(rb_synth_reads_dev(ctx, 7, 20000, 0, n_reads, L, 5000, stride, (uint64_t*)packed));
    int64_t n = 0;
    // uniform layout (prefix k-merizer), twice: the second pass finds every k-mer present and raises counters
    REQ(rb_graph_add_reads_dev(g, (const uint64_t*)packed, nullptr, nullptr, nullptr, n_reads, L, stride, 0, &n));
    REQ(rb_graph_add_reads_dev(g, (const uint64_t*)packed, nullptr, nullptr, nullptr, n_reads, L, stride, 0, &n));
    void* counts = nullptr;
    REQ(rb_dev_alloc(ctx, &counts, n * 4 + 64));
    REQ(rb_graph_count_reads_dev(g, (const uint64_t*)packed, nullptr, nullptr, nullptr, n_reads, L, stride, (float*)counts, nullptr, nullptr, &n));
    // variable-length layout (rolling walker kernels)
    std::vector<int64_t> off((size_t)n_reads);
    std::vector<int32_t> len((size_t)n_reads);
    for (int i = 0; i < n_reads; ++i) { off[(size_t)i] = (int64_t)i * stride; len[(size_t)i] = L - (i % 7); }
    REQ(rb_graph_add_reads_dev(g, (const uint64_t*)packed, nullptr, off.data(), len.data(), n_reads, 0, 0, 0, &n));
    REQ(rb_graph_count_reads_dev(g, (const uint64_t*)packed, nullptr, off.data(), len.data(), n_reads, 0, 0, (float*)counts, nullptr, nullptr, &n));
    double sum = 0;
    for (int64_t i = 0; i < n; ++i) sum += ((float*)counts)[i];
    printf("race_check ok: %lld k-mers, mean count %.3f, %lld launches\n", (long long)n, sum / (double)n, (long long)rb_ctx_kernel_launches(ctx));
    REQ(rb_graph_destroy(g));
    rb_ctx_destroy(ctx);
    return 0;
}
