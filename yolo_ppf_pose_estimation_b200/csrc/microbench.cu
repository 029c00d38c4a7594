// microbench.cu — roofline denominators that MEASURED_PEAKS.json does not carry: the rate of
// shared-memory reductions (red.shared.add.u32 -> ATOMS.POPC.INC), which is what one PPF vote costs.
//   pattern 0  conflict-free: the 32 lanes of a warp hit 32 different banks
//   pattern 1  random words in a 64 KB accumulator (the voting kernel's pattern)
//   pattern 2  all lanes on one word
// Every CTA owns a 64 KB shared array, 512 threads, 2 CTAs per SM as in the voting kernel.
#include "ppf_common.cuh"

namespace b200ppf {

namespace {

constexpr int MB_THREADS = 512;
constexpr int MB_WORDS = 16 * 1024;

__global__ void __launch_bounds__(MB_THREADS, 2)
atoms_rate_kernel(int pattern, uint32_t iters, unsigned long long *sink) {
    extern __shared__ uint32_t acc[];
    for (uint32_t k = threadIdx.x; k < MB_WORDS; k += MB_THREADS) acc[k] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(acc);
    uint32_t state = (blockIdx.x * MB_THREADS + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t word;
            if (pattern == 0) {
                word = ((it * 4 + u) * 37u + warp * 101u) % (MB_WORDS / 32) * 32 + lane;
            } else if (pattern == 1) {
                state = state * 1664525u + 1013904223u;
                word = (state >> 10) % MB_WORDS;
            } else {
                word = warp;
            }
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + word * 4) : "memory");
        }
    }
    __syncthreads();
    unsigned long long s = 0;
    for (uint32_t k = threadIdx.x; k < MB_WORDS; k += MB_THREADS) s += acc[k];
    if (s == 0xFFFFFFFFFFFFFFFFull) *sink = s;  // keep the work observable
}

}  // namespace

int microbench_atoms(b200ppf_ctx *ctx, int pattern, double *atoms_per_sec) {
    *atoms_per_sec = 0.0;
    const size_t smem = MB_WORDS * sizeof(uint32_t);
    PPF_CUDA(ctx, cudaFuncSetAttribute(atoms_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long *sink = nullptr;
    PPF_CUDA(ctx, cudaMallocAsync(&sink, sizeof(unsigned long long), ctx->stream));
    const uint32_t iters = 4096;
    const unsigned grid = (unsigned)ctx->sm_count * 2 * 4;
    PPF_LAUNCH(ctx, atoms_rate_kernel, grid, MB_THREADS, smem, pattern, 64u, sink);  // warm-up
    cudaEventRecord(ctx->ev[0], ctx->stream);
    PPF_LAUNCH(ctx, atoms_rate_kernel, grid, MB_THREADS, smem, pattern, iters, sink);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    cudaFreeAsync(sink, ctx->stream);
    const double total = (double)grid * MB_THREADS * (double)iters * 4.0;
    *atoms_per_sec = total / (ms * 1e-3);
    return B200PPF_OK;
}

}  // namespace b200ppf
