// l2_probe.cu -- measures, on the box it runs on, the two ceilings the Ex05 window kernel can hit:
//   (1) L2 (LTS) throughput: every SM streams a buffer that fits in L2 (ld.global.cg 16 B, 4 in flight per thread);
//   (2) the Ex05 mix: one tile write followed by F reads of the same tile, tiles > L2 in total.
// Prints one JSON line.  Development aid behind MEASURED numbers quoted in DESIGN.md; not part of the product path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ p, size_t nvec, int reps, unsigned long long* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t gsz = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * gsz < nvec; i += 4 * gsz) {
            uint4 a = __ldcg(p + i), b = __ldcg(p + i + gsz), c = __ldcg(p + i + 2 * gsz), d = __ldcg(p + i + 3 * gsz);
            acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y;
        }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc.x;
}
__global__ void __launch_bounds__(256) write_kernel(uint4* __restrict__ p, size_t nvec, uint32_t v) {
    const size_t gsz = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gsz) __stcg(p + i, make_uint4(v, v, v, v));
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t small = 48ull << 20, big = 1ull << 30;
    uint4 *a, *b; unsigned long long* sink;
    cudaMalloc(&a, small); cudaMalloc(&b, big); cudaMalloc(&sink, 8);
    cudaMemset(a, 1, small); cudaMemset(b, 1, big);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = prop.multiProcessorCount * 8;
    float best_l2 = 1e30f, best_dram = 1e30f, best_copy = 1e30f;
    const int reps = 20;
    for (int it = 0; it < 5; ++it) {
        read_kernel<<<grid, 256>>>(a, small / 16, 2, sink);            // warm L2
        cudaEventRecord(e0); read_kernel<<<grid, 256>>>(a, small / 16, reps, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best_l2) best_l2 = ms;
        cudaEventRecord(e0); read_kernel<<<grid, 256>>>(b, big / 16, 1, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best_dram) best_dram = ms;
        cudaEventRecord(e0); write_kernel<<<grid, 256>>>(b, big / 16, it); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best_copy) best_copy = ms;
    }
    const double l2_gbs = (double)small * reps / best_l2 / 1e6, dram_gbs = (double)big / best_dram / 1e6, wr_gbs = (double)big / best_copy / 1e6;
    printf("{\"probe\": \"l2\", \"sms\": %d, \"sm_clock_mhz_attr\": %d, \"l2_read_gbs\": %.1f, \"l2_bytes_per_clk_at_attr_clock\": %.0f, "
           "\"dram_read_gbs\": %.1f, \"dram_write_gbs\": %.1f, \"l2_size_mb\": %d}\n",
           prop.multiProcessorCount, clk_khz / 1000, l2_gbs, l2_gbs * 1e9 / (clk_khz * 1e3), dram_gbs, wr_gbs, prop.l2CacheSize >> 20);
    return 0;
}
