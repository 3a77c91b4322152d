// pb2_dev_utils.cuh -- small sm_100a device helpers shared by the engine kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb2 {

__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_relaxed_gpu(const int32_t* p) {
    int32_t v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(int32_t* p, int32_t v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smid() {
    uint32_t r;
    asm volatile("mov.u32 %0, %smid;" : "=r"(r));
    return r;
}

// 16-byte streaming accesses that bypass L1: tile payloads are produced by other SMs inside the
// same kernel, so they must be read at L2 (the coherence point), and they are touched once.
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ void  st_stream(uint4* p, const uint4& v) { __stcg(p, v); }

// Source may be cudaHostRegister'ed system memory or a peer GPU: plain coherent load, no L1 allocate.
__device__ __forceinline__ uint4 ld_remote(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm volatile("mov.u32 %0, %lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk mover (cp.async.bulk, 1-D): global -> shared -> global through a two-deep shared-memory ring with
// mbarrier completion.  One elected thread drives the whole pipeline, so a copy costs no payload registers and
// keeps 2 x kBulkChunk bytes in flight per CTA whatever the latency of the source (pinned host memory over PCIe,
// a peer GPU over NVLink, local HBM).  Requires 16-byte aligned addresses and a byte count that is a multiple
// of 16; callers fall back to the SIMT loops of pb2_bodies.cuh otherwise.
// This is the device-side replacement of the cudaMemcpyAsync per flow in parsec_default_gpu_stage_in / _stage_out
// (parsec/mca/device/device_gpu.c:1623-1662, :1673-1724).
// ---------------------------------------------------------------------------------------------
#ifndef PB2_BULK_CHUNK
#define PB2_BULK_CHUNK 4096
#endif
constexpr uint32_t kBulkChunk = PB2_BULK_CHUNK;

struct alignas(128) BulkSmem {
    uint8_t  buf[2][kBulkChunk];
    uint64_t bar[2];
    uint32_t parity[2];     // phase each barrier will complete next (persists across copies; owned by thread 0)
};

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_init(BulkSmem& b) {     // thread 0, once per kernel, followed by a barrier
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr_u32(&b.bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr_u32(&b.bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    b.parity[0] = 0; b.parity[1] = 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_addr_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0()  { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_addr_u32(bar)), "r"(parity) : "memory");
    }
}

// Whole CTA calls (uniform arguments); thread 0 moves `bytes` (multiple of 16, both pointers 16-byte aligned).
// On return the bytes are written and ordered before a following __threadfence()/release by any thread of the CTA.
__device__ __forceinline__ void cta_bulk_copy(void* dst, const void* src, size_t bytes, BulkSmem& b) {
    if (threadIdx.x == 0 && bytes) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        const size_t n = (bytes + kBulkChunk - 1) / kBulkChunk;
        uint32_t par0 = b.parity[0], par1 = b.parity[1];
        bulk_g2s(b.buf[0], s, (uint32_t)(bytes < kBulkChunk ? bytes : kBulkChunk), &b.bar[0]);
        for (size_t i = 0; i < n; ++i) {
            const int slot = (int)(i & 1);
            if (i + 1 < n) {
                if (i >= 1) bulk_wait_read0();          // the store that last read the other buffer has drained it
                const size_t off = (i + 1) * kBulkChunk;
                bulk_g2s(b.buf[slot ^ 1], s + off, (uint32_t)(bytes - off < kBulkChunk ? bytes - off : kBulkChunk), &b.bar[slot ^ 1]);
            }
            if (slot == 0) { bulk_bar_wait(&b.bar[0], par0); par0 ^= 1; }
            else           { bulk_bar_wait(&b.bar[1], par1); par1 ^= 1; }
            const size_t off = i * kBulkChunk;
            bulk_s2g(d + off, b.buf[slot], (uint32_t)(bytes - off < kBulkChunk ? bytes - off : kBulkChunk));
        }
        bulk_wait_all0();
        asm volatile("fence.proxy.async;" ::: "memory");
        b.parity[0] = par0; b.parity[1] = par1;
    }
    __syncthreads();
}

}  // namespace pb2
