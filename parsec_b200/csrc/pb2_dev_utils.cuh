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

}  // namespace pb2
