// pb2_gemm.cuh -- tensor-core (tcgen05) engine kernel for PB2_BODY_GEMM_BF16 windows.
#pragma once
#include "pb2_sched.cuh"

namespace pb2 {

static inline int pb2_gemm_nworkers(int sm_count) { return sm_count; }
static inline int pb2_gemm_launch(const WinDev&, int, cudaStream_t) { return PB2_ERR_NOT_IMPLEMENTED; }

}  // namespace pb2
