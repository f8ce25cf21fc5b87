// Tensor-core (tcgen05) GEMM entry points used by the op files; implemented in gemm_tc.cu.
#pragma once
#include "common.cuh"

namespace vqa {

// Returned when the tensor-core path does not cover a shape/flag combination; the caller then
// uses the fp32 CUDA-core kernel (still on the GPU, strictly more precise — never a CPU path).
constexpr int VQA_TC_UNSUPPORTED = 999;

int tc_linear_fwd(const vqa_linear_fwd_params* p, cudaStream_t st);
int tc_linear_bwd(const vqa_linear_bwd_params* p, cudaStream_t st);

}  // namespace vqa
