// Tensor-core (tcgen05) GEMM entry points used by the op files; implemented in gemm_tc.cu.
#pragma once
#include "common.cuh"

namespace vqa {

// A tensor-core math mode that cannot run as asked (unbuilt mode, workspace too small, operand TMA cannot address
// and that cannot be packed) is an error (VQA_EINVAL with a message) — never a silent CUDA-core fallback.

int tc_linear_fwd(const vqa_linear_fwd_params* p, cudaStream_t st);
int tc_linear_bwd(const vqa_linear_bwd_params* p, cudaStream_t st);
int tc_mutan_fwd(const vqa_mutan_fwd_params* p, cudaStream_t st);
int tc_mutan_bwd(const vqa_mutan_bwd_params* p, cudaStream_t st);
size_t tc_mutan_ws(int math, int R, int64_t M, int64_t rows_per, int64_t K1, int64_t K2, int64_t F, int bwd);
int tc_pack_segments(const vqa_pack_segment* segs, int nsegs, cudaStream_t st);
int tc_dropout_bits(float pdrop, uint64_t seed, const uint64_t* seed_dev, uint32_t layer, uint64_t n, uint8_t* out,
                    cudaStream_t st);
int tc_seed_advance(uint64_t* seed_dev, cudaStream_t st);
int tc_dropout_bits_batch(float pdrop, uint64_t seed, const uint64_t* seed_dev, const vqa_bits_segment* segs, int nsegs,
                          cudaStream_t st);
size_t tc_linear_fwd_ws(int math, int groups, int64_t M, int64_t K, int64_t N);
size_t tc_linear_bwd_ws(int math, int groups, int64_t M, int64_t K, int64_t N);

}  // namespace vqa
