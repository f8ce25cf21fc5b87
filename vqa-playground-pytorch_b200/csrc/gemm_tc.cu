// tcgen05 tensor-core GEMM family (placeholder until the TMA/TMEM kernels land in this file).
#include "gemm_tc.h"

namespace vqa {
int tc_linear_fwd(const vqa_linear_fwd_params*, cudaStream_t) { return VQA_TC_UNSUPPORTED; }
int tc_linear_bwd(const vqa_linear_bwd_params*, cudaStream_t) { return VQA_TC_UNSUPPORTED; }
}  // namespace vqa
