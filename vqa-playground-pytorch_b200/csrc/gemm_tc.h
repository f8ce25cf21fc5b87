// Tensor-core (tcgen05) GEMM entry points used by the op files; implemented in gemm_tc.cu.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace vqa {

// A tensor-core math mode that cannot run as asked (unbuilt mode, workspace too small, operand TMA cannot address
// and that cannot be packed) is an error (VQA_EINVAL with a message) — never a silent CUDA-core fallback.

struct LinExt;
struct MutanExt;
int tc_linear_fwd(const vqa_linear_fwd_params* p, cudaStream_t st, const LinExt* ext = nullptr);
int tc_linear_bwd(const vqa_linear_bwd_params* p, cudaStream_t st, const LinExt* ext = nullptr);
int tc_mutan_fwd(const vqa_mutan_fwd_params* p, cudaStream_t st, const MutanExt* ext = nullptr);
int tc_mutan_bwd(const vqa_mutan_bwd_params* p, cudaStream_t st, const MutanExt* ext = nullptr);
size_t tc_mutan_ws(int math, int R, int64_t M, int64_t rows_per, int64_t K1, int64_t K2, int64_t F, int bwd);
int tc_pack_segments(const vqa_pack_segment* segs, int nsegs, cudaStream_t st);
int tc_dropout_bits(float pdrop, uint64_t seed, const uint64_t* seed_dev, uint32_t layer, uint64_t n, uint8_t* out,
                    cudaStream_t st);
int tc_seed_advance(uint64_t* seed_dev, cudaStream_t st);
int tc_dropout_bits_batch(float pdrop, uint64_t seed, const uint64_t* seed_dev, const vqa_bits_segment* segs, int nsegs,
                          cudaStream_t st);
size_t tc_linear_fwd_ws(int math, int groups, int64_t M, int64_t K, int64_t N);

// ---- bf16 operand planes (tc_gemm16.cuh / gemm_tc16.cu) -----------------------------------------------------------
// The bf16 math modes run every LARGE-M GEMM (M >= TC16_MIN_M: the region-side contractions, M = B*N) on the
// bf16-plane kernel and the small-M ones (M = B) on the fp32-operand kernel (3xTF32 for BF16X3, one TF32 pass for BF16).
constexpr int64_t TC16_MIN_M = 1024;
inline bool is_bf16_math(int math) { return math == VQA_MATH_BF16X3 || math == VQA_MATH_BF16; }
inline int small_math_of(int math) { return math == VQA_MATH_BF16X3 ? VQA_MATH_TF32X3 : (math == VQA_MATH_BF16 ? VQA_MATH_TF32 : math); }

struct Planes {                 // bf16 [np][rows][ld], `plane` elements between planes
  const __nv_bfloat16* p = nullptr;
  int64_t ld = 0, plane = 0;
};
// Optional plane hand-offs between the ops of a whole-model plan (single-group launches): operands made earlier in
// the step are not split again, and a forward can emit the planes of its result for the next GEMM.
struct LinExt {
  Planes Xp;                    // planes of dropout(X)  [M][Kp]
  Planes Wp;                    // planes of W           [N_pad][Kp]   (forward)
  Planes WTp;                   // planes of W^T         [K][N_pad]    (dgrad)
  __nv_bfloat16* Yp = nullptr;  // forward: also write the planes of Y [M][ldyp]
  int64_t ldyp = 0, yplane = 0;
  bool raw_dx = false;          // backward: dX = dZ.W as is — the caller's consumer applies the input-dropout mask (and
                                // any pooling addend) while it reads dX (vqa_cor_compound_bwd's dv2_* fields)
};
struct MutanExt {
  int h2_mode = 0;              // forward: 0 = H2 then the X1-side GEMMs; 1 = H2 = X2.W2^T + b2 only (it depends on the
                                // question side alone: a plan runs it early on its side lane); 2 = H2 is already there
  Planes X1p;                   // [M][K1p]
  Planes W1p;                   // stacked [R*Fp][K1p]
  Planes W1Tp;                  // [K1][R*Fp]
};
struct PackPlanesSeg {          // one weight of pack_planes(): K-major planes and / or its slice of transposed planes
  const float* src; int64_t rows, rows_pad, K, Kp;
  __nv_bfloat16* dst; int64_t plane;                       // [np][rows_pad][Kp] or nullptr
  __nv_bfloat16* dstT; int64_t Np, planeT, t_col0;         // [np][K][Np], this weight at columns t_col0.. ; or nullptr
};
struct Mutan16Ops { int np; Planes x1p, w1p, w1tp; __nv_bfloat16* dh1; };

namespace tc16 {
int split_planes(const float* const* X, const int64_t* ldx, int groups, int64_t M, int64_t K, float pdrop, uint64_t seed,
                 const uint64_t* seed_dev, const uint32_t* layer, const uint64_t* base, const uint8_t* const* bits,
                 __nv_bfloat16* const* out, int64_t ldp, int64_t plane, int np, cudaStream_t st);
int pack_planes(const PackPlanesSeg* segs, int nsegs, int np, cudaStream_t st);
}  // namespace tc16
int tc16_linear_fwd(const vqa_linear_fwd_params* p, cudaStream_t st, const LinExt* ext);
int tc16_linear_bwd(const vqa_linear_bwd_params* p, cudaStream_t st, const LinExt* ext);
size_t tc16_linear_fwd_ws(int np, int groups, int64_t M, int64_t K, int64_t N);
size_t tc16_linear_bwd_ws(int np, int groups, int64_t M, int64_t K, int64_t N);
size_t tc16_mutan_ws(int np, int R, int64_t M, int64_t K1, int64_t F, int bwd);
int tc16_mutan_fwd_h1(const vqa_mutan_fwd_params* p, cudaStream_t st, const MutanExt* ext);
int tc16_mutan_bwd_prepare(const vqa_mutan_bwd_params* p, cudaStream_t st, const MutanExt* ext, Mutan16Ops* ops);
int tc16_mutan_bwd_big(const vqa_mutan_bwd_params* p, cudaStream_t st, const Mutan16Ops* ops);
size_t tc_linear_bwd_ws(int math, int groups, int64_t M, int64_t K, int64_t N);

}  // namespace vqa
