// Mutan bilinear fusion forward/backward (include/vqacore.h: vqa_mutan_fwd / vqa_mutan_bwd).
// Replaces MutanFusion.forward and the per-sample bmul loop (putils/__init__.py:205-241, :98-104).
#include "gemm_simt.cuh"
#include "gemm_tc.h"

namespace vqa {

struct RPtr {
  const float* p[VQA_MAX_GROUPS];
};
struct RMut {
  float* p[VQA_MAX_GROUPS];
};

struct PlainLoader {        // A(m,k) = X[m*ld + k]
  static constexpr bool KC = true;
  const float* x; int64_t ld;
  __device__ void select(int) {}
  __device__ float operator()(int64_t m, int64_t k) const { return x[m * ld + k]; }
};
struct WrLoader {           // B(n,k) = W_r[n*K + k], r fixed by set_r (forward) or blockIdx.z
  static constexpr bool KC = true;
  RPtr W; int64_t K; const float* w;
  __device__ void select(int z) { w = W.p[z]; }
  __device__ float operator()(int64_t n, int64_t k) const { return w[n * K + k]; }
};

// Forward: one CTA tile of Y, the rank loop inside, Hadamard-sum epilogue in registers.
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(SIMT_THREADS)
mutan_fwd_kernel(int R, int64_t M, int64_t F, int64_t K1, int64_t rows_per, PlainLoader a, WrLoader b, RPtr b1,
                 const float* __restrict__ H2, float* __restrict__ H1, float* __restrict__ Y, int64_t ldy) {
  __shared__ __align__(16) float As[SIMT_BK * (BM + SIMT_PAD)];
  __shared__ __align__(16) float Bs[SIMT_BK * (BN + SIMT_PAD)];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int ty = threadIdx.x / (BN / TN), tx = threadIdx.x % (BN / TN);
  const int64_t Mh = M / rows_per;
  float out[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) out[i][j] = 0.0f;
  for (int r = 0; r < R; ++r) {
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    b.select(r);
    simt_mainloop<BM, BN, TM, TN>(acc, M, F, K1, m0, n0, a, b, As, Bs);
    const float* bias = b1.p[r];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t m = m0 + ty * TM + i;
      if (m >= M) continue;
      const float* h2 = H2 + ((int64_t)r * Mh + m / rows_per) * F;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int64_t n = n0 + tx * TN + j;
        if (n >= F) continue;
        const float h1 = acc[i][j] + (bias ? bias[n] : 0.0f);
        if (H1) H1[((int64_t)r * M + m) * F + n] = h1;
        out[i][j] = fmaf(h1, h2[n], out[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n < F) Y[m * ldy + n] = out[i][j];
    }
  }
}

// H2_r = X2 . W2_r^T + b2_r   (grouped over r via blockIdx.z)
struct H2Store {
  RPtr b2; float* H2; int64_t Mh, F; const float* bias; float* dst;
  __device__ void select(int z) { bias = b2.p[z]; dst = H2 + (int64_t)z * Mh * F; }
  __device__ void operator()(int64_t m, int64_t n, float acc) const { dst[m * F + n] = acc + (bias ? bias[n] : 0.0f); }
};

// ---- backward -------------------------------------------------------------------------------
// dH2[r,mh,f] = sum_{j<rows_per} dY[mh*rp+j, f] * H1[r, mh*rp+j, f]
__global__ void mutan_dh2_kernel(int64_t M, int64_t F, int64_t rows_per, const float* __restrict__ dY, int64_t lddy,
                                 const float* __restrict__ H1, float* __restrict__ dH2) {
  const int64_t mh = blockIdx.x;
  const int r = blockIdx.y;
  const int64_t Mh = M / rows_per;
  for (int64_t f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.0f;
    for (int64_t j = 0; j < rows_per; ++j) {
      const int64_t m = mh * rows_per + j;
      s = fmaf(dY[m * lddy + f], H1[((int64_t)r * M + m) * F + f], s);
    }
    dH2[((int64_t)r * Mh + mh) * F + f] = s;
  }
}

// dH1_r(m,f) = dY[m,f] * H2[r, m/rp, f]
struct DH1T_Loader {        // wgrad A'(m'=f, k'=m), group = r
  static constexpr bool KC = false;
  const float* dY; int64_t lddy; const float* H2; int64_t Mh, F, rows_per; const float* h2r;
  __device__ void select(int z) { h2r = H2 + (int64_t)z * Mh * F; }
  __device__ float operator()(int64_t f, int64_t m) const { return dY[m * lddy + f] * h2r[(m / rows_per) * F + f]; }
};
struct XT1_Loader {         // wgrad B'(n'=k, k'=m) = X[m,k]; column k == K -> 1 (bias grad)
  static constexpr bool KC = false;
  const float* x; int64_t ld, K;
  __device__ void select(int) {}
  __device__ float operator()(int64_t k, int64_t m) const { return k == K ? 1.0f : x[m * ld + k]; }
};
struct WgradStoreR {
  RMut dW, db; int64_t K; int accumulate; float* w; float* b;
  __device__ void select(int z) { w = dW.p[z]; b = db.p[z]; }
  __device__ void operator()(int64_t n, int64_t k, float acc) const {
    if (k == K) {
      if (b) b[n] = accumulate ? b[n] + acc : acc;
    } else if (w) {
      float* dst = w + n * K + k;
      *dst = accumulate ? *dst + acc : acc;
    }
  }
};
struct DH1_Loader {         // dgrad A(m, k'=r*F+f) = dY[m,f]*H2[r,m/rp,f]
  static constexpr bool KC = true;
  const float* dY; int64_t lddy; const float* H2; int64_t Mh, F, rows_per;
  __device__ void select(int) {}
  __device__ float operator()(int64_t m, int64_t kk) const {
    const int64_t r = kk / F, f = kk - r * F;
    return dY[m * lddy + f] * H2[(r * Mh + m / rows_per) * F + f];
  }
};
struct WrT_Loader {         // dgrad B'(n'=k, k'=r*F+f) = W_r[f,k]
  static constexpr bool KC = false;
  RPtr W; int64_t K, F;
  __device__ void select(int) {}
  __device__ float operator()(int64_t k, int64_t kk) const {
    const int64_t r = kk / F, f = kk - r * F;
    return W.p[r][f * K + k];
  }
};
struct PlainStore {
  float* dst; int64_t ld; int accumulate;
  __device__ void select(int) {}
  __device__ void operator()(int64_t m, int64_t n, float acc) const {
    float* d = dst + m * ld + n;
    *d = accumulate ? *d + acc : acc;
  }
};
struct DH2T_Loader {        // wgrad A'(m'=f, k'=mh) = dH2[r,mh,f]
  static constexpr bool KC = false;
  const float* dH2; int64_t Mh, F; const float* d;
  __device__ void select(int z) { d = dH2 + (int64_t)z * Mh * F; }
  __device__ float operator()(int64_t f, int64_t mh) const { return d[mh * F + f]; }
};
struct DH2_Loader {         // dgrad A(mh, k'=r*F+f) = dH2[r,mh,f]
  static constexpr bool KC = true;
  const float* dH2; int64_t Mh, F;
  __device__ void select(int) {}
  __device__ float operator()(int64_t mh, int64_t kk) const {
    const int64_t r = kk / F, f = kk - r * F;
    return dH2[(r * Mh + mh) * F + f];
  }
};

}  // namespace vqa

using namespace vqa;

extern "C" size_t vqa_mutan_workspace_bytes(int math, int R, int64_t M, int64_t rows_per_h2, int64_t K1, int64_t K2,
                                            int64_t F, int bwd) {
  return tc_mutan_ws(math, R, M, rows_per_h2 > 0 ? rows_per_h2 : 1, K1, K2, F, bwd);
}

extern "C" int vqa_mutan_fwd(const vqa_mutan_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_mutan_fwd: null params");
  VQA_REQUIRE(p->R >= 1 && p->R <= VQA_MAX_GROUPS, "vqa_mutan_fwd: R=%d out of range", p->R);
  VQA_REQUIRE(p->M >= 0 && p->K1 > 0 && p->K2 > 0 && p->F > 0 && p->rows_per_h2 >= 1, "vqa_mutan_fwd: bad shape");
  VQA_REQUIRE(p->M % p->rows_per_h2 == 0, "vqa_mutan_fwd: M=%lld not a multiple of rows_per_h2=%lld",
              (long long)p->M, (long long)p->rows_per_h2);
  VQA_REQUIRE(p->X1 && p->X2 && p->H2 && p->Y, "vqa_mutan_fwd: null pointer");
  for (int r = 0; r < p->R; ++r) VQA_REQUIRE(p->W1[r] && p->W2[r], "vqa_mutan_fwd: null weight for rank %d", r);
  if (p->M == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->math != VQA_MATH_FP32_SIMT) return tc_mutan_fwd(p, st);      // tensor-core modes never fall back to the CUDA-core GEMM
  const int64_t Mh = p->M / p->rows_per_h2;
  {
    PlainLoader a{p->X2, p->ldx2};
    WrLoader b; b.K = p->K2;
    H2Store e; e.H2 = p->H2; e.Mh = Mh; e.F = p->F;
    for (int r = 0; r < VQA_MAX_GROUPS; ++r) {
      const int s = r < p->R ? r : 0;
      b.W.p[r] = p->W2[s]; e.b2.p[r] = p->b2[s];
    }
    VQA_TRY(launch_gemm_simt(Mh, p->F, p->K2, p->R, a, b, e, st, "vqa_mutan_fwd.h2"));
  }
  PlainLoader a{p->X1, p->ldx1};
  WrLoader b; b.K = p->K1;
  RPtr b1;
  for (int r = 0; r < VQA_MAX_GROUPS; ++r) {
    const int s = r < p->R ? r : 0;
    b.W.p[r] = p->W1[s]; b1.p[r] = p->b1[s];
  }
  if (p->M >= 2048) {
    dim3 grid((unsigned)cdiv(p->F, 64), (unsigned)cdiv(p->M, 128));
    mutan_fwd_kernel<128, 64, 8, 4><<<grid, SIMT_THREADS, 0, st>>>(p->R, p->M, p->F, p->K1, p->rows_per_h2, a, b, b1,
                                                                    p->H2, p->H1, p->Y, p->ldy);
  } else {
    dim3 grid((unsigned)cdiv(p->F, 32), (unsigned)cdiv(p->M, 32));
    mutan_fwd_kernel<32, 32, 4, 1><<<grid, SIMT_THREADS, 0, st>>>(p->R, p->M, p->F, p->K1, p->rows_per_h2, a, b, b1,
                                                                   p->H2, p->H1, p->Y, p->ldy);
  }
  return check_launch("vqa_mutan_fwd");
}

extern "C" int vqa_mutan_bwd(const vqa_mutan_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_mutan_bwd: null params");
  VQA_REQUIRE(p->R >= 1 && p->R <= VQA_MAX_GROUPS, "vqa_mutan_bwd: R=%d out of range", p->R);
  VQA_REQUIRE(p->M >= 0 && p->K1 > 0 && p->K2 > 0 && p->F > 0 && p->rows_per_h2 >= 1, "vqa_mutan_bwd: bad shape");
  VQA_REQUIRE(p->M % p->rows_per_h2 == 0, "vqa_mutan_bwd: M not a multiple of rows_per_h2");
  VQA_REQUIRE(p->X1 && p->X2 && p->H1 && p->H2 && p->dY && p->dH2, "vqa_mutan_bwd: null pointer");
  for (int r = 0; r < p->R; ++r) VQA_REQUIRE(p->W1[r] && p->W2[r], "vqa_mutan_bwd: null weight for rank %d", r);
  if (p->M == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->math != VQA_MATH_FP32_SIMT) return tc_mutan_bwd(p, st);      // tensor-core modes never fall back to the CUDA-core GEMM
  const int64_t Mh = p->M / p->rows_per_h2;
  const int64_t RF = (int64_t)p->R * p->F;

  mutan_dh2_kernel<<<dim3((unsigned)Mh, (unsigned)p->R), 256, 0, st>>>(p->M, p->F, p->rows_per_h2, p->dY, p->lddy,
                                                                        p->H1, p->dH2);
  VQA_TRY(check_launch("vqa_mutan_bwd.dh2"));

  {  // dW1_r[f,k] (+ db1_r) = sum_m dH1_r[m,f] * X1[m,k]
    DH1T_Loader a{p->dY, p->lddy, p->H2, Mh, p->F, p->rows_per_h2, nullptr};
    XT1_Loader b{p->X1, p->ldx1, p->K1};
    WgradStoreR e; e.K = p->K1; e.accumulate = p->accumulate_w;
    for (int r = 0; r < VQA_MAX_GROUPS; ++r) {
      const int s = r < p->R ? r : 0;
      e.dW.p[r] = p->dW1[s]; e.db.p[r] = p->db1[s];
    }
    VQA_TRY(launch_gemm_simt(p->F, p->K1 + 1, p->M, p->R, a, b, e, st, "vqa_mutan_bwd.dw1"));
  }
  if (p->dX1) {  // dX1[m,k] = sum_{r,f} dH1_r[m,f] * W1_r[f,k]
    DH1_Loader a{p->dY, p->lddy, p->H2, Mh, p->F, p->rows_per_h2};
    WrT_Loader b; b.K = p->K1; b.F = p->F;
    for (int r = 0; r < VQA_MAX_GROUPS; ++r) b.W.p[r] = p->W1[r < p->R ? r : 0];
    PlainStore e{p->dX1, p->lddx1, p->accumulate_x1};
    VQA_TRY(launch_gemm_simt(p->M, p->K1, RF, 1, a, b, e, st, "vqa_mutan_bwd.dx1"));
  }
  {  // dW2_r[f,k] (+ db2_r) = sum_mh dH2[r,mh,f] * X2[mh,k]
    DH2T_Loader a{p->dH2, Mh, p->F, nullptr};
    XT1_Loader b{p->X2, p->ldx2, p->K2};
    WgradStoreR e; e.K = p->K2; e.accumulate = p->accumulate_w;
    for (int r = 0; r < VQA_MAX_GROUPS; ++r) {
      const int s = r < p->R ? r : 0;
      e.dW.p[r] = p->dW2[s]; e.db.p[r] = p->db2[s];
    }
    VQA_TRY(launch_gemm_simt(p->F, p->K2 + 1, Mh, p->R, a, b, e, st, "vqa_mutan_bwd.dw2"));
  }
  if (p->dX2) {  // dX2[mh,k] = sum_{r,f} dH2[r,mh,f] * W2_r[f,k]
    DH2_Loader a{p->dH2, Mh, p->F};
    WrT_Loader b; b.K = p->K2; b.F = p->F;
    for (int r = 0; r < VQA_MAX_GROUPS; ++r) b.W.p[r] = p->W2[r < p->R ? r : 0];
    PlainStore e{p->dX2, p->lddx2, p->accumulate_x2};
    VQA_TRY(launch_gemm_simt(Mh, p->K2, RF, 1, a, b, e, st, "vqa_mutan_bwd.dx2"));
  }
  return VQA_OK;
}
