// Grouped linear forward/backward (include/vqacore.h: vqa_linear_fwd / vqa_linear_bwd).
// Replaces MyLinear / MyConv1d(k=1) / putils.Linear of the reference and their autograd
// (config/CoR2.py:56-122, config/ODA.py:73-139, putils/__init__.py:16-33).
#include "gemm_simt.cuh"
#include "gemm_tc.h"

namespace vqa {

struct PtrTable {
  const float* p[VQA_MAX_GROUPS];
};
struct MutPtrTable {
  float* p[VQA_MAX_GROUPS];
};
struct I64Table {
  int64_t v[VQA_MAX_GROUPS];
};
struct DropTable {
  uint32_t layer[VQA_MAX_GROUPS];
  uint64_t base[VQA_MAX_GROUPS];
};

// ---- forward: A(m,k) = X[m,k]*mask, B(n,k) = W[n,k], epilogue bias+act -----------------------
struct XDropLoader {
  static constexpr bool KC = true;
  PtrTable X; I64Table ld; DropTable dt; Drop d; int64_t K;
  const float* x; int64_t l;
  __device__ void select(int z) { x = X.p[z]; l = ld.v[z]; d.layer = dt.layer[z]; d.base = dt.base[z]; }
  __device__ float operator()(int64_t m, int64_t k) const {
    const float v = x[m * l + k];
    return d.on ? v * d.mul((uint64_t)(m * K + k)) : v;
  }
};
struct WLoader {            // B(n,k) = W[n*K + k]
  static constexpr bool KC = true;
  PtrTable W; int64_t K; const float* w;
  __device__ void select(int z) { w = W.p[z]; }
  __device__ float operator()(int64_t n, int64_t k) const { return w[n * K + k]; }
};
struct BiasActStore {
  PtrTable b; MutPtrTable Y; I64Table ld; int act;
  const float* bias; float* y; int64_t l;
  __device__ void select(int z) { bias = b.p[z]; y = Y.p[z]; l = ld.v[z]; }
  __device__ void operator()(int64_t m, int64_t n, float acc) const {
    y[m * l + n] = act_apply(act, acc + (bias ? bias[n] : 0.0f));
  }
};

// ---- backward -------------------------------------------------------------------------------
// dz(m,n) = dY[m,n] * act'(Y[m,n])
struct DzT_Loader {         // wgrad A'(m'=n, k'=m): contiguous along m' (=n)
  static constexpr bool KC = false;
  PtrTable dY, Y; I64Table lddy, ldy; int act;
  const float* dy; const float* y; int64_t l1, l2;
  __device__ void select(int z) { dy = dY.p[z]; y = Y.p[z]; l1 = lddy.v[z]; l2 = ldy.v[z]; }
  __device__ float operator()(int64_t n, int64_t m) const {
    const float g = dy[m * l1 + n];
    return act == VQA_ACT_NONE ? g : g * act_grad(act, y[m * l2 + n]);
  }
};
struct XDropT_Loader {      // wgrad B'(n'=k, k'=m) = X~[m,k]; virtual column k == K is all ones (bias grad)
  static constexpr bool KC = false;
  PtrTable X; I64Table ld; DropTable dt; Drop d; int64_t K;
  const float* x; int64_t l;
  __device__ void select(int z) { x = X.p[z]; l = ld.v[z]; d.layer = dt.layer[z]; d.base = dt.base[z]; }
  __device__ float operator()(int64_t k, int64_t m) const {
    if (k == K) return 1.0f;
    const float v = x[m * l + k];
    return d.on ? v * d.mul((uint64_t)(m * K + k)) : v;
  }
};
struct WgradStore {         // out(m'=n, n'=k): dW[n,k] or db[n] when k == K
  MutPtrTable dW, db; int64_t K; int accumulate;
  float* w; float* b;
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
struct Dz_Loader {          // dgrad A(m, k'=n) = dz(m,n): contiguous along k'
  static constexpr bool KC = true;
  PtrTable dY, Y; I64Table lddy, ldy; int act;
  const float* dy; const float* y; int64_t l1, l2;
  __device__ void select(int z) { dy = dY.p[z]; y = Y.p[z]; l1 = lddy.v[z]; l2 = ldy.v[z]; }
  __device__ float operator()(int64_t m, int64_t n) const {
    const float g = dy[m * l1 + n];
    return act == VQA_ACT_NONE ? g : g * act_grad(act, y[m * l2 + n]);
  }
};
struct WT_Loader {          // dgrad B'(n'=k, k'=n) = W[n,k]: contiguous along n'
  static constexpr bool KC = false;
  PtrTable W; int64_t K; const float* w;
  __device__ void select(int z) { w = W.p[z]; }
  __device__ float operator()(int64_t k, int64_t n) const { return w[n * K + k]; }
};
struct DgradStore {
  MutPtrTable dX; I64Table ld; DropTable dt; Drop d; int64_t K; int accumulate;
  const float* pool_alpha; const float* pool_dp; int64_t pool_regions;     // optional pooling-gradient addend
  float* dx; int64_t l;
  __device__ void select(int z) { dx = dX.p[z]; l = ld.v[z]; d.layer = dt.layer[z]; d.base = dt.base[z]; }
  __device__ void operator()(int64_t m, int64_t k, float acc) const {
    if (!dx) return;
    float v = d.on ? acc * d.mul((uint64_t)(m * K + k)) : acc;
    if (pool_alpha) {
      const float* dp = pool_dp + (m / pool_regions) * 4 * K + k;
#pragma unroll
      for (int g = 0; g < 4; ++g) v = fmaf(pool_alpha[m * 4 + g], dp[g * K], v);
    }
    float* dst = dx + m * l + k;
    *dst = accumulate ? *dst + v : v;
  }
};

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_pack_weights(const vqa_pack_segment* segs, int nsegs, void* stream) {
  VQA_REQUIRE(segs != nullptr && nsegs >= 0 && nsegs <= VQA_MAX_PACK_SEGMENTS, "vqa_pack_weights: bad segment list");
  for (int i = 0; i < nsegs; ++i)
    VQA_REQUIRE(segs[i].src && segs[i].dst && segs[i].rows >= 0 && segs[i].rows_pad >= segs[i].rows && segs[i].K > 0,
                "vqa_pack_weights: bad segment %d", i);
  if (nsegs == 0) return VQA_OK;
  return tc_pack_segments(segs, nsegs, (cudaStream_t)stream);
}

extern "C" int vqa_dropout_bits(float p, uint64_t seed, const uint64_t* seed_dev, uint32_t layer, uint64_t n,
                                uint8_t* out, void* stream) {
  VQA_REQUIRE(out != nullptr && p >= 0.0f && p < 1.0f, "vqa_dropout_bits: bad argument");
  VQA_REQUIRE(reinterpret_cast<uintptr_t>(out) % 2 == 0, "vqa_dropout_bits: out must be 2-byte aligned");
  return tc_dropout_bits(p, seed, seed_dev, layer, n, out, (cudaStream_t)stream);
}

extern "C" int vqa_seed_advance(uint64_t* seed_dev, void* stream) {
  VQA_REQUIRE(seed_dev != nullptr, "vqa_seed_advance: null pointer");
  return tc_seed_advance(seed_dev, (cudaStream_t)stream);
}

extern "C" size_t vqa_linear_fwd_workspace_bytes(int math, int groups, int64_t M, int64_t K, int64_t N) {
  return tc_linear_fwd_ws(math, groups, M, K, N);
}
extern "C" size_t vqa_linear_bwd_workspace_bytes(int math, int groups, int64_t M, int64_t K, int64_t N) {
  return tc_linear_bwd_ws(math, groups, M, K, N);
}

extern "C" int vqa_dropout_bits_batch(float p, uint64_t seed, const uint64_t* seed_dev, const vqa_bits_segment* segs,
                                      int nsegs, void* stream) {
  VQA_REQUIRE(segs != nullptr && nsegs >= 0 && nsegs <= VQA_MAX_BITS_SEGMENTS && p >= 0.0f && p < 1.0f,
              "vqa_dropout_bits_batch: bad argument");
  for (int i = 0; i < nsegs; ++i)
    VQA_REQUIRE(segs[i].out != nullptr && reinterpret_cast<uintptr_t>(segs[i].out) % 2 == 0,
                "vqa_dropout_bits_batch: segment %d needs a 2-byte aligned output", i);
  if (nsegs == 0) return VQA_OK;
  return tc_dropout_bits_batch(p, seed, seed_dev, segs, nsegs, (cudaStream_t)stream);
}

extern "C" int vqa_linear_fwd(const vqa_linear_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_linear_fwd: null params");
  VQA_REQUIRE(p->groups >= 1 && p->groups <= VQA_MAX_GROUPS, "vqa_linear_fwd: groups=%d out of range", p->groups);
  VQA_REQUIRE(p->M >= 0 && p->K > 0 && p->N > 0, "vqa_linear_fwd: bad shape M=%lld K=%lld N=%lld", (long long)p->M,
              (long long)p->K, (long long)p->N);
  VQA_REQUIRE(p->p >= 0.0f && p->p < 1.0f, "vqa_linear_fwd: dropout p=%f", p->p);
  for (int g = 0; g < p->groups; ++g)
    VQA_REQUIRE(p->X[g] && p->W[g] && p->Y[g] && p->ldx[g] >= p->K && p->ldy[g] >= p->N,
                "vqa_linear_fwd: group %d has a null pointer or a short leading dimension", g);
  if (p->M == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->math != VQA_MATH_FP32_SIMT) return tc_linear_fwd(p, st);      // tensor-core modes never fall back to the CUDA-core GEMM
  XDropLoader a; WLoader b; BiasActStore e;
  a.K = p->K; b.K = p->K; e.act = p->act;
  a.d = make_drop(p->p, p->seed, 0, 0, 1, p->seed_dev);
  for (int g = 0; g < VQA_MAX_GROUPS; ++g) {
    const int s = g < p->groups ? g : 0;
    a.X.p[g] = p->X[s]; a.ld.v[g] = p->ldx[s]; a.dt.layer[g] = p->layer[s]; a.dt.base[g] = p->drop_index_base[s];
    b.W.p[g] = p->W[s];
    e.b.p[g] = p->b[s]; e.Y.p[g] = p->Y[s]; e.ld.v[g] = p->ldy[s];
  }
  return launch_gemm_simt(p->M, p->N, p->K, p->groups, a, b, e, st, "vqa_linear_fwd");
}

extern "C" int vqa_linear_bwd(const vqa_linear_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_linear_bwd: null params");
  VQA_REQUIRE(p->groups >= 1 && p->groups <= VQA_MAX_GROUPS, "vqa_linear_bwd: groups=%d out of range", p->groups);
  VQA_REQUIRE(p->M >= 0 && p->K > 0 && p->N > 0, "vqa_linear_bwd: bad shape");
  VQA_REQUIRE(p->p >= 0.0f && p->p < 1.0f, "vqa_linear_bwd: dropout p=%f", p->p);
  bool any_w = false, any_x = false;
  for (int g = 0; g < p->groups; ++g) {
    VQA_REQUIRE(p->X[g] && p->W[g] && p->dY[g] && (p->act == VQA_ACT_NONE || p->Y[g]),
                "vqa_linear_bwd: group %d has a null pointer", g);
    any_w |= (p->dW[g] != nullptr) || (p->db[g] != nullptr);
    any_x |= (p->dX[g] != nullptr);
  }
  if (p->pool_alpha)
    VQA_REQUIRE(p->groups == 1 && p->dX[0] && p->pool_dpooled && p->pool_regions >= 1 && p->M % p->pool_regions == 0 &&
                    p->K % 4 == 0,
                "vqa_linear_bwd: the pooling addend needs one group, dX, K %% 4 == 0 and M a multiple of pool_regions");
  if (p->M == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->math != VQA_MATH_FP32_SIMT) return tc_linear_bwd(p, st);      // tensor-core modes never fall back to the CUDA-core GEMM
  if (any_w) {
    DzT_Loader a; XDropT_Loader b; WgradStore e;
    a.act = p->act; b.K = p->K; e.K = p->K; e.accumulate = p->accumulate_w;
    b.d = make_drop(p->p, p->seed, 0, 0, 1, p->seed_dev);
    for (int g = 0; g < VQA_MAX_GROUPS; ++g) {
      const int s = g < p->groups ? g : 0;
      a.dY.p[g] = p->dY[s]; a.Y.p[g] = p->Y[s]; a.lddy.v[g] = p->lddy[s]; a.ldy.v[g] = p->ldy[s];
      b.X.p[g] = p->X[s]; b.ld.v[g] = p->ldx[s]; b.dt.layer[g] = p->layer[s]; b.dt.base[g] = p->drop_index_base[s];
      e.dW.p[g] = p->dW[s]; e.db.p[g] = p->db[s];
    }
    // out[N, K+1] = dZ^T[N, M] . X~[M, K+1]   (column K = ones -> bias gradient)
    VQA_TRY(launch_gemm_simt(p->N, p->K + 1, p->M, p->groups, a, b, e, st, "vqa_linear_bwd.wgrad"));
  }
  if (any_x) {
    Dz_Loader a; WT_Loader b; DgradStore e;
    a.act = p->act; b.K = p->K; e.K = p->K; e.accumulate = p->accumulate_x;
    e.pool_alpha = p->pool_alpha; e.pool_dp = p->pool_dpooled; e.pool_regions = p->pool_regions;
    e.d = make_drop(p->p, p->seed, 0, 0, 1, p->seed_dev);
    for (int g = 0; g < VQA_MAX_GROUPS; ++g) {
      const int s = g < p->groups ? g : 0;
      a.dY.p[g] = p->dY[s]; a.Y.p[g] = p->Y[s]; a.lddy.v[g] = p->lddy[s]; a.ldy.v[g] = p->ldy[s];
      b.W.p[g] = p->W[s];
      e.dX.p[g] = p->dX[s]; e.ld.v[g] = p->lddx[s]; e.dt.layer[g] = p->layer[s]; e.dt.base[g] = p->drop_index_base[s];
    }
    // dX[M, K] = dZ[M, N] . W[N, K]
    VQA_TRY(launch_gemm_simt(p->M, p->K, p->N, p->groups, a, b, e, st, "vqa_linear_bwd.dgrad"));
  }
  return VQA_OK;
}
