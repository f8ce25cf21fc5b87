// Whole-model plans: every kernel of Model.forward / backward enqueued from one C call
// (include/vqacore.h: vqa_cor2_{fwd,bwd}, vqa_oda_{fwd,bwd}).
// Reference: config/CoR2.py:201-237 (+ decare_cat :191-199), config/ODA.py:200-240.
// Parameter tables follow the reference's state_dict order (seq2vec.* excluded).
#include <string.h>

#include "gemm_tc.h"

namespace vqa {

constexpr int64_t H = 310, F = 510, A = 620, AG = 155, Q = 2400, D = 2048;
// padded row strides of intermediates that are TMA operands of the tensor-core GEMMs (16-byte multiples)
constexpr int64_t HP = 320, XP = 512;
constexpr int64_t FPAD = 512;   // roundup(F, 32): row padding of stacked Mutan weights
constexpr float P_DROP = 0.5f;

// ---- state_dict indices -------------------------------------------------------------------------
namespace cor2 {
enum {
  COMPRESS_V = 0, COMPRESS_V2 = 2, COMPRESS_Q = 4,
  VQ1_L1 = 6, VQ1_L2 = 10, ATT1_CONV = 14, ATT1_G = 16,
  VQ2_L1 = 24, VQ2_L2 = 28, ATT2_CONV = 32, ATT2_G = 34,
  LINEAR_Q = 42, FF_L1 = 44, FF_L2 = 48, CLASSIF = 52,
  CQ1 = 54, EQ1 = 56, CQ2 = 58, EQ2 = 60
};
// dropout call sites in forward order (oracle/reasoning_core.py COR2_LAYERS)
enum {
  L_COMPRESS_Q = 0, L_COMPRESS_V = 1, L_ATT1_CONV = 2, L_ATT1_G = 3, L_CQ1 = 7, L_EQ1 = 8, L_CQ2 = 9, L_EQ2 = 10,
  L_COMPRESS_V2 = 11, L_ATT2_CONV = 12, L_ATT2_G = 13, L_LINEAR_Q = 17, L_CLASSIF = 18
};
}  // namespace cor2
namespace oda {
enum { COMPRESS_V = 0, COMPRESS_Q = 2, ATT_CONV = 4, ATT_G = 6, LINEAR_Q = 14, FF_L1 = 16, FF_L2 = 26, CLASSIF = 36 };
enum { L_COMPRESS_V = 0, L_COMPRESS_Q = 1, L_ATT_CONV = 2, L_ATT_G = 3, L_LINEAR_Q = 7, L_CLASSIF = 8 };
}  // namespace oda

// ---- workspace carving ----------------------------------------------------------------------------
struct Carver {
  char* base; size_t off;
  float* take(int64_t n) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += ((size_t)n * sizeof(float) + 255) & ~(size_t)255;
    return p;
  }
};

// ---- internal fork/join lanes -----------------------------------------------------------------------
// A plan may run branches that do not depend on each other (the question-side projections vs the region-side
// GEMMs; in the backward the att1 glimpse linears, the gates and the question projections vs the critical dgrad
// chain) on a library-owned side stream.  Every branch forks from and joins back into the CALLER's stream with
// events before the plan returns, so the external contract (all work ordered on the caller's stream, capturable
// in a CUDA graph — the side stream simply becomes a parallel branch of the graph) is unchanged.
struct Lanes {
  int dev = -1;
  cudaStream_t side = nullptr;
  cudaEvent_t ev[32];
  int next = 0;
  cudaEvent_t record(cudaStream_t s) {
    cudaEvent_t e = ev[next++ & 31];
    cudaEventRecord(e, s);
    return e;
  }
  static void wait(cudaStream_t s, cudaEvent_t e) { cudaStreamWaitEvent(s, e, 0); }
};
static Lanes* get_lanes() {
  static thread_local Lanes L;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (L.dev != dev) {
    if (cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < 32; ++i)
      if (cudaEventCreateWithFlags(&L.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    L.dev = dev;
  }
  return &L;
}

struct Cor2Ws {
  float *ql, *hq1, *hq2, *qf, *g1, *g2, *vl, *f1_H1, *f1_H2, *fuse1, *pooled1, *vf, *v2l, *f2_H1, *f2_H2, *fuse2,
      *pooled2, *ff_H1, *ff_H2, *xf;
  float *dxf, *dvf, *dqf, *d_ff_H2, *dpooled2, *dalpha2, *dz2, *dfuse2, *dv2, *dv2l, *d_f2_H2, *dql, *dg1, *dg2, *dhq1,
      *dhq2, *dalpha_ext, *dpooled1, *dalpha1, *dz1, *dfuse1, *dvl, *d_f1_H2;
  float* lin_ws; size_t lin_ws_bytes;
  float* side_ws; size_t side_ws_bytes;   // scratch of the ops that run on the side lane
  float* dq_parts;                        // [4][B][Q]: per-projection shares of the question-embedding gradient
  uint8_t* bits[32];           // packed dropout keep-bits of every dropout site, by layer id (train mode)
  int64_t bits_n[32];          // element count of each site
  float *vq1_w1p, *vq1_w2p, *vq2_w1p, *vq2_w2p, *ff_w1p, *ff_w2p, *eq1p, *eq2p, *clsp;   // vqa_pack_weights copies
  // bf16 math modes: operand planes [np][rows][ld] handed from producer to consumer GEMMs (gemm_tc.h: Planes)
  __nv_bfloat16 *vp, *v2p, *vlp, *v2lp;                // dropout(v), dropout(v2) [M][2048]; vl, v2l [M][320]
  __nv_bfloat16 *cv_wp, *cv2_wp, *cv2_wtp;             // compress_v / compress_v2 weights [320][2048]; transposed [2048][320]
  __nv_bfloat16 *vq1_wp, *vq1_wtp, *vq2_wp, *vq2_wtp;  // stacked Mutan W1 [2*512][320]; transposed [310][1024]
  size_t bytes;
};

// scratch shared by every linear call of a plan (dZ of the tensor-core backward + padded weight copies)
static int64_t lin_scratch_floats(int64_t B, int64_t N, int64_t C) {
  const int64_t M = B * N;
  const int64_t Cp = ((C + 31) / 32) * 32;
  // linear bwd: dZ + padded weights; Mutan bwd: stacked padded weights + dH1cat [M, R*512] + dH2cat [B, R*512]
  // (bf16 modes: dZ planes in both orientations, 2 x [np = 2][M][320] bf16 = M * 640 floats)
  const int64_t lin = M * 640 + 4 * B * 2048 + B * Cp + 2 * 2048 * 320 + C * 512;
  const int64_t mut = 2 * 5 * 512 * 1240 + M * 1024 + B * 5 * 512 + 65536;
  return (lin > mut ? lin : mut) + 4096;
}

static Cor2Ws carve_cor2(void* base, int64_t B, int64_t N, int64_t C) {
  Carver c{reinterpret_cast<char*>(base), 0};
  const int64_t M = B * N;
  Cor2Ws w;
  w.ql = c.take(B * HP); w.hq1 = c.take(B * HP); w.hq2 = c.take(B * HP); w.qf = c.take(B * HP);
  w.g1 = c.take(B * D); w.g2 = c.take(B * D);
  w.vl = c.take(M * HP);
  w.f1_H1 = c.take(2 * M * F); w.f1_H2 = c.take(2 * B * F); w.fuse1 = c.take(M * F);
  w.pooled1 = c.take(B * G * D);
  w.vf = c.take(B * 2 * A);
  w.v2l = c.take(M * HP);
  w.f2_H1 = c.take(2 * M * F); w.f2_H2 = c.take(2 * B * F); w.fuse2 = c.take(M * F);
  w.pooled2 = c.take(B * G * D);
  w.ff_H1 = c.take(2 * B * F); w.ff_H2 = c.take(2 * B * F); w.xf = c.take(B * XP);
  // backward temporaries
  w.dxf = c.take(B * XP); w.dvf = c.take(B * 2 * A); w.dqf = c.take(B * HP); w.d_ff_H2 = c.take(2 * B * F);
  w.dpooled2 = c.take(B * G * D); w.dalpha2 = c.take(M * G); w.dz2 = c.take(M * G);
  w.dfuse2 = c.take(M * F); w.dv2 = c.take(M * D); w.dv2l = c.take(M * HP); w.d_f2_H2 = c.take(2 * B * F);
  w.dql = c.take(B * HP); w.dg1 = c.take(B * D); w.dg2 = c.take(B * D); w.dhq1 = c.take(B * HP); w.dhq2 = c.take(B * HP);
  w.dalpha_ext = c.take(B);
  w.dpooled1 = c.take(B * G * D); w.dalpha1 = c.take(M * G); w.dz1 = c.take(M * G);
  w.dfuse1 = c.take(M * F); w.dvl = c.take(M * HP); w.d_f1_H2 = c.take(2 * B * F);
  w.lin_ws_bytes = (size_t)lin_scratch_floats(B, N, C) * sizeof(float); w.lin_ws = c.take(lin_scratch_floats(B, N, C));
  w.side_ws_bytes = (size_t)(8 * B * 2048 + 65536) * sizeof(float); w.side_ws = c.take(8 * B * 2048 + 65536);
  w.dq_parts = c.take(4 * B * Q);
  {
    using namespace cor2;
    for (int i = 0; i < 32; ++i) { w.bits[i] = nullptr; w.bits_n[i] = 0; }
    w.bits_n[L_COMPRESS_V] = M * D; w.bits_n[L_COMPRESS_V2] = M * D; w.bits_n[L_ATT1_CONV] = M * F; w.bits_n[L_ATT2_CONV] = M * F;
    for (int l : {L_COMPRESS_Q, L_CQ1, L_CQ2, L_LINEAR_Q}) w.bits_n[l] = B * Q;
    w.bits_n[L_EQ1] = B * H; w.bits_n[L_EQ2] = B * H; w.bits_n[L_CLASSIF] = B * F;
    for (int g = 0; g < G; ++g) { w.bits_n[L_ATT1_G + g] = B * D; w.bits_n[L_ATT2_G + g] = B * D; }
    for (int i = 0; i < 32; ++i)
      if (w.bits_n[i]) w.bits[i] = reinterpret_cast<uint8_t*>(c.take(w.bits_n[i] / 32 + 16));
  }
  w.vq1_w1p = c.take(2 * FPAD * 312); w.vq1_w2p = c.take(2 * FPAD * 312);
  w.vq2_w1p = c.take(2 * FPAD * 312); w.vq2_w2p = c.take(2 * FPAD * 312);
  w.ff_w1p = c.take(2 * FPAD * 2 * A); w.ff_w2p = c.take(2 * FPAD * 312);
  w.eq1p = c.take(D * 312); w.eq2p = c.take(D * 312); w.clsp = c.take(C * 512);
  {  // plane buffers (sized for np = 2; bf16 elements = half floats)
    auto takeh = [&](int64_t n) { return reinterpret_cast<__nv_bfloat16*>(c.take((n + 1) / 2)); };
    w.vp = takeh(2 * M * D); w.v2p = takeh(2 * M * D); w.vlp = takeh(2 * M * HP); w.v2lp = takeh(2 * M * HP);
    w.cv_wp = takeh(2 * HP * D); w.cv2_wp = takeh(2 * HP * D); w.cv2_wtp = takeh(2 * D * HP);
    w.vq1_wp = takeh(2 * 2 * FPAD * HP); w.vq1_wtp = takeh(2 * H * 2 * FPAD);
    w.vq2_wp = takeh(2 * 2 * FPAD * HP); w.vq2_wtp = takeh(2 * H * 2 * FPAD);
  }
  w.bytes = c.off;
  return w;
}

struct OdaWs {
  float *vl, *ql, *qf, *wsum, *pooled, *vf, *ff_H1, *ff_H2, *xf;
  float *dxf, *dvf, *dqf, *d_ff_H2, *dpooled, *dalpha, *dz, *dwsum, *dvl, *dql;
  float* lin_ws; size_t lin_ws_bytes;
  float* side_ws; size_t side_ws_bytes;
  float* dq_parts;                                     // [2][B][Q]
  uint8_t* bits[32];
  int64_t bits_n[32];
  float *ff_w1p, *ff_w2p, *clsp;
  __nv_bfloat16 *vp, *cv_wp;                           // bf16 modes: planes of dropout(v) and of compress_v's weight
  void* pair_ws; size_t pair_ws_bytes;                 // pairwise attention scratch: keep bits of the [B,N,N*H] site + partial logits
  size_t bytes;
};

static OdaWs carve_oda(void* base, int64_t B, int64_t N, int64_t C) {
  Carver c{reinterpret_cast<char*>(base), 0};
  const int64_t M = B * N;
  OdaWs w;
  w.vl = c.take(M * H); w.ql = c.take(B * H); w.qf = c.take(B * HP); w.wsum = c.take(G * H);
  w.pooled = c.take(B * G * D); w.vf = c.take(B * A);
  w.ff_H1 = c.take(5 * B * F); w.ff_H2 = c.take(5 * B * F); w.xf = c.take(B * XP);
  w.dxf = c.take(B * XP); w.dvf = c.take(B * A); w.dqf = c.take(B * HP); w.d_ff_H2 = c.take(5 * B * F);
  w.dpooled = c.take(B * G * D); w.dalpha = c.take(M * G); w.dz = c.take(M * G); w.dwsum = c.take(G * H);
  w.dvl = c.take(M * H); w.dql = c.take(B * H);
  w.lin_ws_bytes = (size_t)lin_scratch_floats(B, N, C) * sizeof(float); w.lin_ws = c.take(lin_scratch_floats(B, N, C));
  w.side_ws_bytes = (size_t)(8 * B * 2048 + 65536) * sizeof(float); w.side_ws = c.take(8 * B * 2048 + 65536);
  w.dq_parts = c.take(2 * B * Q);
  {
    using namespace oda;
    for (int i = 0; i < 32; ++i) { w.bits[i] = nullptr; w.bits_n[i] = 0; }
    w.bits_n[L_COMPRESS_V] = M * D; w.bits_n[L_COMPRESS_Q] = B * Q; w.bits_n[L_LINEAR_Q] = B * Q; w.bits_n[L_CLASSIF] = B * F;
    for (int g = 0; g < G; ++g) w.bits_n[L_ATT_G + g] = B * D;
    for (int i = 0; i < 32; ++i)
      if (w.bits_n[i]) w.bits[i] = reinterpret_cast<uint8_t*>(c.take(w.bits_n[i] / 32 + 16));
    // the pairwise site's keep bits sit at the head of the pair kernels' own scratch (vqa_oda_pair_attn_workspace_bytes)
    w.pair_ws_bytes = vqa_oda_pair_attn_workspace_bytes(B, N, H);
    w.pair_ws = c.take((int64_t)(w.pair_ws_bytes / sizeof(float)) + 64);
    w.bits_n[L_ATT_CONV] = M * N * H; w.bits[L_ATT_CONV] = reinterpret_cast<uint8_t*>(w.pair_ws);
  }
  w.ff_w1p = c.take(5 * FPAD * A); w.ff_w2p = c.take(5 * FPAD * 312); w.clsp = c.take(C * 512);
  w.vp = reinterpret_cast<__nv_bfloat16*>(c.take(M * D)); w.cv_wp = reinterpret_cast<__nv_bfloat16*>(c.take(HP * D));
  w.bytes = c.off;
  return w;
}

// dst[i] = sum_{k < parts} src[k*n + i]   (the question-embedding gradient: one share per 2400->310 projection)
__global__ void sum_parts_kernel(int64_t n, int parts, const float* __restrict__ src, float* __restrict__ dst) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 s = *reinterpret_cast<const float4*>(src + i);
  for (int k = 1; k < parts; ++k) {
    const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)k * n + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *reinterpret_cast<float4*>(dst + i) = s;
}
static int sum_parts(int64_t n, int parts, const float* src, float* dst, cudaStream_t st) {
  sum_parts_kernel<<<(unsigned)cdiv(n / 4, 256), 256, 0, st>>>(n, parts, src, dst);
  return check_launch("sum_parts");
}

// ---- small builders ---------------------------------------------------------------------------------
struct Ctx {
  const vqa_model_fwd_params* p;
  void* stream;
  const float* const* W;      // parameter table
  float* const* dW;           // gradient table (backward) or nullptr
  int accumulate;
  void* lin_ws; size_t lin_ws_bytes;
  int packed;                 // padded weight copies in the workspace are valid (tensor-core math)
  uint8_t* const* bits;       // keep-bits by dropout layer id (train mode), else nullptr
  float pdrop() const { return p->train ? P_DROP : 0.0f; }
  float* grad(int idx) const { return dW ? dW[idx] : nullptr; }
};

// single or grouped linear forward; weight index widx[g] (bias = widx[g]+1)
static int lin_fwd(const Ctx& c, int groups, int64_t M, int64_t K, int64_t N, int act, const float* const* X,
                   const int64_t* ldx, const int* widx, float* const* Y, const int64_t* ldy, const uint32_t* layer,
                   const float* const* Wp = nullptr, const LinExt* ext = nullptr) {
  vqa_linear_fwd_params lp = {};
  lp.groups = groups; lp.M = M; lp.K = K; lp.N = N; lp.act = act; lp.math = c.p->math;
  lp.p = c.pdrop(); lp.seed = c.p->seed; lp.seed_dev = c.p->seed_dev;
  for (int g = 0; g < groups; ++g) {
    lp.X[g] = X[g]; lp.ldx[g] = ldx[g]; lp.W[g] = c.W[widx[g]]; lp.b[g] = c.W[widx[g] + 1];
    lp.Y[g] = Y[g]; lp.ldy[g] = ldy[g]; lp.layer[g] = layer[g]; lp.drop_index_base[g] = 0;
    lp.drop_bits[g] = c.bits ? c.bits[layer[g]] : nullptr;
    lp.Wp[g] = (Wp && c.packed) ? Wp[g] : nullptr;
  }
  lp.workspace = c.lin_ws; lp.workspace_bytes = c.lin_ws_bytes;
  if (ext) return tc_linear_fwd(&lp, (cudaStream_t)c.stream, ext);      // bf16 modes: operand planes handed over
  return vqa_linear_fwd(&lp, c.stream);
}

static int lin_bwd(const Ctx& c, int groups, int64_t M, int64_t K, int64_t N, int act, const float* const* X,
                   const int64_t* ldx, const int* widx, const float* const* Y, const int64_t* ldy,
                   const float* const* dY, const int64_t* lddy, float* const* dX, const int64_t* lddx, int accumulate_x,
                   const uint32_t* layer, const float* const* Wp = nullptr,
                   const float* pool_alpha = nullptr, const float* pool_dpooled = nullptr, int64_t pool_regions = 0,
                   const LinExt* ext = nullptr) {
  vqa_linear_bwd_params lp = {};
  lp.groups = groups; lp.M = M; lp.K = K; lp.N = N; lp.act = act; lp.math = c.p->math;
  lp.p = c.pdrop(); lp.seed = c.p->seed; lp.seed_dev = c.p->seed_dev; lp.accumulate_w = c.accumulate; lp.accumulate_x = accumulate_x;
  for (int g = 0; g < groups; ++g) {
    lp.X[g] = X[g]; lp.ldx[g] = ldx[g]; lp.W[g] = c.W[widx[g]];
    lp.Y[g] = Y[g]; lp.ldy[g] = ldy[g]; lp.dY[g] = dY[g]; lp.lddy[g] = lddy[g];
    lp.dW[g] = c.grad(widx[g]); lp.db[g] = c.grad(widx[g] + 1);
    lp.dX[g] = dX ? dX[g] : nullptr; lp.lddx[g] = lddx ? lddx[g] : 0;
    lp.layer[g] = layer[g]; lp.drop_index_base[g] = 0;
    lp.drop_bits[g] = c.bits ? c.bits[layer[g]] : nullptr;
    lp.Wp[g] = (Wp && c.packed) ? Wp[g] : nullptr;
  }
  lp.workspace = c.lin_ws; lp.workspace_bytes = c.lin_ws_bytes;
  lp.pool_alpha = pool_alpha; lp.pool_dpooled = pool_dpooled; lp.pool_regions = pool_regions;
  if (ext) return tc_linear_bwd(&lp, (cudaStream_t)c.stream, ext);
  return vqa_linear_bwd(&lp, c.stream);
}

static int mutan_fwd(const Ctx& c, int R, int64_t M, int64_t K1, int64_t K2, int64_t rows_per, const float* X1,
                     int64_t ldx1, const float* X2, int64_t ldx2, int l1, int l2, float* H1, float* H2, float* Y,
                     int64_t ldy, const float* W1p = nullptr, const float* W2p = nullptr, const MutanExt* ext = nullptr) {
  vqa_mutan_fwd_params mp = {};
  mp.R = R; mp.M = M; mp.K1 = K1; mp.K2 = K2; mp.F = F; mp.rows_per_h2 = rows_per; mp.math = c.p->math;
  mp.X1 = X1; mp.ldx1 = ldx1; mp.X2 = X2; mp.ldx2 = ldx2;
  for (int r = 0; r < R; ++r) {
    mp.W1[r] = c.W[l1 + 2 * r]; mp.b1[r] = c.W[l1 + 2 * r + 1];
    mp.W2[r] = c.W[l2 + 2 * r]; mp.b2[r] = c.W[l2 + 2 * r + 1];
  }
  mp.H1 = H1; mp.H2 = H2; mp.Y = Y; mp.ldy = ldy;
  if (c.packed) { mp.W1p = W1p; mp.W2p = W2p; }
  mp.workspace = c.lin_ws; mp.workspace_bytes = c.lin_ws_bytes;
  if (ext) return tc_mutan_fwd(&mp, (cudaStream_t)c.stream, ext);
  return vqa_mutan_fwd(&mp, c.stream);
}

static int mutan_bwd(const Ctx& c, int R, int64_t M, int64_t K1, int64_t K2, int64_t rows_per, const float* X1,
                     int64_t ldx1, const float* X2, int64_t ldx2, int l1, int l2, const float* H1, const float* H2,
                     const float* dY, int64_t lddy, float* dH2, float* dX1, int64_t lddx1, float* dX2, int64_t lddx2,
                     int accumulate_x2, const float* W1p = nullptr, const float* W2p = nullptr,
                     const MutanExt* ext = nullptr) {
  vqa_mutan_bwd_params mp = {};
  mp.R = R; mp.M = M; mp.K1 = K1; mp.K2 = K2; mp.F = F; mp.rows_per_h2 = rows_per; mp.math = c.p->math;
  mp.accumulate_w = c.accumulate; mp.accumulate_x1 = 0; mp.accumulate_x2 = accumulate_x2;
  mp.X1 = X1; mp.ldx1 = ldx1; mp.X2 = X2; mp.ldx2 = ldx2;
  for (int r = 0; r < R; ++r) {
    mp.W1[r] = c.W[l1 + 2 * r]; mp.W2[r] = c.W[l2 + 2 * r];
    mp.dW1[r] = c.grad(l1 + 2 * r); mp.db1[r] = c.grad(l1 + 2 * r + 1);
    mp.dW2[r] = c.grad(l2 + 2 * r); mp.db2[r] = c.grad(l2 + 2 * r + 1);
  }
  mp.H1 = H1; mp.H2 = H2; mp.dY = dY; mp.lddy = lddy; mp.dH2 = dH2;
  mp.dX1 = dX1; mp.lddx1 = lddx1; mp.dX2 = dX2; mp.lddx2 = lddx2;
  if (c.packed) { mp.W1p = W1p; mp.W2p = W2p; }
  mp.workspace = c.lin_ws; mp.workspace_bytes = c.lin_ws_bytes;
  if (ext) return tc_mutan_bwd(&mp, (cudaStream_t)c.stream, ext);
  return vqa_mutan_bwd(&mp, c.stream);
}

// one launch for the keep-bits of every dropout site of a plan
// `first` >= 0: only that layer (the site the main lane needs at once) / every layer but that one (the rest, made on
// the side lane under the first GEMM)
static int make_bits(const vqa_model_fwd_params* p, uint8_t* const* bits, const int64_t* n, void* stream, int first,
                     bool only_first) {
  vqa_bits_segment s[VQA_MAX_BITS_SEGMENTS];
  int k = 0;
  for (int i = 0; i < 32; ++i)
    if (bits[i] && n[i] > 0 && ((i == first) == only_first)) {
      s[k].layer = (uint32_t)i; s[k].n = (uint64_t)n[i]; s[k].out = bits[i]; ++k;
    }
  return vqa_dropout_bits_batch(P_DROP, p->seed, p->seed_dev, s, k, stream);
}

struct PackList {
  vqa_pack_segment s[VQA_MAX_PACK_SEGMENTS];
  int n = 0;
  void add(const float* src, float* dst, int64_t rows, int64_t rows_pad, int64_t K) {
    s[n].src = src; s[n].dst = dst; s[n].rows = rows; s[n].rows_pad = rows_pad; s[n].K = K; ++n;
  }
  // the R matrices of one Mutan side, stacked with FPAD rows each
  void add_mutan(const float* const* W, int l, int R, float* dst, int64_t K) {
    const int64_t Kp = (K + 3) / 4 * 4;
    for (int r = 0; r < R; ++r) add(W[l + 2 * r], dst + (int64_t)r * FPAD * Kp, F, FPAD, K);
  }
};

// MyATT glimpse linears: pooled[B,G,D] -> vf[:, col0 + g*155 ...]
static int glimpse_fwd(const Ctx& c, int64_t B, const float* pooled, int w0, float* vf, int64_t ldvf, int64_t col0,
                       uint32_t layer0) {
  const float* X[G]; int64_t ldx[G]; int widx[G]; float* Y[G]; int64_t ldy[G]; uint32_t layer[G];
  for (int g = 0; g < G; ++g) {
    X[g] = pooled + g * D; ldx[g] = G * D; widx[g] = w0 + 2 * g; Y[g] = vf + col0 + g * AG; ldy[g] = ldvf;
    layer[g] = layer0 + g;
  }
  return lin_fwd(c, G, B, D, AG, VQA_ACT_RELU, X, ldx, widx, Y, ldy, layer);
}
static int glimpse_bwd(const Ctx& c, int64_t B, const float* pooled, int w0, const float* vf, const float* dvf,
                       int64_t ldvf, int64_t col0, float* dpooled, uint32_t layer0) {
  const float* X[G]; int64_t ldx[G]; int widx[G]; const float* Y[G]; int64_t ldy[G]; const float* dY[G];
  int64_t lddy[G]; float* dX[G]; int64_t lddx[G]; uint32_t layer[G];
  for (int g = 0; g < G; ++g) {
    X[g] = pooled + g * D; ldx[g] = G * D; widx[g] = w0 + 2 * g;
    Y[g] = vf + col0 + g * AG; ldy[g] = ldvf; dY[g] = dvf + col0 + g * AG; lddy[g] = ldvf;
    dX[g] = dpooled + g * D; lddx[g] = G * D; layer[g] = layer0 + g;
  }
  return lin_bwd(c, G, B, D, AG, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, dX, lddx, 0, layer);
}

static int check_model(const vqa_model_fwd_params* p, size_t need, const char* who, bool cor2) {
  VQA_REQUIRE(p != nullptr, "%s: null params", who);
  VQA_REQUIRE(p->B >= 1 && p->N >= 1 && p->C >= 1, "%s: bad shape B=%lld N=%lld C=%lld", who, (long long)p->B,
              (long long)p->N, (long long)p->C);
  VQA_REQUIRE(p->v && p->q && p->params && p->logits && p->alpha1 && p->workspace, "%s: null pointer", who);
  if (cor2) VQA_REQUIRE(p->alpha2 && p->v2, "%s: alpha2 / v2 output buffers required", who);
  if (p->workspace_bytes < need) {
    set_error("%s: workspace has %zu bytes, %zu needed", who, p->workspace_bytes, need);
    return VQA_EWORKSPACE;
  }
  return vqa_device_check();
}

}  // namespace vqa

using namespace vqa;

// ====================================================================================== introspection
// Location of a forward stash tensor inside the caller's workspace (used by the parity tests to read the
// ReLU activation patterns the kernels actually produced).
extern "C" int vqa_stash_info(int model, const char* name, int64_t B, int64_t N, int64_t C, size_t* offset_bytes,
                              int64_t* rows, int64_t* cols, int64_t* ld) {
  VQA_REQUIRE(name && offset_bytes && rows && cols && ld, "vqa_stash_info: null pointer");
  const int64_t M = B * N;
  char* base = reinterpret_cast<char*>(0x1000);   // fake base: only offsets are used
  struct Ent { const char* n; const float* p; int64_t r, c, l; };
  if (model == 0) {
    Cor2Ws w = carve_cor2(base, B, N, C);
    const Ent tab[] = {{"compress_v", w.vl, M, H, HP}, {"compress_v2", w.v2l, M, H, HP}, {"compress_q", w.ql, B, H, HP},
                       {"compress_q_1", w.hq1, B, H, HP}, {"compress_q_2", w.hq2, B, H, HP}, {"linear_q", w.qf, B, H, HP},
                       {"glimpses", w.vf, B, 2 * A, 2 * A}};
    for (const Ent& e : tab)
      if (strcmp(e.n, name) == 0) {
        *offset_bytes = (size_t)(reinterpret_cast<const char*>(e.p) - base); *rows = e.r; *cols = e.c; *ld = e.l;
        return VQA_OK;
      }
  } else {
    OdaWs w = carve_oda(base, B, N, C);
    const Ent tab[] = {{"compress_v", w.vl, M, H, H}, {"compress_q", w.ql, B, H, H}, {"linear_q", w.qf, B, H, HP},
                       {"glimpses", w.vf, B, A, A}};
    for (const Ent& e : tab)
      if (strcmp(e.n, name) == 0) {
        *offset_bytes = (size_t)(reinterpret_cast<const char*>(e.p) - base); *rows = e.r; *cols = e.c; *ld = e.l;
        return VQA_OK;
      }
  }
  set_error("vqa_stash_info: unknown stash tensor '%s'", name);
  return VQA_EINVAL;
}

// ====================================================================================== CoR2
extern "C" size_t vqa_cor2_workspace_bytes(int64_t B, int64_t N, int64_t C) {
  return carve_cor2(nullptr, B, N, C).bytes;
}

extern "C" int vqa_cor2_fwd(const vqa_model_fwd_params* p, void* stream) {
  VQA_TRY(check_model(p, p ? carve_cor2(nullptr, p->B, p->N, p->C).bytes : 0, "vqa_cor2_fwd", true));
  using namespace cor2;
  const int64_t B = p->B, N = p->N, M = B * N;
  Cor2Ws w = carve_cor2(p->workspace, B, N, p->C);
  Ctx c{p, stream, p->params, nullptr, 0, w.lin_ws, w.lin_ws_bytes, p->math != VQA_MATH_FP32_SIMT, p->train ? w.bits : nullptr};
  const float* eqp[2] = {w.eq1p, w.eq2p}; const float* clp[1] = {w.clsp}; (void)eqp; (void)clp;
  // Head of the step.  Main lane: keep-bits of compress_v's input, then straight into compress_v (its weight needs
  // no packing).  Side lane: the keep-bits of every other dropout site (the question projections below read theirs
  // in stream order; the main lane reads its next ones only after waiting for e_ql), the packed copies of every
  // weight whose rows TMA cannot address — made once for this step's forward AND backward — then the
  // question-side chain (four projections -> gates).  The main lane first touches a packed weight after e_ql too.
  Lanes* L = get_lanes();
  VQA_REQUIRE(L != nullptr, "vqa_cor2_fwd: cannot create the internal side stream");
  cudaStream_t ms = (cudaStream_t)stream, ss = L->side;
  Ctx cs = c; cs.stream = ss; cs.lin_ws = w.side_ws; cs.lin_ws_bytes = w.side_ws_bytes;
  cudaEvent_t e_ql, e_gates;
  // bf16 math modes: the large GEMMs read bf16 operand planes (gemm_tc.h).  compress_v's input planes are made here
  // with the dropout mask applied (Philox in registers: no keep-bit cache for this site, the planes ARE its stash).
  const bool b16 = is_bf16_math(p->math) && M >= TC16_MIN_M;
  const int np = p->math == VQA_MATH_BF16X3 ? 2 : 1;
  LinExt x_cv, x_cv2;
  MutanExt x_vq1, x_vq2;
  if (b16) {
    x_cv.Xp = Planes{w.vp, D, M * D}; x_cv.Wp = Planes{w.cv_wp, D, HP * D};
    x_cv.Yp = w.vlp; x_cv.ldyp = HP; x_cv.yplane = M * HP;
    x_cv2.Xp = Planes{w.v2p, D, M * D}; x_cv2.Wp = Planes{w.cv2_wp, D, HP * D};
    x_cv2.Yp = w.v2lp; x_cv2.ldyp = HP; x_cv2.yplane = M * HP;
    x_vq1.X1p = Planes{w.vlp, HP, M * HP}; x_vq1.W1p = Planes{w.vq1_wp, HP, 2 * FPAD * HP};
    x_vq2.X1p = Planes{w.v2lp, HP, M * HP}; x_vq2.W1p = Planes{w.vq2_wp, HP, 2 * FPAD * HP};
  }
  if (p->train && !b16) {                          // the mask compress_v needs right away
    ProfScope ps_(stream, "dropout_bits");
    VQA_TRY(make_bits(p, w.bits, w.bits_n, stream, L_COMPRESS_V, true));
  }
  if (b16) {
    ProfScope ps_(stream, "planes.v");
    PackPlanesSeg sg = {};
    sg.src = c.W[COMPRESS_V]; sg.rows = H; sg.rows_pad = HP; sg.K = D; sg.Kp = D; sg.dst = w.cv_wp; sg.plane = HP * D;
    VQA_TRY(tc16::pack_planes(&sg, 1, np, ms));
    const float* X[1] = {p->v}; int64_t ldx[1] = {D}; uint32_t layer[1] = {L_COMPRESS_V}; uint64_t base[1] = {0};
    __nv_bfloat16* out[1] = {w.vp};
    VQA_TRY(tc16::split_planes(X, ldx, 1, M, D, c.pdrop(), p->seed, p->seed_dev, layer, base, nullptr, out, D, M * D, np, ms));
  }
  Lanes::wait(ss, L->record(ms));                  // fork
  if (p->train) {                                  // every other site: first read after the main lane's e_ql wait
    ProfScope ps_(ss, "dropout_bits.rest");
    VQA_TRY(make_bits(p, w.bits, w.bits_n, ss, L_COMPRESS_V, false));
  }
  if (c.packed) {
    ProfScope ps_(ss, "pack_weights");
    PackList pl;
    pl.add_mutan(c.W, VQ1_L1, 2, w.vq1_w1p, H); pl.add_mutan(c.W, VQ1_L2, 2, w.vq1_w2p, H);
    pl.add_mutan(c.W, VQ2_L1, 2, w.vq2_w1p, H); pl.add_mutan(c.W, VQ2_L2, 2, w.vq2_w2p, H);
    pl.add_mutan(c.W, FF_L1, 2, w.ff_w1p, 2 * A); pl.add_mutan(c.W, FF_L2, 2, w.ff_w2p, H);
    pl.add(c.W[EQ1], w.eq1p, D, D, H); pl.add(c.W[EQ2], w.eq2p, D, D, H);
    pl.add(c.W[CLASSIF], w.clsp, p->C, p->C, F);
    VQA_TRY(vqa_pack_weights(pl.s, pl.n, ss));
    if (b16) {      // operand planes of the remaining large-GEMM weights, both orientations, one launch
      PackPlanesSeg sg[5] = {};
      sg[0].src = c.W[COMPRESS_V2]; sg[0].rows = H; sg[0].rows_pad = HP; sg[0].K = D; sg[0].Kp = D;
      sg[0].dst = w.cv2_wp; sg[0].plane = HP * D; sg[0].dstT = w.cv2_wtp; sg[0].Np = HP; sg[0].planeT = D * HP; sg[0].t_col0 = 0;
      for (int r = 0; r < 2; ++r) {
        PackPlanesSeg& a = sg[1 + r]; PackPlanesSeg& b = sg[3 + r];
        a.src = c.W[VQ1_L1 + 2 * r]; b.src = c.W[VQ2_L1 + 2 * r];
        a.rows = b.rows = F; a.rows_pad = b.rows_pad = FPAD; a.K = b.K = H; a.Kp = b.Kp = HP;
        a.dst = w.vq1_wp + (int64_t)r * FPAD * HP; b.dst = w.vq2_wp + (int64_t)r * FPAD * HP;
        a.plane = b.plane = 2 * FPAD * HP;
        a.dstT = w.vq1_wtp; b.dstT = w.vq2_wtp; a.Np = b.Np = 2 * FPAD; a.planeT = b.planeT = H * 2 * FPAD;
        a.t_col0 = b.t_col0 = (int64_t)r * FPAD;
      }
      VQA_TRY(tc16::pack_planes(sg, 5, np, ss));
    }
  }
  {  // four 2400->310 question projections in one launch (config/CoR2.py:211,195,196,228)
    const float* X[4] = {p->q, p->q, p->q, p->q}; int64_t ldx[4] = {Q, Q, Q, Q};
    int widx[4] = {COMPRESS_Q, CQ1, CQ2, LINEAR_Q}; float* Y[4] = {w.ql, w.hq1, w.hq2, w.qf};
    int64_t ldy[4] = {HP, HP, HP, HP}; uint32_t layer[4] = {L_COMPRESS_Q, L_CQ1, L_CQ2, L_LINEAR_Q};
    { ProfScope ps_(ss, "q_proj4.fwd"); VQA_TRY(lin_fwd(cs, 4, B, Q, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, layer)); }
    if (b16) {     // the question halves H2 of both region-question fusions, off the main lane
      ProfScope ps_(ss, "fusion_vq.h2");
      MutanExt h1 = x_vq1, h2 = x_vq2;
      h1.h2_mode = h2.h2_mode = 1;
      VQA_TRY(mutan_fwd(cs, 2, M, H, H, N, w.vl, HP, w.ql, HP, VQ1_L1, VQ1_L2, w.f1_H1, w.f1_H2, w.fuse1, F, w.vq1_w1p, w.vq1_w2p, &h1));
      VQA_TRY(mutan_fwd(cs, 2, M, H, H, N, w.v2l, HP, w.ql, HP, VQ2_L1, VQ2_L2, w.f2_H1, w.f2_H2, w.fuse2, F, w.vq2_w1p, w.vq2_w2p, &h2));
      x_vq1.h2_mode = x_vq2.h2_mode = 2;
    }
    e_ql = L->record(ss);
  }
  {  // gates g1, g2 = sigmoid(310->2048) (config/CoR2.py:195-196)
    const float* X[2] = {w.hq1, w.hq2}; int64_t ldx[2] = {HP, HP}; int widx[2] = {EQ1, EQ2};
    float* Y[2] = {w.g1, w.g2}; int64_t ldy[2] = {D, D}; uint32_t layer[2] = {L_EQ1, L_EQ2};
    { ProfScope ps_(ss, "gates.fwd"); VQA_TRY(lin_fwd(cs, 2, B, H, D, VQA_ACT_SIGMOID, X, ldx, widx, Y, ldy, layer, eqp)); }
    e_gates = L->record(ss);
  }
  {  // compress_v (config/CoR2.py:213)
    const float* X[1] = {p->v}; int64_t ldx[1] = {D}; int widx[1] = {COMPRESS_V}; float* Y[1] = {w.vl};
    int64_t ldy[1] = {HP}; uint32_t layer[1] = {L_COMPRESS_V};
    { ProfScope ps_(stream, "compress_v.fwd"); VQA_TRY(lin_fwd(c, 1, M, D, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, layer, nullptr, b16 ? &x_cv : nullptr)); }
  }
  Lanes::wait(ms, e_ql);      // ql / qf ready
  { ProfScope ps_(stream, "fusion_vq1.fwd"); VQA_TRY(mutan_fwd(c, 2, M, H, H, N, w.vl, HP, w.ql, HP, VQ1_L1, VQ1_L2, w.f1_H1, w.f1_H2, w.fuse1, F, w.vq1_w1p, w.vq1_w2p, b16 ? &x_vq1 : nullptr)); }   // fusion_vq1 :214
  {  // att1 on raw v (:214)
    vqa_region_softmax_pool_fwd_params ap = {};
    ap.B = B; ap.N = N; ap.Ff = F; ap.D = D;
    ap.drop.p = c.pdrop(); ap.drop.layer = L_ATT1_CONV; ap.drop.seed = p->seed; ap.drop.seed_dev = p->seed_dev;
    ap.drop_bits = p->train ? w.bits[L_ATT1_CONV] : nullptr;
    ap.fuse = w.fuse1; ap.Wc = c.W[ATT1_CONV]; ap.bc = c.W[ATT1_CONV + 1]; ap.x = p->v;
    ap.alpha = p->alpha1; ap.pooled = w.pooled1;
    { ProfScope ps_(stream, "att1.pool.fwd"); VQA_TRY(vqa_region_softmax_pool_fwd(&ap, stream)); }
  }
  { ProfScope ps_(stream, "att1.glimpse.fwd"); VQA_TRY(glimpse_fwd(c, B, w.pooled1, ATT1_G, w.vf, 2 * A, 0, L_ATT1_G)); }
  Lanes::wait(ms, e_gates);   // join: g1, g2 ready (last work of the side lane)
  {  // compound objects (:215-216)
    vqa_cor_compound_fwd_params cp = {};
    cp.B = B; cp.N = N; cp.D = D; cp.x = p->v; cp.pooled = w.pooled1; cp.alpha = p->alpha1; cp.g1 = w.g1; cp.g2 = w.g2;
    cp.v2 = p->v2;
    if (b16) {     // the operand planes of dropout(v2) for compress_v2's forward and weight-gradient GEMMs, same pass
      cp.v2_planes = w.v2p; cp.v2_nplanes = np; cp.v2_plane_stride = M * D;
      cp.v2_keep_bits = p->train ? w.bits[L_COMPRESS_V2] : nullptr; cp.v2_keep_scale = 1.0f / (1.0f - P_DROP);
    }
    { ProfScope ps_(stream, "compound.fwd"); VQA_TRY(vqa_cor_compound_fwd(&cp, stream)); }
  }
  {  // compress_v2 (:218)
    const float* X[1] = {p->v2}; int64_t ldx[1] = {D}; int widx[1] = {COMPRESS_V2}; float* Y[1] = {w.v2l};
    int64_t ldy[1] = {HP}; uint32_t layer[1] = {L_COMPRESS_V2};
    { ProfScope ps_(stream, "compress_v2.fwd"); VQA_TRY(lin_fwd(c, 1, M, D, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, layer, nullptr, b16 ? &x_cv2 : nullptr)); }
  }
  { ProfScope ps_(stream, "fusion_vq2.fwd"); VQA_TRY(mutan_fwd(c, 2, M, H, H, N, w.v2l, HP, w.ql, HP, VQ2_L1, VQ2_L2, w.f2_H1, w.f2_H2, w.fuse2, F, w.vq2_w1p, w.vq2_w2p, b16 ? &x_vq2 : nullptr)); }  // fusion_vq2 :219
  {  // att2 on v2 (:219)
    vqa_region_softmax_pool_fwd_params ap = {};
    ap.B = B; ap.N = N; ap.Ff = F; ap.D = D;
    ap.drop.p = c.pdrop(); ap.drop.layer = L_ATT2_CONV; ap.drop.seed = p->seed; ap.drop.seed_dev = p->seed_dev;
    ap.drop_bits = p->train ? w.bits[L_ATT2_CONV] : nullptr;
    ap.fuse = w.fuse2; ap.Wc = c.W[ATT2_CONV]; ap.bc = c.W[ATT2_CONV + 1]; ap.x = p->v2;
    ap.alpha = p->alpha2; ap.pooled = w.pooled2;
    { ProfScope ps_(stream, "att2.pool.fwd"); VQA_TRY(vqa_region_softmax_pool_fwd(&ap, stream)); }
  }
  { ProfScope ps_(stream, "att2.glimpse.fwd"); VQA_TRY(glimpse_fwd(c, B, w.pooled2, ATT2_G, w.vf, 2 * A, A, L_ATT2_G)); }
  { ProfScope ps_(stream, "fusion_final.fwd"); VQA_TRY(mutan_fwd(c, 2, B, 2 * A, H, 1, w.vf, 2 * A, w.qf, HP, FF_L1, FF_L2, w.ff_H1, w.ff_H2, w.xf, XP, w.ff_w1p, w.ff_w2p)); }    // fusion_final :233
  {  // linear_classif (:236)
    const float* X[1] = {w.xf}; int64_t ldx[1] = {XP}; int widx[1] = {CLASSIF}; float* Y[1] = {p->logits};
    int64_t ldy[1] = {p->C}; uint32_t layer[1] = {L_CLASSIF};
    { ProfScope ps_(stream, "classif.fwd"); VQA_TRY(lin_fwd(c, 1, B, F, p->C, VQA_ACT_NONE, X, ldx, widx, Y, ldy, layer, clp)); }
  }
  return VQA_OK;
}

extern "C" int vqa_cor2_bwd(const vqa_model_bwd_params* bp, void* stream) {
  VQA_REQUIRE(bp != nullptr, "vqa_cor2_bwd: null params");
  const vqa_model_fwd_params* p = &bp->fwd;
  VQA_TRY(check_model(p, carve_cor2(nullptr, p->B, p->N, p->C).bytes, "vqa_cor2_bwd", true));
  VQA_REQUIRE(bp->dlogits && bp->grads, "vqa_cor2_bwd: null dlogits / grads");
  using namespace cor2;
  const int64_t B = p->B, N = p->N, M = B * N;
  Cor2Ws w = carve_cor2(p->workspace, B, N, p->C);
  int acc = bp->accumulate;
  if (!acc && bp->grads_flat && bp->grads_flat_bytes) {   // one zero-fill for every gradient tensor, then everything accumulates
    cudaMemsetAsync(bp->grads_flat, 0, bp->grads_flat_bytes, (cudaStream_t)stream);
    acc = 1;
  }
  Ctx c{p, stream, p->params, bp->grads, acc, w.lin_ws, w.lin_ws_bytes, p->math != VQA_MATH_FP32_SIMT, p->train ? w.bits : nullptr};
  const float* eqp[2] = {w.eq1p, w.eq2p}; const float* clp[1] = {w.clsp}; (void)eqp; (void)clp;
  // Side lane: att1's glimpse linears, the gates and the question projections are off the critical dgrad chain
  // (classif -> fusion_final -> att2 -> fusion_vq2 -> compress_v2 -> compound -> att1 -> fusion_vq1 -> compress_v).
  Lanes* L = get_lanes();
  VQA_REQUIRE(L != nullptr, "vqa_cor2_bwd: cannot create the internal side stream");
  cudaStream_t ms = (cudaStream_t)stream, ss = L->side;
  Ctx cs = c; cs.stream = ss; cs.lin_ws = w.side_ws; cs.lin_ws_bytes = w.side_ws_bytes;
  const bool b16 = is_bf16_math(p->math) && M >= TC16_MIN_M;          // operand planes left by the forward (see vqa_cor2_fwd)
  LinExt x_cv, x_cv2;
  MutanExt x_vq1, x_vq2;
  if (b16) {
    x_cv.Xp = Planes{w.vp, D, M * D};
    x_cv2.Xp = Planes{w.v2p, D, M * D}; x_cv2.WTp = Planes{w.cv2_wtp, HP, D * HP};
    x_cv2.raw_dx = true;       // dv2 is finished (mask + att2 pooling gradient) by compound.bwd while it reads it
    x_vq1.X1p = Planes{w.vlp, HP, M * HP}; x_vq1.W1Tp = Planes{w.vq1_wtp, 2 * FPAD, H * 2 * FPAD};
    x_vq2.X1p = Planes{w.v2lp, HP, M * HP}; x_vq2.W1Tp = Planes{w.vq2_wtp, 2 * FPAD, H * 2 * FPAD};
  }
  {  // linear_classif
    const float* X[1] = {w.xf}; int64_t ldx[1] = {XP}; int widx[1] = {CLASSIF}; const float* Y[1] = {p->logits};
    int64_t ldy[1] = {p->C}; const float* dY[1] = {bp->dlogits}; int64_t lddy[1] = {p->C};
    float* dX[1] = {w.dxf}; int64_t lddx[1] = {XP}; uint32_t layer[1] = {L_CLASSIF};
    { ProfScope ps_(stream, "classif.bwd"); VQA_TRY(lin_bwd(c, 1, B, F, p->C, VQA_ACT_NONE, X, ldx, widx, Y, ldy, dY, lddy, dX, lddx, 0, layer, clp)); }
  }
  // gradient groups (vqa_grad_groups): an event per group lets the caller reduce finished buckets early
  auto mark = [&](int group, cudaStream_t s) {
    if (bp->group_events[group]) cudaEventRecord((cudaEvent_t)bp->group_events[group], s);
  };
  mark(0, ms);
  { ProfScope ps_(stream, "fusion_final.bwd"); VQA_TRY(mutan_bwd(c, 2, B, 2 * A, H, 1, w.vf, 2 * A, w.qf, HP, FF_L1, FF_L2, w.ff_H1, w.ff_H2, w.dxf, XP, w.d_ff_H2, w.dvf, 2 * A, w.dqf, HP, 0, w.ff_w1p, w.ff_w2p)); }
  mark(1, ms);
  const cudaEvent_t e_ff = L->record(ms);          // dvf, dqf ready
  Lanes::wait(ss, e_ff);
  // The four 2400->310 question projections are differentiated as soon as THEIR output gradient exists — linear_q
  // here (dqf), compress_q_1/2 after the gates, only compress_q (dql) at the very end — so that three quarters of
  // their 12 MB of weight gradients reach the all-reduce early instead of forming its exposed tail.
  float* dqp[4] = {w.dq_parts, w.dq_parts + B * Q, w.dq_parts + 2 * B * Q, w.dq_parts + 3 * B * Q};
  {
    const float* X[1] = {p->q}; int64_t ldx[1] = {Q}; int widx[1] = {LINEAR_Q}; const float* Y[1] = {w.qf};
    int64_t ldy[1] = {HP}; const float* dY[1] = {w.dqf}; int64_t lddy[1] = {HP}; uint32_t layer[1] = {L_LINEAR_Q};
    float* dX[1] = {dqp[3]}; int64_t lddx[1] = {Q};
    { ProfScope ps_(ss, "linear_q.bwd"); VQA_TRY(lin_bwd(cs, 1, B, Q, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, bp->dq ? dX : nullptr, bp->dq ? lddx : nullptr, 0, layer)); }
  }
  mark(2, ss);
  { ProfScope ps_(ss, "att1.glimpse.bwd"); VQA_TRY(glimpse_bwd(cs, B, w.pooled1, ATT1_G, w.vf, w.dvf, 2 * A, 0, w.dpooled1, L_ATT1_G)); }
  mark(3, ss);
  const cudaEvent_t e_g1 = L->record(ss);          // dpooled1 initialised
  // ---- att2 branch
  { ProfScope ps_(stream, "att2.glimpse.bwd"); VQA_TRY(glimpse_bwd(c, B, w.pooled2, ATT2_G, w.vf, w.dvf, 2 * A, A, w.dpooled2, L_ATT2_G)); }
  mark(4, ms);
  {
    vqa_region_softmax_pool_bwd_params ap = {};
    ap.B = B; ap.N = N; ap.Ff = F; ap.D = D;
    ap.drop.p = c.pdrop(); ap.drop.layer = L_ATT2_CONV; ap.drop.seed = p->seed; ap.drop.seed_dev = p->seed_dev;
    ap.drop_bits = p->train ? w.bits[L_ATT2_CONV] : nullptr;
    ap.accumulate_w = acc; ap.accumulate_x = 0;
    ap.fuse = w.fuse2; ap.Wc = c.W[ATT2_CONV]; ap.x = p->v2; ap.alpha = p->alpha2; ap.dpooled = w.dpooled2;
    ap.dalpha0_ext = nullptr; ap.dalpha = w.dalpha2; ap.dz = w.dz2;
    ap.dWc = c.grad(ATT2_CONV); ap.dbc = c.grad(ATT2_CONV + 1); ap.dfuse = w.dfuse2;
    ap.dx = nullptr;             // the pooling's share of dv2 is added by compress_v2's dgrad epilogue below
    { ProfScope ps_(stream, "att2.pool.bwd"); VQA_TRY(vqa_region_softmax_pool_bwd(&ap, stream)); }
  }
  mark(5, ms);
  { ProfScope ps_(stream, "fusion_vq2.bwd"); VQA_TRY(mutan_bwd(c, 2, M, H, H, N, w.v2l, HP, w.ql, HP, VQ2_L1, VQ2_L2, w.f2_H1, w.f2_H2, w.dfuse2, F, w.d_f2_H2, w.dv2l, HP, w.dql, HP, 0, w.vq2_w1p, w.vq2_w2p, b16 ? &x_vq2 : nullptr)); }
  mark(6, ms);
  {  // compress_v2: v2 feeds both compress_v2 and att2's pooling; the dgrad store adds sum_g alpha2 * dpooled2
    const float* X[1] = {p->v2}; int64_t ldx[1] = {D}; int widx[1] = {COMPRESS_V2}; const float* Y[1] = {w.v2l};
    int64_t ldy[1] = {HP}; const float* dY[1] = {w.dv2l}; int64_t lddy[1] = {HP};
    float* dX[1] = {w.dv2}; int64_t lddx[1] = {D}; uint32_t layer[1] = {L_COMPRESS_V2};
    { ProfScope ps_(stream, "compress_v2.bwd"); VQA_TRY(lin_bwd(c, 1, M, D, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, dX, lddx, 0, layer, nullptr, p->alpha2, w.dpooled2, N, b16 ? &x_cv2 : nullptr)); }
  }
  mark(7, ms);
  // ---- att1 branch: the glimpse linears (side lane) initialise dpooled1, then the compound objects add to it
  Lanes::wait(ms, e_g1);
  {
    vqa_cor_compound_bwd_params cp = {};
    cp.B = B; cp.N = N; cp.D = D; cp.x = p->v; cp.pooled = w.pooled1; cp.alpha = p->alpha1; cp.g1 = w.g1; cp.g2 = w.g2;
    cp.dv2 = w.dv2; cp.dg1 = w.dg1; cp.dg2 = w.dg2; cp.dpooled = w.dpooled1; cp.dalpha0_ext = w.dalpha_ext;
    if (b16) {                 // w.dv2 holds the raw dZ.W of compress_v2's dgrad
      cp.dv2_keep_bits = p->train ? w.bits[L_COMPRESS_V2] : nullptr; cp.dv2_keep_scale = 1.0f / (1.0f - P_DROP);
      cp.dv2_pool_alpha = p->alpha2; cp.dv2_pool_dpooled = w.dpooled2;
    }
    { ProfScope ps_(stream, "compound.bwd"); VQA_TRY(vqa_cor_compound_bwd(&cp, stream)); }
  }
  {  // gates
    const float* X[2] = {w.hq1, w.hq2}; int64_t ldx[2] = {HP, HP}; int widx[2] = {EQ1, EQ2};
    const float* Y[2] = {w.g1, w.g2}; int64_t ldy[2] = {D, D}; const float* dY[2] = {w.dg1, w.dg2};
    int64_t lddy[2] = {D, D}; float* dX[2] = {w.dhq1, w.dhq2}; int64_t lddx[2] = {HP, HP};
    uint32_t layer[2] = {L_EQ1, L_EQ2};
    Lanes::wait(ss, L->record(ms));                // dg1, dg2 ready
    { ProfScope ps_(ss, "gates.bwd"); VQA_TRY(lin_bwd(cs, 2, B, H, D, VQA_ACT_SIGMOID, X, ldx, widx, Y, ldy, dY, lddy, dX, lddx, 0, layer, eqp)); }
  }
  mark(8, ss);
  {  // compress_q_1 / compress_q_2 (the gates' first layers): dhq1, dhq2 were just written on this lane
    const float* X[2] = {p->q, p->q}; int64_t ldx[2] = {Q, Q}; int widx[2] = {CQ1, CQ2};
    const float* Y[2] = {w.hq1, w.hq2}; int64_t ldy[2] = {HP, HP}; const float* dY[2] = {w.dhq1, w.dhq2};
    int64_t lddy[2] = {HP, HP}; uint32_t layer[2] = {L_CQ1, L_CQ2};
    float* dX[2] = {dqp[1], dqp[2]}; int64_t lddx[2] = {Q, Q};
    { ProfScope ps_(ss, "compress_q12.bwd"); VQA_TRY(lin_bwd(cs, 2, B, Q, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, bp->dq ? dX : nullptr, bp->dq ? lddx : nullptr, 0, layer)); }
  }
  mark(9, ss);
  {
    vqa_region_softmax_pool_bwd_params ap = {};
    ap.B = B; ap.N = N; ap.Ff = F; ap.D = D;
    ap.drop.p = c.pdrop(); ap.drop.layer = L_ATT1_CONV; ap.drop.seed = p->seed; ap.drop.seed_dev = p->seed_dev;
    ap.drop_bits = p->train ? w.bits[L_ATT1_CONV] : nullptr;
    ap.accumulate_w = acc; ap.accumulate_x = 0;
    ap.fuse = w.fuse1; ap.Wc = c.W[ATT1_CONV]; ap.x = p->v; ap.alpha = p->alpha1; ap.dpooled = w.dpooled1;
    ap.dalpha0_ext = w.dalpha_ext; ap.dalpha = w.dalpha1; ap.dz = w.dz1;
    ap.dWc = c.grad(ATT1_CONV); ap.dbc = c.grad(ATT1_CONV + 1); ap.dfuse = w.dfuse1; ap.dx = nullptr;
    { ProfScope ps_(stream, "att1.pool.bwd"); VQA_TRY(vqa_region_softmax_pool_bwd(&ap, stream)); }
  }
  mark(10, ms);
  { ProfScope ps_(stream, "fusion_vq1.bwd"); VQA_TRY(mutan_bwd(c, 2, M, H, H, N, w.vl, HP, w.ql, HP, VQ1_L1, VQ1_L2, w.f1_H1, w.f1_H2, w.dfuse1, F, w.d_f1_H2, w.dvl, HP, w.dql, HP, 1, w.vq1_w1p, w.vq1_w2p, b16 ? &x_vq1 : nullptr)); }
  mark(11, ms);
  Lanes::wait(ss, L->record(ms));                  // dql complete (fusion_vq2 + fusion_vq1)
  {  // compress_v: v is a graph input, no dgrad
    const float* X[1] = {p->v}; int64_t ldx[1] = {D}; int widx[1] = {COMPRESS_V}; const float* Y[1] = {w.vl};
    int64_t ldy[1] = {HP}; const float* dY[1] = {w.dvl}; int64_t lddy[1] = {HP}; uint32_t layer[1] = {L_COMPRESS_V};
    { ProfScope ps_(stream, "compress_v.bwd"); VQA_TRY(lin_bwd(c, 1, M, D, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, nullptr, nullptr, 0, layer, nullptr, nullptr, nullptr, 0, b16 ? &x_cv : nullptr)); }
  }
  mark(12, ms);
  {  // compress_q: its output gradient dql is complete only now (fusion_vq2 + fusion_vq1)
    const float* X[1] = {p->q}; int64_t ldx[1] = {Q}; int widx[1] = {COMPRESS_Q}; const float* Y[1] = {w.ql};
    int64_t ldy[1] = {HP}; const float* dY[1] = {w.dql}; int64_t lddy[1] = {HP}; uint32_t layer[1] = {L_COMPRESS_Q};
    float* dX[1] = {dqp[0]}; int64_t lddx[1] = {Q};
    { ProfScope ps_(ss, "compress_q.bwd"); VQA_TRY(lin_bwd(cs, 1, B, Q, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, bp->dq ? dX : nullptr, bp->dq ? lddx : nullptr, 0, layer)); }
    if (bp->dq) VQA_TRY(sum_parts(B * Q, 4, w.dq_parts, bp->dq, ss));
  }
  mark(13, ss);
  Lanes::wait(ms, L->record(ss));                  // join
  return VQA_OK;
}

// Completion groups of the two backward plans (the mark(k, ...) calls above/below), by state_dict index.
extern "C" int vqa_grad_groups(int model, int* group_of_param, int n_params) {
  if (!group_of_param) return -1;
  if (model == 0) {
    if (n_params != 62) return -1;
    auto set = [&](int a, int b, int g) { for (int i = a; i < b; ++i) group_of_param[i] = g; };
    set(52, 54, 0); set(44, 52, 1); set(42, 44, 2); set(16, 24, 3); set(34, 42, 4); set(32, 34, 5); set(24, 32, 6);
    set(2, 4, 7); set(56, 58, 8); set(60, 62, 8); set(54, 56, 9); set(58, 60, 9); set(14, 16, 10); set(6, 14, 11);
    set(0, 2, 12); set(4, 6, 13);
    return 14;
  }
  if (model == 1) {
    if (n_params != 38) return -1;
    auto set = [&](int a, int b, int g) { for (int i = a; i < b; ++i) group_of_param[i] = g; };
    set(36, 38, 0); set(16, 36, 1); set(6, 14, 2); set(4, 6, 3); set(0, 2, 4); set(2, 4, 5); set(14, 16, 5);
    return 6;
  }
  return -1;
}

// ====================================================================================== ODA
extern "C" size_t vqa_oda_workspace_bytes(int64_t B, int64_t N, int64_t C) {
  return carve_oda(nullptr, B, N, C).bytes;
}

extern "C" int vqa_oda_fwd(const vqa_model_fwd_params* p, void* stream) {
  VQA_TRY(check_model(p, p ? carve_oda(nullptr, p->B, p->N, p->C).bytes : 0, "vqa_oda_fwd", false));
  using namespace oda;
  const int64_t B = p->B, N = p->N, M = B * N;
  OdaWs w = carve_oda(p->workspace, B, N, p->C);
  Ctx c{p, stream, p->params, nullptr, 0, w.lin_ws, w.lin_ws_bytes, p->math != VQA_MATH_FP32_SIMT, p->train ? w.bits : nullptr};
  const float* clp[1] = {w.clsp}; (void)clp;
  // Head of the step as in vqa_cor2_fwd: bits on the main lane, fork, weight packing + question projections on the
  // side lane (joined before the pairwise attention; the packed weights are first used after that join).
  Lanes* L = get_lanes();
  VQA_REQUIRE(L != nullptr, "vqa_oda_fwd: cannot create the internal side stream");
  cudaStream_t ms = (cudaStream_t)stream, ss = L->side;
  Ctx cs = c; cs.stream = ss; cs.lin_ws = w.side_ws; cs.lin_ws_bytes = w.side_ws_bytes;
  const bool b16 = is_bf16_math(p->math) && M >= TC16_MIN_M;       // bf16 modes: compress_v reads operand planes (see vqa_cor2_fwd)
  const int np = p->math == VQA_MATH_BF16X3 ? 2 : 1;
  LinExt x_cv;
  if (b16) { x_cv.Xp = Planes{w.vp, D, M * D}; x_cv.Wp = Planes{w.cv_wp, D, HP * D}; }
  if (p->train && !b16) {
    ProfScope ps_(stream, "dropout_bits");
    VQA_TRY(make_bits(p, w.bits, w.bits_n, stream, L_COMPRESS_V, true));
  }
  if (b16) {
    ProfScope ps_(stream, "planes.v");
    PackPlanesSeg sg = {};
    sg.src = c.W[COMPRESS_V]; sg.rows = H; sg.rows_pad = HP; sg.K = D; sg.Kp = D; sg.dst = w.cv_wp; sg.plane = HP * D;
    VQA_TRY(tc16::pack_planes(&sg, 1, np, ms));
    const float* X[1] = {p->v}; int64_t ldx[1] = {D}; uint32_t layer[1] = {L_COMPRESS_V}; uint64_t base[1] = {0};
    __nv_bfloat16* out[1] = {w.vp};
    VQA_TRY(tc16::split_planes(X, ldx, 1, M, D, c.pdrop(), p->seed, p->seed_dev, layer, base, nullptr, out, D, M * D, np, ms));
  }
  Lanes::wait(ss, L->record(ms));               // fork
  if (p->train) {
    ProfScope ps_(ss, "dropout_bits.rest");
    VQA_TRY(make_bits(p, w.bits, w.bits_n, ss, L_COMPRESS_V, false));
  }
  if (c.packed) {
    ProfScope ps_(ss, "pack_weights");
    PackList pl;
    pl.add_mutan(c.W, FF_L1, 5, w.ff_w1p, A); pl.add_mutan(c.W, FF_L2, 5, w.ff_w2p, H);
    pl.add(c.W[CLASSIF], w.clsp, p->C, p->C, F);
    VQA_TRY(vqa_pack_weights(pl.s, pl.n, ss));
  }
  {  // compress_v (config/ODA.py:211)
    const float* X[1] = {p->v}; int64_t ldx[1] = {D}; int widx[1] = {COMPRESS_V}; float* Y[1] = {w.vl};
    int64_t ldy[1] = {H}; uint32_t layer[1] = {L_COMPRESS_V};
    { ProfScope ps_(stream, "compress_v.fwd"); VQA_TRY(lin_fwd(c, 1, M, D, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, layer, nullptr, b16 ? &x_cv : nullptr)); }
  }
  {  // compress_q + linear_q (:214, :233)
    const float* X[2] = {p->q, p->q}; int64_t ldx[2] = {Q, Q}; int widx[2] = {COMPRESS_Q, LINEAR_Q};
    float* Y[2] = {w.ql, w.qf}; int64_t ldy[2] = {H, HP}; uint32_t layer[2] = {L_COMPRESS_Q, L_LINEAR_Q};
    { ProfScope ps_(ss, "q_proj2.fwd"); VQA_TRY(lin_fwd(cs, 2, B, Q, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, layer)); }
    Lanes::wait(ms, L->record(ss));             // join
  }
  {  // pairwise differences + conv_att + softmax + pooling (:216-226)
    vqa_oda_pair_attn_fwd_params ap = {};
    ap.B = B; ap.N = N; ap.H = H; ap.D = D; ap.train = p->train;
    ap.drop.p = c.pdrop(); ap.drop.layer = L_ATT_CONV; ap.drop.seed = p->seed; ap.drop.seed_dev = p->seed_dev;
    ap.vl = w.vl; ap.ql = w.ql; ap.W = c.W[ATT_CONV]; ap.bc = c.W[ATT_CONV + 1]; ap.x = p->v; ap.wsum = w.wsum;
    ap.alpha = p->alpha1; ap.pooled = w.pooled;
    ap.workspace = w.pair_ws; ap.workspace_bytes = w.pair_ws_bytes; ap.keep_bits_ready = 1;   // drawn by make_bits above
    { ProfScope ps_(stream, "oda_pair_attn.fwd"); VQA_TRY(vqa_oda_pair_attn_fwd(&ap, stream)); }
  }
  { ProfScope ps_(stream, "att.glimpse.fwd"); VQA_TRY(glimpse_fwd(c, B, w.pooled, ATT_G, w.vf, A, 0, L_ATT_G)); }
  { ProfScope ps_(stream, "fusion_final.fwd"); VQA_TRY(mutan_fwd(c, 5, B, A, H, 1, w.vf, A, w.qf, HP, FF_L1, FF_L2, w.ff_H1, w.ff_H2, w.xf, XP, w.ff_w1p, w.ff_w2p)); }       // fusion_final :236
  {  // linear_classif (:239)
    const float* X[1] = {w.xf}; int64_t ldx[1] = {XP}; int widx[1] = {CLASSIF}; float* Y[1] = {p->logits};
    int64_t ldy[1] = {p->C}; uint32_t layer[1] = {L_CLASSIF};
    { ProfScope ps_(stream, "classif.fwd"); VQA_TRY(lin_fwd(c, 1, B, F, p->C, VQA_ACT_NONE, X, ldx, widx, Y, ldy, layer, clp)); }
  }
  return VQA_OK;
}

extern "C" int vqa_oda_bwd(const vqa_model_bwd_params* bp, void* stream) {
  VQA_REQUIRE(bp != nullptr, "vqa_oda_bwd: null params");
  const vqa_model_fwd_params* p = &bp->fwd;
  VQA_TRY(check_model(p, carve_oda(nullptr, p->B, p->N, p->C).bytes, "vqa_oda_bwd", false));
  VQA_REQUIRE(bp->dlogits && bp->grads, "vqa_oda_bwd: null dlogits / grads");
  using namespace oda;
  const int64_t B = p->B, N = p->N, M = B * N;
  OdaWs w = carve_oda(p->workspace, B, N, p->C);
  int acc = bp->accumulate;
  if (!acc && bp->grads_flat && bp->grads_flat_bytes) {   // one zero-fill for every gradient tensor, then everything accumulates
    cudaMemsetAsync(bp->grads_flat, 0, bp->grads_flat_bytes, (cudaStream_t)stream);
    acc = 1;
  }
  Ctx c{p, stream, p->params, bp->grads, acc, w.lin_ws, w.lin_ws_bytes, p->math != VQA_MATH_FP32_SIMT, p->train ? w.bits : nullptr};
  const float* clp[1] = {w.clsp}; (void)clp;
  {
    const float* X[1] = {w.xf}; int64_t ldx[1] = {XP}; int widx[1] = {CLASSIF}; const float* Y[1] = {p->logits};
    int64_t ldy[1] = {p->C}; const float* dY[1] = {bp->dlogits}; int64_t lddy[1] = {p->C};
    float* dX[1] = {w.dxf}; int64_t lddx[1] = {XP}; uint32_t layer[1] = {L_CLASSIF};
    { ProfScope ps_(stream, "classif.bwd"); VQA_TRY(lin_bwd(c, 1, B, F, p->C, VQA_ACT_NONE, X, ldx, widx, Y, ldy, dY, lddy, dX, lddx, 0, layer, clp)); }
  }
  auto mark = [&](int group) {
    if (bp->group_events[group]) cudaEventRecord((cudaEvent_t)bp->group_events[group], (cudaStream_t)stream);
  };
  mark(0);
  { ProfScope ps_(stream, "fusion_final.bwd"); VQA_TRY(mutan_bwd(c, 5, B, A, H, 1, w.vf, A, w.qf, HP, FF_L1, FF_L2, w.ff_H1, w.ff_H2, w.dxf, XP, w.d_ff_H2, w.dvf, A, w.dqf, HP, 0, w.ff_w1p, w.ff_w2p)); }
  mark(1);
  { ProfScope ps_(stream, "att.glimpse.bwd"); VQA_TRY(glimpse_bwd(c, B, w.pooled, ATT_G, w.vf, w.dvf, A, 0, w.dpooled, L_ATT_G)); }
  mark(2);
  {
    vqa_oda_pair_attn_bwd_params ap = {};
    ap.B = B; ap.N = N; ap.H = H; ap.D = D; ap.train = p->train;
    ap.drop.p = c.pdrop(); ap.drop.layer = L_ATT_CONV; ap.drop.seed = p->seed; ap.drop.seed_dev = p->seed_dev;
    ap.accumulate_w = acc;
    ap.vl = w.vl; ap.ql = w.ql; ap.W = c.W[ATT_CONV]; ap.x = p->v; ap.alpha = p->alpha1; ap.wsum = w.wsum;
    ap.dpooled = w.dpooled; ap.dalpha = w.dalpha; ap.dz = w.dz; ap.dwsum = w.dwsum;
    ap.dW = c.grad(ATT_CONV); ap.dbc = c.grad(ATT_CONV + 1); ap.dvl = w.dvl; ap.dql = w.dql;
    ap.workspace = w.pair_ws; ap.workspace_bytes = w.pair_ws_bytes;
    { ProfScope ps_(stream, "oda_pair_attn.bwd"); VQA_TRY(vqa_oda_pair_attn_bwd(&ap, stream)); }
  }
  mark(3);
  {
    const float* X[1] = {p->v}; int64_t ldx[1] = {D}; int widx[1] = {COMPRESS_V}; const float* Y[1] = {w.vl};
    int64_t ldy[1] = {H}; const float* dY[1] = {w.dvl}; int64_t lddy[1] = {H}; uint32_t layer[1] = {L_COMPRESS_V};
    LinExt x_cv;
    const bool b16 = is_bf16_math(p->math) && M >= TC16_MIN_M;
    if (b16) x_cv.Xp = Planes{w.vp, D, M * D};
    { ProfScope ps_(stream, "compress_v.bwd"); VQA_TRY(lin_bwd(c, 1, M, D, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, nullptr, nullptr, 0, layer, nullptr, nullptr, nullptr, 0, b16 ? &x_cv : nullptr)); }
  }
  mark(4);
  {
    const float* X[2] = {p->q, p->q}; int64_t ldx[2] = {Q, Q}; int widx[2] = {COMPRESS_Q, LINEAR_Q};
    const float* Y[2] = {w.ql, w.qf}; int64_t ldy[2] = {H, HP}; const float* dY[2] = {w.dql, w.dqf};
    int64_t lddy[2] = {H, HP}; uint32_t layer[2] = {L_COMPRESS_Q, L_LINEAR_Q};
    float* dX[2] = {w.dq_parts, w.dq_parts + B * Q};
    int64_t lddx[2] = {Q, Q};
    { ProfScope ps_(stream, "q_proj2.bwd"); VQA_TRY(lin_bwd(c, 2, B, Q, H, VQA_ACT_RELU, X, ldx, widx, Y, ldy, dY, lddy, bp->dq ? dX : nullptr, bp->dq ? lddx : nullptr, 0, layer)); }
    if (bp->dq) VQA_TRY(sum_parts(B * Q, 2, w.dq_parts, bp->dq, (cudaStream_t)stream));
  }
  mark(5);
  return VQA_OK;
}
