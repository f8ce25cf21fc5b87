// SkipThoughts question encoder pieces (include/vqacore.h: vqa_seq_dropout_masks, vqa_gru_*).
// Reference: putils/__init__.py — SkipThoughts.forward :975-982 (embedding -> BayesianGRU -> hidden state at the last
// non-PAD token), BayesianGRUCell.forward :622-637, SequentialDropout :503-539 (ONE Bernoulli mask per sequence and
// call site, shared by all time steps), BayesianGRU.forward :689-741.
// The dense contractions (input projections of all time steps at once, the three recurrent projections of a step,
// their dgrad / wgrad) are the grouped tensor-core linears of linear.cu / gemm_tc*.cu; this file holds what sits
// between them: the sequence-tied masks, the embedding gather / scatter, the fused gate math and its backward, and
// the last-token selection.
#include "common.cuh"

namespace vqa {

// out[m][b][f] = keep(seed, layer0 + m, b*dim + f) ? 1/(1-p) : 0     (p = 0: all ones)
__global__ void seq_masks_kernel(Drop d, uint32_t layer0, int64_t per_mask, int64_t total, float* __restrict__ out) {
  const uint64_t seed = d.key();
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 16 < total; g += (int64_t)gridDim.x * blockDim.x) {
    // 16 consecutive outputs; a run may straddle two masks (and per_mask need not be a multiple of 16)
    const int64_t e0 = g * 16;
    uint4 r = make_uint4(0, 0, 0, 0);
    int64_t have_m = -1, have_g = -1;
    for (int u = 0; u < 16 && e0 + u < total; ++u) {
      const int64_t e = e0 + u, m = e / per_mask, idx = e - m * per_mask;
      if (!d.on) { out[e] = 1.0f; continue; }
      if (m != have_m || (idx >> 4) != have_g) {
        r = philox_group(seed, layer0 + (uint32_t)m, (uint64_t)(idx >> 4));
        have_m = m; have_g = idx >> 4;
      }
      const uint32_t by = (pick_word(r, ((uint32_t)idx >> 2) & 3u) >> (8u * ((uint32_t)idx & 3u))) & 0xFFu;
      out[e] = by >= d.thr ? d.scale : 0.0f;
    }
  }
}

// X_g[(t*B + b)][f] = E[idx[b,t]][f] * mask_g[b][f],  g < 3   (time-major rows: a step's slice is contiguous)
__global__ void gru_embed_fwd_kernel(int64_t B, int64_t T, int64_t I, const int64_t* __restrict__ idx,
                                     const float* __restrict__ emb, const float* __restrict__ masks,
                                     float* __restrict__ out) {
  const int64_t row = blockIdx.x;                 // t*B + b
  const int64_t t = row / B, b = row - t * B;
  const float* e = emb + idx[b * T + t] * I;
  for (int64_t f = threadIdx.x; f < I; f += blockDim.x) {
    const float x = __ldg(e + f);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float m = masks ? __ldg(masks + ((int64_t)g * B + b) * I + f) : 1.0f;
      out[((int64_t)g * T * B + row) * I + f] = x * m;
    }
  }
}

// dE[idx[b,t]][f] += sum_g dX_g[(t*B+b)][f] * mask_g[b][f]   (nn.Embedding(padding_idx=0): row 0 receives nothing)
__global__ void gru_embed_bwd_kernel(int64_t B, int64_t T, int64_t I, const int64_t* __restrict__ idx,
                                     const float* __restrict__ masks, const float* __restrict__ dX,
                                     float* __restrict__ demb) {
  const int64_t row = blockIdx.x;
  const int64_t t = row / B, b = row - t * B;
  const int64_t w = idx[b * T + t];
  if (w == 0) return;
  for (int64_t f = threadIdx.x; f < I; f += blockDim.x) {
    float s = 0.0f;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float m = masks ? __ldg(masks + ((int64_t)g * B + b) * I + f) : 1.0f;
      s = fmaf(dX[((int64_t)g * T * B + row) * I + f], m, s);
    }
    atomicAdd(demb + w * I + f, s);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// One time step of BayesianGRUCell.forward (:630-636) after the six projections:
//   r = sigmoid(gi_r + gh_r), i = sigmoid(gi_i + gh_i), n = af(gi_n + r * gh_n), h' = (1 - i) n + i h
// and the three masked copies of h' that the next step's recurrent projections read (drop_hr / drop_hi / drop_hn).
// gh_* = NULL at the first step (h = 0, no hidden bias).  4 elements per thread.
__global__ void gru_gate_fwd_kernel(vqa_gru_gate_fwd_params p) {
  const int64_t total = p.B * p.H;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= total) return;
  auto ld = [&](const float* q) { return q ? *reinterpret_cast<const float4*>(q + e) : make_float4(0.f, 0.f, 0.f, 0.f); };
  const float4 gir = ld(p.gi[0]), gii = ld(p.gi[1]), gin = ld(p.gi[2]);
  const float4 ghr = ld(p.gh[0]), ghi = ld(p.gh[1]), ghn = ld(p.gh[2]);
  const float4 hp = ld(p.h_prev);
  const float a_r[4] = {gir.x + ghr.x, gir.y + ghr.y, gir.z + ghr.z, gir.w + ghr.w};
  const float a_i[4] = {gii.x + ghi.x, gii.y + ghi.y, gii.z + ghi.z, gii.w + ghi.w};
  const float g_n[4] = {gin.x, gin.y, gin.z, gin.w}, h_n[4] = {ghn.x, ghn.y, ghn.z, ghn.w};
  const float h0[4] = {hp.x, hp.y, hp.z, hp.w};
  float r[4], ii[4], n[4], h[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    r[u] = sigmoidf_(a_r[u]);
    ii[u] = sigmoidf_(a_i[u]);
    const float pre = fmaf(r[u], h_n[u], g_n[u]);
    n[u] = p.act == VQA_ACT_RELU ? fmaxf(pre, 0.0f) : tanhf(pre);
    h[u] = fmaf(ii[u], h0[u] - n[u], n[u]);        // (1 - i) n + i h
  }
  auto st = [&](float* q, const float (&v)[4]) { if (q) *reinterpret_cast<float4*>(q + e) = make_float4(v[0], v[1], v[2], v[3]); };
  st(p.r, r); st(p.i, ii); st(p.n, n); st(p.h, h);
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    if (!p.hm[g]) continue;
    float4 m = p.hmask[g] ? *reinterpret_cast<const float4*>(p.hmask[g] + e) : make_float4(1.f, 1.f, 1.f, 1.f);
    *reinterpret_cast<float4*>(p.hm[g] + e) = make_float4(h[0] * m.x, h[1] * m.y, h[2] * m.z, h[3] * m.w);
  }
}

// Backward of one step.  dh = dh_partial + sum_g dhm_g * hmask_g + (t is sample b's last token ? dx_last[b] : 0);
//   dn = dh (1 - i), di = dh (h_prev - n), dh_partial_out = dh i
//   da_n = dn af'(n), dr = da_n gh_n, dgh_n = da_n r, da_r = dr r (1 - r), da_i = di i (1 - i)
// da_* are the gradients of the three input projections AND of gh_r / gh_i; dgh_n that of gh_n.
__global__ void gru_gate_bwd_kernel(vqa_gru_gate_bwd_params p) {
  const int64_t total = p.B * p.H;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= total) return;
  const int64_t b = e / p.H;
  auto ld = [&](const float* q) { return q ? *reinterpret_cast<const float4*>(q + e) : make_float4(0.f, 0.f, 0.f, 0.f); };
  float4 dh = ld(p.dh_partial);
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    if (!p.dhm[g]) continue;
    const float4 d = ld(p.dhm[g]);
    const float4 m = p.hmask[g] ? *reinterpret_cast<const float4*>(p.hmask[g] + e) : make_float4(1.f, 1.f, 1.f, 1.f);
    dh.x = fmaf(d.x, m.x, dh.x); dh.y = fmaf(d.y, m.y, dh.y); dh.z = fmaf(d.z, m.z, dh.z); dh.w = fmaf(d.w, m.w, dh.w);
  }
  if (p.dx_last && p.last_pos[b] == p.t) {
    const float4 d = ld(p.dx_last);
    dh.x += d.x; dh.y += d.y; dh.z += d.z; dh.w += d.w;
  }
  const float4 r4 = ld(p.r), i4 = ld(p.i), n4 = ld(p.n), g4 = ld(p.gh_n), h4 = ld(p.h_prev);
  const float dhv[4] = {dh.x, dh.y, dh.z, dh.w}, r[4] = {r4.x, r4.y, r4.z, r4.w}, ii[4] = {i4.x, i4.y, i4.z, i4.w};
  const float n[4] = {n4.x, n4.y, n4.z, n4.w}, ghn[4] = {g4.x, g4.y, g4.z, g4.w}, h0[4] = {h4.x, h4.y, h4.z, h4.w};
  float dar[4], dai[4], dan[4], dghn[4], dhp[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float dn = dhv[u] * (1.0f - ii[u]);
    const float di = dhv[u] * (h0[u] - n[u]);
    dhp[u] = dhv[u] * ii[u];
    dan[u] = dn * (p.act == VQA_ACT_RELU ? (n[u] > 0.0f ? 1.0f : 0.0f) : (1.0f - n[u] * n[u]));
    const float dr = dan[u] * ghn[u];
    dghn[u] = dan[u] * r[u];
    dar[u] = dr * r[u] * (1.0f - r[u]);
    dai[u] = di * ii[u] * (1.0f - ii[u]);
  }
  auto st = [&](float* q, const float (&v)[4]) { if (q) *reinterpret_cast<float4*>(q + e) = make_float4(v[0], v[1], v[2], v[3]); };
  st(p.da[0], dar); st(p.da[1], dai); st(p.da[2], dan); st(p.dgh_n, dghn); st(p.dh_partial_out, dhp);
}

// last_pos[b] = (#non-PAD tokens of question b) - 1, wrapped like the reference's mask[i][lengths[i] - 1] (-1 -> T - 1)
__global__ void gru_last_pos_kernel(int64_t B, int64_t T, const int64_t* __restrict__ idx, int64_t* __restrict__ last_pos) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t len = 0;
  for (int64_t t = 0; t < T; ++t) len += idx[b * T + t] != 0;
  last_pos[b] = len > 0 ? len - 1 : T - 1;
}

// out[b][:] = hs[last_pos[b]][b][:]
__global__ void gru_select_last_kernel(int64_t B, int64_t H, const float* __restrict__ hs,
                                       const int64_t* __restrict__ last_pos, float* __restrict__ out) {
  const int64_t b = blockIdx.x;
  const float* src = hs + (last_pos[b] * B + b) * H;
  for (int64_t f = threadIdx.x * 4; f < H; f += blockDim.x * 4)
    *reinterpret_cast<float4*>(out + b * H + f) = *reinterpret_cast<const float4*>(src + f);
}

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_seq_dropout_masks(float p, uint64_t seed, const uint64_t* seed_dev, uint32_t layer0, int64_t B,
                                     int64_t dim, int nmasks, float* out, void* stream) {
  VQA_REQUIRE(B >= 0 && dim >= 1 && nmasks >= 1 && out, "vqa_seq_dropout_masks: bad arguments");
  VQA_REQUIRE(p >= 0.0f && p < 1.0f, "vqa_seq_dropout_masks: p = %f out of [0,1)", (double)p);
  if (B == 0) return VQA_OK;
  const int64_t per = B * dim, total = per * nmasks;
  Drop d = make_drop(p, seed, layer0, 0, 1, seed_dev);
  const int64_t groups = cdiv(total, 16);
  seq_masks_kernel<<<(unsigned)(cdiv(groups, 256) > 4096 ? 4096 : cdiv(groups, 256)), 256, 0, (cudaStream_t)stream>>>(
      d, layer0, per, total, out);
  return check_launch("seq_masks");
}

extern "C" int vqa_gru_embed_fwd(int64_t B, int64_t T, int64_t I, const int64_t* idx, const float* emb,
                                 const float* masks, float* out, void* stream) {
  VQA_REQUIRE(B >= 0 && T >= 1 && I >= 1 && idx && emb && out, "vqa_gru_embed_fwd: bad arguments");
  if (B == 0) return VQA_OK;
  gru_embed_fwd_kernel<<<(unsigned)(B * T), 160, 0, (cudaStream_t)stream>>>(B, T, I, idx, emb, masks, out);
  return check_launch("gru_embed_fwd");
}

extern "C" int vqa_gru_embed_bwd(int64_t B, int64_t T, int64_t I, const int64_t* idx, const float* masks,
                                 const float* dX, float* demb, void* stream) {
  VQA_REQUIRE(B >= 0 && T >= 1 && I >= 1 && idx && dX && demb, "vqa_gru_embed_bwd: bad arguments");
  if (B == 0) return VQA_OK;
  gru_embed_bwd_kernel<<<(unsigned)(B * T), 160, 0, (cudaStream_t)stream>>>(B, T, I, idx, masks, dX, demb);
  return check_launch("gru_embed_bwd");
}

static int gate_shape_ok(int64_t B, int64_t H, int act, const char* who) {
  VQA_REQUIRE(B >= 0 && H >= 4 && H % 4 == 0, "%s: bad shape B=%lld H=%lld (H must be a multiple of 4)", who, (long long)B,
              (long long)H);
  VQA_REQUIRE(act == VQA_ACT_RELU || act == VQA_ACT_TANH, "%s: activation must be relu or tanh", who);
  return VQA_OK;
}

extern "C" int vqa_gru_gate_fwd(const vqa_gru_gate_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_gru_gate_fwd: null params");
  VQA_TRY(gate_shape_ok(p->B, p->H, p->act, "vqa_gru_gate_fwd"));
  VQA_REQUIRE(p->gi[0] && p->gi[1] && p->gi[2] && p->h, "vqa_gru_gate_fwd: null pointer");
  VQA_REQUIRE((p->gh[0] != nullptr) == (p->h_prev != nullptr), "vqa_gru_gate_fwd: gh and h_prev come together");
  if (p->B == 0) return VQA_OK;
  gru_gate_fwd_kernel<<<(unsigned)cdiv(p->B * p->H / 4, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("gru_gate_fwd");
}

extern "C" int vqa_gru_gate_bwd(const vqa_gru_gate_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_gru_gate_bwd: null params");
  VQA_TRY(gate_shape_ok(p->B, p->H, p->act, "vqa_gru_gate_bwd"));
  VQA_REQUIRE(p->r && p->i && p->n && p->da[0] && p->da[1] && p->da[2], "vqa_gru_gate_bwd: null pointer");
  VQA_REQUIRE(!p->dx_last || p->last_pos, "vqa_gru_gate_bwd: dx_last needs last_pos");
  if (p->B == 0) return VQA_OK;
  gru_gate_bwd_kernel<<<(unsigned)cdiv(p->B * p->H / 4, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("gru_gate_bwd");
}

extern "C" int vqa_gru_last_pos(int64_t B, int64_t T, const int64_t* idx, int64_t* last_pos, void* stream) {
  VQA_REQUIRE(B >= 0 && T >= 1 && idx && last_pos, "vqa_gru_last_pos: bad arguments");
  if (B == 0) return VQA_OK;
  gru_last_pos_kernel<<<(unsigned)cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(B, T, idx, last_pos);
  return check_launch("gru_last_pos");
}

extern "C" int vqa_gru_select_last(int64_t B, int64_t H, const float* hs, const int64_t* last_pos, float* out,
                                   void* stream) {
  VQA_REQUIRE(B >= 0 && H >= 4 && H % 4 == 0 && hs && last_pos && out, "vqa_gru_select_last: bad arguments");
  if (B == 0) return VQA_OK;
  gru_select_last_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(B, H, hs, last_pos, out);
  return check_launch("gru_select_last");
}
