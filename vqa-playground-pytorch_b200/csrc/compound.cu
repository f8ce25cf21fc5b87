// CoR2 compound objects (include/vqacore.h: vqa_cor_compound_{fwd,bwd}).
// Replaces decare_cat + the alpha-weighted sum over i (config/CoR2.py:191-199, :215-216):
//   v2[b,j,:] = vt[b,:]*g1[b,:] + s[b]*v[b,j,:]*g2[b,:],  vt = pooled[b,0,:], s = sum_i alpha[b,i,0].
// HBM-bound: forward reads x once and writes v2 once (2*N*D*4 bytes per sample).
#include "common.cuh"

namespace vqa {

constexpr int CMP_THREADS = 128;

__global__ void __launch_bounds__(CMP_THREADS)
cor_compound_fwd_kernel(int64_t N, int64_t D, const float* __restrict__ x, const float* __restrict__ pooled,
                        const float* __restrict__ alpha, const float* __restrict__ g1, const float* __restrict__ g2,
                        float* __restrict__ v2) {
  const int64_t b = blockIdx.y;
  const int64_t c = ((int64_t)blockIdx.x * CMP_THREADS + threadIdx.x) * 4;
  if (c >= D) return;
  float s = 0.0f;
  for (int64_t i = 0; i < N; ++i) s += __ldg(&alpha[(b * N + i) * G]);
  const float4 vt = *reinterpret_cast<const float4*>(&pooled[b * G * D + c]);
  const float4 a1 = *reinterpret_cast<const float4*>(&g1[b * D + c]);
  float4 a2 = *reinterpret_cast<const float4*>(&g2[b * D + c]);
  const float4 base = make_float4(vt.x * a1.x, vt.y * a1.y, vt.z * a1.z, vt.w * a1.w);
  a2.x *= s; a2.y *= s; a2.z *= s; a2.w *= s;
  const float* xb = x + b * N * D + c;
  float* ob = v2 + b * N * D + c;
  constexpr int UNROLL = 6;
  int64_t j = 0;
  for (; j + UNROLL <= N; j += UNROLL) {
    float4 xv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) xv[u] = ld_stream4(xb + (j + u) * D);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float4 o;
      o.x = fmaf(xv[u].x, a2.x, base.x); o.y = fmaf(xv[u].y, a2.y, base.y);
      o.z = fmaf(xv[u].z, a2.z, base.z); o.w = fmaf(xv[u].w, a2.w, base.w);
      *reinterpret_cast<float4*>(ob + (j + u) * D) = o;
    }
  }
  for (; j < N; ++j) {
    const float4 xv = ld_stream4(xb + j * D);
    float4 o;
    o.x = fmaf(xv.x, a2.x, base.x); o.y = fmaf(xv.y, a2.y, base.y);
    o.z = fmaf(xv.z, a2.z, base.z); o.w = fmaf(xv.w, a2.w, base.w);
    *reinterpret_cast<float4*>(ob + j * D) = o;
  }
}

// One CTA per sample (deterministic ds reduction).
constexpr int CMPB_THREADS = 512;
__global__ void __launch_bounds__(CMPB_THREADS)
cor_compound_bwd_kernel(int64_t N, int64_t D, const float* __restrict__ x, const float* __restrict__ pooled,
                        const float* __restrict__ alpha, const float* __restrict__ g1, const float* __restrict__ g2,
                        const float* __restrict__ dv2, float* __restrict__ dg1, float* __restrict__ dg2,
                        float* __restrict__ dpooled, float* __restrict__ dalpha0_ext) {
  __shared__ float red[CMPB_THREADS / 32];
  const int64_t b = blockIdx.x;
  float s = 0.0f;
  for (int64_t i = 0; i < N; ++i) s += __ldg(&alpha[(b * N + i) * G]);
  float ds = 0.0f;
  for (int64_t c = (int64_t)threadIdx.x * 4; c < D; c += CMPB_THREADS * 4) {
    float4 dbar = make_float4(0.f, 0.f, 0.f, 0.f), xd = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xb = x + b * N * D + c;
    const float* db = dv2 + b * N * D + c;
#pragma unroll 4
    for (int64_t j = 0; j < N; ++j) {
      const float4 xv = ld_stream4(xb + j * D);
      const float4 dv = ld_stream4(db + j * D);
      dbar.x += dv.x; dbar.y += dv.y; dbar.z += dv.z; dbar.w += dv.w;
      xd.x = fmaf(xv.x, dv.x, xd.x); xd.y = fmaf(xv.y, dv.y, xd.y);
      xd.z = fmaf(xv.z, dv.z, xd.z); xd.w = fmaf(xv.w, dv.w, xd.w);
    }
    const float4 vt = *reinterpret_cast<const float4*>(&pooled[b * G * D + c]);
    const float4 a1 = *reinterpret_cast<const float4*>(&g1[b * D + c]);
    const float4 a2 = *reinterpret_cast<const float4*>(&g2[b * D + c]);
    *reinterpret_cast<float4*>(&dg1[b * D + c]) = make_float4(vt.x * dbar.x, vt.y * dbar.y, vt.z * dbar.z, vt.w * dbar.w);
    *reinterpret_cast<float4*>(&dg2[b * D + c]) = make_float4(s * xd.x, s * xd.y, s * xd.z, s * xd.w);
    float4* dp = reinterpret_cast<float4*>(&dpooled[b * G * D + c]);
    float4 o = *dp;
    o.x = fmaf(a1.x, dbar.x, o.x); o.y = fmaf(a1.y, dbar.y, o.y);
    o.z = fmaf(a1.z, dbar.z, o.z); o.w = fmaf(a1.w, dbar.w, o.w);
    *dp = o;
    ds += a2.x * xd.x + a2.y * xd.y + a2.z * xd.z + a2.w * xd.w;
  }
  ds = warp_sum(ds);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < CMPB_THREADS / 32 ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) dalpha0_ext[b] = v;
  }
}

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_cor_compound_fwd(const vqa_cor_compound_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_cor_compound_fwd: null params");
  VQA_REQUIRE(p->B >= 0 && p->N >= 1 && p->D >= 4 && p->D % 4 == 0, "vqa_cor_compound_fwd: bad shape");
  VQA_REQUIRE(p->x && p->pooled && p->alpha && p->g1 && p->g2 && p->v2, "vqa_cor_compound_fwd: null pointer");
  if (p->B == 0) return VQA_OK;
  dim3 grid((unsigned)cdiv(p->D / 4, CMP_THREADS), (unsigned)p->B);
  KProf kp_(stream, "cor_compound_fwd", "hbm", 8.0 * (double)p->B * p->N * p->D);
  cor_compound_fwd_kernel<<<grid, CMP_THREADS, 0, (cudaStream_t)stream>>>(p->N, p->D, p->x, p->pooled, p->alpha, p->g1,
                                                                          p->g2, p->v2);
  return check_launch("cor_compound_fwd");
}

extern "C" int vqa_cor_compound_bwd(const vqa_cor_compound_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_cor_compound_bwd: null params");
  VQA_REQUIRE(p->B >= 0 && p->N >= 1 && p->D >= 4 && p->D % 4 == 0, "vqa_cor_compound_bwd: bad shape");
  VQA_REQUIRE(p->x && p->pooled && p->alpha && p->g1 && p->g2 && p->dv2 && p->dg1 && p->dg2 && p->dpooled &&
                  p->dalpha0_ext,
              "vqa_cor_compound_bwd: null pointer");
  if (p->B == 0) return VQA_OK;
  KProf kp_(stream, "cor_compound_bwd", "hbm", 8.0 * (double)p->B * p->N * p->D);
  cor_compound_bwd_kernel<<<(unsigned)p->B, CMPB_THREADS, 0, (cudaStream_t)stream>>>(
      p->N, p->D, p->x, p->pooled, p->alpha, p->g1, p->g2, p->dv2, p->dg1, p->dg2, p->dpooled, p->dalpha0_ext);
  return check_launch("cor_compound_bwd");
}
