// CoR2 compound objects (include/vqacore.h: vqa_cor_compound_{fwd,bwd}).
// Replaces decare_cat + the alpha-weighted sum over i (config/CoR2.py:191-199, :215-216):
//   v2[b,j,:] = vt[b,:]*g1[b,:] + s[b]*v[b,j,:]*g2[b,:],  vt = pooled[b,0,:], s = sum_i alpha[b,i,0].
// HBM-bound: forward reads x once and writes v2 once (2*N*D*4 bytes per sample).
#include <cuda_bf16.h>

#include "common.cuh"

namespace vqa {

constexpr int CMP_THREADS = 128;

__global__ void __launch_bounds__(CMP_THREADS)
cor_compound_fwd_kernel(int64_t N, int64_t D, const float* __restrict__ x, const float* __restrict__ pooled,
                        const float* __restrict__ alpha, const float* __restrict__ g1, const float* __restrict__ g2,
                        float* __restrict__ v2, __nv_bfloat16* __restrict__ planes, int np, int64_t plane_stride,
                        const uint8_t* __restrict__ keep_bits, float keep_scale) {
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
  __nv_bfloat16* pb = planes ? planes + b * N * D + c : nullptr;
  const uint8_t* kb = (planes && keep_bits) ? keep_bits + ((b * N * D + c) >> 3) : nullptr;   // rows D/8 bytes apart
  const uint32_t ksh = (uint32_t)(c & 4);
  // the operand planes of dropout(v2) for compress_v2's GEMMs: mask, split into bf16 hi (+ lo), 8-byte stores
  auto emit_planes = [&](int64_t j, const float4& o, uint32_t nb) {
    const float m[4] = {(nb & 1u) ? o.x * keep_scale : 0.0f, (nb & 2u) ? o.y * keep_scale : 0.0f,
                        (nb & 4u) ? o.z * keep_scale : 0.0f, (nb & 8u) ? o.w * keep_scale : 0.0f};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { h[e] = __float2bfloat16_rn(m[e]); l[e] = __float2bfloat16_rn(m[e] - __bfloat162float(h[e])); }
    uint2 hi, lo;
    hi.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hi.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lo.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lo.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(pb + j * D) = hi;
    if (np == 2) *reinterpret_cast<uint2*>(pb + plane_stride + j * D) = lo;
  };
  constexpr int UNROLL = 6;
  int64_t j = 0;
  for (; j + UNROLL <= N; j += UNROLL) {
    float4 xv[UNROLL];
    uint32_t nb[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) xv[u] = ld_stream4(xb + (j + u) * D);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) nb[u] = kb ? (uint32_t)__ldg(kb + (j + u) * (D >> 3)) >> ksh : 0xFu;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float4 o;
      o.x = fmaf(xv[u].x, a2.x, base.x); o.y = fmaf(xv[u].y, a2.y, base.y);
      o.z = fmaf(xv[u].z, a2.z, base.z); o.w = fmaf(xv[u].w, a2.w, base.w);
      *reinterpret_cast<float4*>(ob + (j + u) * D) = o;
      if (pb) emit_planes(j + u, o, nb[u]);
    }
  }
  for (; j < N; ++j) {
    const float4 xv = ld_stream4(xb + j * D);
    const uint32_t nb = kb ? (uint32_t)__ldg(kb + j * (D >> 3)) >> ksh : 0xFu;
    float4 o;
    o.x = fmaf(xv.x, a2.x, base.x); o.y = fmaf(xv.y, a2.y, base.y);
    o.z = fmaf(xv.z, a2.z, base.z); o.w = fmaf(xv.w, a2.w, base.w);
    *reinterpret_cast<float4*>(ob + j * D) = o;
    if (pb) emit_planes(j, o, nb);
  }
}

// One CTA per sample (deterministic ds reduction).
constexpr int CMPB_THREADS = 512;
// __launch_bounds__(.., 2): <= 64 registers, two CTAs per SM, so that the 256 samples of the benchmark batch are ONE wave
// (ncu r2: 126 registers, one CTA per SM, two waves, 37 % of the HBM rate).
__global__ void __launch_bounds__(CMPB_THREADS, 2)
cor_compound_bwd_kernel(int64_t N, int64_t D, const float* __restrict__ x, const float* __restrict__ pooled,
                        const float* __restrict__ alpha, const float* __restrict__ g1, const float* __restrict__ g2,
                        const float* __restrict__ dv2, float* __restrict__ dg1, float* __restrict__ dg2,
                        float* __restrict__ dpooled, float* __restrict__ dalpha0_ext,
                        const uint8_t* __restrict__ keep_bits, float keep_scale, const float* __restrict__ pool_alpha,
                        const float* __restrict__ pool_dp) {
  __shared__ float red[CMPB_THREADS / 32];
  extern __shared__ __align__(16) float cmpb_smem[];
  float* dp2_s = cmpb_smem;                         // [G][D] pool_dp of this sample (kept out of the register budget)
  float* al2_s = cmpb_smem + (pool_alpha ? G * D : 0);   // [N*G] pool_alpha of this sample (fused dv2 finish)
  const int64_t b = blockIdx.x;
  if (pool_alpha) {
    for (int64_t t = threadIdx.x; t < N * G; t += CMPB_THREADS) al2_s[t] = pool_alpha[b * N * G + t];
    for (int64_t t = (int64_t)threadIdx.x * 4; t < G * D; t += CMPB_THREADS * 4)
      *reinterpret_cast<float4*>(dp2_s + t) = __ldg(reinterpret_cast<const float4*>(pool_dp + b * G * D + t));
    __syncthreads();
  }
  float s = 0.0f;
  for (int64_t i = 0; i < N; ++i) s += __ldg(&alpha[(b * N + i) * G]);
  float ds = 0.0f;
  const int Ni = (int)N, Di = (int)D;                 // 32-bit row / column offsets below (one sample is N*D < 2^31 floats)
  for (int c = (int)threadIdx.x * 4; c < Di; c += CMPB_THREADS * 4) {
    float4 dbar = make_float4(0.f, 0.f, 0.f, 0.f), xd = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xb = x + b * N * D + c;
    const float* db = dv2 + b * N * D + c;
    const uint8_t* kb = keep_bits ? keep_bits + ((b * N * D + c) >> 3) : nullptr;      // D % 8 == 0: rows are D/8 bytes apart
    const uint32_t ksh = (uint32_t)(c & 4);
    const int Db = Di >> 3;
    constexpr int U = 3;                             // rows in flight per thread: every load of a batch before its first use
    for (int j0 = 0; j0 < Ni; j0 += U) {
      float4 xs[U], ds[U];
      uint32_t ns[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = j0 + u < Ni;
        xs[u] = in ? ld_stream4(xb + (j0 + u) * Di) : make_float4(0.f, 0.f, 0.f, 0.f);
        ds[u] = in ? ld_stream4(db + (j0 + u) * Di) : make_float4(0.f, 0.f, 0.f, 0.f);
        ns[u] = (in && kb) ? (uint32_t)__ldg(kb + (j0 + u) * Db) >> ksh : 0xFu;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
      const int j = j0 + u;
      if (j >= Ni) break;
      const float4 xv = xs[u];
      float4 dv = ds[u];
      if (kb) {                                     // the raw GEMM term gets compress_v2's input-dropout mask here
        const uint32_t nb = ns[u];
        dv.x = (nb & 1u) ? dv.x * keep_scale : 0.0f; dv.y = (nb & 2u) ? dv.y * keep_scale : 0.0f;
        dv.z = (nb & 4u) ? dv.z * keep_scale : 0.0f; dv.w = (nb & 8u) ? dv.w * keep_scale : 0.0f;
      }
      if (pool_alpha) {                             // + the gradient of att2's pooling over the same v2
        const float4 a4 = *reinterpret_cast<const float4*>(&al2_s[j * G]);
        const float av[G] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 dp = *reinterpret_cast<const float4*>(dp2_s + g * Di + c);
          dv.x = fmaf(av[g], dp.x, dv.x); dv.y = fmaf(av[g], dp.y, dv.y);
          dv.z = fmaf(av[g], dp.z, dv.z); dv.w = fmaf(av[g], dp.w, dv.w);
        }
      }
      dbar.x += dv.x; dbar.y += dv.y; dbar.z += dv.z; dbar.w += dv.w;
      xd.x = fmaf(xv.x, dv.x, xd.x); xd.y = fmaf(xv.y, dv.y, xd.y);
      xd.z = fmaf(xv.z, dv.z, xd.z); xd.w = fmaf(xv.w, dv.w, xd.w);
      }
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
  VQA_REQUIRE(!p->v2_planes || ((p->v2_nplanes == 1 || p->v2_nplanes == 2) && p->D % 8 == 0),
              "vqa_cor_compound_fwd: v2_planes needs v2_nplanes in {1, 2} and D %% 8 == 0");
  KProf kp_(stream, "cor_compound_fwd", "hbm", (8.0 + (p->v2_planes ? 2.0 * p->v2_nplanes : 0.0)) * (double)p->B * p->N * p->D);
  cor_compound_fwd_kernel<<<grid, CMP_THREADS, 0, (cudaStream_t)stream>>>(
      p->N, p->D, p->x, p->pooled, p->alpha, p->g1, p->g2, p->v2, reinterpret_cast<__nv_bfloat16*>(p->v2_planes),
      p->v2_nplanes, p->v2_plane_stride, p->v2_keep_bits, p->v2_keep_bits ? p->v2_keep_scale : 1.0f);
  return check_launch("cor_compound_fwd");
}

extern "C" int vqa_cor_compound_bwd(const vqa_cor_compound_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_cor_compound_bwd: null params");
  VQA_REQUIRE(p->B >= 0 && p->N >= 1 && p->D >= 4 && p->D % 4 == 0, "vqa_cor_compound_bwd: bad shape");
  VQA_REQUIRE(p->x && p->pooled && p->alpha && p->g1 && p->g2 && p->dv2 && p->dg1 && p->dg2 && p->dpooled &&
                  p->dalpha0_ext,
              "vqa_cor_compound_bwd: null pointer");
  VQA_REQUIRE(!p->dv2_pool_alpha || p->dv2_pool_dpooled, "vqa_cor_compound_bwd: dv2_pool_alpha needs dv2_pool_dpooled");
  VQA_REQUIRE(!p->dv2_keep_bits || p->D % 8 == 0, "vqa_cor_compound_bwd: dv2_keep_bits needs D %% 8 == 0");
  if (p->B == 0) return VQA_OK;
  KProf kp_(stream, "cor_compound_bwd", "hbm", 8.0 * (double)p->B * p->N * p->D);
  const size_t smem = p->dv2_pool_alpha ? (size_t)(p->N * G + G * p->D) * sizeof(float) : 0;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(cor_compound_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    set_error("vqa_cor_compound_bwd: cannot reserve %zu bytes of shared memory", smem);
    return VQA_ECUDA;
  }
  cor_compound_bwd_kernel<<<(unsigned)p->B, CMPB_THREADS, smem, (cudaStream_t)stream>>>(
      p->N, p->D, p->x, p->pooled, p->alpha, p->g1, p->g2, p->dv2, p->dg1, p->dg2, p->dpooled, p->dalpha0_ext,
      p->dv2_keep_bits, p->dv2_keep_bits ? p->dv2_keep_scale : 1.0f, p->dv2_pool_alpha, p->dv2_pool_dpooled);
  return check_launch("cor_compound_bwd");
}
