// vqa_region_softmax_pool_{fwd,bwd}: the MyATT core (config/CoR2.py:137-154 == config/ODA.py:154-171).
#include "attention.cuh"

namespace vqa {

__global__ void softmax_regions_kernel(int64_t N, float* __restrict__ alpha) {
  extern __shared__ float z_s[];
  const int64_t b = blockIdx.x;
  for (int64_t t = threadIdx.x; t < N * G; t += blockDim.x) z_s[t] = alpha[b * N * G + t];
  __syncthreads();
  softmax_regions_smem(z_s, (int)N);
  __syncthreads();
  for (int64_t t = threadIdx.x; t < N * G; t += blockDim.x) alpha[b * N * G + t] = z_s[t];
}

// pooled[b,g,c] = sum_i alpha[b,i,g]*x[b,i,c].  HBM-bound: x is read exactly once, 128-bit loads,
// every thread keeps UNROLL independent loads in flight.  grid = (cdiv(D/4,128), B).
constexpr int POOL_THREADS = 128;
__global__ void __launch_bounds__(POOL_THREADS)
pool_fwd_kernel(int64_t N, int64_t D, const float* __restrict__ x, const float* __restrict__ alpha,
                float* __restrict__ pooled) {
  extern __shared__ float al_s[];   // [N*G]
  const int64_t b = blockIdx.y;
  for (int64_t t = threadIdx.x; t < N * G; t += POOL_THREADS) al_s[t] = alpha[b * N * G + t];
  __syncthreads();
  const int64_t c = ((int64_t)blockIdx.x * POOL_THREADS + threadIdx.x) * 4;
  if (c >= D) return;
  float4 acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* xb = x + b * N * D + c;
  constexpr int UNROLL = 6;
  int64_t i = 0;
  for (; i + UNROLL <= N; i += UNROLL) {
    float4 xv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) xv[u] = ld_stream4(xb + (i + u) * D);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const float4 a = *reinterpret_cast<const float4*>(&al_s[(i + u) * G]);
      const float av[G] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int g = 0; g < G; ++g) {
        acc[g].x = fmaf(av[g], xv[u].x, acc[g].x);
        acc[g].y = fmaf(av[g], xv[u].y, acc[g].y);
        acc[g].z = fmaf(av[g], xv[u].z, acc[g].z);
        acc[g].w = fmaf(av[g], xv[u].w, acc[g].w);
      }
    }
  }
  for (; i < N; ++i) {
    const float4 xv = ld_stream4(xb + i * D);
    const float4 a = *reinterpret_cast<const float4*>(&al_s[i * G]);
    const float av[G] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int g = 0; g < G; ++g) {
      acc[g].x = fmaf(av[g], xv.x, acc[g].x);
      acc[g].y = fmaf(av[g], xv.y, acc[g].y);
      acc[g].z = fmaf(av[g], xv.z, acc[g].z);
      acc[g].w = fmaf(av[g], xv.w, acc[g].w);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) *reinterpret_cast<float4*>(&pooled[(b * G + g) * D + c]) = acc[g];
}

int launch_pool_fwd(int64_t B, int64_t N, int64_t D, const float* x, const float* alpha, float* pooled,
                    cudaStream_t st) {
  dim3 grid((unsigned)cdiv(D / 4, POOL_THREADS), (unsigned)B);
  KProf kp_(st, "pool_fwd", "hbm", 4.0 * ((double)B * N * D + (double)B * G * D + (double)B * N * G));
  pool_fwd_kernel<<<grid, POOL_THREADS, (size_t)N * G * sizeof(float), st>>>(N, D, x, alpha, pooled);
  return check_launch("pool_fwd");
}

// dalpha[b,i,g] = <dpooled[b,g,:], x[b,i,:]> (+ ext for g=0);  dx[b,i,:] (+)= sum_g alpha[b,i,g] dpooled[b,g,:].
// The second (and last) pass over x in the backward.  One warp owns RW consecutive regions of one sample and walks the
// feature axis once: every dpooled vector it fetches (L1/L2-resident, 4*D floats per sample) is used for RW rows, so the
// streaming read of x, not the re-read of dpooled, sets the pace.  grid = (cdiv(N, warps*RW), B); the launcher picks
// the warp count so that one CTA covers a whole sample when N <= 48 (no nearly-empty tail CTA).
constexpr int POOL_BWD_RW = 4;
constexpr int POOL_BWD_MAX_WARPS = 12;
// The sample's dpooled [G][D] (32 KB at D = 2048) is staged in shared memory once per CTA: holding the 8 float4 of it a
// lane needs per iteration in registers made the kernel a 128-register one (one 9-warp CTA per SM, two waves at batch
// 256, 38 % of the HBM rate in ncu r2); from shared memory it is ~70 registers and three CTAs per SM.
template <bool DX>      // DX = false: only dalpha is produced (x is a graph input, or its pooling gradient is added elsewhere)
__global__ void __launch_bounds__(POOL_BWD_MAX_WARPS * 32, DX ? 2 : 3)
pool_bwd_kernel(int64_t N, int64_t D, const float* __restrict__ x, const float* __restrict__ alpha,
                const float* __restrict__ dpooled, const float* __restrict__ dalpha0_ext,
                const float* __restrict__ dalpha_ext, float* __restrict__ dalpha, float* __restrict__ dx,
                int accumulate_x) {
  constexpr int RW = POOL_BWD_RW;
  extern __shared__ __align__(16) float dp_s[];       // [G][D]
  const int64_t b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t t = (int64_t)threadIdx.x * 4; t < G * D; t += (int64_t)blockDim.x * 4)
    *reinterpret_cast<float4*>(dp_s + t) = __ldg(reinterpret_cast<const float4*>(dpooled + b * G * D + t));
  __syncthreads();
  const int64_t i0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * RW;
  if (i0 >= N) return;
  const int nr = N - i0 < RW ? (int)(N - i0) : RW;
  float a[RW][G];
  float dot[RW][G];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      a[r][g] = (DX && r < nr) ? alpha[(b * N + i0 + r) * G + g] : 0.0f;
      dot[r][g] = 0.0f;
    }
  const float* xr = x + (b * N + i0) * D;
  const int Di = (int)D;
  constexpr int CU = 1;                              // column chunks in flight per lane (registers: 3 CTAs per SM instead)
  for (int c0 = lane * 4; c0 < Di; c0 += 128 * CU) {
    float4 xv[CU][RW];
#pragma unroll
    for (int u = 0; u < CU; ++u) {
      const int c = c0 + u * 128;
#pragma unroll
      for (int r = 0; r < RW; ++r)
        xv[u][r] = (r < nr && c < Di) ? ld_stream4(xr + r * Di + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < CU; ++u) {
      const int c = c0 + u * 128;
      float4 d[G];
#pragma unroll
      for (int g = 0; g < G; ++g)
        d[g] = c < Di ? *reinterpret_cast<const float4*>(dp_s + g * Di + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 dg = d[g], xr4 = xv[u][r];
          dot[r][g] = fmaf(xr4.x, dg.x, fmaf(xr4.y, dg.y, fmaf(xr4.z, dg.z, fmaf(xr4.w, dg.w, dot[r][g]))));
          if (DX) {
            o.x = fmaf(a[r][g], dg.x, o.x); o.y = fmaf(a[r][g], dg.y, o.y);
            o.z = fmaf(a[r][g], dg.z, o.z); o.w = fmaf(a[r][g], dg.w, o.w);
          }
        }
        if (DX && dx && r < nr && c < Di) {
          float4* dst = reinterpret_cast<float4*>(dx + (b * N + i0 + r) * D + c);
          if (accumulate_x) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int g = 0; g < G; ++g) dot[r][g] = warp_sum(dot[r][g]);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      if (r >= nr) break;
      if (dalpha0_ext) dot[r][0] += dalpha0_ext[b];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int64_t o = (b * N + i0 + r) * G + g;
        dalpha[o] = dalpha_ext ? dot[r][g] + dalpha_ext[o] : dot[r][g];
      }
    }
  }
}

int launch_pool_bwd(int64_t B, int64_t N, int64_t D, const float* x, const float* alpha, const float* dpooled,
                    const float* dalpha0_ext, float* dalpha, float* dx, int accumulate_x, cudaStream_t st,
                    const float* dalpha_ext) {
  int64_t warps = cdiv(N, POOL_BWD_RW);
  if (warps > POOL_BWD_MAX_WARPS) warps = 8;
  dim3 grid((unsigned)cdiv(N, warps * POOL_BWD_RW), (unsigned)B);
  KProf kp_(st, "pool_bwd", "hbm", 4.0 * ((double)B * N * D * (dx ? 2 : 1) + (double)B * G * D));
  const size_t smem = (size_t)G * D * sizeof(float);
  if (smem > 200 * 1024) { set_error("pool_bwd: D=%lld too large for the shared dpooled stage", (long long)D); return VQA_EINVAL; }
  auto kern = dx ? pool_bwd_kernel<true> : pool_bwd_kernel<false>;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    set_error("pool_bwd: cannot reserve %zu bytes of shared memory", smem);
    return VQA_ECUDA;
  }
  kern<<<grid, (unsigned)(warps * 32), smem, st>>>(N, D, x, alpha, dpooled, dalpha0_ext, dalpha_ext, dalpha, dx, accumulate_x);
  return check_launch("pool_bwd");
}

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_region_softmax_pool_fwd(const vqa_region_softmax_pool_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_region_softmax_pool_fwd: null params");
  VQA_REQUIRE(p->B >= 0 && p->N >= 1 && p->Ff >= 1 && p->D >= 4 && p->D % 4 == 0,
              "vqa_region_softmax_pool_fwd: bad shape B=%lld N=%lld Ff=%lld D=%lld (D must be a multiple of 4)",
              (long long)p->B, (long long)p->N, (long long)p->Ff, (long long)p->D);
  VQA_REQUIRE(p->fuse && p->Wc && p->bc && p->x && p->alpha && p->pooled, "vqa_region_softmax_pool_fwd: null pointer");
  VQA_REQUIRE(p->drop.p >= 0.0f && p->drop.p < 1.0f, "vqa_region_softmax_pool_fwd: dropout p");
  const size_t smem = (size_t)(G * p->Ff + p->N * G) * sizeof(float);
  VQA_REQUIRE(smem <= 200 * 1024, "vqa_region_softmax_pool_fwd: Ff=%lld too large for shared memory", (long long)p->Ff);
  if (p->B == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FuseGeneric fs{p->fuse, p->N, p->Ff, make_drop(p->drop.p, p->drop.seed, p->drop.layer, 0, 1, p->drop.seed_dev),
                 p->drop.p > 0.0f ? p->drop_bits : nullptr};
  auto kern = att_logits_softmax_kernel<FuseGeneric>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  {
    KProf kp_(st, "att_logits_softmax", "hbm", 4.0 * ((double)p->B * p->N * p->Ff + (double)p->B * p->N * G));
    const unsigned S = p->N >= 16 ? 4 : 1;
    kern<<<dim3((unsigned)p->B, S), ATT_THREADS, smem, st>>>(fs, p->N, p->Ff, p->Wc, p->bc, p->alpha);
    VQA_TRY(check_launch("att_logits_softmax"));
    if (S > 1) {
      softmax_regions_kernel<<<(unsigned)p->B, 128, (size_t)p->N * G * sizeof(float), st>>>(p->N, p->alpha);
      VQA_TRY(check_launch("softmax_regions"));
    }
  }
  return launch_pool_fwd(p->B, p->N, p->D, p->x, p->alpha, p->pooled, st);
}

extern "C" int vqa_region_softmax_pool_bwd(const vqa_region_softmax_pool_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_region_softmax_pool_bwd: null params");
  VQA_REQUIRE(p->B >= 0 && p->N >= 1 && p->Ff >= 1 && p->D >= 4 && p->D % 4 == 0,
              "vqa_region_softmax_pool_bwd: bad shape");
  VQA_REQUIRE(p->fuse && p->Wc && p->x && p->alpha && p->dpooled && p->dalpha && p->dz && p->dWc,
              "vqa_region_softmax_pool_bwd: null pointer");
  if (p->B == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VQA_TRY(launch_pool_bwd(p->B, p->N, p->D, p->x, p->alpha, p->dpooled, p->dalpha0_ext, p->dalpha, p->dx,
                          p->accumulate_x, st, p->dalpha_ext));
  if (!p->accumulate_w) {
    cudaMemsetAsync(p->dWc, 0, (size_t)G * p->Ff * sizeof(float), st);
    if (p->dbc) cudaMemsetAsync(p->dbc, 0, G * sizeof(float), st);
  }
  FuseGeneric fs{p->fuse, p->N, p->Ff, make_drop(p->drop.p, p->drop.seed, p->drop.layer, 0, 1, p->drop.seed_dev),
                 p->drop.p > 0.0f ? p->drop_bits : nullptr};
  const int64_t groups = p->B < 2 * (int64_t)sm_count() ? p->B : 2 * (int64_t)sm_count();
  dim3 grid((unsigned)groups, (unsigned)cdiv(p->Ff, ATT_THREADS));
  const size_t smem = (size_t)2 * p->N * G * sizeof(float);
  KProf kp_(st, "att_logits_softmax_bwd", "hbm", 4.0 * (double)p->B * p->N * p->Ff * (p->dfuse ? 2 : 1));
  att_logits_softmax_bwd_kernel<FuseGeneric, false><<<grid, ATT_THREADS, smem, st>>>(
      fs, p->B, p->N, p->Ff, p->Wc, p->alpha, p->dalpha, p->dz, p->dWc, p->dbc, p->dfuse, nullptr);
  return check_launch("att_logits_softmax_bwd");
}
