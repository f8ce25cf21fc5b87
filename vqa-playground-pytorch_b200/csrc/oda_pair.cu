// ODA object-difference attention (include/vqacore.h: vqa_oda_pair_attn_{fwd,bwd}).
// Replaces the 36x36 Python pair loop, the stack/transpose/contiguous copy and the 11160-channel
// conv_att of the reference (config/ODA.py:216-226, :192).  The [B,N,N*H] tensor never exists:
//   eval : z[b,i,g] = sum_k ql[b,k] vl[b,i,k] Wsum[g,k] (+ terms constant in i, which the region
//          softmax cancels) — attention.cuh kernels with the FuseOdaEval source;
//   train: every (i,j,k) term is formed in registers with its Philox keep-bit.
#include "attention.cuh"

namespace vqa {

// wsum[g,k] = sum_j W[g, j*H + k].  grid = G
__global__ void oda_wsum_kernel(int64_t N, int64_t H, const float* __restrict__ W, float* __restrict__ wsum) {
  const int g = blockIdx.x;
  for (int64_t k = threadIdx.x; k < H; k += blockDim.x) {
    float s = 0.0f;
    for (int64_t j = 0; j < N; ++j) s += W[(g * N + j) * H + k];
    wsum[g * H + k] = s;
  }
}

// dW[g, j*H+k] (+)= dwsum[g,k] for every j (eval-mode gradient of the factorised form).
__global__ void oda_dw_broadcast_kernel(int64_t N, int64_t H, const float* __restrict__ dwsum, float* __restrict__ dW,
                                        int accumulate) {
  const int64_t total = (int64_t)G * N * H;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t % H, g = t / (N * H);
    const float v = dwsum[g * H + k];
    dW[t] = accumulate ? dW[t] + v : v;
  }
}

// ---- train mode ------------------------------------------------------------------------------
// z[b,i,g] = bc[g] + scale * sum_{e=(j,k)} keep(b,i,e) W[g,e] (vl[b,i,k]-vl[b,j,k]) ql[b,k].
// grid = (N, B); each thread walks aligned quads of the flat (j,k) row so one Philox call
// yields four keep-bits.  Requires (N*H) % 4 == 0.
constexpr int ODA_THREADS = 256;
__global__ void __launch_bounds__(ODA_THREADS)
oda_pair_logits_train_kernel(int64_t N, int64_t H, Drop d, const float* __restrict__ vl, const float* __restrict__ ql,
                             const float* __restrict__ W, const float* __restrict__ bc, float* __restrict__ z) {
  extern __shared__ float sm[];
  float* vi_s = sm;          // vl[b,i,:]
  float* ql_s = sm + H;      // ql[b,:]
  __shared__ float red[G][ODA_THREADS / 32];
  const int64_t i = blockIdx.x, b = blockIdx.y;
  const int64_t NH = N * H;
  for (int64_t k = threadIdx.x; k < H; k += ODA_THREADS) {
    vi_s[k] = vl[(b * N + i) * H + k];
    ql_s[k] = ql[b * H + k];
  }
  __syncthreads();
  const float* vb = vl + b * N * H;
  const uint64_t base = d.base + (uint64_t)((b * N + i) * NH);
  float acc[G] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t t = threadIdx.x; t < NH / 4; t += ODA_THREADS) {
    const int64_t e0 = t * 4;
    const uint32_t bt = philox_bytes4(d.key(), d.layer, base + (uint64_t)e0);
    const uint32_t wd[4] = {bt & 0xFFu, (bt >> 8) & 0xFFu, (bt >> 16) & 0xFFu, bt >> 24};
    int64_t j = e0 / H, k = e0 - j * H;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (wd[u] >= d.thr) {
        const float delta = (vi_s[k] - vb[j * H + k]) * ql_s[k];
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = fmaf(W[g * NH + e0 + u], delta, acc[g]);
      }
      if (++k == H) { k = 0; ++j; }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float v = warp_sum(acc[g]);
    if (lane == 0) red[g][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float s = 0.0f;
    for (int w = 0; w < ODA_THREADS / 32; ++w) s += red[threadIdx.x][w];
    z[(b * N + i) * G + threadIdx.x] = fmaf(s, d.scale, bc[threadIdx.x]);
  }
}

// dz = alpha (.) (dalpha - <alpha, dalpha>), dbc += sum dz.  grid = B
__global__ void softmax_regions_bwd_kernel(int64_t N, const float* __restrict__ alpha, const float* __restrict__ dalpha,
                                           float* __restrict__ dz, float* __restrict__ dbc) {
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= G) return;
  float s = 0.0f;
  for (int64_t i = lane; i < N; i += 32) s = fmaf(alpha[(b * N + i) * G + warp], dalpha[(b * N + i) * G + warp], s);
  s = warp_sum(s);
  float tot = 0.0f;
  for (int64_t i = lane; i < N; i += 32) {
    const int64_t o = (b * N + i) * G + warp;
    const float v = alpha[o] * (dalpha[o] - s);
    dz[o] = v;
    tot += v;
  }
  tot = warp_sum(tot);
  if (lane == 0 && dbc) atomicAdd(&dbc[warp], tot);
}

// dvl / dql in train mode.  grid = B; thread k owns column k of dvl[b] (held in shared memory), so
// the +(i,k) / -(j,k) scatter needs no atomics and the result is deterministic.
//   u(i,j,k) = scale*keep*sum_g dz[i,g] W[g,j,k];  dvl[i,k] += u*ql[k];  dvl[j,k] -= u*ql[k];
//   dql[k] += u*(vl[i,k]-vl[j,k]).
__global__ void oda_pair_bwd_train_dv_kernel(int64_t N, int64_t H, Drop d, const float* __restrict__ vl,
                                             const float* __restrict__ ql, const float* __restrict__ W,
                                             const float* __restrict__ dz, float* __restrict__ dvl,
                                             float* __restrict__ dql) {
  extern __shared__ float sm[];
  float* dz_s = sm;              // [N*G]
  float* col_s = sm + N * G;     // [N*H]
  const int64_t b = blockIdx.x;
  const int64_t NH = N * H;
  for (int64_t t = threadIdx.x; t < N * G; t += blockDim.x) dz_s[t] = dz[b * N * G + t];
  for (int64_t t = threadIdx.x; t < NH; t += blockDim.x) col_s[t] = 0.0f;
  __syncthreads();
  const float* vb = vl + b * N * H;
  for (int64_t k = threadIdx.x; k < H; k += blockDim.x) {
    float dq = 0.0f;
    for (int64_t j = 0; j < N; ++j) {
      const float vj = vb[j * H + k];
      float w[G];
#pragma unroll
      for (int g = 0; g < G; ++g) w[g] = W[g * NH + j * H + k];
      float colj = 0.0f;
      for (int64_t i = 0; i < N; ++i) {
        const uint64_t idx = (uint64_t)((b * N + i) * NH + j * H + k);
        if (philox_byte(d.key(), d.layer, d.base + idx) >= d.thr) {
          const float* zz = dz_s + i * G;
          const float u = (zz[0] * w[0] + zz[1] * w[1] + zz[2] * w[2] + zz[3] * w[3]) * d.scale;
          colj -= u;
          dq = fmaf(u, vb[i * H + k] - vj, dq);
          col_s[i * H + k] += u;
        }
      }
      col_s[j * H + k] += colj;
    }
    const float qk = ql[b * H + k];
    for (int64_t i = 0; i < N; ++i) dvl[(b * N + i) * H + k] = col_s[i * H + k] * qk;
    dql[b * H + k] = dq;
  }
}

// dW[g,e] += scale * sum_{b,i} dz[b,i,g] keep(b,i,e) (vl[b,i,k]-vl[b,j,k]) ql[b,k].
// grid = (cdiv(NH/4,128), b-chunks); a thread owns one aligned quad of e for its chunk of samples.
constexpr int ODA_BCHUNK = 4;
__global__ void __launch_bounds__(128)
oda_pair_bwd_train_dw_kernel(int64_t B, int64_t N, int64_t H, Drop d, const float* __restrict__ vl,
                             const float* __restrict__ ql, const float* __restrict__ dz, float* __restrict__ dW) {
  const int64_t NH = N * H;
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t >= NH / 4) return;
  const int64_t e0 = t * 4;
  int64_t jj[4], kk[4];
  {
    int64_t j = e0 / H, k = e0 - j * H;
    for (int u = 0; u < 4; ++u) {
      jj[u] = j; kk[u] = k;
      if (++k == H) { k = 0; ++j; }
    }
  }
  float acc[4][G];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int g = 0; g < G; ++g) acc[u][g] = 0.0f;
  const int64_t b0 = (int64_t)blockIdx.y * ODA_BCHUNK;
  const int64_t b1 = b0 + ODA_BCHUNK < B ? b0 + ODA_BCHUNK : B;
  for (int64_t b = b0; b < b1; ++b) {
    const float* vb = vl + b * N * H;
    float qv[4], vj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { qv[u] = ql[b * H + kk[u]]; vj[u] = vb[jj[u] * H + kk[u]]; }
    for (int64_t i = 0; i < N; ++i) {
      const uint64_t idx = d.base + (uint64_t)((b * N + i) * NH + e0);
      const uint32_t bt = philox_bytes4(d.key(), d.layer, idx);
      const uint32_t wd[4] = {bt & 0xFFu, (bt >> 8) & 0xFFu, (bt >> 16) & 0xFFu, bt >> 24};
      const float4 z4 = *reinterpret_cast<const float4*>(&dz[(b * N + i) * G]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (wd[u] >= d.thr) {
          const float delta = (vb[i * H + kk[u]] - vj[u]) * qv[u];
          acc[u][0] = fmaf(z4.x, delta, acc[u][0]);
          acc[u][1] = fmaf(z4.y, delta, acc[u][1]);
          acc[u][2] = fmaf(z4.z, delta, acc[u][2]);
          acc[u][3] = fmaf(z4.w, delta, acc[u][3]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int g = 0; g < G; ++g) atomicAdd(&dW[g * NH + e0 + u], acc[u][g] * d.scale);
}

}  // namespace vqa

using namespace vqa;

static int oda_check(int64_t B, int64_t N, int64_t H, int64_t D, const char* who) {
  VQA_REQUIRE(B >= 0 && N >= 1 && H >= 1 && D >= 4 && D % 4 == 0, "%s: bad shape B=%lld N=%lld H=%lld D=%lld", who,
              (long long)B, (long long)N, (long long)H, (long long)D);
  VQA_REQUIRE((N * H) % 4 == 0, "%s: N*H=%lld must be a multiple of 4", who, (long long)(N * H));
  return VQA_OK;
}

extern "C" int vqa_oda_pair_attn_fwd(const vqa_oda_pair_attn_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_oda_pair_attn_fwd: null params");
  VQA_TRY(oda_check(p->B, p->N, p->H, p->D, "vqa_oda_pair_attn_fwd"));
  VQA_REQUIRE(p->vl && p->ql && p->W && p->bc && p->x && p->alpha && p->pooled, "vqa_oda_pair_attn_fwd: null pointer");
  if (p->B == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool train = p->train && p->drop.p > 0.0f;
  if (!train) {
    VQA_REQUIRE(p->wsum != nullptr, "vqa_oda_pair_attn_fwd: wsum workspace required in eval mode");
    oda_wsum_kernel<<<G, 256, 0, st>>>(p->N, p->H, p->W, p->wsum);
    VQA_TRY(check_launch("oda_wsum"));
    FuseOdaEval fs{p->vl, p->ql, p->N, p->H};
    const size_t smem = (size_t)(G * p->H + p->N * G) * sizeof(float);
    att_logits_softmax_kernel<FuseOdaEval><<<(unsigned)p->B, ATT_THREADS, smem, st>>>(fs, p->N, p->H, p->wsum, p->bc,
                                                                                     p->alpha);
    VQA_TRY(check_launch("oda_logits_eval"));
  } else {
    Drop d = make_drop(p->drop.p, p->drop.seed, p->drop.layer, 0, 1, p->drop.seed_dev);
    dim3 grid((unsigned)p->N, (unsigned)p->B);
    oda_pair_logits_train_kernel<<<grid, ODA_THREADS, (size_t)2 * p->H * sizeof(float), st>>>(
        p->N, p->H, d, p->vl, p->ql, p->W, p->bc, p->alpha);
    VQA_TRY(check_launch("oda_pair_logits_train"));
    softmax_regions_kernel<<<(unsigned)p->B, 128, (size_t)p->N * G * sizeof(float), st>>>(p->N, p->alpha);
    VQA_TRY(check_launch("softmax_regions"));
  }
  return launch_pool_fwd(p->B, p->N, p->D, p->x, p->alpha, p->pooled, st);
}

extern "C" int vqa_oda_pair_attn_bwd(const vqa_oda_pair_attn_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_oda_pair_attn_bwd: null params");
  VQA_TRY(oda_check(p->B, p->N, p->H, p->D, "vqa_oda_pair_attn_bwd"));
  VQA_REQUIRE(p->vl && p->ql && p->W && p->x && p->alpha && p->dpooled && p->dalpha && p->dz && p->dW && p->dvl &&
                  p->dql,
              "vqa_oda_pair_attn_bwd: null pointer");
  if (p->B == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t NH = p->N * p->H;
  VQA_TRY(launch_pool_bwd(p->B, p->N, p->D, p->x, p->alpha, p->dpooled, nullptr, p->dalpha, nullptr, 0, st));
  if (!p->accumulate_w && p->dbc) cudaMemsetAsync(p->dbc, 0, G * sizeof(float), st);
  const bool train = p->train && p->drop.p > 0.0f;
  if (!train) {
    VQA_REQUIRE(p->wsum && p->dwsum, "vqa_oda_pair_attn_bwd: wsum/dwsum workspaces required in eval mode");
    cudaMemsetAsync(p->dwsum, 0, (size_t)G * p->H * sizeof(float), st);
    FuseOdaEval fs{p->vl, p->ql, p->N, p->H};
    const int64_t groups = p->B < 2 * (int64_t)sm_count() ? p->B : 2 * (int64_t)sm_count();
    dim3 grid((unsigned)groups, (unsigned)cdiv(p->H, ATT_THREADS));
    att_logits_softmax_bwd_kernel<FuseOdaEval, true><<<grid, ATT_THREADS, (size_t)2 * p->N * G * sizeof(float), st>>>(
        fs, p->B, p->N, p->H, p->wsum, p->alpha, p->dalpha, p->dz, p->dwsum, p->dbc, p->dvl, p->dql);
    VQA_TRY(check_launch("oda_logits_eval_bwd"));
    oda_dw_broadcast_kernel<<<(unsigned)cdiv(G * NH, 256), 256, 0, st>>>(p->N, p->H, p->dwsum, p->dW, p->accumulate_w);
    return check_launch("oda_dw_broadcast");
  }
  Drop d = make_drop(p->drop.p, p->drop.seed, p->drop.layer, 0, 1, p->drop.seed_dev);
  softmax_regions_bwd_kernel<<<(unsigned)p->B, 128, 0, st>>>(p->N, p->alpha, p->dalpha, p->dz, p->dbc);
  VQA_TRY(check_launch("softmax_regions_bwd"));
  if (!p->accumulate_w) cudaMemsetAsync(p->dW, 0, (size_t)G * NH * sizeof(float), st);
  {
    const int threads = (int)(p->H >= 512 ? 512 : ((p->H + 31) / 32) * 32);
    const size_t smem = (size_t)(p->N * G + NH) * sizeof(float);
    VQA_REQUIRE(smem <= 220 * 1024, "vqa_oda_pair_attn_bwd: N*H=%lld too large for shared memory", (long long)NH);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(oda_pair_bwd_train_dv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    oda_pair_bwd_train_dv_kernel<<<(unsigned)p->B, threads, smem, st>>>(p->N, p->H, d, p->vl, p->ql, p->W, p->dz,
                                                                       p->dvl, p->dql);
    VQA_TRY(check_launch("oda_pair_bwd_train_dv"));
  }
  {
    dim3 grid((unsigned)cdiv(NH / 4, 128), (unsigned)cdiv(p->B, ODA_BCHUNK));
    oda_pair_bwd_train_dw_kernel<<<grid, 128, 0, st>>>(p->B, p->N, p->H, d, p->vl, p->ql, p->dz, p->dW);
    VQA_TRY(check_launch("oda_pair_bwd_train_dw"));
  }
  return VQA_OK;
}
