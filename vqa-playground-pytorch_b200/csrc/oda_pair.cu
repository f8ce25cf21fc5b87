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
// The [B,N,N*H] tensor of the reference (config/ODA.py:222) is indexed by (b, i, e) with e = j*H + k.
// z[b,i,g] = bc[g] + scale * sum_e keep(b,i,e) W[g,e] (vl[b,i,k]-vl[b,j,k]) ql[b,k].
//
// "e-mapping" kernels: a thread owns EPT = 8 consecutive e of one sample — their W[g,e], vl[b,j,k], ql[b,k] stay
// in registers — and loops over all regions i, drawing the 8 keep-bytes of (b,i,e..e+7) with one Philox call
// (N*H is a multiple of 8, so the 8 elements never straddle a 16-element Philox group).  W is read once per
// (sample, e) instead of once per (sample, i, e): 100x less L2 traffic at N = 100.
constexpr int ODA_EPT = 8;
constexpr int ODA_THREADS = 256;

// v[e] = row[kk[e]], e < 8.  kk holds consecutive feature indices unless the chunk wraps into the next region (or runs
// past the end); consecutive and 8-byte aligned (even start, even row length) -> four 8-byte loads instead of eight
// 4-byte ones: lanes sit 32 bytes apart, so every load instruction costs 8 L1 wavefronts whatever its width.
__device__ __forceinline__ void load_row8(const float* __restrict__ row, const int (&kk)[8], bool contiguous,
                                          float (&v)[8]) {
  if (contiguous) {
    const float2* p = reinterpret_cast<const float2*>(row + kk[0]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = __ldg(p + e);
      v[2 * e] = t.x; v[2 * e + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldg(row + kk[e]);
  }
}

// keep flags (1.0 / 0.0) of 8 consecutive indices starting at idx.  Fast path idx % 8 == 0: one Philox call.
__device__ __forceinline__ void keep8(const Drop& d, uint64_t seed, uint64_t idx, float (&keep)[8]) {
  const uint4 r = philox_group(seed, d.layer, idx >> 4);
  if ((idx & 7) == 0) {
    const uint32_t lo = (idx & 8) ? r.z : r.x, hi = (idx & 8) ? r.w : r.y;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      keep[e] = ((lo >> (8 * e)) & 0xFFu) >= d.thr ? 1.0f : 0.0f;
      keep[4 + e] = ((hi >> (8 * e)) & 0xFFu) >= d.thr ? 1.0f : 0.0f;
    }
    return;
  }
  const uint4 r1 = philox_group(seed, d.layer, (idx >> 4) + 1);     // unaligned rows (N*H not a multiple of 8)
  const uint32_t wd[8] = {r.x, r.y, r.z, r.w, r1.x, r1.y, r1.z, r1.w};
  const uint32_t off = (uint32_t)idx & 15u;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t pos = off + e;
    keep[e] = ((wd[pos >> 2] >> (8u * (pos & 3u))) & 0xFFu) >= d.thr ? 1.0f : 0.0f;
  }
}

// grid = (cdiv(NH/8, 256), B).  z must hold bc[g] on entry (oda_init_logits_kernel); partial sums are added.
__global__ void __launch_bounds__(ODA_THREADS)
oda_pair_logits_train_kernel(int64_t N, int64_t H, Drop d, const float* __restrict__ vl, const float* __restrict__ ql,
                             const float* __restrict__ W, float* __restrict__ z) {
  extern __shared__ float part[];                       // [warps][N][G]
  const int64_t b = blockIdx.y;
  const int64_t NH = N * H;
  const int64_t e0 = ((int64_t)blockIdx.x * ODA_THREADS + threadIdx.x) * ODA_EPT;
  const bool active = e0 < NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t seed = d.key();
  const float* vb = vl + b * NH;
  float w[G][ODA_EPT], vj[ODA_EPT], qk[ODA_EPT];
  int kk[ODA_EPT];
  if (active) {
    int64_t j = e0 / H, k = e0 - j * H;
#pragma unroll
    for (int e = 0; e < ODA_EPT; ++e) {
      const bool in = e0 + e < NH;                      // ragged tail when N*H is not a multiple of 8
      kk[e] = in ? (int)k : 0;
      vj[e] = in ? vb[j * H + k] : 0.0f;
      qk[e] = in ? ql[b * H + k] * d.scale : 0.0f;
#pragma unroll
      for (int g = 0; g < G; ++g) w[g][e] = in ? W[g * NH + e0 + e] : 0.0f;
      if (++k == H) { k = 0; ++j; }
    }
  }
  // the 8 features are consecutive in every row (no wrap, no ragged tail) and 8-byte aligned there
  const bool contig = active && e0 + ODA_EPT <= NH && kk[ODA_EPT - 1] == kk[0] + ODA_EPT - 1 && (kk[0] & 1) == 0 &&
                      (H & 1) == 0 && (reinterpret_cast<uintptr_t>(vl) & 7) == 0;
  for (int64_t i = 0; i < N; ++i) {
    float acc[G] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      float keep[ODA_EPT], vi8[ODA_EPT];
      keep8(d, seed, d.base + (uint64_t)((b * N + i) * NH + e0), keep);
      load_row8(vb + i * H, kk, contig, vi8);
#pragma unroll
      for (int e = 0; e < ODA_EPT; ++e) {          // branch-free: a 50 % mask would diverge on every element
        const float delta = (vi8[e] - vj[e]) * (qk[e] * keep[e]);
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = fmaf(w[g][e], delta, acc[g]);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = warp_sum(acc[g]);
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < G; ++g) part[(warp * N + i) * G + g] = acc[g];
    }
  }
  __syncthreads();
  for (int64_t t = threadIdx.x; t < N * G; t += ODA_THREADS) {
    float s = 0.0f;
#pragma unroll
    for (int wv = 0; wv < ODA_THREADS / 32; ++wv) s += part[wv * N * G + t];
    atomicAdd(&z[b * N * G + t], s);
  }
}

// z[b,i,g] = bc[g]
__global__ void oda_init_logits_kernel(int64_t total, const float* __restrict__ bc, float* __restrict__ z) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) z[t] = bc[t % G];
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

// Backward, e-mapping (thread owns 8 consecutive e = fixed (j,k) pairs, loops over the samples of its chunk and
// over i).  With u(b,i,e) = scale * keep * sum_g dz[b,i,g] W[g,e]:
//   dvl[b,j,k]  = -ql[b,k] * sum_i u              (this thread is the only writer of (b,j,k): plain store)
//   dql[b,k]   += sum_i u * (vl[b,i,k]-vl[b,j,k])  (atomic: other j share k)
//   dW[g,e]    += scale * sum_{b,i} dz[b,i,g] keep (vl[b,i,k]-vl[b,j,k]) ql[b,k]   (registers over the chunk, then atomic)
// grid = (cdiv(NH/8, 128), cdiv(B, ODA_BCHUNK))
constexpr int ODA_BCHUNK = 8;
__global__ void __launch_bounds__(128)
oda_pair_bwd_train_e_kernel(int64_t B, int64_t N, int64_t H, Drop d, const float* __restrict__ vl,
                            const float* __restrict__ ql, const float* __restrict__ W, const float* __restrict__ dz,
                            float* __restrict__ dW, float* __restrict__ dvl, float* __restrict__ dql) {
  const int64_t NH = N * H;
  const int64_t e0 = ((int64_t)blockIdx.x * 128 + threadIdx.x) * ODA_EPT;
  if (e0 >= NH) return;
  const uint64_t seed = d.key();
  float w[G][ODA_EPT], dwacc[G][ODA_EPT];
  int jj[ODA_EPT], kk[ODA_EPT];
  {
    int64_t j = e0 / H, k = e0 - j * H;
#pragma unroll
    for (int e = 0; e < ODA_EPT; ++e) {
      const bool in = e0 + e < NH;
      jj[e] = in ? (int)j : 0; kk[e] = in ? (int)k : 0;
#pragma unroll
      for (int g = 0; g < G; ++g) { w[g][e] = in ? W[g * NH + e0 + e] : 0.0f; dwacc[g][e] = 0.0f; }
      if (++k == H) { k = 0; ++j; }
    }
  }
  const bool contig = e0 + ODA_EPT <= NH && kk[ODA_EPT - 1] == kk[0] + ODA_EPT - 1 && (kk[0] & 1) == 0 && (H & 1) == 0 &&
                      (reinterpret_cast<uintptr_t>(vl) & 7) == 0;
  const int64_t b0 = (int64_t)blockIdx.y * ODA_BCHUNK;
  const int64_t b1 = b0 + ODA_BCHUNK < B ? b0 + ODA_BCHUNK : B;
  for (int64_t b = b0; b < b1; ++b) {
    const float* vb = vl + b * NH;
    float vj[ODA_EPT], qk[ODA_EPT], minus[ODA_EPT], dq[ODA_EPT];
#pragma unroll
    for (int e = 0; e < ODA_EPT; ++e) {
      vj[e] = vb[(int64_t)jj[e] * H + kk[e]];
      qk[e] = ql[b * H + kk[e]];
      minus[e] = 0.0f; dq[e] = 0.0f;
    }
#pragma unroll 2                                   // two regions in flight: their Philox chains interleave (8 warps per SM)
    for (int64_t i = 0; i < N; ++i) {
      float keep[ODA_EPT];
      keep8(d, seed, d.base + (uint64_t)((b * N + i) * NH + e0), keep);
      const float4 z4 = __ldg(reinterpret_cast<const float4*>(dz + (b * N + i) * G));
      const float zg[G] = {z4.x * d.scale, z4.y * d.scale, z4.z * d.scale, z4.w * d.scale};
      float vi8[ODA_EPT];
      load_row8(vb + i * H, kk, contig, vi8);
#pragma unroll
      for (int e = 0; e < ODA_EPT; ++e) {          // branch-free
        const float diff = (vi8[e] - vj[e]) * keep[e];
        const float u = (zg[0] * w[0][e] + zg[1] * w[1][e] + zg[2] * w[2][e] + zg[3] * w[3][e]) * keep[e];
        minus[e] += u;
        dq[e] = fmaf(u, diff, dq[e]);
        const float dl = diff * qk[e];
#pragma unroll
        for (int g = 0; g < G; ++g) dwacc[g][e] = fmaf(zg[g], dl, dwacc[g][e]);
      }
    }
#pragma unroll
    for (int e = 0; e < ODA_EPT; ++e) {
      if (e0 + e < NH) {
        dvl[b * NH + e0 + e] = -minus[e] * qk[e];
        atomicAdd(&dql[b * H + kk[e]], dq[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < ODA_EPT; ++e)
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (e0 + e < NH) atomicAdd(&dW[g * NH + e0 + e], dwacc[g][e]);
}

// Backward, (i,k)-mapping: thread (i, 16-column slot) loops over j and adds the "+" term
//   dvl[b,i,k] += ql[b,k] * sum_j u(b,i,(j,k))
// to the value the e-kernel stored (it is the only "+" writer of (b,i,k)).  W rows of a block of ODA_JB regions j
// are staged in shared memory and shared by the ODA_IC rows i of the CTA.  A thread reads the 16 features
// k0 .. k0+15 of its slot; the row is stored slot-interleaved, feature k at (k % 16) * KS + k / 16, so that the
// lanes of a warp (consecutive slots) read consecutive words — the plain layout put them 16 words apart, a 16-way
// bank conflict on every read.  grid = (cdiv(N, ODA_IC), B); threads = ODA_IC * cdiv(H,16) rounded up to a warp.
// KSC: slots per row as a compile-time constant (20 for the models' H = 310) so that every shared-memory offset of
// the inner loop is an immediate; 0 = computed from H at run time.
constexpr int ODA_IC = 16, ODA_JB = 8;
template <int KSC>
__global__ void __launch_bounds__(320, 2)
oda_pair_bwd_train_plus_kernel(int64_t N, int64_t H, Drop d, const float* __restrict__ ql,
                                               const float* __restrict__ W, const float* __restrict__ dz,
                                               float* __restrict__ dvl) {
  extern __shared__ float w_s[];                        // [G][ODA_JB][Hp], Hp = 16 * KS
  const int64_t NH = N * H;
  const int KS = KSC > 0 ? KSC : (int)((H + 15) / 16);
  const int Hp = 16 * KS;
  const int64_t b = blockIdx.y;
  const int ii = threadIdx.x / KS, ks = threadIdx.x % KS;
  const int64_t i = (int64_t)blockIdx.x * ODA_IC + ii;
  const bool active = ii < ODA_IC && i < N;
  const int k0 = ks * 16;
  const int nk = (int)(H - k0 < 16 ? H - k0 : 16);
  const uint64_t seed = d.key();
  float zg[G] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float4 z4 = __ldg(reinterpret_cast<const float4*>(dz + (b * N + i) * G));
    zg[0] = z4.x * d.scale; zg[1] = z4.y * d.scale; zg[2] = z4.z * d.scale; zg[3] = z4.w * d.scale;
  }
  float plus[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) plus[e] = 0.0f;
  for (int64_t jb = 0; jb < N; jb += ODA_JB) {
    const int64_t nj = N - jb < ODA_JB ? N - jb : ODA_JB;
    __syncthreads();
    // (loops instead of one flat index: the flat form spent three 64-bit divisions per staged element — two thirds of
    // the kernel's instructions)
    for (int gj = 0; gj < G * (int)nj; ++gj) {
      const int g = gj / (int)nj, jl = gj - g * (int)nj;
      const float* src = W + g * NH + (jb + jl) * H;
      float* dst = w_s + (g * ODA_JB + jl) * Hp;
      for (int k = threadIdx.x; k < (int)H; k += blockDim.x) dst[(k & 15) * KS + (k >> 4)] = __ldg(src + k);
    }
    __syncthreads();
    if (!active) continue;
    for (int64_t jl = 0; jl < nj; ++jl) {
      const uint64_t idx0 = d.base + (uint64_t)((b * N + i) * NH + (jb + jl) * H + k0);
      const uint4 r0 = philox_group(seed, d.layer, idx0 >> 4);
      const uint32_t off = (uint32_t)idx0 & 15u;
      uint4 r1 = r0;
      if (off) r1 = philox_group(seed, d.layer, (idx0 >> 4) + 1);
      // the 16 keep-bytes start at byte `off` of the 32-byte pair (r0, r1): rotate by whole words with predicated
      // moves (static register indexing), then funnel-shift the remaining 0..3 bytes
      uint32_t a[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      if (off & 4u) {
#pragma unroll
        for (int t = 0; t < 7; ++t) a[t] = a[t + 1];
      }
      if (off & 8u) {
#pragma unroll
        for (int t = 0; t < 6; ++t) a[t] = a[t + 2];
      }
      const uint32_t sh = 8u * (off & 3u);
      uint32_t by[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) by[t] = __funnelshift_r(a[t], a[t + 1], sh);
      const float* wrow = w_s + jl * Hp + ks;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float m = (e < nk && ((by[e >> 2] >> (8 * (e & 3))) & 0xFFu) >= d.thr) ? 1.0f : 0.0f;
        const float u = e < nk ? zg[0] * wrow[e * KS] + zg[1] * wrow[ODA_JB * Hp + e * KS] +
                                     zg[2] * wrow[2 * ODA_JB * Hp + e * KS] + zg[3] * wrow[3 * ODA_JB * Hp + e * KS]
                               : 0.0f;
        plus[e] = fmaf(m, u, plus[e]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (e < nk) dvl[(b * N + i) * H + k0 + e] += plus[e] * ql[b * H + k0 + e];
  }
}

}  // namespace vqa

using namespace vqa;

static int oda_check(int64_t B, int64_t N, int64_t H, int64_t D, const char* who) {
  VQA_REQUIRE(B >= 0 && N >= 1 && H >= 1 && D >= 4 && D % 4 == 0, "%s: bad shape B=%lld N=%lld H=%lld D=%lld", who,
              (long long)B, (long long)N, (long long)H, (long long)D);
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
    att_logits_softmax_kernel<FuseOdaEval><<<dim3((unsigned)p->B, 1), ATT_THREADS, smem, st>>>(fs, p->N, p->H, p->wsum,
                                                                                               p->bc, p->alpha);
    VQA_TRY(check_launch("oda_logits_eval"));
  } else {
    Drop d = make_drop(p->drop.p, p->drop.seed, p->drop.layer, 0, 1, p->drop.seed_dev);
    const int64_t NH = p->N * p->H;
    oda_init_logits_kernel<<<(unsigned)cdiv(p->B * p->N * G, 256), 256, 0, st>>>(p->B * p->N * G, p->bc, p->alpha);
    VQA_TRY(check_launch("oda_init_logits"));
    dim3 grid((unsigned)cdiv(cdiv(NH, ODA_EPT), ODA_THREADS), (unsigned)p->B);
    const size_t smem = (size_t)(ODA_THREADS / 32) * p->N * G * sizeof(float);
    VQA_REQUIRE(smem <= 48 * 1024, "vqa_oda_pair_attn_fwd: N=%lld too large", (long long)p->N);
    KProf kp_(st, "oda_pair_logits_train", "flop", (double)p->B * p->N * p->N * p->H * (2.0 * G + 2.0));
    oda_pair_logits_train_kernel<<<grid, ODA_THREADS, smem, st>>>(p->N, p->H, d, p->vl, p->ql, p->W, p->alpha);
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
  cudaMemsetAsync(p->dql, 0, (size_t)p->B * p->H * sizeof(float), st);
  {
    dim3 grid((unsigned)cdiv(cdiv(NH, ODA_EPT), 128), (unsigned)cdiv(p->B, ODA_BCHUNK));
    KProf kp_(st, "oda_pair_bwd_train_e", "flop", (double)p->B * p->N * p->N * p->H * (4.0 * G + 4.0));
    oda_pair_bwd_train_e_kernel<<<grid, 128, 0, st>>>(p->B, p->N, p->H, d, p->vl, p->ql, p->W, p->dz, p->dW, p->dvl,
                                                      p->dql);
    VQA_TRY(check_launch("oda_pair_bwd_train_e"));
  }
  {
    const int KS = (int)((p->H + 15) / 16);
    const int threads = ((ODA_IC * KS + 31) / 32) * 32;
    VQA_REQUIRE(threads <= 1024, "vqa_oda_pair_attn_bwd: H=%lld too large", (long long)p->H);
    const size_t smem = (size_t)G * ODA_JB * 16 * KS * sizeof(float);        // slot-interleaved rows of 16 * KS words
    auto kern = KS == 20 ? oda_pair_bwd_train_plus_kernel<20> : oda_pair_bwd_train_plus_kernel<0>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)cdiv(p->N, ODA_IC), (unsigned)p->B);
    KProf kp_(st, "oda_pair_bwd_train_plus", "flop", (double)p->B * p->N * p->N * p->H * (2.0 * G + 1.0));
    kern<<<grid, threads, smem, st>>>(p->N, p->H, d, p->ql, p->W, p->dz, p->dvl);
    VQA_TRY(check_launch("oda_pair_bwd_train_plus"));
  }
  return VQA_OK;
}
