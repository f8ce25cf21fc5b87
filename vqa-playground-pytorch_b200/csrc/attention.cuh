// Region-softmax attention + pooling kernels shared by MyATT (both models) and ODA's
// object-difference attention.  See include/vqacore.h for the maths and the reference lines.
#pragma once
#include "common.cuh"

namespace vqa {

constexpr int ATT_THREADS = 256;

// ---- sources of the "fuse" row the attention logits are taken from ---------------------------
// Generic: a materialised [B,N,Ff] tensor with input dropout (MyATT.conv_att, config/CoR2.py:140).
struct FuseGeneric {
  const float* fuse; int64_t N, Ff; Drop d;
  const uint8_t* bits;        // optional packed keep-bits (vqa_dropout_bits) of the same mask
  __device__ __forceinline__ float mul(int64_t b, int64_t i, int64_t c) const {
    const uint64_t e = (uint64_t)((b * N + i) * Ff + c);
    if (bits) return ((__ldg(bits + (e >> 3)) >> (e & 7)) & 1u) ? d.scale : 0.0f;
    return d.mul(e);
  }
  __device__ __forceinline__ float raw(int64_t b, int64_t i, int64_t c) const { return fuse[(b * N + i) * Ff + c]; }
  // Row-relative access (the row base is computed once, columns are 32-bit offsets): rowbase = element index of
  // (b, i, 0); the kernels' inner loops were dominated by 64-bit index arithmetic before.
  __device__ __forceinline__ void begin_sample(int64_t) {}
  __device__ __forceinline__ int64_t rowbase(int64_t b, int64_t i) const { return (b * N + i) * Ff; }
  // value at element index base + off; `col` = its column (only the ODA source needs it)
  __device__ __forceinline__ float raw_at(int64_t base, int off, int) const { return __ldg(fuse + base + off); }
  __device__ __forceinline__ float mul_at(int64_t rb, int c) const {
    const uint64_t e = (uint64_t)(rb + c);
    if (bits) return ((__ldg(bits + (e >> 3)) >> (e & 7)) & 1u) ? d.scale : 0.0f;
    return d.mul(e);
  }
  // keep flags of the 4 consecutive elements c..c+3 (bit j = element c+j); the multiplier of a kept element is scale()
  __device__ __forceinline__ uint32_t keep4_at(int64_t rb, int c) const {
    const uint64_t e = (uint64_t)(rb + c);
    if (bits) {
      const uint32_t sh = (uint32_t)(e & 7);
      uint32_t w = __ldg(bits + (e >> 3));
      if (sh > 4) w |= (uint32_t)__ldg(bits + (e >> 3) + 1) << 8;      // the 4 bits straddle two bytes
      return (w >> sh) & 0xFu;
    }
    float m[4];
    d.mul4(e, m);
    return (m[0] != 0.0f ? 1u : 0u) | (m[1] != 0.0f ? 2u : 0u) | (m[2] != 0.0f ? 4u : 0u) | (m[3] != 0.0f ? 8u : 0u);
  }
  __device__ __forceinline__ float scale() const { return d.scale; }
  // multipliers of 4 consecutive elements c..c+3 of row (b,i)
  __device__ __forceinline__ void mul4(int64_t b, int64_t i, int64_t c, float (&m)[4]) const {
    const uint64_t e = (uint64_t)((b * N + i) * Ff + c);
    if (bits) {
      const uint32_t sh = (uint32_t)(e & 7);
      uint32_t w = __ldg(bits + (e >> 3));
      if (sh > 4) w |= (uint32_t)__ldg(bits + (e >> 3) + 1) << 8;      // the 4 bits straddle two bytes
      w >>= sh;
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = ((w >> j) & 1u) ? d.scale : 0.0f;
      return;
    }
    d.mul4(e, m);
  }
};
// ODA, eval mode: fuse_eff[b,i,k] = vl[b,i,k]*ql[b,k] against Wsum[g,k] = sum_j W[g,j*H+k]
// (the -vl[b,j,k] part is constant over i and cancels in the region softmax).
struct FuseOdaEval {
  const float* vl; const float* ql; int64_t N, Ff;   // Ff == H
  __device__ __forceinline__ float mul(int64_t, int64_t, int64_t) const { return 1.0f; }
  __device__ __forceinline__ void mul4(int64_t, int64_t, int64_t, float (&m)[4]) const { m[0] = m[1] = m[2] = m[3] = 1.0f; }
  __device__ __forceinline__ float raw(int64_t b, int64_t i, int64_t c) const {
    return vl[(b * N + i) * Ff + c] * ql[b * Ff + c];
  }
  __device__ __forceinline__ int64_t rowbase(int64_t b, int64_t i) const { return (b * N + i) * Ff; }
  int64_t qoff;                                        // b * Ff of the sample being processed (begin_sample)
  __device__ __forceinline__ void begin_sample(int64_t b) { qoff = b * Ff; }
  __device__ __forceinline__ float raw_at(int64_t base, int off, int col) const {
    return __ldg(vl + base + off) * __ldg(ql + qoff + col);
  }
  __device__ __forceinline__ float mul_at(int64_t, int) const { return 1.0f; }
  __device__ __forceinline__ uint32_t keep4_at(int64_t, int) const { return 0xFu; }
  __device__ __forceinline__ float scale() const { return 1.0f; }
};

// softmax over regions of z[N][G] held in shared memory, one warp per glimpse, in place
__device__ __forceinline__ void softmax_regions_smem(float* z, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < G) {
    float mx = -INFINITY;
    for (int i = lane; i < N; i += 32) mx = fmaxf(mx, z[i * G + warp]);
    mx = warp_max(mx);
    float s = 0.0f;
    for (int i = lane; i < N; i += 32) {
      const float e = expf(z[i * G + warp] - mx);
      z[i * G + warp] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int i = lane; i < N; i += 32) z[i * G + warp] *= inv;
  }
}

// z[b,i,g] = sum_c Wc[g,c]*fuse~[b,i,c] + bc[g]; alpha = softmax_i z.   grid = (B, S).  S = 1: logits and softmax in
// one CTA per sample.  S > 1: CTA (b, s) takes the rows i = s, s + S, ... and writes the raw logits; the softmax over
// regions follows in softmax_regions_kernel (one CTA per sample keeps only ~14 warps per SM resident at B = 256 and
// left the kernel latency-bound at 13 % of the HBM rate).
// dynamic smem: G*Ff (weights) + N*G (logits).  One warp per region row; a lane owns quads of 4 consecutive
// elements (one Philox call per quad when the row start is 4-aligned, two otherwise) and issues all of its
// loads for the row before using them.
template <class FS>
__global__ void __launch_bounds__(ATT_THREADS, 4)     // <= 64 registers: 116 left two CTAs per SM and 3.5 waves (ncu r2)
att_logits_softmax_kernel(FS fs, int64_t N, int64_t Ff, const float* __restrict__ Wc, const float* __restrict__ bc,
                          float* __restrict__ alpha) {
  extern __shared__ float smem[];
  float* w_s = smem;
  float* z_s = smem + G * Ff;
  const int64_t b = blockIdx.x;
  fs.begin_sample(b);
  {  // weights -> shared memory, every load of a thread in flight before its first store
    const int total = (int)(G * Ff);
    constexpr int WU = 8;
    float wv[WU];
#pragma unroll
    for (int u = 0; u < WU; ++u) {
      const int t = threadIdx.x + u * ATT_THREADS;
      wv[u] = t < total ? __ldg(Wc + t) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < WU; ++u) {
      const int t = threadIdx.x + u * ATT_THREADS;
      if (t < total) w_s[t] = wv[u];
    }
    for (int t = threadIdx.x + WU * ATT_THREADS; t < total; t += ATT_THREADS) w_s[t] = __ldg(Wc + t);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int QPL = 4;                       // quads per lane per pass: 32 lanes * 4 quads * 4 = 512 elements
  const int S = gridDim.y;
  for (int64_t i = (int64_t)blockIdx.y + (int64_t)warp * S; i < N; i += (int64_t)(ATT_THREADS / 32) * S) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int64_t rb = fs.rowbase(b, i);
    const int F = (int)Ff;
    for (int c0 = 0; c0 < F; c0 += 32 * QPL * 4) {
      float f[QPL][4];
      uint32_t keep[QPL];
#pragma unroll
      for (int j = 0; j < QPL; ++j) {
        const int c = c0 + (lane + 32 * j) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) f[j][e] = (c + e < F) ? fs.raw_at(rb, c + e, c + e) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < QPL; ++j) {      // the mask loads are issued before the first use of anything above
        const int c = c0 + (lane + 32 * j) * 4;
        keep[j] = c < F ? fs.keep4_at(rb, c) : 0u;
      }
      const float sc = fs.scale();
#pragma unroll
      for (int j = 0; j < QPL; ++j) {
        const int c = c0 + (lane + 32 * j) * 4;
        if (c < F) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (c + e < F) {
              const float v = (keep[j] >> e) & 1u ? f[j][e] * sc : 0.0f;
              a0 = fmaf(w_s[c + e], v, a0);
              a1 = fmaf(w_s[F + c + e], v, a1);
              a2 = fmaf(w_s[2 * F + c + e], v, a2);
              a3 = fmaf(w_s[3 * F + c + e], v, a3);
            }
          }
        }
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) {
      if (S > 1) {
        *reinterpret_cast<float4*>(&alpha[(b * N + i) * G]) = make_float4(a0 + bc[0], a1 + bc[1], a2 + bc[2], a3 + bc[3]);
      } else {
        z_s[i * G + 0] = a0 + bc[0];
        z_s[i * G + 1] = a1 + bc[1];
        z_s[i * G + 2] = a2 + bc[2];
        z_s[i * G + 3] = a3 + bc[3];
      }
    }
  }
  if (S > 1) return;
  __syncthreads();
  softmax_regions_smem(z_s, (int)N);
  __syncthreads();
  for (int64_t t = threadIdx.x; t < N * G; t += ATT_THREADS) alpha[b * N * G + t] = z_s[t];
}

// in-place z -> alpha for logits already in global memory (ODA train path). grid = B, smem N*G
__global__ void softmax_regions_kernel(int64_t N, float* __restrict__ alpha);

int launch_pool_fwd(int64_t B, int64_t N, int64_t D, const float* x, const float* alpha, float* pooled,
                    cudaStream_t st);
// dalpha (into dz buffer) and optional dx
int launch_pool_bwd(int64_t B, int64_t N, int64_t D, const float* x, const float* alpha, const float* dpooled,
                    const float* dalpha0_ext, float* dalpha, float* dx, int accumulate_x, cudaStream_t st,
                    const float* dalpha_ext = nullptr);

// ---- backward of logits+softmax --------------------------------------------------------------
// grid = (sample groups, cdiv(Ff, ATT_THREADS)); each thread owns one column c for its samples.
// dalpha is read by every CTA; dz_out is written by the blockIdx.y == 0 CTAs only (separate buffers).
// ODA_EVAL: dfuse is redirected into dvl/dql.
template <class FS, bool ODA_EVAL>
__global__ void __launch_bounds__(ATT_THREADS)
att_logits_softmax_bwd_kernel(FS fs, int64_t B, int64_t N, int64_t Ff, const float* __restrict__ Wc,
                              const float* __restrict__ alpha, const float* __restrict__ dalpha,
                              float* __restrict__ dz_out, float* __restrict__ dWc,
                              float* __restrict__ dbc, float* __restrict__ dfuse, float* __restrict__ dql) {
  extern __shared__ float smem[];
  float* dz_s = smem;            // [N*G]
  float* al_s = smem + N * G;    // [N*G]
  __shared__ float dot_s[G];
  const int64_t c = (int64_t)blockIdx.y * ATT_THREADS + threadIdx.x;
  const bool active = c < Ff;
  float w[G] = {0.f, 0.f, 0.f, 0.f}, accw[G] = {0.f, 0.f, 0.f, 0.f};
  float accb = 0.0f;
  if (active)
#pragma unroll
    for (int g = 0; g < G; ++g) w[g] = Wc[g * Ff + c];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    fs.begin_sample(b);
    __syncthreads();
    for (int64_t t = threadIdx.x; t < N * G; t += ATT_THREADS) {
      dz_s[t] = dalpha[b * N * G + t];
      al_s[t] = alpha[b * N * G + t];
    }
    __syncthreads();
    if (warp < G) {
      float s = 0.0f;
      for (int i = lane; i < N; i += 32) s = fmaf(al_s[i * G + warp], dz_s[i * G + warp], s);
      s = warp_sum(s);
      if (lane == 0) dot_s[warp] = s;
    }
    __syncthreads();
    for (int64_t t = threadIdx.x; t < N * G; t += ATT_THREADS) dz_s[t] = al_s[t] * (dz_s[t] - dot_s[t % G]);
    __syncthreads();
    if (blockIdx.y == 0) {
      for (int64_t t = threadIdx.x; t < N * G; t += ATT_THREADS) {
        dz_out[b * N * G + t] = dz_s[t];
        if (t < G) {
          float s = 0.0f;
          for (int64_t i = 0; i < N; ++i) s += dz_s[i * G + t];
          accb += s;      // thread t < G accumulates dbc[t]
        }
      }
    }
    if (active) {
      float dq = 0.0f;
      constexpr int U = 6;
      const int64_t sb = fs.rowbase(b, 0) + c;          // element (b, 0, c); rows are Ff apart (32-bit offsets below)
      const int F = (int)Ff, Ni = (int)N;
      for (int i0 = 0; i0 < Ni; i0 += U) {
        float fr[U], mu[U];
#pragma unroll
        for (int u = 0; u < U; ++u) fr[u] = (i0 + u < Ni) ? fs.raw_at(sb, (i0 + u) * F, (int)c) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u) mu[u] = (i0 + u < Ni) ? fs.mul_at(sb, (i0 + u) * F) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t i = i0 + u;
          if (i < N) {
            const float4 z4 = *reinterpret_cast<const float4*>(&dz_s[i * G]);
            const float f = fr[u] * mu[u];
            accw[0] = fmaf(z4.x, f, accw[0]);
            accw[1] = fmaf(z4.y, f, accw[1]);
            accw[2] = fmaf(z4.z, f, accw[2]);
            accw[3] = fmaf(z4.w, f, accw[3]);
            const float df = (z4.x * w[0] + z4.y * w[1] + z4.z * w[2] + z4.w * w[3]) * mu[u];
            if constexpr (ODA_EVAL) {
              // fuse_eff = vl*ql: dvl = df*ql, dql += df*vl
              const float qv = fs.ql[b * Ff + c], vv = fs.vl[(b * N + i) * Ff + c];
              dfuse[(b * N + i) * Ff + c] = df * qv;
              dq = fmaf(df, vv, dq);
            } else {
              if (dfuse) dfuse[(b * N + i) * Ff + c] = df;
            }
          }
        }
      }
      if constexpr (ODA_EVAL) dql[b * Ff + c] = dq;
    }
  }
  if (active)
#pragma unroll
    for (int g = 0; g < G; ++g) atomicAdd(&dWc[g * Ff + c], accw[g]);
  if (blockIdx.y == 0 && threadIdx.x < G && dbc) atomicAdd(&dbc[threadIdx.x], accb);
}

}  // namespace vqa
