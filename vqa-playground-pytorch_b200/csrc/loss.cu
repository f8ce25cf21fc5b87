// KLDivLoss(size_average=False)(log_softmax(x,1), a) fused with its gradient
// (include/vqacore.h: vqa_kld_logsoftmax_fwd_bwd; reference train.py:536-544).
#include <cuda_bf16.h>

#include "common.cuh"

namespace vqa {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < LOSS_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// one CTA per row
__global__ void __launch_bounds__(LOSS_THREADS)
kld_logsoftmax_kernel(int64_t C, float grad_scale, const float* __restrict__ logits, const float* __restrict__ target,
                      float* __restrict__ loss_rows, float* __restrict__ dlogits) {
  __shared__ float red[LOSS_THREADS / 32];
  const int64_t b = blockIdx.x;
  const float* x = logits + b * C;
  const float* a = target + b * C;
  float mx = -INFINITY;
  for (int64_t c = threadIdx.x; c < C; c += LOSS_THREADS) mx = fmaxf(mx, x[c]);
  mx = block_reduce(mx, red, true);
  float se = 0.0f, sa = 0.0f;
  for (int64_t c = threadIdx.x; c < C; c += LOSS_THREADS) {
    se += expf(x[c] - mx);
    sa += a[c];
  }
  se = block_reduce(se, red, false);
  sa = block_reduce(sa, red, false);
  const float lse = mx + logf(se);
  float l = 0.0f;
  for (int64_t c = threadIdx.x; c < C; c += LOSS_THREADS) {
    const float t = a[c], lp = x[c] - lse;
    if (t > 0.0f) l += t * (logf(t) - lp);        // kl_div: 0 where target == 0
    if (dlogits) dlogits[b * C + c] = grad_scale * (expf(lp) * sa - t);
  }
  l = block_reduce(l, red, false);
  if (threadIdx.x == 0) loss_rows[b] = l;
}

// loss = sum_b loss_rows[b] (deterministic single-CTA tree);  dlogits *= *scale (device scalar, no host sync)
__global__ void __launch_bounds__(LOSS_THREADS) sum_rows_kernel(int64_t B, const float* __restrict__ rows, float* __restrict__ out) {
  __shared__ float red[LOSS_THREADS / 32];
  float s = 0.0f;
  for (int64_t b = threadIdx.x; b < B; b += LOSS_THREADS) s += rows[b];
  s = block_reduce(s, red, false);
  if (threadIdx.x == 0) *out = s;
}
__global__ void scale_by_device_scalar_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ scale,
                                              float* __restrict__ y) {
  const float sc = __ldg(scale);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    y[t] = x[t] * sc;
}

// Prediction tail of the reference's eval loop (train.py:146-169): pred[b] = argmax_c logits[b, c] over all answers
// (OpenEnded) or over the candidate list mc_idx[b, :] (MultipleChoice; -1 = padding; no candidate -> -1), the FIRST
// maximum on ties, as torch.max / the reference's `ans_prob < out[j][k]` scan give.  One warp per row.
__global__ void __launch_bounds__(LOSS_THREADS)
argmax_rows_kernel(int64_t B, int64_t C, const float* __restrict__ logits, const int64_t* __restrict__ mc_idx,
                   int64_t n_mc, int64_t* __restrict__ pred, float* __restrict__ best) {
  const int64_t b = (int64_t)blockIdx.x * (LOSS_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* x = logits + b * C;
  float bv = -INFINITY;
  int64_t bi = -1;
  auto take = [&](float v, int64_t i) {
    if (i >= 0 && (bi < 0 || v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
  };
  if (mc_idx) {
    for (int64_t t = lane; t < n_mc; t += 32) {
      const int64_t k = mc_idx[b * n_mc + t];
      if (k >= 0 && k < C) take(x[k], k);
    }
  } else {
    for (int64_t c = lane; c < C; c += 32) take(x[c], c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    take(ov, oi);
  }
  if (lane == 0) {
    pred[b] = bi;
    if (best) best[b] = bv;
  }
}

// dst[i] = float(src[i]), bf16 -> fp32 (exact).  8 elements per thread, 16-byte loads.
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(int64_t n, const __nv_bfloat16* __restrict__ src, float* __restrict__ dst) {
  const int64_t n8 = n / 8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += (int64_t)gridDim.x * blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + t);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float4 a, b;
    a.x = __uint_as_float(w[0] << 16); a.y = __uint_as_float(w[0] & 0xFFFF0000u);
    a.z = __uint_as_float(w[1] << 16); a.w = __uint_as_float(w[1] & 0xFFFF0000u);
    b.x = __uint_as_float(w[2] << 16); b.y = __uint_as_float(w[2] & 0xFFFF0000u);
    b.z = __uint_as_float(w[3] << 16); b.w = __uint_as_float(w[3] & 0xFFFF0000u);
    reinterpret_cast<float4*>(dst)[2 * t] = a;
    reinterpret_cast<float4*>(dst)[2 * t + 1] = b;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __bfloat162float(src[i]);
}

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_cast_bf16_f32(int64_t n, const void* src, float* dst, void* stream) {
  VQA_REQUIRE(n >= 0 && src && dst, "vqa_cast_bf16_f32: bad argument");
  VQA_REQUIRE(reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0,
              "vqa_cast_bf16_f32: buffers must be 16-byte aligned");
  if (n == 0) return VQA_OK;
  int64_t blocks = cdiv(n / 8 + 1, 256);
  if (blocks > 8 * (int64_t)sm_count()) blocks = 8 * (int64_t)sm_count();
  KProf kp_(stream, "cast_bf16_f32", "hbm", 6.0 * (double)n);
  cast_bf16_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, reinterpret_cast<const __nv_bfloat16*>(src), dst);
  return check_launch("cast_bf16_f32");
}

extern "C" int vqa_argmax_rows(int64_t B, int64_t C, const float* logits, const int64_t* mc_idx, int64_t n_mc,
                               int64_t* pred, float* best, void* stream) {
  VQA_REQUIRE(B >= 0 && C >= 1 && logits && pred && (mc_idx == nullptr || n_mc >= 1), "vqa_argmax_rows: bad argument");
  if (B == 0) return VQA_OK;
  argmax_rows_kernel<<<(unsigned)cdiv(B, LOSS_THREADS / 32), LOSS_THREADS, 0, (cudaStream_t)stream>>>(B, C, logits, mc_idx,
                                                                                                    n_mc, pred, best);
  return check_launch("argmax_rows");
}

extern "C" int vqa_kld_logsoftmax_fwd_bwd(const vqa_kld_logsoftmax_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_kld_logsoftmax_fwd_bwd: null params");
  VQA_REQUIRE(p->B >= 0 && p->C >= 1, "vqa_kld_logsoftmax_fwd_bwd: bad shape");
  VQA_REQUIRE(p->logits && p->target && p->loss_rows, "vqa_kld_logsoftmax_fwd_bwd: null pointer");
  if (p->B == 0) return VQA_OK;
  KProf kp_(stream, "kld_logsoftmax", "hbm", 12.0 * (double)p->B * p->C);
  kld_logsoftmax_kernel<<<(unsigned)p->B, LOSS_THREADS, 0, (cudaStream_t)stream>>>(p->C, p->grad_scale, p->logits,
                                                                                  p->target, p->loss_rows, p->dlogits);
  return check_launch("kld_logsoftmax");
}

extern "C" int vqa_sum_rows(int64_t n, const float* rows, float* out, void* stream) {
  VQA_REQUIRE(n >= 0 && rows && out, "vqa_sum_rows: bad argument");
  sum_rows_kernel<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>(n, rows, out);
  return check_launch("sum_rows");
}

extern "C" int vqa_scale_by_device_scalar(int64_t n, const float* x, const float* scale, float* y, void* stream) {
  VQA_REQUIRE(n >= 0 && x && scale && y, "vqa_scale_by_device_scalar: bad argument");
  if (n == 0) return VQA_OK;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 2048) blocks = 2048;
  scale_by_device_scalar_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, x, scale, y);
  return check_launch("scale_by_device_scalar");
}
