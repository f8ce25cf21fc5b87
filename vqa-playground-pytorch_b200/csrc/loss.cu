// KLDivLoss(size_average=False)(log_softmax(x,1), a) fused with its gradient
// (include/vqacore.h: vqa_kld_logsoftmax_fwd_bwd; reference train.py:536-544).
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

}  // namespace vqa

using namespace vqa;

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
