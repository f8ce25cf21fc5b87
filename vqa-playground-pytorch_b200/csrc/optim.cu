// Global-norm gradient clipping + Adam over the flat data-parallel gradient buffer (include/vqacore.h:
// vqa_clip_adam_step).  Replaces nn.utils.clip_grad_norm_(model.parameters(), 0.25) followed by
// torch.optim.Adam.step() of the reference's train step (train.py:82-86, optimizer built at train.py:292) — two
// launches instead of the several dozen small ATen kernels of the per-tensor / foreach implementations.
#include "common.cuh"

namespace vqa {

constexpr int OPT_THREADS = 256;

// Sum of squares of g[0..n) with a FIXED reduction order (bit-identical on every data-parallel rank and every run:
// a float atomicAdd per block made the clip coefficient, and after one step the parameters, differ in the last bit
// between ranks).  Stage 1: block b writes its partial to part[1 + b].  Stage 2 (sumsq_final_kernel, one warp-tree
// over the <= OPT_MAX_PARTIALS partials): part[0] = total.
constexpr int OPT_MAX_PARTIALS = VQA_CLIP_SCRATCH_FLOATS - 1;
__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(int64_t n, const float* __restrict__ g, float* __restrict__ part) {
  __shared__ float red[OPT_THREADS / 32];
  float s = 0.0f;
  const int64_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t t = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; t < n4; t += (int64_t)gridDim.x * OPT_THREADS) {
    const float4 v = g4[t];
    s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
  }
  if (blockIdx.x == 0)
    for (int64_t t = n4 * 4 + threadIdx.x; t < n; t += OPT_THREADS) s = fmaf(g[t], g[t], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.0f;
    t = warp_sum(t);
    if (threadIdx.x == 0) part[1 + blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(OPT_THREADS) sumsq_final_kernel(int nparts, float* __restrict__ part) {
  __shared__ float red[OPT_THREADS / 32];
  float s = 0.0f;
  for (int t = threadIdx.x; t < nparts; t += OPT_THREADS) s += part[1 + t];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.0f;
    t = warp_sum(t);
    if (threadIdx.x == 0) part[0] = t;
  }
}

struct SegTable { vqa_param_segment s[VQA_MAX_PARAM_SEGMENTS]; };

// grid = (blocks, segments).  Same arithmetic, in the same order, as torch.optim.Adam's single-tensor path:
//   g *= clip;  m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// *step += 1 and, with a decay factor, *lr *= gamma: the reference's scheduler.step() (ExponentialLR) runs BEFORE
// optimizer.step() (train.py:75-76, :296), so the decayed rate is the one this update uses.
__global__ void adam_tick_kernel(int64_t* __restrict__ step, double* __restrict__ lr, double gamma) {
  *step += 1;
  if (lr && gamma > 0.0) *lr *= gamma;
}

// step_dev / lr_dev non-null: the step count (already ticked) and the learning rate are read from device memory, so the
// launch is identical every step and can be replayed from a CUDA graph.
__global__ void __launch_bounds__(OPT_THREADS)
clip_adam_kernel(SegTable tab, float* __restrict__ grads, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                 const float* __restrict__ sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                 float bias_c1, float sqrt_bias_c2, int write_grads, const int64_t* __restrict__ step_dev,
                 const double* __restrict__ lr_dev) {
  const vqa_param_segment sg = tab.s[blockIdx.y];
  if (step_dev) {
    const double t = (double)__ldg(step_dev);
    bias_c1 = (float)(1.0 - pow((double)beta1, t));
    sqrt_bias_c2 = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  if (lr_dev) lr = (float)__ldg(lr_dev);
  float clip = 1.0f;
  if (max_norm > 0.0f) {
    const float c = max_norm / (sqrtf(__ldg(sumsq)) + 1e-6f);      // clip_grad_norm_: coef clamped to 1
    clip = c < 1.0f ? c : 1.0f;
  }
  const float step_size = lr / bias_c1;
  for (int64_t t = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; t < sg.numel; t += (int64_t)gridDim.x * OPT_THREADS) {
    const int64_t f = sg.offset + t;
    const float g = grads[f] * clip;
    float m = exp_avg[f], v = exp_avg_sq[f];
    m = m + (g - m) * (1.0f - beta1);                                // lerp_(grad, 1 - beta1)
    v = v * beta2 + (1.0f - beta2) * g * g;
    exp_avg[f] = m; exp_avg_sq[f] = v;
    if (write_grads) grads[f] = g;
    const float denom = sqrtf(v) / sqrt_bias_c2 + eps;
    sg.param[t] -= step_size * (m / denom);
  }
}

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_clip_adam_step(const vqa_clip_adam_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_clip_adam_step: null params");
  VQA_REQUIRE(p->nsegs >= 0 && p->nsegs <= VQA_MAX_PARAM_SEGMENTS && p->segs, "vqa_clip_adam_step: bad segment table");
  VQA_REQUIRE(p->grads_flat && p->exp_avg && p->exp_avg_sq && p->total >= 0, "vqa_clip_adam_step: null buffer");
  VQA_REQUIRE((p->step >= 1 || p->step_dev) && p->beta1 >= 0.0f && p->beta1 < 1.0f && p->beta2 >= 0.0f &&
                  p->beta2 < 1.0f && p->eps >= 0.0f,
              "vqa_clip_adam_step: bad hyper-parameter (step counts from 1)");
  VQA_REQUIRE(p->max_norm <= 0.0f || p->scratch, "vqa_clip_adam_step: clipping needs the VQA_CLIP_SCRATCH_FLOATS scratch");
  VQA_REQUIRE(reinterpret_cast<uintptr_t>(p->grads_flat) % 16 == 0, "vqa_clip_adam_step: grads_flat must be 16-byte aligned");
  int64_t biggest = 0;
  SegTable tab = {};
  for (int i = 0; i < p->nsegs; ++i) {
    VQA_REQUIRE(p->segs[i].param && p->segs[i].numel >= 0 && p->segs[i].offset >= 0 &&
                    p->segs[i].offset + p->segs[i].numel <= p->total,
                "vqa_clip_adam_step: segment %d lies outside the flat buffers", i);
    tab.s[i] = p->segs[i];
    if (p->segs[i].numel > biggest) biggest = p->segs[i].numel;
  }
  if (p->nsegs == 0 || biggest == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->max_norm > 0.0f) {
    int64_t blocks = cdiv(p->total / 4 + 1, OPT_THREADS);
    if (blocks > 4 * (int64_t)sm_count()) blocks = 4 * (int64_t)sm_count();
    if (blocks > OPT_MAX_PARTIALS) blocks = OPT_MAX_PARTIALS;
    sumsq_kernel<<<(unsigned)blocks, OPT_THREADS, 0, st>>>(p->total, p->grads_flat, p->scratch);
    VQA_TRY(check_launch("sumsq"));
    sumsq_final_kernel<<<1, OPT_THREADS, 0, st>>>((int)blocks, p->scratch);
    VQA_TRY(check_launch("sumsq_final"));
  }
  if (p->step_dev) {
    adam_tick_kernel<<<1, 1, 0, st>>>(p->step_dev, p->lr_dev, p->lr_gamma);
    VQA_TRY(check_launch("adam_tick"));
  }
  const double bc1 = p->step_dev ? 1.0 : 1.0 - pow((double)p->beta1, (double)p->step);
  const double bc2 = p->step_dev ? 1.0 : 1.0 - pow((double)p->beta2, (double)p->step);
  int64_t blocks = cdiv(biggest, OPT_THREADS * 4);
  if (blocks > 1024) blocks = 1024;
  clip_adam_kernel<<<dim3((unsigned)blocks, (unsigned)p->nsegs), OPT_THREADS, 0, st>>>(
      tab, p->grads_flat, p->exp_avg, p->exp_avg_sq, p->scratch, p->max_norm, p->lr, p->beta1, p->beta2, p->eps,
      (float)bc1, (float)sqrt(bc2), p->write_clipped_grads, p->step_dev, p->lr_dev);
  return check_launch("clip_adam");
}
