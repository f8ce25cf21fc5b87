// Epilogue functors shared by the two tcgen05 GEMM kernels (tc_gemm.cuh: fp32 operands, TF32 math, in-kernel operand
// transform; tc_gemm16.cuh: bf16 operand planes).  A "staged" functor sees the accumulator tile row-major through
// shared memory: a thread owns one group of 4 consecutive columns (`column`), issues every global read of a batch of
// rows first (`preload`) and then combines and stores (`row4`).  The transposed functor (wgrad) owns 4 consecutive tile
// ROWS (`rowquad`) and adds one column at a time (`col4`).
#pragma once
#include <type_traits>

#include <cuda_bf16.h>

#include "tc_gemm.cuh"

namespace vqa {
namespace tc {

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): |x - hi - lo| <= 2^-16 |x| (two round-to-nearest steps of 2^-8)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// 4 consecutive values -> 4 bf16 hi (8 bytes) and 4 bf16 lo (8 bytes)
__device__ __forceinline__ void split4_bf16(const float (&o)[4], uint2& hi, uint2& lo) {
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_bf16(o[e], h[e], l[e]);
  hi.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
  hi.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
  lo.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
  lo.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
}

// Each receives 32 consecutive accumulator columns [n0, n0+32) of row m.

// 4 consecutive row elements with whatever vector width the destination alignment allows; nv = valid count
__device__ __forceinline__ void store4(float* dst, const float (&o)[4], int nv) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
  if (nv == 4 && (a & 15) == 0) {
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
  } else if (nv == 4 && (a & 7) == 0) {
    *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
    *reinterpret_cast<float2*>(dst + 2) = make_float2(o[2], o[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < nv) dst[e] = o[e];
  }
}
// o[0..nv) added to dst[0..nv) with reductions that return nothing: one 16-byte red when the quad is whole and aligned
__device__ __forceinline__ void red_add4(float* dst, const float (&o)[4], int nv) {
  if (nv == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3])
                 : "memory");
  } else if (nv == 4 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(dst), "f"(o[0]), "f"(o[1]) : "memory");
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(dst + 2), "f"(o[2]), "f"(o[3]) : "memory");
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < nv) atomicAdd(dst + e, o[e]);
  }
}
__device__ __forceinline__ void load4(const float* src, float (&o)[4], int nv) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  if (nv == 4 && (a & 15) == 0) {
    const float4 t = *reinterpret_cast<const float4*>(src);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else if (nv == 4 && (a & 7) == 0) {
    const float2 t0 = *reinterpret_cast<const float2*>(src), t1 = *reinterpret_cast<const float2*>(src + 2);
    o[0] = t0.x; o[1] = t0.y; o[2] = t1.x; o[3] = t1.y;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = e < nv ? src[e] : 0.0f;
  }
}

// y = act(acc + bias), row-major store.  With k-splits (atomic != 0) the partial sums are accumulated into a
// zeroed Y with red.global.add, split 0 contributes the bias, and the activation is applied afterwards by
// act_inplace_kernel.
struct EpiBiasAct {
  static constexpr bool kOcc2 = false;
  static constexpr bool kStaged = true;
  static constexpr int kBatch = 8;
  float* Y[MAXG];
  const float* bias[MAXG];
  int64_t ld[MAXG];
  int act;
  int atomic;
  // optional: the result also as bf16 operand planes for the next GEMM (tc_gemm16.cuh): plane 0 = bf16(y), plane 1 =
  // bf16(y - plane 0) when planes == 2; [planes][rows][ldp], planes plane_stride elements apart.  Non-atomic only.
  __nv_bfloat16* Yp[MAXG];
  int64_t ldp, plane_stride;
  int planes;
  struct Col { float* y; int64_t ld; float b[4]; int nv; __nv_bfloat16* yp; };
  struct Pre {};
  __device__ __forceinline__ void column(Col& c, int g, int split, int n, int N) const {
    c.y = Y[g] + n; c.ld = ld[g]; c.nv = N - n < 4 ? N - n : 4;
    c.yp = Yp[g] ? Yp[g] + n : nullptr;
    const float* bp = bias[g];
#pragma unroll
    for (int e = 0; e < 4; ++e) c.b[e] = (bp && e < c.nv && (!atomic || split == 0)) ? __ldg(bp + n + e) : 0.0f;
  }
  __device__ __forceinline__ void preload(Pre&, const Col&, int) const {}
  __device__ __forceinline__ void row4(const Col& c, int m, const float4 v, const Pre&) const {
    float* y = c.y + (int64_t)m * c.ld;
    float o[4] = {v.x + c.b[0], v.y + c.b[1], v.z + c.b[2], v.w + c.b[3]};
    if (atomic) {
      red_add4(y, o, c.nv);
      return;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = e < c.nv ? act_apply(act, o[e]) : 0.0f;
    store4(y, o, c.nv);
    if (c.yp) {                      // pad columns (>= N) of a quad are written as zeros
      uint2 hi, lo;
      split4_bf16(o, hi, lo);
      __nv_bfloat16* d = c.yp + (int64_t)m * ldp;
      *reinterpret_cast<uint2*>(d) = hi;
      if (planes == 2) *reinterpret_cast<uint2*>(d + plane_stride) = lo;
    }
  }
};

// wgrad: D'(m' = input feature k, n' = output feature n) accumulated into dW[n, k] (row stride ldw) with
// red.global.add (split-K partials and the "+=" of a flat gradient buffer are the same operation).
struct EpiWgradT {
  static constexpr bool kOcc2 = false;
  static constexpr bool kStaged = false;
  static constexpr int kBatch = 8;
  float* dW[MAXG];
  int64_t ldw;
  struct Row { float* w; int nv; };
  __device__ __forceinline__ void rowquad(Row& r, int g, int m, int M) const {
    r.w = dW[g] ? dW[g] + m : nullptr;             // consecutive tile rows m are consecutive addresses of dW[n, :]
    r.nv = M - m < 4 ? M - m : 4;
  }
  __device__ __forceinline__ void col4(const Row& r, int n, const float4 v) const {
    if (!r.w) return;
    const float o[4] = {v.x, v.y, v.z, v.w};
    red_add4(r.w + (int64_t)n * ldw, o, r.nv);
  }
};

// dgrad: dX[m, n] (=|+=) acc * mask(m*drop_ld + n) / (1-p)
// POOL: additionally dX[m, n] += sum_g alpha[m, g] * dpooled[m / regions, g, n] — the gradient of an attention pooling
// over the same X (MyATT's bmatmul over v2, config/CoR2.py), which would otherwise cost a full write of dX by the
// pooling backward and a read-modify-write here.
template <bool POOL>
struct EpiDgradT {
  static constexpr bool kOcc2 = false;
  static constexpr bool kStaged = true;
  static constexpr int kBatch = POOL ? 4 : 8;
  float* dX[MAXG];
  int64_t ld[MAXG];
  int accumulate;
  int atomic;       // k-splits: every partial is masked and added with red.global.add (dX zeroed or "+=")
  int drop_on;
  Drop drop;
  GroupDrop gd;
  int64_t drop_ld;
  int wide_bits;               // drop_ld % 4 != 0: a quad's mask bits may run into the next byte
  const uint8_t* bits[MAXG];
  const float* pool_alpha;      // [M, 4]
  const float* pool_dp;         // [M / pool_regions, 4, pool_ld]
  int64_t pool_regions, pool_ld;
  struct Col { float* x; int64_t ld; int n, nv, first; const uint8_t* bits; const float* dp; uint32_t layer; uint64_t base; };
  struct PreBase { float old[4]; uint32_t byte, byte_hi; };
  struct PrePool : PreBase { float4 al; float4 dp[4]; };
  using Pre = typename std::conditional<POOL, PrePool, PreBase>::type;
  __device__ __forceinline__ void column(Col& c, int g, int split, int n, int N) const {
    c.first = split == 0;
    c.x = dX[g] ? dX[g] + n : nullptr; c.ld = ld[g]; c.n = n; c.nv = N - n < 4 ? N - n : 4;
    c.bits = drop_on ? bits[g] : nullptr; c.layer = gd.layer[g]; c.base = gd.base[g];
    c.dp = POOL ? pool_dp + n : nullptr;
  }
  __device__ __forceinline__ void preload(Pre& r, const Col& c, int m) const {
    if (!c.x) return;
    if (c.bits) {            // the quad's 4 mask bits start at bit (e & 7) and may run into the next byte
      const uint64_t e = (uint64_t)m * (uint64_t)drop_ld + (uint64_t)c.n;
      r.byte = __ldg(c.bits + (e >> 3));
      if (wide_bits) r.byte_hi = __ldg(c.bits + (e >> 3) + 1);      // combined at use: no load is waited for here
    }
    if (accumulate && !atomic) load4(c.x + (int64_t)m * c.ld, r.old, c.nv);
    if constexpr (POOL) {
      r.al = __ldg(reinterpret_cast<const float4*>(pool_alpha) + m);
      const float* dp = c.dp + (int64_t)((uint32_t)m / (uint32_t)pool_regions) * 4 * pool_ld;   // N % 4 == 0 (host check)
#pragma unroll
      for (int j = 0; j < 4; ++j) r.dp[j] = __ldg(reinterpret_cast<const float4*>(dp + j * pool_ld));
    }
  }
  __device__ __forceinline__ void row4(const Col& c, int m, const float4 v, const Pre& pre) const {
    if (!c.x) return;
    float* x = c.x + (int64_t)m * c.ld;
    float o[4] = {v.x, v.y, v.z, v.w};
    if (c.bits) {
      const uint32_t nb = (wide_bits ? (pre.byte | (pre.byte_hi << 8)) : pre.byte) >>
                          (((uint32_t)m * (uint32_t)drop_ld + (uint32_t)c.n) & 7u);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = ((nb >> j) & 1u) ? o[j] * drop.scale : 0.0f;
    } else if (drop_on) {
      const uint32_t bt = philox_bytes4(drop.key(), c.layer, c.base + ((uint64_t)m * (uint64_t)drop_ld + (uint64_t)c.n));
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = ((bt >> (8 * e)) & 0xFFu) >= drop.thr ? o[e] * drop.scale : 0.0f;
    }
    if constexpr (POOL) {
      if (c.first) {
        const float al[4] = {pre.al.x, pre.al.y, pre.al.z, pre.al.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[0] = fmaf(al[j], pre.dp[j].x, o[0]); o[1] = fmaf(al[j], pre.dp[j].y, o[1]);
          o[2] = fmaf(al[j], pre.dp[j].z, o[2]); o[3] = fmaf(al[j], pre.dp[j].w, o[3]);
        }
      }
    }
    if (atomic) {
      red_add4(x, o, c.nv);
      return;
    }
    if (accumulate) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] += pre.old[e];
    }
    store4(x, o, c.nv);
  }
};
using EpiDgrad = EpiDgradT<false>;

// forward epilogue for rank r = group: h1 = acc + b1_r;  H1_r[m,n] = h1;  Y[m,n] += h1 * H2_r[m / rows_per, n].
// Non-atomic mode runs rank by rank in stream order (r == 0 stores, r > 0 read-modify-writes Y);
// atomic mode (k-splits, small M) accumulates every partial into zeroed Y / H1 with red.global.add.
struct EpiMutan {
  static constexpr bool kOcc2 = true;
  static constexpr bool kStaged = true;
  static constexpr int kBatch = 8;
  const float* bias[MAXG]; const float* H2[MAXG]; float* H1[MAXG]; float* Y;
  int64_t ldh, ldy, rows_per; int accumulate; int atomic;
  struct Col { const float* h2; float* h1; float* y; float b[4]; int nv; };
  struct Pre { float h2[4]; float old[4]; };
  __device__ __forceinline__ void column(Col& c, int g, int split, int n, int N) const {
    c.nv = N - n < 4 ? N - n : 4;
    c.h2 = H2[g] + n; c.h1 = H1[g] ? H1[g] + n : nullptr; c.y = Y + n;
    const float* bp = bias[g];
#pragma unroll
    for (int e = 0; e < 4; ++e) c.b[e] = (bp && e < c.nv && split == 0) ? __ldg(bp + n + e) : 0.0f;
  }
  __device__ __forceinline__ void preload(Pre& r, const Col& c, int m) const {
    load4(c.h2 + (int64_t)((uint32_t)m / (uint32_t)rows_per) * ldh, r.h2, c.nv);
    if (accumulate && !atomic) load4(c.y + (int64_t)m * ldy, r.old, c.nv);
  }
  __device__ __forceinline__ void row4(const Col& c, int m, const float4 v, const Pre& pre) const {
    float h[4] = {v.x + c.b[0], v.y + c.b[1], v.z + c.b[2], v.w + c.b[3]}, o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = e < c.nv ? h[e] * pre.h2[e] : 0.0f;
    float* y = c.y + (int64_t)m * ldy;
    float* h1 = c.h1 ? c.h1 + (int64_t)m * ldh : nullptr;
    if (atomic) {
      if (h1) red_add4(h1, h, c.nv);
      red_add4(y, o, c.nv);
      return;
    }
    if (h1) store4(h1, h, c.nv);
    if (accumulate) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] += pre.old[e];
    }
    store4(y, o, c.nv);
  }
};


}  // namespace tc
}  // namespace vqa
