// Shared device/host helpers for libvqacore_sm100a.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vqacore.h"

namespace vqa {

// ------------------------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int check_launch(const char* what);          // cudaGetLastError -> VQA_ECUDA
int sm_count();

// Times one op of a whole-model plan when vqa_profile_begin() is active (no-op otherwise).
bool prof_active();
struct ProfScope {
  int idx; void* st;
  ProfScope(void* stream, const char* name);
  ~ProfScope();
};

// Per-KERNEL record for bench.py's roofline (only while vqa_profile_begin() is active): "k:<name> <kind>=<work>" with
// kind = hbm (algorithmic bytes of the launch) or flop (algorithmic fp32 CUDA-core flops).  The tensor-core GEMMs
// write their own "k:<what> M.. N.. K.. g.. s.." records.
struct KProf {
  char label[128];
  ProfScope ps;
  static const char* make(char* buf, size_t cap, const char* name, const char* kind, double work) {
    if (!prof_active()) return "k:";
    snprintf(buf, cap, "k:%s %s=%.0f", name, kind, work);
    return buf;
  }
  KProf(void* stream, const char* name, const char* kind, double work)
      : ps(stream, make(label, sizeof(label), name, kind, work)) {}
};

#define VQA_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      vqa::set_error(__VA_ARGS__);        \
      return VQA_EINVAL;                  \
    }                                     \
  } while (0)

#define VQA_TRY(expr)                     \
  do {                                    \
    int _rc = (expr);                     \
    if (_rc != VQA_OK) return _rc;        \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int G = VQA_GLIMPSES;

// ------------------------------------------------------------------------------- Philox4x32-10
// Twin of oracle/philox.py (test infrastructure); the contract is in include/vqacore.h.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// One Philox call covers a GROUP of 16 consecutive linear indices: element idx owns byte (idx & 15) of the
// 128-bit output (little-endian over the words x,y,z,w).  keep(idx) = byte >= thr8, thr8 = floor(p * 256)
// (p = 0.5, the only rate the reference uses, is exact).  Contract: include/vqacore.h; twin: oracle/philox.py.
__device__ __forceinline__ uint4 philox_group(uint64_t seed, uint32_t layer, uint64_t g) {
  return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), layer, 0u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ uint32_t pick_word(const uint4& r, uint32_t i) {
  return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w));
}
__device__ __forceinline__ uint32_t philox_byte(uint64_t seed, uint32_t layer, uint64_t idx) {
  const uint4 r = philox_group(seed, layer, idx >> 4);
  return (pick_word(r, ((uint32_t)idx >> 2) & 3u) >> (8u * ((uint32_t)idx & 3u))) & 0xFFu;
}
// the four bytes (packed, element e in bits [8e, 8e+8)) of 4 consecutive indices idx..idx+3
__device__ __forceinline__ uint32_t philox_bytes4(uint64_t seed, uint32_t layer, uint64_t idx) {
  if ((idx & 3) == 0) return pick_word(philox_group(seed, layer, idx >> 4), ((uint32_t)idx >> 2) & 3u);
  // unaligned start (row length not a multiple of 4): the run straddles two words, possibly two groups
  const uint4 r0 = philox_group(seed, layer, idx >> 4);
  const uint64_t last = idx + 3;
  const uint4 r1 = (last >> 4) == (idx >> 4) ? r0 : philox_group(seed, layer, last >> 4);
  uint32_t out = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint64_t i = idx + e;
    const uint4& r = (i >> 4) == (idx >> 4) ? r0 : r1;
    out |= ((pick_word(r, ((uint32_t)i >> 2) & 3u) >> (8u * ((uint32_t)i & 3u))) & 0xFFu) << (8 * e);
  }
  return out;
}

// Device-side view of one dropout call site.
struct Drop {
  uint64_t seed;
  const uint64_t* seed_ptr;   // when non-null the key is read from device memory (CUDA-graph replay, fresh mask per step)
  uint64_t base;      // added to the element index
  uint32_t layer;
  uint32_t thr;       // keep iff byte >= thr (0..255)
  float scale;        // 1/(1-p)
  int on;
  __device__ __forceinline__ uint64_t key() const { return seed_ptr ? __ldg(seed_ptr) : seed; }
  // multiplier (0 or scale) for logical element idx
  __device__ __forceinline__ float mul(uint64_t idx) const {
    if (!on) return 1.0f;
    return philox_byte(key(), layer, base + idx) >= thr ? scale : 0.0f;
  }
  // multipliers for 4 consecutive elements idx..idx+3
  __device__ __forceinline__ void mul4(uint64_t idx, float (&m)[4]) const {
    if (!on) { m[0] = m[1] = m[2] = m[3] = 1.0f; return; }
    const uint32_t b = philox_bytes4(key(), layer, base + idx);
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = ((b >> (8 * e)) & 0xFFu) >= thr ? scale : 0.0f;
  }
};

static inline uint32_t drop_threshold(float p) {
  double t = (double)p * 256.0;
  if (t < 0) t = 0;
  if (t > 255.0) t = 255.0;
  return (uint32_t)t;   // floor(p * 256)
}

static inline Drop make_drop(float p, uint64_t seed, uint32_t layer, uint64_t base, int train_on = 1,
                             const uint64_t* seed_ptr = nullptr) {
  Drop d;
  d.seed = seed;
  d.seed_ptr = seed_ptr;
  d.base = base;
  d.layer = layer;
  d.on = (train_on && p > 0.0f) ? 1 : 0;
  d.thr = drop_threshold(p);
  d.scale = d.on ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}

// ------------------------------------------------------------------------------- small device utils
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float act_apply(int act, float z) {
  if (act == VQA_ACT_RELU) return fmaxf(z, 0.0f);
  if (act == VQA_ACT_SIGMOID) return 1.0f / (1.0f + __expf(-z));
  return z;
}
// derivative expressed through the OUTPUT y = act(z)
__device__ __forceinline__ float act_grad(int act, float y) {
  if (act == VQA_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  if (act == VQA_ACT_SIGMOID) return y * (1.0f - y);
  return 1.0f;
}

// streaming 128-bit load that does not allocate in L1 (data read once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace vqa
