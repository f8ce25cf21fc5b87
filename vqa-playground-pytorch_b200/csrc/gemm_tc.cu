// Host side of the tcgen05 GEMM family: TMA tensor-map construction, epilogue functors, launch plumbing
// for the grouped linear forward / wgrad / dgrad (include/vqacore.h: vqa_linear_fwd / vqa_linear_bwd with
// math = VQA_MATH_TF32X3 or VQA_MATH_TF32).  Kernel: tc_gemm.cuh.
#include "gemm_tc.h"

#include <stdlib.h>

#include <mutex>


#include "tc_epilogues.cuh"

namespace vqa {
namespace tc {

// ------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D fp32 tensor [rows, cols] row-major with row stride ld (elements); box = box_cols x box_rows, 128B swizzle.
// Out-of-bounds parts of a box are zero-filled, which is what pads ragged M / N / K edges.
static int make_tmap(CUtensorMap* out, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                     int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VQA_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for ptr=%p rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
              (const void*)ptr, (long long)rows, (long long)cols, (long long)ld, box_cols, box_rows);
    return VQA_ECUDA;
  }
  return VQA_OK;
}

static inline bool tma_ok(const void* ptr, int64_t ld) {
  return (reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && (ld % 4 == 0);
}

// Thread-block clusters with TMA multicast are implemented (tc_gemm.cuh) but OFF by default: on this path they
// measured slower (the tiles must then be fetched as 4 KB boxes and the kernel is bound by DRAM row locality of the
// strided A tiles rather than by L2 bandwidth).  VQA_TC_CM / VQA_TC_CN >= 2 enable them for experiments.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int max_cm() { static int v = env_int("VQA_TC_CM", 1); return v; }
static int max_cn() { static int v = env_int("VQA_TC_CN", 1); return v; }
static bool cluster_enabled() { return max_cm() > 1 || max_cn() > 1; }

// K-major operand: source is [rows(M or N), K]; MN-major operand: source is [K, rows(M or N)].
static int operand_tmap(CUtensorMap* out, const float* ptr, bool mn_major, int64_t mn_extent, int64_t k_extent,
                        int64_t ld, int tile_mn) {
  if (!mn_major) return make_tmap(out, ptr, mn_extent, k_extent, ld, BK, cluster_enabled() ? 32 : tile_mn, CU_TENSOR_MAP_SWIZZLE_128B);
  return make_tmap(out, ptr, k_extent, mn_extent, ld, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

// y = act(y) in place over a [M, N] window of row stride ld; grid = (column chunks, row chunks, groups): no index
// divisions, 128-bit accesses where the window allows them
struct ActArgs { float* Y[MAXG]; int64_t ld[MAXG]; };
__global__ void __launch_bounds__(256) act_inplace_kernel(ActArgs a, int64_t M, int64_t N, int act, int rows_per_cta) {
  float* __restrict__ y = a.Y[blockIdx.z];
  const int64_t ld = a.ld[blockIdx.z];
  const int64_t n = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  const bool vec = n + 4 <= N && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
  for (int64_t m = r0; m < r1; ++m) {
    float* p = y + m * ld + n;
    if (vec) {
      float4 v = *reinterpret_cast<float4*>(p);
      v.x = act_apply(act, v.x); v.y = act_apply(act, v.y); v.z = act_apply(act, v.z); v.w = act_apply(act, v.w);
      *reinterpret_cast<float4*>(p) = v;
    } else {
      for (int e = 0; e < 4 && n + e < N; ++e) p[e] = act_apply(act, p[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------ launch
static int occ2_mode() {                 // VQA_TC_OCC2=0 disables the two-CTAs-per-SM variant (experiments)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VQA_TC_OCC2");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

template <int BN, bool X3, class Epi, int OCC>
static int launch_occ(const Params<Epi>& p, int groups, cudaStream_t st, const char* what) {
  using C = Cfg<BN, X3, OCC>;
  auto kern = tc_gemm_kernel<BN, X3, Epi, OCC>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
    return check_launch(what);
  if (OCC == 2) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  dim3 grid((unsigned)cdiv(p.M, BM), (unsigned)cdiv(p.N, BN), (unsigned)(groups * p.k_splits));
  // cluster = CM adjacent m-tiles x CN adjacent n-tiles sharing operand tiles by TMA multicast
  unsigned cm = 1, cn = 1;
  for (unsigned c = 4; c >= 2; c >>= 1) if ((int)c <= max_cm() && grid.x % c == 0) { cm = c; break; }
  if (max_cn() >= 2 && grid.y % 2 == 0) cn = 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cm; attr[0].val.clusterDim.y = cn; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // per-kernel record for bench.py's roofline ("k:" entries of vqa_profile_end carry the GEMM shape)
  char label[160] = "k:";
  if (prof_active())
    snprintf(label, sizeof(label), "k:%s M%d N%d K%d g%d s%d", what, p.M, p.N, p.K, groups, p.k_splits);
  ProfScope ps_(st, label);
  if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) return check_launch(what);
  return check_launch(what);
}

template <int BN, bool X3, class Epi>
static int launch_cfg(const Params<Epi>& p, int groups, cudaStream_t st, const char* what) {
  // more than one wave of CTAs: two smaller CTAs per SM overlap each other's prologue / epilogue
  // (measured: Mutan forward -8 us per launch; the dgrad epilogues spill at 102 registers and lose 50 us, so the
  // variant is only built for the epilogues that declare kOcc2)
  if constexpr (Epi::kOcc2) {
    const int64_t ctas = cdiv(p.M, BM) * cdiv(p.N, BN) * groups * p.k_splits;
    if (ctas > sm_count() && occ2_mode() && max_cm() <= 1 && max_cn() <= 1)
      return launch_occ<BN, X3, Epi, 2>(p, groups, st, what);
  }
  return launch_occ<BN, X3, Epi, 1>(p, groups, st, what);
}

static inline int pick_bn(int64_t N) {
  const int64_t w160 = cdiv(N, 160) * 160 - N, w128 = cdiv(N, 128) * 128 - N;
  return w160 < w128 ? 160 : 128;
}

static void zero_window(float* y, int64_t ld, int64_t M, int64_t N, cudaStream_t st) {
  if (ld == N) cudaMemsetAsync(y, 0, (size_t)M * N * sizeof(float), st);
  else cudaMemset2DAsync(y, (size_t)ld * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st);
}

static int rewrite_hi_flag() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VQA_TC_REWRITE_HI");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

// k-splits for problems whose output tiles alone cannot fill the chip.  One CTA per SM is resident (shared
// memory), so the CTA count should not spill into a mostly empty second wave: tiles * splits <= #SMs.
static int pick_splits(int64_t tiles, int64_t K) {
  const int64_t kb = cdiv(K, BK);
  const int64_t sms = sm_count();
  if (tiles * 2 > sms) return 1;
  int64_t s = sms / tiles;
  if (s > kb / 4) s = kb / 4;
  return (int)(s < 1 ? 1 : s);
}

// wgrad always reduces over the (long) row dimension: fill one wave exactly, never a partial second one
static int pick_splits_wgrad(int64_t tiles, int64_t K) {
  const int64_t kb = cdiv(K, BK);
  const int64_t sms = sm_count();
  int64_t s = tiles >= sms ? 1 : sms / tiles;
  if (s > kb / 4) s = kb / 4;
  return (int)(s < 1 ? 1 : s);
}

template <class Epi>
static int launch(Params<Epi> p, int groups, bool x3, cudaStream_t st, const char* what) {
  p.rewrite_hi = rewrite_hi_flag();
  p.box_split = cluster_enabled() ? 1 : 0;
  {
    const char* e = getenv("VQA_TC_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  const int bn = pick_bn(p.N);
  if (bn == 160) return x3 ? launch_cfg<160, true>(p, groups, st, what) : launch_cfg<160, false>(p, groups, st, what);
  return x3 ? launch_cfg<128, true>(p, groups, st, what) : launch_cfg<128, false>(p, groups, st, what);
}

// the keep-bit cache of a group is indexed from element 0 of X_g: only usable when the group's mask is, too
static inline const uint8_t* cached_bits(const uint8_t* const* bits, const uint64_t* base, int g) {
  return base[g] == 0 ? bits[g] : nullptr;
}

static void fill_drop(Drop& d, GroupDrop& gd, float pdrop, uint64_t seed, const uint32_t* layer, const uint64_t* base,
                      int groups, const uint64_t* seed_dev = nullptr) {
  d = make_drop(pdrop, seed, 0, 0, 1, seed_dev);
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < groups ? g : 0;
    gd.layer[g] = layer[s];
    gd.base[g] = base[s];
  }
}

// dZ = dY (.) act'(Y) into a padded [M, ldz] buffer (pad columns zero) and db (+)= colsum(dZ).
// grid = (cdiv(N,32), row chunks, groups); 256 threads = 8 rows x 32 columns.
struct DzArgs {
  const float* dY[MAXG]; const float* Y[MAXG]; float* dZ[MAXG]; float* db[MAXG];
  int64_t lddy[MAXG], ldy[MAXG];
};
__global__ void __launch_bounds__(256)
dz_colsum_kernel(DzArgs a, int64_t M, int64_t N, int64_t ldz, int act, int rows_per_cta) {
  __shared__ float red[8][33];
  const int g = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  const float* dy = a.dY[g];
  const float* y = a.Y[g];
  float* dz = a.dZ[g];
  float s = 0.0f;
  for (int64_t m = r0 + ty; m < r1; m += 8) {
    float v = 0.0f;
    if (n < N) {
      v = dy[m * a.lddy[g] + n];
      if (act != VQA_ACT_NONE) v *= act_grad(act, y[m * a.ldy[g] + n]);
    }
    if (dz && n < ldz) dz[m * ldz + n] = v;
    s += v;
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N && a.db[g]) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(a.db[g] + n, t);
  }
}

// Padded copy of weights whose row stride (K*4 bytes) is not a multiple of 16 and therefore cannot be
// addressed by TMA: dst[g][r, 0..Kp) = src[g][r, 0..K) | 0 for r < rows, zero rows up to rows_pad.
// grid = (blocks, groups)
struct PackArgs { const float* src[MAXG]; int64_t ld[MAXG]; };
__global__ void pack_rows_kernel(PackArgs a, float* __restrict__ dst, int64_t rows, int64_t rows_pad, int64_t K,
                                 int64_t Kp) {
  const int g = blockIdx.y;
  const float* src = a.src[g];
  const int64_t ld = a.ld[g];
  float* d = dst + (size_t)g * rows_pad * Kp;
  const int64_t total = rows_pad * Kp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / Kp, k = t - r * Kp;
    d[t] = (k < K && r < rows) ? src[r * ld + k] : 0.0f;
  }
}

static inline int64_t roundup(int64_t x, int64_t m) { return cdiv(x, m) * m; }
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int pack_weights(const float* const* W, int groups, int64_t rows, int64_t rows_pad, int64_t K, float* dst,
                        cudaStream_t st, const int64_t* ld = nullptr) {
  PackArgs a = {};
  for (int g = 0; g < MAXG; ++g) { a.src[g] = W[g < groups ? g : 0]; a.ld[g] = ld ? ld[g < groups ? g : 0] : K; }
  const int64_t Kp = roundup(K, 4);
  int64_t blocks = cdiv(rows_pad * Kp, 256);
  if (blocks > 4096) blocks = 4096;
  pack_rows_kernel<<<dim3((unsigned)blocks, (unsigned)groups), 256, 0, st>>>(a, dst, rows, rows_pad, K, Kp);
  return check_launch("pack_rows");
}

// One launch for a list of weights (vqa_pack_weights). grid = (blocks, nsegs)
struct PackSegs { vqa_pack_segment s[VQA_MAX_PACK_SEGMENTS]; };
__global__ void pack_segments_kernel(PackSegs a) {
  const vqa_pack_segment sg = a.s[blockIdx.y];
  const int64_t Kp = (sg.K + 3) / 4 * 4;
  const int64_t total = sg.rows_pad * Kp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / Kp, k = t - r * Kp;
    sg.dst[t] = (k < sg.K && r < sg.rows) ? sg.src[r * sg.K + k] : 0.0f;
  }
}

// bit j of the result = (byte j of w >= threshold): per-byte compare, then the four byte MSBs are gathered with one
// multiply (bit 7 -> 28, 15 -> 29, 23 -> 30, 31 -> 31; the partial products never overlap)
__device__ __forceinline__ uint32_t keep_nibble(uint32_t w, uint32_t thr4) {
  return ((__vcmpgeu4(w, thr4) & 0x80808080u) * 0x00204081u) >> 28;
}

// keep-bits: thread per group of 16 elements -> 2 bytes
__global__ void dropout_bits_kernel(uint64_t seed, const uint64_t* seed_ptr, uint32_t layer, uint32_t thr,
                                    uint64_t ngroups, uint16_t* __restrict__ out) {
  if (seed_ptr) seed = __ldg(seed_ptr);
  const uint32_t thr4 = thr * 0x01010101u;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox_group(seed, layer, g);
    out[g] = (uint16_t)(keep_nibble(r.x, thr4) | (keep_nibble(r.y, thr4) << 4) | (keep_nibble(r.z, thr4) << 8) |
                        (keep_nibble(r.w, thr4) << 12));
  }
}

struct BitSegs {
  vqa_bits_segment s[VQA_MAX_BITS_SEGMENTS];
  uint64_t first[VQA_MAX_BITS_SEGMENTS + 1];     // running count of 16-element groups: segment i owns [first[i], first[i+1])
  int n;
};
// one thread per 16-element group of the concatenated segments
__global__ void dropout_bits_batch_kernel(uint64_t seed, const uint64_t* seed_ptr, uint32_t thr, BitSegs a) {
  if (seed_ptr) seed = __ldg(seed_ptr);
  const uint32_t thr4 = thr * 0x01010101u;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.first[a.n]; t += (uint64_t)gridDim.x * blockDim.x) {
    int i = 0;
    while (t >= a.first[i + 1]) ++i;
    const vqa_bits_segment sg = a.s[i];
    const uint64_t g = t - a.first[i];
    uint16_t* out = reinterpret_cast<uint16_t*>(sg.out);
    const uint4 r = philox_group(seed, sg.layer, g);
    out[g] = (uint16_t)(keep_nibble(r.x, thr4) | (keep_nibble(r.y, thr4) << 4) | (keep_nibble(r.z, thr4) << 8) |
                        (keep_nibble(r.w, thr4) << 12));
  }
}

// ------------------------------------------------------------------------------------------ Mutan pieces
struct DbTable { float* p[MAXG]; };

// Both Mutan gradient operands in one pass over dY (one CTA per (h2 row, rank)):
//   dH1cat[m, r*Fp + f]  = dY[m,f] * H2[r, mh, f]                     (zero in the pad columns f >= F)
//   dH2cat[mh, r*Fp + f] = sum_{j<rows_per} dY[mh*rp+j, f] * H1[r, mh*rp+j, f]
//   db1_r[f] += sum_j dH1cat,  db2_r[f] += dH2cat
// VEC = 2: columns handled in pairs (needs F, lddy even and 8-byte aligned bases), else one column per thread.
template <int VEC>
__global__ void __launch_bounds__(256)
mutan_dh_kernel(int64_t M, int64_t F, int64_t Fp, int64_t rows_per, int R, const float* __restrict__ dY, int64_t lddy,
                const float* __restrict__ H1, const float* __restrict__ H2, float* __restrict__ dH1cat,
                float* __restrict__ dH2cat, DbTable db1, DbTable db2, __nv_bfloat16* __restrict__ dH1p,
                int64_t dh1_plane, int np) {
  const int64_t mh = blockIdx.x;
  const int r = blockIdx.y;
  const int64_t Mh = M / rows_per, RF = (int64_t)R * Fp;
  const float* h1r = H1 + (int64_t)r * M * F;
  constexpr int U = 6;
  for (int64_t f = (int64_t)threadIdx.x * VEC; f < Fp; f += 256 * VEC) {
    const bool ok = f < F;
    float h2[VEC], acc2[VEC], accb[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      h2[e] = ok ? H2[((int64_t)r * Mh + mh) * F + f + e] : 0.0f;
      acc2[e] = 0.0f; accb[e] = 0.0f;
    }
    for (int64_t j0 = 0; j0 < rows_per; j0 += U) {
      float dy[U][VEC], h1[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t m = mh * rows_per + j0 + u;
        const bool live = ok && j0 + u < rows_per;
        if (VEC == 2) {
          const float2 a = live ? *reinterpret_cast<const float2*>(dY + m * lddy + f) : make_float2(0.f, 0.f);
          const float2 b = live ? *reinterpret_cast<const float2*>(h1r + m * F + f) : make_float2(0.f, 0.f);
          dy[u][0] = a.x; dy[u][VEC - 1] = a.y; h1[u][0] = b.x; h1[u][VEC - 1] = b.y;
        } else {
          dy[u][0] = live ? dY[m * lddy + f] : 0.0f;
          h1[u][0] = live ? h1r[m * F + f] : 0.0f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u >= rows_per) break;
        const int64_t m = mh * rows_per + j0 + u;
        float v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          v[e] = dy[u][e] * h2[e];
          accb[e] += v[e];
          acc2[e] = fmaf(dy[u][e], h1[u][e], acc2[e]);
        }
        const int64_t oi = m * RF + (int64_t)r * Fp + f;
        if (dH1p) {                 // bf16 operand planes for the bf16-plane GEMMs instead of the fp32 tensor
          __nv_bfloat16 hi[VEC], lo[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) split_bf16(v[e], hi[e], lo[e]);
          if (VEC == 2) {
            *reinterpret_cast<__nv_bfloat162*>(dH1p + oi) = __halves2bfloat162(hi[0], hi[VEC - 1]);
            if (np == 2) *reinterpret_cast<__nv_bfloat162*>(dH1p + dh1_plane + oi) = __halves2bfloat162(lo[0], lo[VEC - 1]);
          } else {
            dH1p[oi] = hi[0];
            if (np == 2) dH1p[dh1_plane + oi] = lo[0];
          }
        } else {
          float* o = dH1cat + oi;
          if (VEC == 2) *reinterpret_cast<float2*>(o) = make_float2(v[0], v[VEC - 1]);
          else o[0] = v[0];
        }
      }
    }
    float* o2 = dH2cat + mh * RF + (int64_t)r * Fp + f;
    if (VEC == 2) *reinterpret_cast<float2*>(o2) = make_float2(acc2[0], acc2[VEC - 1]);
    else o2[0] = acc2[0];
    if (ok) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (db1.p[r]) atomicAdd(db1.p[r] + f + e, accb[e]);
        if (db2.p[r]) atomicAdd(db2.p[r] + f + e, acc2[e]);
      }
    }
  }
}

}  // namespace tc

// ============================================================================================ linear fwd
// Operands TMA cannot address (row stride not a multiple of 16 bytes: K = 310 / 510 activations handed over as
// contiguous [M, K] tensors) are copied into padded scratch first — the tensor-core path is never silently
// replaced by the CUDA-core one; what cannot be packed either is an error.
static int tc_fail(const char* who, const char* why) {
  set_error("%s: a tensor-core math mode was requested but %s (there is no silent CUDA-core fallback; use "
            "math = VQA_MATH_FP32_SIMT explicitly)", who, why);
  return VQA_EINVAL;
}

int tc_linear_fwd(const vqa_linear_fwd_params* p, cudaStream_t st, const LinExt* ext) {
  using namespace tc;
  if (is_bf16_math(p->math)) {
    if (p->M >= TC16_MIN_M) return tc16_linear_fwd(p, st, ext);
    vqa_linear_fwd_params q = *p;
    q.math = small_math_of(p->math);
    return tc_linear_fwd(&q, st, nullptr);
  }
  if (p->math != VQA_MATH_TF32X3 && p->math != VQA_MATH_TF32) return tc_fail("vqa_linear_fwd", "this math mode is not built for this op");
  if (p->M > INT32_MAX || p->K > INT32_MAX || p->N > INT32_MAX) return tc_fail("vqa_linear_fwd", "a dimension exceeds 2^31");
  bool pack = false, packx = false;
  for (int g = 0; g < p->groups; ++g) {
    packx |= !tma_ok(p->X[g], p->ldx[g]);
    pack |= !tma_ok(p->W[g], p->K);
  }
  const int64_t Kp = roundup(p->K, 4);
  bool prepacked = pack;
  for (int g = 0; g < p->groups; ++g) prepacked &= p->Wp[g] != nullptr;
  const size_t w_bytes = (pack && !prepacked) ? align256((size_t)p->groups * p->N * Kp * sizeof(float)) : 0;
  const size_t x_bytes = packx ? align256((size_t)p->groups * p->M * Kp * sizeof(float)) : 0;
  if (w_bytes + x_bytes) {
    if (!p->workspace || p->workspace_bytes < w_bytes + x_bytes || reinterpret_cast<uintptr_t>(p->workspace) % 16 != 0)
      return tc_fail("vqa_linear_fwd", "the workspace is missing, misaligned or smaller than vqa_linear_fwd_workspace_bytes()");
  }
  float* wpk = reinterpret_cast<float*>(p->workspace);
  float* xpk = reinterpret_cast<float*>(reinterpret_cast<char*>(p->workspace) + w_bytes);
  if (w_bytes) VQA_TRY(pack_weights(p->W, p->groups, p->N, p->N, p->K, wpk, st));
  if (x_bytes) VQA_TRY(pack_weights(p->X, p->groups, p->M, p->M, p->K, xpk, st, p->ldx));
  const int bn = pick_bn(p->N);
  Params<EpiBiasAct> q = {};
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < p->groups ? g : 0;
    if (packx) VQA_TRY(operand_tmap(&q.tmA[g], xpk + (size_t)s * p->M * Kp, false, p->M, p->K, Kp, BM));
    else VQA_TRY(operand_tmap(&q.tmA[g], p->X[s], false, p->M, p->K, p->ldx[s], BM));
    if (prepacked) VQA_TRY(operand_tmap(&q.tmB[g], p->Wp[s], false, p->N, p->K, Kp, bn));
    else if (pack) VQA_TRY(operand_tmap(&q.tmB[g], wpk + (size_t)s * p->N * Kp, false, p->N, p->K, Kp, bn));
    else VQA_TRY(operand_tmap(&q.tmB[g], p->W[s], false, p->N, p->K, p->K, bn));
    q.epi.Y[g] = p->Y[s]; q.epi.bias[g] = p->b[s]; q.epi.ld[g] = p->ldy[s];
  }
  q.M = (int)p->M; q.N = (int)p->N; q.K = (int)p->K; q.a_mn = 0; q.b_mn = 0;
  q.k_splits = pick_splits(cdiv(p->M, BM) * cdiv(p->N, bn) * p->groups, p->K);
  q.drop_on = p->p > 0.0f;
  fill_drop(q.drop, q.gd, p->p, p->seed, p->layer, p->drop_index_base, p->groups, p->seed_dev);
  q.drop_ld = p->K; q.drop_rows = p->M;
  for (int g = 0; g < MAXG; ++g) q.drop_bits[g] = cached_bits(p->drop_bits, p->drop_index_base, g < p->groups ? g : 0);
  q.epi.act = p->act;
  q.epi.atomic = q.k_splits > 1;
  if (q.epi.atomic)
    for (int g = 0; g < p->groups; ++g) zero_window(p->Y[g], p->ldy[g], p->M, p->N, st);
  VQA_TRY(launch(q, p->groups, p->math == VQA_MATH_TF32X3, st, "tc_linear_fwd"));
  if (q.epi.atomic && p->act != VQA_ACT_NONE) {
    ActArgs a = {};
    for (int g = 0; g < MAXG; ++g) { const int s = g < p->groups ? g : 0; a.Y[g] = p->Y[s]; a.ld[g] = p->ldy[s]; }
    const int rows_per_cta = 8;
    act_inplace_kernel<<<dim3((unsigned)cdiv(p->N, 1024), (unsigned)cdiv(p->M, rows_per_cta), (unsigned)p->groups), 256, 0, st>>>(
        a, p->M, p->N, p->act, rows_per_cta);
    VQA_TRY(check_launch("tc_linear_fwd.act"));
  }
  return VQA_OK;
}

// ============================================================================================ linear bwd
template <bool POOL>
static int dgrad_launch(const vqa_linear_bwd_params* p, float* dz, float* wpk, int64_t ldz, int64_t Kp, bool pack,
                        bool prepacked, bool x3, cudaStream_t st) {
  using namespace tc;
  Params<EpiDgradT<POOL>> q = {};
  const int bn = pick_bn(p->K);
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < p->groups ? g : 0;
    VQA_TRY(operand_tmap(&q.tmA[g], dz + (size_t)s * p->M * ldz, false, p->M, p->N, ldz, BM));
    if (prepacked) VQA_TRY(operand_tmap(&q.tmB[g], p->Wp[s], true, p->K, p->N, Kp, bn));
    else if (pack) VQA_TRY(operand_tmap(&q.tmB[g], wpk + (size_t)s * p->N * Kp, true, p->K, p->N, Kp, bn));
    else VQA_TRY(operand_tmap(&q.tmB[g], p->W[s], true, p->K, p->N, p->K, bn));
    q.epi.dX[g] = p->dX[s]; q.epi.ld[g] = p->lddx[s];
  }
  q.M = (int)p->M; q.N = (int)p->K; q.K = (int)p->N; q.a_mn = 0; q.b_mn = 1;
  q.k_splits = pick_splits(cdiv(p->M, BM) * cdiv(p->K, bn) * p->groups, p->N);
  q.drop_on = 0;
  fill_drop(q.drop, q.gd, 0.0f, p->seed, p->layer, p->drop_index_base, p->groups);
  q.epi.accumulate = p->accumulate_x;
  q.epi.atomic = q.k_splits > 1;
  if (q.epi.atomic && !p->accumulate_x)
    for (int g = 0; g < p->groups; ++g)
      if (p->dX[g]) zero_window(p->dX[g], p->lddx[g], p->M, p->K, st);
  q.epi.drop_on = p->p > 0.0f;
  fill_drop(q.epi.drop, q.epi.gd, p->p, p->seed, p->layer, p->drop_index_base, p->groups, p->seed_dev);
  q.epi.drop_ld = p->K; q.epi.wide_bits = (p->K & 3) != 0;
  for (int g = 0; g < MAXG; ++g) q.epi.bits[g] = cached_bits(p->drop_bits, p->drop_index_base, g < p->groups ? g : 0);
  if constexpr (POOL) {
    q.epi.pool_alpha = p->pool_alpha; q.epi.pool_dp = p->pool_dpooled;
    q.epi.pool_regions = p->pool_regions; q.epi.pool_ld = p->K;
  }
  return launch(q, p->groups, x3, st, "tc_linear_bwd.dgrad");
}

int tc_linear_bwd(const vqa_linear_bwd_params* p, cudaStream_t st, const LinExt* ext) {
  using namespace tc;
  if (is_bf16_math(p->math)) {
    if (p->M >= TC16_MIN_M) return tc16_linear_bwd(p, st, ext);
    vqa_linear_bwd_params q = *p;
    q.math = small_math_of(p->math);
    return tc_linear_bwd(&q, st, nullptr);
  }
  if (p->math != VQA_MATH_TF32X3 && p->math != VQA_MATH_TF32) return tc_fail("vqa_linear_bwd", "this math mode is not built for this op");
  if (p->M > INT32_MAX || p->K > INT32_MAX || p->N > INT32_MAX) return tc_fail("vqa_linear_bwd", "a dimension exceeds 2^31");
  const bool x3 = p->math == VQA_MATH_TF32X3;
  const int64_t ldz = roundup(p->N, 32);
  const int64_t Kp = roundup(p->K, 4);
  const size_t dz_bytes = align256((size_t)p->groups * p->M * ldz * sizeof(float));
  bool any_w = false, any_x = false, pack = false, packx = false;
  for (int g = 0; g < p->groups; ++g) {
    any_w |= p->dW[g] != nullptr || p->db[g] != nullptr;
    if (p->dX[g]) {
      any_x = true;
      pack |= !tma_ok(p->W[g], p->K);
    }
  }
  if (any_w)
    for (int g = 0; g < p->groups; ++g) packx |= !tma_ok(p->X[g], p->ldx[g]);
  bool prepacked = pack;
  for (int g = 0; g < p->groups; ++g) prepacked &= p->Wp[g] != nullptr;
  const size_t w_bytes = (pack && !prepacked) ? align256((size_t)p->groups * p->N * Kp * sizeof(float)) : 0;
  const size_t x_bytes = packx ? align256((size_t)p->groups * p->M * Kp * sizeof(float)) : 0;
  if (!p->workspace || p->workspace_bytes < dz_bytes + w_bytes + x_bytes || reinterpret_cast<uintptr_t>(p->workspace) % 16 != 0)
    return tc_fail("vqa_linear_bwd", "the workspace is missing, misaligned or smaller than vqa_linear_bwd_workspace_bytes()");
  float* dz = reinterpret_cast<float*>(p->workspace);
  float* wpk = reinterpret_cast<float*>(reinterpret_cast<char*>(p->workspace) + dz_bytes);
  float* xpk = reinterpret_cast<float*>(reinterpret_cast<char*>(p->workspace) + dz_bytes + w_bytes);
  if (w_bytes) VQA_TRY(pack_weights(p->W, p->groups, p->N, p->N, p->K, wpk, st));
  if (x_bytes) VQA_TRY(pack_weights(p->X, p->groups, p->M, p->M, p->K, xpk, st, p->ldx));

  // 1. dZ (padded) + bias gradient
  {
    DzArgs a = {};
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < p->groups ? g : 0;
      a.dY[g] = p->dY[s]; a.Y[g] = p->Y[s]; a.lddy[g] = p->lddy[s]; a.ldy[g] = p->ldy[s];
      a.dZ[g] = dz + (size_t)s * p->M * ldz; a.db[g] = p->db[s];
    }
    if (!p->accumulate_w)
      for (int g = 0; g < p->groups; ++g)
        if (p->db[g]) cudaMemsetAsync(p->db[g], 0, (size_t)p->N * sizeof(float), st);
    const int rows_per_cta = p->M <= 1024 ? 16 : 64;          // small M: more CTAs (the kernel is latency-bound there)
    dim3 grid((unsigned)cdiv(ldz, 32), (unsigned)cdiv(p->M, rows_per_cta), (unsigned)p->groups);
    KProf kp_(st, "dz_colsum", "hbm", 4.0 * (double)p->groups * p->M * p->N * (p->act != VQA_ACT_NONE ? 3 : 2));
    dz_colsum_kernel<<<grid, 256, 0, st>>>(a, p->M, p->N, ldz, p->act, rows_per_cta);
    VQA_TRY(check_launch("tc_linear_bwd.dz"));
  }
  // 2. wgrad: D'[K_in, N_out] = X~^T . dZ  (both operands MN-major views of row-major [M, .] tensors)
  if (any_w) {
    Params<EpiWgradT> q = {};
    const int bn = pick_bn(p->N);
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < p->groups ? g : 0;
      if (packx) VQA_TRY(operand_tmap(&q.tmA[g], xpk + (size_t)s * p->M * Kp, true, p->K, p->M, Kp, BM));
      else VQA_TRY(operand_tmap(&q.tmA[g], p->X[s], true, p->K, p->M, p->ldx[s], BM));
      VQA_TRY(operand_tmap(&q.tmB[g], dz + (size_t)s * p->M * ldz, true, p->N, p->M, ldz, bn));
      q.epi.dW[g] = p->dW[s];
    }
    q.epi.ldw = p->K;
    q.M = (int)p->K; q.N = (int)p->N; q.K = (int)p->M; q.a_mn = 1; q.b_mn = 1;
    q.drop_on = p->p > 0.0f;
    fill_drop(q.drop, q.gd, p->p, p->seed, p->layer, p->drop_index_base, p->groups, p->seed_dev);
    q.drop_ld = p->K; q.drop_rows = p->M;
    for (int g = 0; g < MAXG; ++g) q.drop_bits[g] = cached_bits(p->drop_bits, p->drop_index_base, g < p->groups ? g : 0);
    // split the reduction (over the M rows) so that the grid covers the chip
    q.k_splits = pick_splits_wgrad(cdiv(p->K, BM) * cdiv(p->N, bn) * p->groups, p->M);
    if (!p->accumulate_w)
      for (int g = 0; g < p->groups; ++g)
        if (p->dW[g]) cudaMemsetAsync(p->dW[g], 0, (size_t)p->N * p->K * sizeof(float), st);
    VQA_TRY(launch(q, p->groups, x3, st, "tc_linear_bwd.wgrad"));
  }
  // 3. dgrad: dX[M, K_in] = dZ[M, N_out] . W[N_out, K_in]   (W is the MN-major B operand as stored)
  if (any_x) return p->pool_alpha ? dgrad_launch<true>(p, dz, wpk, ldz, Kp, pack, prepacked, x3, st)
                                  : dgrad_launch<false>(p, dz, wpk, ldz, Kp, pack, prepacked, x3, st);
  return VQA_OK;
}

// ============================================================================================ Mutan
// Workspace layout (floats): W1pk [R*Fp, K1p] | W2pk [R*Fp, K2p] | dH1cat [M, R*Fp] | dH2cat [Mh, R*Fp]
struct MutanWs {
  float *w1pk, *w2pk, *dh1, *dh2, *x1pk, *x2pk;
  size_t bytes;
};
// packx1 / packx2: X1 / X2 are not TMA-addressable as given and get a padded copy
// big16: the X1-side GEMMs run on the bf16-plane kernel, which carves its own scratch behind this one
static MutanWs mutan_ws(void* base, int R, int64_t M, int64_t Mh, int64_t K1, int64_t K2, int64_t F, bool bwd,
                        bool packx1 = false, bool packx2 = false, bool big16 = false) {
  using namespace tc;
  const int64_t Fp = roundup(F, 32), K1p = roundup(K1, 4), K2p = roundup(K2, 4);
  char* b = reinterpret_cast<char*>(base);
  size_t off = 0;
  MutanWs w;
  auto take = [&](int64_t n) { float* p = b ? reinterpret_cast<float*>(b + off) : nullptr; off += align256((size_t)n * 4); return p; };
  w.w1pk = big16 ? nullptr : take(R * Fp * K1p);
  w.w2pk = take(R * Fp * K2p);
  w.dh1 = (bwd && !big16) ? take(M * R * Fp) : nullptr;
  w.dh2 = bwd ? take(Mh * R * Fp) : nullptr;
  w.x1pk = (packx1 && !big16) ? take(M * K1p) : nullptr;
  w.x2pk = packx2 ? take(Mh * K2p) : nullptr;
  w.bytes = off;
  return w;
}
size_t tc_mutan_ws(int math, int R, int64_t M, int64_t rows_per, int64_t K1, int64_t K2, int64_t F, int bwd) {
  if (math == VQA_MATH_FP32_SIMT) return 0;
  if (is_bf16_math(math) && M >= TC16_MIN_M)
    return mutan_ws(nullptr, R, M, M / rows_per, K1, K2, F, bwd != 0, false, K2 % 4 != 0, true).bytes +
           tc16_mutan_ws(math == VQA_MATH_BF16X3 ? 2 : 1, R, M, K1, F, bwd);
  return mutan_ws(nullptr, R, M, M / rows_per, K1, K2, F, bwd != 0, K1 % 4 != 0, K2 % 4 != 0).bytes;   // upper bound
}

int tc_mutan_fwd(const vqa_mutan_fwd_params* p, cudaStream_t st, const MutanExt* ext) {
  using namespace tc;
  const bool big16 = is_bf16_math(p->math) && p->M >= TC16_MIN_M;
  const int math = small_math_of(p->math);
  if (math != VQA_MATH_TF32X3 && math != VQA_MATH_TF32) return tc_fail("vqa_mutan_fwd", "this math mode is not built for this op");
  if (p->M > INT32_MAX) return tc_fail("vqa_mutan_fwd", "a dimension exceeds 2^31");
  const int64_t Mh = p->M / p->rows_per_h2;
  const bool packx1 = !big16 && !tma_ok(p->X1, p->ldx1), packx2 = !tma_ok(p->X2, p->ldx2);
  MutanWs w = mutan_ws(p->workspace, p->R, p->M, Mh, p->K1, p->K2, p->F, false, packx1, packx2, big16);
  if (!p->workspace || p->workspace_bytes < w.bytes || reinterpret_cast<uintptr_t>(p->workspace) % 256 != 0)
    return tc_fail("vqa_mutan_fwd", "the workspace is missing, not 256-byte aligned or smaller than vqa_mutan_workspace_bytes()");
  const bool x3 = math == VQA_MATH_TF32X3;
  const int64_t Fp = roundup(p->F, 32), K1p = roundup(p->K1, 4), K2p = roundup(p->K2, 4);
  const float* X1 = p->X1; const float* X2 = p->X2;
  int64_t ldx1 = p->ldx1, ldx2 = p->ldx2;
  if (packx1) { VQA_TRY(pack_weights(&p->X1, 1, p->M, p->M, p->K1, w.x1pk, st, &p->ldx1)); X1 = w.x1pk; ldx1 = K1p; }
  if (packx2) { VQA_TRY(pack_weights(&p->X2, 1, Mh, Mh, p->K2, w.x2pk, st, &p->ldx2)); X2 = w.x2pk; ldx2 = K2p; }
  if (p->W1p && p->W2p) {
    w.w1pk = const_cast<float*>(p->W1p); w.w2pk = const_cast<float*>(p->W2p);
  } else if (p->W2p && big16) {
    w.w2pk = const_cast<float*>(p->W2p);
  } else {
    if (!big16) VQA_TRY(pack_weights(p->W1, p->R, p->F, Fp, p->K1, w.w1pk, st));
    VQA_TRY(pack_weights(p->W2, p->R, p->F, Fp, p->K2, w.w2pk, st));
  }
  const int bn = pick_bn(p->F);
  if (!(ext && ext->h2_mode == 2)) {  // H2_r = X2 . W2_r^T + b2_r  (grouped over r)
    Params<EpiBiasAct> q = {};
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < p->R ? g : 0;
      VQA_TRY(operand_tmap(&q.tmA[g], X2, false, Mh, p->K2, ldx2, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], w.w2pk + (size_t)s * Fp * K2p, false, p->F, p->K2, K2p, bn));
      q.epi.Y[g] = p->H2 + (size_t)s * Mh * p->F; q.epi.bias[g] = p->b2[s]; q.epi.ld[g] = p->F;
    }
    q.M = (int)Mh; q.N = (int)p->F; q.K = (int)p->K2;
    q.k_splits = pick_splits(cdiv(Mh, BM) * cdiv(p->F, bn) * p->R, p->K2);
    q.epi.act = VQA_ACT_NONE; q.epi.atomic = q.k_splits > 1;
    if (q.epi.atomic) cudaMemsetAsync(p->H2, 0, (size_t)p->R * Mh * p->F * sizeof(float), st);
    VQA_TRY(launch(q, p->R, x3, st, "tc_mutan_fwd.h2"));
  }
  if (ext && ext->h2_mode == 1) return VQA_OK;
  if (big16) {      // the X1-side GEMMs on the bf16-plane kernel, its scratch behind this function's
    vqa_mutan_fwd_params q = *p;
    q.workspace = reinterpret_cast<char*>(p->workspace) + w.bytes;
    q.workspace_bytes = p->workspace_bytes - w.bytes;
    return tc16_mutan_fwd_h1(&q, st, ext);
  }
  const int splits = pick_splits(cdiv(p->M, BM) * cdiv(p->F, bn) * p->R, p->K1);
  auto fill = [&](Params<EpiMutan>& q, int r0) -> int {
    for (int g = 0; g < MAXG; ++g) {
      const int r = (r0 + g) < p->R ? (r0 + g) : r0;
      VQA_TRY(operand_tmap(&q.tmA[g], X1, false, p->M, p->K1, ldx1, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], w.w1pk + (size_t)r * Fp * K1p, false, p->F, p->K1, K1p, bn));
      q.epi.bias[g] = p->b1[r]; q.epi.H2[g] = p->H2 + (size_t)r * Mh * p->F;
      q.epi.H1[g] = p->H1 ? p->H1 + (size_t)r * p->M * p->F : nullptr;
    }
    q.M = (int)p->M; q.N = (int)p->F; q.K = (int)p->K1;
    q.epi.Y = p->Y; q.epi.ldh = p->F; q.epi.ldy = p->ldy; q.epi.rows_per = p->rows_per_h2;
    return VQA_OK;
  };
  if (splits > 1) {
    // small M: all ranks and k-splits in one launch, partials accumulated atomically into zeroed Y / H1
    Params<EpiMutan> q = {};
    VQA_TRY(fill(q, 0));
    q.k_splits = splits; q.epi.atomic = 1; q.epi.accumulate = 1;
    zero_window(p->Y, p->ldy, p->M, p->F, st);
    if (p->H1) cudaMemsetAsync(p->H1, 0, (size_t)p->R * p->M * p->F * sizeof(float), st);
    VQA_TRY(launch(q, p->R, x3, st, "tc_mutan_fwd.h1"));
  } else {
    for (int r = 0; r < p->R; ++r) {   // rank by rank: Y is read-modify-written in stream order
      Params<EpiMutan> q = {};
      VQA_TRY(fill(q, r));
      q.k_splits = 1; q.epi.atomic = 0; q.epi.accumulate = r > 0;
      VQA_TRY(launch(q, 1, x3, st, "tc_mutan_fwd.h1"));
    }
  }
  return VQA_OK;
}

int tc_mutan_bwd(const vqa_mutan_bwd_params* p, cudaStream_t st, const MutanExt* ext) {
  using namespace tc;
  const bool big16 = is_bf16_math(p->math) && p->M >= TC16_MIN_M;
  const int math = small_math_of(p->math);
  if (math != VQA_MATH_TF32X3 && math != VQA_MATH_TF32) return tc_fail("vqa_mutan_bwd", "this math mode is not built for this op");
  if (p->M > INT32_MAX) return tc_fail("vqa_mutan_bwd", "a dimension exceeds 2^31");
  const int64_t Mh = p->M / p->rows_per_h2;
  const bool packx1 = !big16 && !tma_ok(p->X1, p->ldx1), packx2 = !tma_ok(p->X2, p->ldx2);
  MutanWs w = mutan_ws(p->workspace, p->R, p->M, Mh, p->K1, p->K2, p->F, true, packx1, packx2, big16);
  if (!p->workspace || p->workspace_bytes < w.bytes || reinterpret_cast<uintptr_t>(p->workspace) % 256 != 0)
    return tc_fail("vqa_mutan_bwd", "the workspace is missing, not 256-byte aligned or smaller than vqa_mutan_workspace_bytes()");
  const bool x3 = math == VQA_MATH_TF32X3;
  Mutan16Ops ops16 = {};
  ops16.np = 1;
  if (big16) {      // operand planes of the X1-side GEMMs (scratch behind this function's)
    vqa_mutan_bwd_params q = *p;
    q.workspace = reinterpret_cast<char*>(p->workspace) + w.bytes;
    q.workspace_bytes = p->workspace_bytes - w.bytes;
    VQA_TRY(tc16_mutan_bwd_prepare(&q, st, ext, &ops16));
  }
  const int R = p->R;
  const int64_t Fp = roundup(p->F, 32), K1p = roundup(p->K1, 4), K2p = roundup(p->K2, 4), RF = R * Fp;
  const float* X1 = p->X1; const float* X2 = p->X2;
  int64_t ldx1 = p->ldx1, ldx2 = p->ldx2;
  if (packx1) { VQA_TRY(pack_weights(&p->X1, 1, p->M, p->M, p->K1, w.x1pk, st, &p->ldx1)); X1 = w.x1pk; ldx1 = K1p; }
  if (packx2) { VQA_TRY(pack_weights(&p->X2, 1, Mh, Mh, p->K2, w.x2pk, st, &p->ldx2)); X2 = w.x2pk; ldx2 = K2p; }
  if (p->W1p && p->W2p) {
    w.w1pk = const_cast<float*>(p->W1p); w.w2pk = const_cast<float*>(p->W2p);
  } else if (p->W2p && big16) {
    w.w2pk = const_cast<float*>(p->W2p);
  } else {
    if (!big16) VQA_TRY(pack_weights(p->W1, R, p->F, Fp, p->K1, w.w1pk, st));
    VQA_TRY(pack_weights(p->W2, R, p->F, Fp, p->K2, w.w2pk, st));
  }
  if (!p->accumulate_w)
    for (int r = 0; r < R; ++r) {
      if (p->db1[r]) cudaMemsetAsync(p->db1[r], 0, (size_t)p->F * 4, st);
      if (p->db2[r]) cudaMemsetAsync(p->db2[r], 0, (size_t)p->F * 4, st);
      if (p->dW1[r]) cudaMemsetAsync(p->dW1[r], 0, (size_t)p->F * p->K1 * 4, st);
      if (p->dW2[r]) cudaMemsetAsync(p->dW2[r], 0, (size_t)p->F * p->K2 * 4, st);
    }
  DbTable db1 = {}, db2 = {};
  for (int r = 0; r < R; ++r) { db1.p[r] = p->db1[r]; db2.p[r] = p->db2[r]; }
  {
    const bool vec2 = p->F % 2 == 0 && p->lddy % 2 == 0 && reinterpret_cast<uintptr_t>(p->dY) % 8 == 0 &&
                      reinterpret_cast<uintptr_t>(p->H1) % 8 == 0;
    const dim3 grid((unsigned)Mh, (unsigned)R);
    KProf kp_(st, "mutan_dh", "hbm", 4.0 * (double)p->M * p->F * (1.0 + 2.0 * R));
    if (vec2)
      mutan_dh_kernel<2><<<grid, 256, 0, st>>>(p->M, p->F, Fp, p->rows_per_h2, R, p->dY, p->lddy, p->H1, p->H2, w.dh1,
                                              w.dh2, db1, db2, big16 ? ops16.dh1 : nullptr, p->M * RF, ops16.np);
    else
      mutan_dh_kernel<1><<<grid, 256, 0, st>>>(p->M, p->F, Fp, p->rows_per_h2, R, p->dY, p->lddy, p->H1, p->H2, w.dh1,
                                              w.dh2, db1, db2, big16 ? ops16.dh1 : nullptr, p->M * RF, ops16.np);
    VQA_TRY(check_launch("tc_mutan_bwd.dh"));
  }
  auto wgrad = [&](const float* X, int64_t ldx, int64_t Krows, int64_t Kin, const float* dHcat, float* const* dW,
                   const char* what) -> int {
    // D'[Kin, F] = X^T[Kin, Krows] . dH_r[Krows, F], grouped over r
    Params<EpiWgradT> q = {};
    const int bn = pick_bn(p->F);
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < R ? g : 0;
      VQA_TRY(operand_tmap(&q.tmA[g], X, true, Kin, Krows, ldx, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], dHcat + (size_t)s * Fp, true, p->F, Krows, RF, bn));
      q.epi.dW[g] = dW[s];
    }
    q.epi.ldw = Kin;
    q.M = (int)Kin; q.N = (int)p->F; q.K = (int)Krows; q.a_mn = 1; q.b_mn = 1;
    q.k_splits = pick_splits_wgrad(cdiv(Kin, BM) * cdiv(p->F, bn) * R, Krows);
    return launch(q, R, x3, st, what);
  };
  auto dgrad = [&](const float* dHcat, int64_t rows, const float* Wpk, int64_t Kin, int64_t Kinp, float* dX,
                   int64_t lddx, int accumulate, const char* what) -> int {
    // dX[rows, Kin] = dHcat[rows, R*Fp] . Wpk[R*Fp, Kin]
    Params<EpiDgrad> q = {};
    const int bn = pick_bn(Kin);
    for (int g = 0; g < MAXG; ++g) {
      VQA_TRY(operand_tmap(&q.tmA[g], dHcat, false, rows, RF, RF, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], Wpk, true, Kin, RF, Kinp, bn));
      q.epi.dX[g] = dX; q.epi.ld[g] = lddx;
    }
    q.M = (int)rows; q.N = (int)Kin; q.K = (int)RF; q.a_mn = 0; q.b_mn = 1;
    q.k_splits = pick_splits(cdiv(rows, BM) * cdiv(Kin, bn), RF);
    q.epi.accumulate = accumulate; q.epi.drop_on = 0; q.epi.atomic = q.k_splits > 1;
    if (q.epi.atomic && !accumulate) zero_window(dX, lddx, rows, Kin, st);
    return launch(q, 1, x3, st, what);
  };
  if (big16) {
    VQA_TRY(tc16_mutan_bwd_big(p, st, &ops16));
  } else {
    VQA_TRY(wgrad(X1, ldx1, p->M, p->K1, w.dh1, p->dW1, "tc_mutan_bwd.dw1"));
    if (p->dX1) VQA_TRY(dgrad(w.dh1, p->M, w.w1pk, p->K1, K1p, p->dX1, p->lddx1, p->accumulate_x1, "tc_mutan_bwd.dx1"));
  }
  VQA_TRY(wgrad(X2, ldx2, Mh, p->K2, w.dh2, p->dW2, "tc_mutan_bwd.dw2"));
  if (p->dX2) VQA_TRY(dgrad(w.dh2, Mh, w.w2pk, p->K2, K2p, p->dX2, p->lddx2, p->accumulate_x2, "tc_mutan_bwd.dx2"));
  return VQA_OK;
}

int tc_pack_segments(const vqa_pack_segment* segs, int nsegs, cudaStream_t st) {
  tc::PackSegs a = {};
  int64_t biggest = 1;
  for (int i = 0; i < nsegs; ++i) {
    a.s[i] = segs[i];
    const int64_t n = segs[i].rows_pad * ((segs[i].K + 3) / 4 * 4);
    if (n > biggest) biggest = n;
  }
  int64_t blocks = cdiv(biggest, 256 * 4);
  if (blocks > 1024) blocks = 1024;
  tc::pack_segments_kernel<<<dim3((unsigned)blocks, (unsigned)nsegs), 256, 0, st>>>(a);
  return check_launch("pack_segments");
}

// *seed += 1 (graph-captured at the head of every replayed step)
__global__ void seed_advance_kernel(uint64_t* seed) { *seed += 1; }
int tc_seed_advance(uint64_t* seed_dev, cudaStream_t st) {
  seed_advance_kernel<<<1, 1, 0, st>>>(seed_dev);
  return check_launch("seed_advance");
}

int tc_dropout_bits(float pdrop, uint64_t seed, const uint64_t* seed_dev, uint32_t layer, uint64_t n, uint8_t* out,
                    cudaStream_t st) {
  const uint64_t ngroups = (n + 15) / 16;
  if (ngroups == 0) return VQA_OK;
  uint64_t blocks = (ngroups + 255) / 256;
  if (blocks > 65535) blocks = 65535;
  tc::dropout_bits_kernel<<<(unsigned)blocks, 256, 0, st>>>(seed, seed_dev, layer, drop_threshold(pdrop), ngroups,
                                                            reinterpret_cast<uint16_t*>(out));
  return check_launch("dropout_bits");
}

int tc_dropout_bits_batch(float pdrop, uint64_t seed, const uint64_t* seed_dev, const vqa_bits_segment* segs, int nsegs,
                          cudaStream_t st) {
  tc::BitSegs a = {};
  a.n = nsegs;
  for (int i = 0; i < nsegs; ++i) {
    a.s[i] = segs[i];
    a.first[i + 1] = a.first[i] + (segs[i].n + 15) / 16;
  }
  if (a.first[nsegs] == 0) return VQA_OK;
  uint64_t blocks = (a.first[nsegs] + 255) / 256;
  if (blocks > 16384) blocks = 16384;
  KProf kp_(st, "dropout_bits_batch", "hbm", (double)a.first[nsegs] * 2.0);
  tc::dropout_bits_batch_kernel<<<(unsigned)blocks, 256, 0, st>>>(seed, seed_dev, drop_threshold(pdrop), a);
  return check_launch("dropout_bits_batch");
}

// Upper bounds: a K that is not a multiple of 4 floats makes W (and a contiguous [M, K] X) un-addressable by TMA.
size_t tc_linear_fwd_ws(int math, int groups, int64_t M, int64_t K, int64_t N) {
  if (is_bf16_math(math)) {
    if (M >= TC16_MIN_M) return tc16_linear_fwd_ws(math == VQA_MATH_BF16X3 ? 2 : 1, groups, M, K, N);
    math = small_math_of(math);
  }
  if (math == VQA_MATH_FP32_SIMT || K % 4 == 0) return 0;
  return tc::align256((size_t)groups * N * tc::roundup(K, 4) * sizeof(float)) +
         tc::align256((size_t)groups * M * tc::roundup(K, 4) * sizeof(float));
}
size_t tc_linear_bwd_ws(int math, int groups, int64_t M, int64_t K, int64_t N) {
  if (is_bf16_math(math)) {
    if (M >= TC16_MIN_M) return tc16_linear_bwd_ws(math == VQA_MATH_BF16X3 ? 2 : 1, groups, M, K, N);
    math = small_math_of(math);
  }
  if (math == VQA_MATH_FP32_SIMT) return 0;
  size_t b = tc::align256((size_t)groups * M * tc::roundup(N, 32) * sizeof(float));
  if (K % 4 != 0)
    b += tc::align256((size_t)groups * N * tc::roundup(K, 4) * sizeof(float)) +
         tc::align256((size_t)groups * M * tc::roundup(K, 4) * sizeof(float));
  return b;
}

}  // namespace vqa
