// Host side of the bf16-plane tcgen05 GEMM (tc_gemm16.cuh): 3-D TMA tensor maps over operand planes, the kernels
// that PRODUCE planes (input split + dropout mask, weight pack in both orientations, dZ = dY (.) act'(Y) in both
// orientations), launch plumbing, and the large-M linear / Mutan forward and backward built from them
// (VQA_MATH_BF16X3: fp32-parity arithmetic, x = hi + lo planes, three bf16 MMAs per product;
//  VQA_MATH_BF16: one plane, plain bf16 operands).  Small-M problems (M < TC16_MIN_M) stay on tc_gemm.cuh.
#include "gemm_tc.h"

#include <stdlib.h>

#include <mutex>

#include "tc_gemm16.cuh"

namespace vqa {
namespace tc16 {

using tc::EpiBiasAct;
using tc::EpiDgradT;
using tc::EpiMutan;
using tc::EpiWgradT;

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

// bf16 planes [np][rows][ld] (row-major, `plane` elements between planes) as a 3-D tensor (cols, rows, plane);
// box = (box_cols, box_rows, np), 128B swizzle.  Out-of-bounds parts of a box are zero-filled: ragged M / N / K edges
// need no padded copies.
static int make_tmap3(CUtensorMap* out, const Planes& t, int np, int64_t rows, int64_t cols, int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VQA_ECUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)np};
  cuuint64_t strides[2] = {(cuuint64_t)t.ld * 2, (cuuint64_t)(np > 1 ? t.plane : t.ld * rows) * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)np};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(t.p), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 planes) failed (%d) for ptr=%p rows=%lld cols=%lld ld=%lld plane=%lld box=%dx%d",
              (int)r, (const void*)t.p, (long long)rows, (long long)cols, (long long)t.ld, (long long)t.plane, box_cols,
              box_rows);
    return VQA_ECUDA;
  }
  return VQA_OK;
}
// K-major operand: source planes are [rows(M or N), K]; MN-major operand: source planes are [K, rows(M or N)].
static int operand_tmap(CUtensorMap* out, const Planes& t, int np, bool mn_major, int64_t mn_extent, int64_t k_extent,
                        int tile_mn) {
  if (!mn_major) return make_tmap3(out, t, np, mn_extent, k_extent, BK, tile_mn);
  return make_tmap3(out, t, np, k_extent, mn_extent, 64, BK);
}

static inline int64_t roundup(int64_t x, int64_t m) { return cdiv(x, m) * m; }
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------ launch
static inline int pick_bn(int64_t N) {
  const int64_t w160 = cdiv(N, 160) * 160 - N, w128 = cdiv(N, 128) * 128 - N;
  return w160 < w128 ? 160 : 128;
}

// k-splits of a persistent launch: the work-item count that best fills whole rounds of the SMs (fewest splits on a
// tie: every split adds one reduction pass over the output), with at least 4 k-blocks per split.
static int pick_splits(int64_t tiles, int64_t K) {
  const int64_t kb = cdiv(K, BK), sms = sm_count();
  int best = 1;
  double best_eff = 0.0;
  for (int64_t s = 1; s <= 32 && s * 4 <= kb; ++s) {
    const int64_t per = cdiv(kb, s);
    if ((s - 1) * per >= kb) continue;                 // an empty last split
    const int64_t items = tiles * s;
    const double eff = (double)items / (double)(cdiv(items, sms) * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best = (int)s; }
  }
  return best;
}

template <int BN, int NP, class Epi>
static int launch_cfg(const Params<Epi>& p, cudaStream_t st, const char* what) {
  using C = Cfg<BN, NP>;
  auto kern = gemm16_kernel<BN, NP, Epi>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
    return check_launch(what);
  const int64_t work = (int64_t)p.groups * p.k_splits * cdiv(p.M, BM) * cdiv(p.N, BN);
  const unsigned grid = (unsigned)(work < sm_count() ? work : sm_count());
  char label[160] = "k:";
  if (prof_active())
    snprintf(label, sizeof(label), "k:%s@%s M%d N%d K%d g%d s%d", what, NP == 2 ? "bf16x3" : "bf16", p.M, p.N, p.K, p.groups,
             p.k_splits);
  ProfScope ps_(st, label);
  kern<<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(p);
  return check_launch(what);
}

// bn: 128 or 160 (K-major B) — an MN-major B needs 128
template <class Epi>
static int launch(const Params<Epi>& p, int np, int bn, cudaStream_t st, const char* what) {
  if (p.b_mn && bn % 64 != 0) {
    set_error("%s: an MN-major B operand needs a tile width that is a multiple of 64 (got %d)", what, bn);
    return VQA_EINVAL;
  }
  if (bn == 160) return np == 2 ? launch_cfg<160, 2>(p, st, what) : launch_cfg<160, 1>(p, st, what);
  return np == 2 ? launch_cfg<128, 2>(p, st, what) : launch_cfg<128, 1>(p, st, what);
}

// ------------------------------------------------------------------------------------------ plane producers
// 4 consecutive values -> bf16 plane(s); 8-byte stores
__device__ __forceinline__ void store_planes4(__nv_bfloat16* dst, int64_t plane, int np, const float (&o)[4]) {
  uint2 hi, lo;
  tc::split4_bf16(o, hi, lo);
  *reinterpret_cast<uint2*>(dst) = hi;
  if (np == 2) *reinterpret_cast<uint2*>(dst + plane) = lo;
}

// Planes of dropout(X): out[.][m][k] = X[m,k] * keep(base + m*K + k) / (1-p) for k < K, 0 for K <= k < ldp.
// A thread makes 4 consecutive k.  The mask comes from the packed keep-bits when given, else from Philox in registers
// (include/vqacore.h contract) — the planes ARE the stash of the mask for the forward and the weight-gradient GEMM.
struct SplitArgs {
  const float* X[MAXG]; int64_t ldx[MAXG];
  __nv_bfloat16* out[MAXG];
  const uint8_t* bits[MAXG];
  uint32_t layer[MAXG]; uint64_t base[MAXG];
};
// grid = (column chunks of 1024, row chunks, groups); a thread makes one quad of `rows_per_cta` consecutive rows
__global__ void __launch_bounds__(256)
split_planes_kernel(SplitArgs a, int64_t M, int64_t K, int64_t ldp, int64_t plane, int np, Drop d, int rows_per_cta) {
  const int g = blockIdx.z;
  const float* __restrict__ x = a.X[g];
  const int64_t ldx = a.ldx[g];
  __nv_bfloat16* out = a.out[g];
  const uint8_t* __restrict__ bits = a.bits[g];
  const uint64_t seed = d.on ? d.key() : 0;
  const bool vec = (ldx % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
  const int64_t k = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (k >= ldp) return;
  const int64_t m_begin = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t m_end = m_begin + rows_per_cta < M ? m_begin + rows_per_cta : M;
#pragma unroll 4
  for (int64_t m = m_begin; m < m_end; ++m) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (k + 4 <= K && vec) {
      const float4 v = ld_stream4(x + m * ldx + k);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (k + e < K) o[e] = x[m * ldx + k + e];
    }
    if (d.on && k < K) {
      const uint64_t idx = a.base[g] + (uint64_t)(m * K + k);
      uint32_t keep;                         // bit e = keep element e
      if (bits) {
        const uint32_t sh = (uint32_t)(idx & 7);
        uint32_t w = __ldg(bits + (idx >> 3));
        if (sh > 4) w |= (uint32_t)__ldg(bits + (idx >> 3) + 1) << 8;
        keep = w >> sh;
      } else {
        const uint32_t bt = philox_bytes4(seed, a.layer[g], idx);
        keep = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) keep |= (((bt >> (8 * e)) & 0xFFu) >= d.thr ? 1u : 0u) << e;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = ((keep >> e) & 1u) ? o[e] * d.scale : 0.0f;
    }
    store_planes4(out + m * ldp + k, plane, np, o);
  }
}

int split_planes(const float* const* X, const int64_t* ldx, int groups, int64_t M, int64_t K, float pdrop, uint64_t seed,
                 const uint64_t* seed_dev, const uint32_t* layer, const uint64_t* base, const uint8_t* const* bits,
                 __nv_bfloat16* const* out, int64_t ldp, int64_t plane, int np, cudaStream_t st) {
  SplitArgs a = {};
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < groups ? g : 0;
    a.X[g] = X[s]; a.ldx[g] = ldx[s]; a.out[g] = out[s];
    a.bits[g] = (bits && base && base[s] == 0) ? bits[s] : nullptr;
    a.layer[g] = layer ? layer[s] : 0; a.base[g] = base ? base[s] : 0;
  }
  Drop d = make_drop(pdrop, seed, 0, 0, 1, seed_dev);
  const int rows_per_cta = M >= 4096 ? 8 : 1;
  KProf kp_(st, "split_planes", "hbm", (double)groups * M * K * (4.0 + 2.0 * np));
  split_planes_kernel<<<dim3((unsigned)cdiv(ldp, 1024), (unsigned)cdiv(M, rows_per_cta), (unsigned)groups), 256, 0, st>>>(
      a, M, K, ldp, plane, np, d, rows_per_cta);
  return check_launch("split_planes");
}

// Weight planes in both orientations, several weights per launch (vqa_pack_weights' bf16 twin):
//   dst  [np][rows_pad][Kp]  = W[r,k]          (K-major B of the forward GEMM; rows >= rows and k >= K zero)
//   dstT [np][Kt][Np]        = W[n,k] at (k,n) (K-major B of the dgrad GEMM: rows = input features)
struct PackPlaneSegs { PackPlanesSeg s[VQA_MAX_PACK_SEGMENTS]; };
__global__ void __launch_bounds__(256) pack_planes_kernel(PackPlaneSegs a, int np) {
  const PackPlanesSeg sg = a.s[blockIdx.y];
  if (sg.dst) {
    const int64_t total = sg.rows_pad * sg.Kp;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = t / sg.Kp, k = t - r * sg.Kp;
      const float v = (k < sg.K && r < sg.rows) ? sg.src[r * sg.K + k] : 0.0f;
      __nv_bfloat16 hi, lo;
      tc::split_bf16(v, hi, lo);
      sg.dst[t] = hi;
      if (np == 2) sg.dst[t + sg.plane] = lo;
    }
  }
  if (sg.dstT) {          // this segment owns columns [t_col0, t_col0 + rows_pad) of the (possibly stacked) transposed planes
    const int64_t total = sg.K * sg.rows_pad;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      const int64_t k = t / sg.rows_pad, c = t - k * sg.rows_pad;
      const float v = c < sg.rows ? sg.src[c * sg.K + k] : 0.0f;
      __nv_bfloat16 hi, lo;
      tc::split_bf16(v, hi, lo);
      const int64_t o = k * sg.Np + sg.t_col0 + c;
      sg.dstT[o] = hi;
      if (np == 2) sg.dstT[o + sg.planeT] = lo;
    }
  }
}
int pack_planes(const PackPlanesSeg* segs, int nsegs, int np, cudaStream_t st) {
  if (nsegs <= 0) return VQA_OK;
  if (nsegs > VQA_MAX_PACK_SEGMENTS) {
    set_error("pack_planes: too many segments (%d)", nsegs);
    return VQA_EINVAL;
  }
  PackPlaneSegs a = {};
  int64_t biggest = 1;
  for (int i = 0; i < nsegs; ++i) {
    a.s[i] = segs[i];
    const int64_t n = segs[i].rows_pad * segs[i].Kp;
    if (n > biggest) biggest = n;
  }
  int64_t blocks = cdiv(biggest, 256 * 4);
  if (blocks > 1024) blocks = 1024;
  pack_planes_kernel<<<dim3((unsigned)blocks, (unsigned)nsegs), 256, 0, st>>>(a, np);
  return check_launch("pack_planes");
}

// dZ = dY (.) act'(Y) as planes in the two orientations the backward GEMMs read, and db (+)= colsum(dZ):
//   dZp [np][M][ldz]   rows = samples  (K-major A of dgrad)          — optional
//   dZt [np][Nt][Mp]   rows = output features (K-major B of wgrad)   — optional
// grid = (cdiv(ldz, 32), cdiv(M, 64), groups); 256 threads = 8 x 32; pad columns N..ldz are written as zeros.
struct Dz16Args {
  const float* dY[MAXG]; const float* Y[MAXG]; float* db[MAXG];
  int64_t lddy[MAXG], ldy[MAXG];
  __nv_bfloat16* dZp[MAXG]; __nv_bfloat16* dZt[MAXG];
};
__global__ void __launch_bounds__(256)
dz_planes_kernel(Dz16Args a, int64_t M, int64_t N, int64_t ldz, int64_t plane_z, int64_t Mp, int64_t plane_t, int np,
                 int act) {
  __shared__ float tile[64][33];
  __shared__ float red[8][33];
  const int g = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * 32, n = n0 + tx;
  const int64_t m0 = (int64_t)blockIdx.y * 64;
  const float* __restrict__ dy = a.dY[g];
  const float* __restrict__ y = a.Y[g];
  __nv_bfloat16* dzp = a.dZp[g];
  __nv_bfloat16* dzt = a.dZt[g];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty + 8 * i;
    const int64_t m = m0 + r;
    float v = 0.0f;
    if (m < M && n < N) {
      v = dy[m * a.lddy[g] + n];
      if (act != VQA_ACT_NONE) v *= act_grad(act, y[m * a.ldy[g] + n]);
    }
    tile[r][tx] = v;
    s += v;
    if (dzp && m < M && n < ldz) {
      __nv_bfloat16 hi, lo;
      tc::split_bf16(v, hi, lo);
      dzp[m * ldz + n] = hi;
      if (np == 2) dzp[plane_z + m * ldz + n] = lo;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N && a.db[g]) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(a.db[g] + n, t);
  }
  if (dzt) {
    // transposed store: thread (tx, ty) writes column n0 + ty + 8j at rows m0 + tx and m0 + 32 + tx
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ty + 8 * j;
      const int64_t nn = n0 + c;
      if (nn >= ldz) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t m = m0 + 32 * h + tx;
        if (m >= Mp) continue;
        __nv_bfloat16 hi, lo;
        tc::split_bf16(m < M ? tile[32 * h + tx][c] : 0.0f, hi, lo);
        dzt[nn * Mp + m] = hi;
        if (np == 2) dzt[plane_t + nn * Mp + m] = lo;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ workspace carving
struct Carver {
  char* base; size_t off;
  __nv_bfloat16* take(int64_t n_bf16) {
    __nv_bfloat16* p = base ? reinterpret_cast<__nv_bfloat16*>(base + off) : nullptr;
    off += align256((size_t)n_bf16 * 2);
    return p;
  }
};

}  // namespace tc16

using tc16::Carver;
using tc16::roundup;

static int fail16(const char* who, const char* why) {
  set_error("%s: %s (bf16 tensor-core mode; there is no silent fallback)", who, why);
  return VQA_EINVAL;
}

// ============================================================================================ linear fwd
// Workspace (bf16 elements): X planes [g][np][M][Kp8] (unless handed over) | W planes [g][np][Np][Kp8] (unless handed over)
size_t tc16_linear_fwd_ws(int np, int groups, int64_t M, int64_t K, int64_t N) {
  Carver c{nullptr, 0};
  const int64_t Kp = roundup(K, 8), Np = roundup(N, 32);
  c.take((int64_t)groups * np * M * Kp);
  c.take((int64_t)groups * np * Np * Kp);
  return c.off + 256;
}

int tc16_linear_fwd(const vqa_linear_fwd_params* p, cudaStream_t st, const LinExt* ext) {
  using namespace tc16;
  const int np = p->math == VQA_MATH_BF16X3 ? 2 : 1;
  if (p->M > INT32_MAX || p->K > INT32_MAX || p->N > INT32_MAX) return fail16("vqa_linear_fwd", "a dimension exceeds 2^31");
  if (ext && p->groups != 1) return fail16("vqa_linear_fwd", "plane hand-offs are for single-group launches");
  const int64_t Kp = roundup(p->K, 8), Np = roundup(p->N, 32);
  Carver c{reinterpret_cast<char*>(p->workspace), 0};
  if (reinterpret_cast<uintptr_t>(p->workspace) % 256 != 0) c.off = 256 - reinterpret_cast<uintptr_t>(p->workspace) % 256;
  Planes xp[MAXG], wp[MAXG];
  const bool have_x = ext && ext->Xp.p, have_w = ext && ext->Wp.p;
  __nv_bfloat16* xbase = have_x ? nullptr : c.take((int64_t)p->groups * np * p->M * Kp);
  __nv_bfloat16* wbase = have_w ? nullptr : c.take((int64_t)p->groups * np * Np * Kp);
  if ((!have_x || !have_w) && (!p->workspace || c.off > p->workspace_bytes))
    return fail16("vqa_linear_fwd", "the workspace is missing or smaller than vqa_linear_fwd_workspace_bytes()");
  if (have_x) xp[0] = ext->Xp;
  else {
    __nv_bfloat16* outs[MAXG];
    for (int g = 0; g < p->groups; ++g) {
      outs[g] = xbase + (int64_t)g * np * p->M * Kp;
      xp[g] = Planes{outs[g], Kp, p->M * Kp};
    }
    VQA_TRY(split_planes(p->X, p->ldx, p->groups, p->M, p->K, p->p, p->seed, p->seed_dev, p->layer, p->drop_index_base,
                         p->drop_bits, outs, Kp, p->M * Kp, np, st));
  }
  if (have_w) wp[0] = ext->Wp;
  else {
    PackPlanesSeg segs[MAXG] = {};
    for (int g = 0; g < p->groups; ++g) {
      __nv_bfloat16* d = wbase + (int64_t)g * np * Np * Kp;
      segs[g].src = p->W[g]; segs[g].rows = p->N; segs[g].rows_pad = Np; segs[g].K = p->K; segs[g].Kp = Kp;
      segs[g].dst = d; segs[g].plane = Np * Kp;
      wp[g] = Planes{d, Kp, Np * Kp};
    }
    VQA_TRY(pack_planes(segs, p->groups, np, st));
  }
  const int bn = pick_bn(p->N);
  Params<EpiBiasAct> q = {};
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < p->groups ? g : 0;
    VQA_TRY(operand_tmap(&q.tmA[g], xp[s], np, false, p->M, p->K, BM));
    VQA_TRY(operand_tmap(&q.tmB[g], wp[s], np, false, p->N, p->K, bn));
    q.epi.Y[g] = p->Y[s]; q.epi.bias[g] = p->b[s]; q.epi.ld[g] = p->ldy[s];
  }
  q.M = (int)p->M; q.N = (int)p->N; q.K = (int)p->K; q.groups = p->groups; q.k_splits = 1; q.a_mn = 0; q.b_mn = 0;
  q.epi.act = p->act; q.epi.atomic = 0;
  if (ext && ext->Yp) { q.epi.Yp[0] = ext->Yp; q.epi.ldp = ext->ldyp; q.epi.plane_stride = ext->yplane; q.epi.planes = np; }
  return launch(q, np, bn, st, "tc16_linear_fwd");
}

// ============================================================================================ linear bwd
// Workspace: dZp [g][np][M][ldz] | dZt [g][np][ldz][Mp] | X planes [g][np][M][Kp8] | W^T planes [g][np][K][Np]
size_t tc16_linear_bwd_ws(int np, int groups, int64_t M, int64_t K, int64_t N) {
  Carver c{nullptr, 0};
  const int64_t Kp = roundup(K, 8), ldz = roundup(N, 32), Mp = roundup(M, 8);
  c.take((int64_t)groups * np * M * ldz);
  c.take((int64_t)groups * np * ldz * Mp);
  c.take((int64_t)groups * np * M * Kp);
  c.take((int64_t)groups * np * K * ldz);
  return c.off + 256;
}

template <bool POOL>
static int dgrad16_launch(const vqa_linear_bwd_params* p, const Planes* dzp, const Planes* wtp, int np, cudaStream_t st,
                          bool raw = false) {
  using namespace tc16;
  Params<EpiDgradT<POOL>> q = {};
  const int bn = pick_bn(p->K);
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < p->groups ? g : 0;
    VQA_TRY(operand_tmap(&q.tmA[g], dzp[s], np, false, p->M, p->N, BM));
    VQA_TRY(operand_tmap(&q.tmB[g], wtp[s], np, false, p->K, p->N, bn));
    q.epi.dX[g] = p->dX[s]; q.epi.ld[g] = p->lddx[s];
  }
  q.M = (int)p->M; q.N = (int)p->K; q.K = (int)p->N; q.groups = p->groups; q.k_splits = 1; q.a_mn = 0; q.b_mn = 0;
  q.epi.accumulate = p->accumulate_x; q.epi.atomic = 0;
  q.epi.drop_on = p->p > 0.0f && !raw;
  q.epi.drop = make_drop(raw ? 0.0f : p->p, p->seed, 0, 0, 1, p->seed_dev);
  for (int g = 0; g < MAXG; ++g) {
    const int s = g < p->groups ? g : 0;
    q.epi.gd.layer[g] = p->layer[s]; q.epi.gd.base[g] = p->drop_index_base[s];
    q.epi.bits[g] = p->drop_index_base[s] == 0 ? p->drop_bits[s] : nullptr;
  }
  q.epi.drop_ld = p->K; q.epi.wide_bits = (p->K & 3) != 0;
  if constexpr (POOL) {
    q.epi.pool_alpha = p->pool_alpha; q.epi.pool_dp = p->pool_dpooled;
    q.epi.pool_regions = p->pool_regions; q.epi.pool_ld = p->K;
  }
  return launch(q, np, bn, st, "tc16_linear_bwd.dgrad");
}

int tc16_linear_bwd(const vqa_linear_bwd_params* p, cudaStream_t st, const LinExt* ext) {
  using namespace tc16;
  const int np = p->math == VQA_MATH_BF16X3 ? 2 : 1;
  if (p->M > INT32_MAX || p->K > INT32_MAX || p->N > INT32_MAX) return fail16("vqa_linear_bwd", "a dimension exceeds 2^31");
  if (ext && p->groups != 1) return fail16("vqa_linear_bwd", "plane hand-offs are for single-group launches");
  const int64_t Kp = roundup(p->K, 8), ldz = roundup(p->N, 32), Mp = roundup(p->M, 8);
  bool any_w = false, any_x = false;
  for (int g = 0; g < p->groups; ++g) {
    any_w |= p->dW[g] != nullptr;
    any_x |= p->dX[g] != nullptr;
  }
  bool any_b = false;
  for (int g = 0; g < p->groups; ++g) any_b |= p->db[g] != nullptr;
  const bool have_x = ext && ext->Xp.p, have_wt = ext && ext->WTp.p;
  Carver c{reinterpret_cast<char*>(p->workspace), 0};
  if (reinterpret_cast<uintptr_t>(p->workspace) % 256 != 0) c.off = 256 - reinterpret_cast<uintptr_t>(p->workspace) % 256;
  __nv_bfloat16* dzp_base = any_x ? c.take((int64_t)p->groups * np * p->M * ldz) : nullptr;
  __nv_bfloat16* dzt_base = any_w ? c.take((int64_t)p->groups * np * ldz * Mp) : nullptr;
  __nv_bfloat16* x_base = (any_w && !have_x) ? c.take((int64_t)p->groups * np * p->M * Kp) : nullptr;
  __nv_bfloat16* wt_base = (any_x && !have_wt) ? c.take((int64_t)p->groups * np * p->K * ldz) : nullptr;
  if (!p->workspace || c.off > p->workspace_bytes)
    return fail16("vqa_linear_bwd", "the workspace is missing or smaller than vqa_linear_bwd_workspace_bytes()");

  Planes dzp[MAXG], dzt[MAXG], xp[MAXG], wtp[MAXG];
  // 1. dZ planes (both orientations) + bias gradient
  {
    Dz16Args a = {};
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < p->groups ? g : 0;
      a.dY[g] = p->dY[s]; a.Y[g] = p->Y[s]; a.lddy[g] = p->lddy[s]; a.ldy[g] = p->ldy[s]; a.db[g] = p->db[s];
      a.dZp[g] = dzp_base ? dzp_base + (int64_t)s * np * p->M * ldz : nullptr;
      a.dZt[g] = dzt_base ? dzt_base + (int64_t)s * np * ldz * Mp : nullptr;
      dzp[g] = Planes{a.dZp[g], ldz, p->M * ldz};
      dzt[g] = Planes{a.dZt[g], Mp, ldz * Mp};
    }
    if (!p->accumulate_w)
      for (int g = 0; g < p->groups; ++g)
        if (p->db[g]) cudaMemsetAsync(p->db[g], 0, (size_t)p->N * sizeof(float), st);
    if (any_w || any_x || any_b) {
      dim3 grid((unsigned)cdiv(ldz, 32), (unsigned)cdiv(p->M, 64), (unsigned)p->groups);
      KProf kp_(st, "dz_planes", "hbm", (double)p->groups * p->M * p->N * ((p->act != VQA_ACT_NONE ? 8.0 : 4.0) +
                                                                          2.0 * np * ((any_w ? 1 : 0) + (any_x ? 1 : 0))));
      dz_planes_kernel<<<grid, 256, 0, st>>>(a, p->M, p->N, ldz, p->M * ldz, Mp, ldz * Mp, np, p->act);
      VQA_TRY(check_launch("tc16_linear_bwd.dz"));
    }
  }
  // 2. wgrad: D'[K_in, N_out] = X~^T . dZ   (A: MN-major view of the X~ planes; B: the transposed dZ planes, K-major)
  if (any_w) {
    if (have_x) xp[0] = ext->Xp;
    else {
      __nv_bfloat16* outs[MAXG];
      for (int g = 0; g < p->groups; ++g) {
        outs[g] = x_base + (int64_t)g * np * p->M * Kp;
        xp[g] = Planes{outs[g], Kp, p->M * Kp};
      }
      VQA_TRY(split_planes(p->X, p->ldx, p->groups, p->M, p->K, p->p, p->seed, p->seed_dev, p->layer, p->drop_index_base,
                           p->drop_bits, outs, Kp, p->M * Kp, np, st));
    }
    Params<EpiWgradT> q = {};
    const int bn = pick_bn(p->N);
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < p->groups ? g : 0;
      VQA_TRY(operand_tmap(&q.tmA[g], xp[s], np, true, p->K, p->M, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], dzt[s], np, false, p->N, p->M, bn));
      q.epi.dW[g] = p->dW[s];
    }
    q.epi.ldw = p->K;
    q.M = (int)p->K; q.N = (int)p->N; q.K = (int)p->M; q.groups = p->groups; q.a_mn = 1; q.b_mn = 0;
    q.k_splits = pick_splits(cdiv(p->K, BM) * cdiv(p->N, bn) * p->groups, p->M);
    if (!p->accumulate_w)
      for (int g = 0; g < p->groups; ++g)
        if (p->dW[g]) cudaMemsetAsync(p->dW[g], 0, (size_t)p->N * p->K * sizeof(float), st);
    VQA_TRY(launch(q, np, bn, st, "tc16_linear_bwd.wgrad"));
  }
  // 3. dgrad: dX[M, K_in] = dZ[M, N_out] . W[N_out, K_in]   (B: the W^T planes [K_in rows, N_out], K-major)
  if (any_x) {
    if (have_wt) wtp[0] = ext->WTp;
    else {
      PackPlanesSeg segs[MAXG] = {};
      for (int g = 0; g < p->groups; ++g) {
        __nv_bfloat16* d = wt_base + (int64_t)g * np * p->K * ldz;
        segs[g].src = p->W[g]; segs[g].rows = p->N; segs[g].rows_pad = ldz; segs[g].K = p->K; segs[g].Kp = Kp;
        segs[g].dstT = d; segs[g].Np = ldz; segs[g].planeT = p->K * ldz; segs[g].t_col0 = 0;
        wtp[g] = Planes{d, ldz, p->K * ldz};
      }
      VQA_TRY(pack_planes(segs, p->groups, np, st));
    }
    if (ext && ext->raw_dx) return dgrad16_launch<false>(p, dzp, wtp, np, st, true);
    return p->pool_alpha ? dgrad16_launch<true>(p, dzp, wtp, np, st) : dgrad16_launch<false>(p, dzp, wtp, np, st);
  }
  return VQA_OK;
}

// ============================================================================================ Mutan
// Workspace: X1 planes [np][M][K1p8] | W1 planes [np][R*Fp][K1p8] | W1^T planes [np][K1][R*Fp] | dH1cat planes [np][M][R*Fp]
size_t tc16_mutan_ws(int np, int R, int64_t M, int64_t K1, int64_t F, int bwd) {
  Carver c{nullptr, 0};
  const int64_t K1p = roundup(K1, 8), Fp = roundup(F, 32), RF = R * Fp;
  c.take(np * M * K1p);
  c.take(np * RF * K1p);
  if (bwd) { c.take(np * K1 * RF); c.take(np * M * RF); }
  return c.off + 256;
}

// x1p: always; w1p / w1tp / dh1: only when the pointer is given
static int mutan16_operands(const char* who, int np, int R, int64_t M, int64_t K1, int64_t F, const float* X1, int64_t ldx1,
                            const float* const* W1, bool bwd, void* ws, size_t ws_bytes, const MutanExt* ext, Planes* x1p,
                            Planes* w1p, Planes* w1tp, __nv_bfloat16** dh1, cudaStream_t st) {
  using namespace tc16;
  const int64_t K1p = roundup(K1, 8), Fp = roundup(F, 32), RF = R * Fp;
  Carver c{reinterpret_cast<char*>(ws), 0};
  if (reinterpret_cast<uintptr_t>(ws) % 256 != 0) c.off = 256 - reinterpret_cast<uintptr_t>(ws) % 256;
  const bool have_x = ext && ext->X1p.p, have_w = !w1p || (ext && ext->W1p.p), have_wt = ext && ext->W1Tp.p;
  __nv_bfloat16* xb = have_x ? nullptr : c.take(np * M * K1p);
  __nv_bfloat16* wb = have_w ? nullptr : c.take(np * RF * K1p);
  __nv_bfloat16* wtb = (bwd && w1tp && !have_wt) ? c.take(np * K1 * RF) : nullptr;
  if (bwd && dh1) *dh1 = c.take(np * M * RF);
  if (!ws || c.off > ws_bytes) return fail16(who, "the workspace is missing or smaller than vqa_mutan_workspace_bytes()");
  if (have_x) *x1p = ext->X1p;
  else {
    __nv_bfloat16* outs[1] = {xb};
    *x1p = Planes{xb, K1p, M * K1p};
    VQA_TRY(split_planes(&X1, &ldx1, 1, M, K1, 0.0f, 0, nullptr, nullptr, nullptr, nullptr, outs, K1p, M * K1p, np, st));
  }
  if (w1p) {
    if (have_w) *w1p = ext->W1p;
    else *w1p = Planes{wb, K1p, RF * K1p};
  }
  if (w1tp) {
    if (have_wt) *w1tp = ext->W1Tp;
    else *w1tp = Planes{wtb, RF, K1 * RF};
  }
  if (!have_w || (w1tp && !have_wt)) {
    PackPlanesSeg segs[MAXG] = {};
    for (int r = 0; r < R; ++r) {
      segs[r].src = W1[r]; segs[r].rows = F; segs[r].rows_pad = Fp; segs[r].K = K1; segs[r].Kp = K1p;
      if (!have_w) { segs[r].dst = wb + (int64_t)r * Fp * K1p; segs[r].plane = RF * K1p; }
      if (w1tp && !have_wt) { segs[r].dstT = wtb; segs[r].Np = RF; segs[r].planeT = K1 * RF; segs[r].t_col0 = (int64_t)r * Fp; }
    }
    VQA_TRY(pack_planes(segs, R, np, st));
  }
  return VQA_OK;
}

// h1 GEMMs of the forward (the small H2 = X2.W2^T GEMM is done by the caller): rank by rank in stream order,
// Y (=|+=) (X1.W1_r^T + b1_r) (.) H2_r
int tc16_mutan_fwd_h1(const vqa_mutan_fwd_params* p, cudaStream_t st, const MutanExt* ext) {
  using namespace tc16;
  const int np = p->math == VQA_MATH_BF16X3 ? 2 : 1;
  const int64_t Mh = p->M / p->rows_per_h2, Fp = roundup(p->F, 32);
  Planes x1p, w1p;
  VQA_TRY(mutan16_operands("vqa_mutan_fwd", np, p->R, p->M, p->K1, p->F, p->X1, p->ldx1, p->W1, false, p->workspace,
                           p->workspace_bytes, ext, &x1p, &w1p, nullptr, nullptr, st));
  const int bn = pick_bn(p->F);
  for (int r = 0; r < p->R; ++r) {
    Params<EpiMutan> q = {};
    Planes wr = w1p;
    wr.p = w1p.p + (int64_t)r * Fp * w1p.ld;
    for (int g = 0; g < MAXG; ++g) {
      VQA_TRY(operand_tmap(&q.tmA[g], x1p, np, false, p->M, p->K1, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], wr, np, false, p->F, p->K1, bn));
      q.epi.bias[g] = p->b1[r]; q.epi.H2[g] = p->H2 + (size_t)r * Mh * p->F;
      q.epi.H1[g] = p->H1 ? p->H1 + (size_t)r * p->M * p->F : nullptr;
    }
    q.M = (int)p->M; q.N = (int)p->F; q.K = (int)p->K1; q.groups = 1; q.k_splits = 1; q.a_mn = 0; q.b_mn = 0;
    q.epi.Y = p->Y; q.epi.ldh = p->F; q.epi.ldy = p->ldy; q.epi.rows_per = p->rows_per_h2;
    q.epi.atomic = 0; q.epi.accumulate = r > 0;
    VQA_TRY(launch(q, np, bn, st, "tc16_mutan_fwd.h1"));
  }
  return VQA_OK;
}

// dW1 and dX1 of the backward from the dH1cat planes (made by the caller's fused mutan_dh kernel into *dh1)
int tc16_mutan_bwd_prepare(const vqa_mutan_bwd_params* p, cudaStream_t st, const MutanExt* ext, Mutan16Ops* ops) {
  const int np = p->math == VQA_MATH_BF16X3 ? 2 : 1;
  ops->np = np;
  return mutan16_operands("vqa_mutan_bwd", np, p->R, p->M, p->K1, p->F, p->X1, p->ldx1, p->W1, true, p->workspace,
                          p->workspace_bytes, ext, &ops->x1p, nullptr, p->dX1 ? &ops->w1tp : nullptr, &ops->dh1, st);
}

int tc16_mutan_bwd_big(const vqa_mutan_bwd_params* p, cudaStream_t st, const Mutan16Ops* ops) {
  using namespace tc16;
  const int np = ops->np, R = p->R;
  const int64_t Fp = roundup(p->F, 32), RF = R * Fp;
  const Planes dh{ops->dh1, RF, p->M * RF};
  {  // dW1_r[f, k] += sum_m dH1_r[m, f] X1[m, k]:  D'[K1, F] = X1^T . dH1_r, both MN-major views, grouped over r
    Params<EpiWgradT> q = {};
    const int bn = 128;
    for (int g = 0; g < MAXG; ++g) {
      const int s = g < R ? g : 0;
      Planes dr = dh;
      dr.p = dh.p + (int64_t)s * Fp;
      VQA_TRY(operand_tmap(&q.tmA[g], ops->x1p, np, true, p->K1, p->M, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], dr, np, true, p->F, p->M, bn));
      q.epi.dW[g] = p->dW1[s];
    }
    q.epi.ldw = p->K1;
    q.M = (int)p->K1; q.N = (int)p->F; q.K = (int)p->M; q.groups = R; q.a_mn = 1; q.b_mn = 1;
    q.k_splits = pick_splits(cdiv(p->K1, BM) * cdiv(p->F, bn) * R, p->M);
    VQA_TRY(launch(q, np, bn, st, "tc16_mutan_bwd.dw1"));
  }
  if (p->dX1) {  // dX1[M, K1] = dH1cat[M, R*Fp] . W1pk[R*Fp, K1]   (B: the W1^T planes [K1 rows, R*Fp], K-major)
    Params<EpiDgradT<false>> q = {};
    const int bn = pick_bn(p->K1);
    for (int g = 0; g < MAXG; ++g) {
      VQA_TRY(operand_tmap(&q.tmA[g], dh, np, false, p->M, RF, BM));
      VQA_TRY(operand_tmap(&q.tmB[g], ops->w1tp, np, false, p->K1, RF, bn));
      q.epi.dX[g] = p->dX1; q.epi.ld[g] = p->lddx1;
    }
    q.M = (int)p->M; q.N = (int)p->K1; q.K = (int)RF; q.groups = 1; q.k_splits = 1; q.a_mn = 0; q.b_mn = 0;
    q.epi.accumulate = p->accumulate_x1; q.epi.drop_on = 0; q.epi.atomic = 0;
    VQA_TRY(launch(q, np, bn, st, "tc16_mutan_bwd.dx1"));
  }
  return VQA_OK;
}

}  // namespace vqa
