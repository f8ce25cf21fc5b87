// tcgen05 / TMEM / TMA GEMM for sm_100a:   D[M,N] = A[M,K] . B[N,K]^T   fp32 storage, TF32 tensor-core
// math with fp32 accumulation in TMEM, optionally error-compensated 3xTF32 (x = hi + lo, three MMAs),
// which is what the "fp32 parity" mode needs (single TF32 misses the 1e-4 tolerance).
//
// One CTA computes a 128 x BN tile (of one k-split).  Warp roles (320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      TMEM allocator + the single thread that issues tcgen05.mma and tcgen05.commit
//   warps 2-9   operand transform in shared memory between TMA landing and MMA issue (input dropout on A — from the
//               keep-bit cache, fetched one k-block ahead, or Philox regenerated in registers — and the hi/lo split
//               for 3xTF32; every thread keeps several independent 16-byte chunks in flight), then the epilogue:
//               tcgen05.ld the accumulator rows (two warps per TMEM lane quadrant, splitting the columns), stage the
//               tile in shared memory, and let every thread walk one 4-column group (or, for wgrad, one 4-row group
//               of the transposed tile) with all of a batch's global reads issued before their first use
// Operands are described by TMA tensor maps and may be K-major (row-major [rows,K]) or MN-major
// (row-major [K,rows], i.e. the transposed view used by wgrad/dgrad) — no transposed copies are made.
// Shared-memory/instruction descriptor layouts follow cute/arch/mma_sm100_desc.hpp and
// cute/atom/mma_traits_sm100.hpp (make_umma_desc).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vqa {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                    // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BM * 128;
constexpr int ATOM_BYTES = BK * 128;      // one MN-major swizzle atom column: 32 k-rows x 128 B
constexpr int XFORM_THREADS = 256;
constexpr int NUM_THREADS = 64 + XFORM_THREADS;
constexpr int MAXG = VQA_MAX_GROUPS;

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-space 128-bit accesses (generic LD/ST on a casted pointer take the slow global path)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// same load, delivered to the same shared-memory offset (and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_ctaid_y() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.y;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_y() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctaid.y;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem], TF32 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// commit that arrives on the same-offset mbarrier of every CTA in cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------ descriptors
// UMMA shared-memory descriptor, SWIZZLE_128B, sm_100 version bit set.
//   K-major : rows of 128 B (32 tf32 along K), 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: 32-bit operands must use SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major tf32
//             operands, SW128_32B is the only available smem layout"): rows of 128 B (32 elements along MN),
//             32-byte chunks XOR-ed with (row & 3); 4-row k-groups 512 B apart (SBO), next 32 MN elements
//             ATOM_BYTES apart (LBO).  TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)(mn_major ? (ATOM_BYTES >> 4) : 1u) << 16;
  d |= (uint64_t)(mn_major ? (512u >> 4) : (1024u >> 4)) << 32;
  d |= 1ull << 46;                      // descriptor version (Blackwell)
  d |= (mn_major ? 1ull : 2ull) << 61;  // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return d;
}
// kind::tf32 instruction descriptor: fp32 accumulate, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int bn, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ params
struct GroupDrop {
  uint32_t layer[MAXG];
  uint64_t base[MAXG];
};

template <class Epi>
struct Params {
  CUtensorMap tmA[MAXG];
  CUtensorMap tmB[MAXG];
  int M, N, K;            // D is MxN, reduction length K
  int k_splits;           // blockIdx.z = group * k_splits + split
  int a_mn, b_mn;         // operand major-ness
  int box_split;          // K-major tiles are fetched as 32-row boxes (needed for cluster multicast) instead of one box
  int debug;              // timing experiments only (VQA_TC_DEBUG): 1 skip A transform, 2 skip B transform, 4 no transform stage
  int rewrite_hi;         // 3xTF32: store the truncated hi part back (0: rely on the MMA ignoring the low 13 bits)
  int drop_on;            // Philox dropout on the A operand
  Drop drop;              // seed / thr / scale (layer + base per group below)
  GroupDrop gd;
  int64_t drop_ld;        // row length of the logical tensor the mask is indexed in
  int64_t drop_rows;      // its row count (bounds for the bit-mask loads)
  const uint8_t* drop_bits[MAXG];   // optional packed keep-bits (vqa_dropout_bits); else Philox in registers
  Epi epi;
};

// Shared-memory rings.  The TMA-landed ("raw") tiles are prefetched NS_RAW deep to cover the L2/HBM latency;
// the 3xTF32 residual ("lo") tiles are produced by the transform warps just ahead of the MMA and only need
// NS_LO = 2 slots, which is what lets the raw ring be deep despite the 227 KB limit.
// OCC = 2 is the two-CTAs-per-SM variant for launches of more than one wave: half the shared memory each (2 raw
// stages + 1 residual slot), so that the prologue/epilogue of one CTA overlaps the main loop of its neighbour; the
// SM as a whole still has 4 raw stages in flight.
template <int BN, bool X3, int OCC = 1>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int RAW_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int LO_BYTES = X3 ? RAW_BYTES : 0;
  static constexpr int NS_LO = X3 ? (OCC == 2 ? 1 : 2) : 0;
  static constexpr int BUDGET = OCC == 2 ? 110 * 1024 : 220 * 1024;
  static constexpr int NS_RAW_ = (BUDGET - NS_LO * LO_BYTES) / RAW_BYTES;
  static constexpr int NS_RAW = NS_RAW_ > 6 ? 6 : NS_RAW_;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : (BN <= 256 ? 256 : 512)));
  static constexpr int TILE_BYTES = NS_RAW * RAW_BYTES + NS_LO * LO_BYTES;
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(NS_RAW >= 2, "tile too large");
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
  static_assert(SMEM_BYTES <= 227 * 1024 / OCC, "shared memory budget");
  static_assert(OCC == 1 || TMEM_COLS <= 256, "two resident CTAs share the 512 TMEM columns");
};

// split a 16-byte chunk in place into tf32-representable hi and the fp32 residual lo
__device__ __forceinline__ void split4(float4& v, float4& lo) {
  float4 hi;
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
  v = hi;
}

// ------------------------------------------------------------------------------------------ kernel
template <int BN, bool X3, class Epi, int OCC = 1>
__global__ void __launch_bounds__(NUM_THREADS, OCC) tc_gemm_kernel(const __grid_constant__ Params<Epi> p) {
  using C = Cfg<BN, X3, OCC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int NR = C::NS_RAW, NL = C::NS_LO > 0 ? C::NS_LO : 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::TILE_BYTES);
  uint64_t* full = bars;                      // [NR] TMA bytes landed in raw slot
  uint64_t* empty = bars + NR;                // [NR] MMAs that read raw slot have completed
  uint64_t* ready = bars + 2 * NR;            // [NR] transform of the k-block in raw slot done
  uint64_t* lo_empty = bars + 3 * NR;         // [NL] MMAs that read lo slot have completed
  uint64_t* accum_full = bars + 3 * NR + NL;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NR + NL + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z / p.k_splits, split = blockIdx.z % p.k_splits;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
  const int kb_begin = split * kb_per;
  const int kb_end = min(kb_total, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);
  const bool need_xform = (X3 || p.drop_on) && !(p.debug & 4);

  // Thread-block cluster (CM x CN CTAs = CM adjacent m-tiles x CN adjacent n-tiles of the same k-split): the 4 KB
  // boxes of an A tile are fetched once per cluster ROW and multicast to its CN CTAs, the boxes of a B tile once
  // per cluster COLUMN and multicast to its CM CTAs, which divides the L2->SM operand traffic by CN resp. CM.
  const uint32_t CM = cluster_nctaid_x(), CN = cluster_nctaid_y();
  const uint32_t mr = cluster_ctaid_x(), nr = cluster_ctaid_y();
  uint16_t mask_a = 0, mask_b = 0;                      // cluster ranks are x + y*CM
  for (uint32_t j = 0; j < CN; ++j) mask_a |= (uint16_t)(1u << (mr + j * CM));
  for (uint32_t i = 0; i < CM; ++i) mask_b |= (uint16_t)(1u << (i + nr * CM));
  const uint16_t mask_all = mask_a | mask_b;
  const bool clustered = CM * CN > 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NR; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], XFORM_THREADS);
      mbar_init(&empty[s], (uint32_t)__popc((uint32_t)mask_all));   // every CTA that reads what this CTA loads
    }
    for (int s = 0; s < NL; ++s) mbar_init(&lo_empty[s], 1);
    mbar_init(accum_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA[g]);
    tma_prefetch_desc(&p.tmB[g]);
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();          // peers' barriers are initialised before anything is multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_a = [&](int s) { return smem + s * C::RAW_BYTES; };
  auto stage_b = [&](int s) { return smem + s * C::RAW_BYTES + A_TILE_BYTES; };
  auto stage_alo = [&](int l) { return smem + NR * C::RAW_BYTES + l * C::LO_BYTES; };
  auto stage_blo = [&](int l) { return smem + NR * C::RAW_BYTES + l * C::LO_BYTES + A_TILE_BYTES; };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      for (int it = 0; it < nkb; ++it) {
        const int s = it % NR;
        const uint32_t ph = (it / NR) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_TILE_BYTES + C::B_TILE_BYTES);
        const int k0 = (kb_begin + it) * BK;
        // MN-major tiles (and K-major ones under a cluster) are fetched as 4 KB boxes; box j is issued by one CTA
        // of the cluster row (A) or column (B) and multicast to the others.  Un-clustered K-major tiles are one
        // box each: fewer, larger TMA requests measured 2x faster on the 1-pass kernel.
        if (!p.a_mn && !p.box_split) {
          tma_load_2d(stage_a(s), &p.tmA[g], &full[s], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) {
            if ((uint32_t)j % CN != nr) continue;
            const int x = p.a_mn ? m0 + j * 32 : k0, y = p.a_mn ? k0 : m0 + j * 32;
            if (CN > 1) tma_load_2d_mc(stage_a(s) + j * ATOM_BYTES, &p.tmA[g], &full[s], x, y, mask_a);
            else tma_load_2d(stage_a(s) + j * ATOM_BYTES, &p.tmA[g], &full[s], x, y);
          }
        }
        if (!p.b_mn && !p.box_split) {
          tma_load_2d(stage_b(s), &p.tmB[g], &full[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            if ((uint32_t)j % CM != mr) continue;
            const int x = p.b_mn ? n0 + j * 32 : k0, y = p.b_mn ? k0 : n0 + j * 32;
            if (CM > 1) tma_load_2d_mc(stage_b(s) + j * ATOM_BYTES, &p.tmB[g], &full[s], x, y, mask_b);
            else tma_load_2d(stage_b(s) + j * ATOM_BYTES, &p.tmB[g], &full[s], x, y);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN, p.a_mn != 0, p.b_mn != 0);
      const uint32_t a_step = p.a_mn ? 1024u : 32u;   // bytes per k-step of 8 tf32
      const uint32_t b_step = p.b_mn ? 1024u : 32u;
      for (int it = 0; it < nkb; ++it) {
        const int s = it % NR, l = it % NL;
        const uint32_t ph = (it / NR) & 1;
        mbar_wait(need_xform ? &ready[s] : &full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(stage_a(s)), b_addr = smem_u32(stage_b(s));
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t da = make_smem_desc(a_addr + k * a_step, p.a_mn != 0);
          const uint64_t db = make_smem_desc(b_addr + k * b_step, p.b_mn != 0);
          umma_tf32(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          if (X3) {
            const uint64_t dal = make_smem_desc(smem_u32(stage_alo(l)) + k * a_step, p.a_mn != 0);
            const uint64_t dbl = make_smem_desc(smem_u32(stage_blo(l)) + k * b_step, p.b_mn != 0);
            umma_tf32(tmem_base, da, dbl, idesc, 1u);
            umma_tf32(tmem_base, dal, db, idesc, 1u);
          }
        }
        if (clustered) umma_commit_mc(&empty[s], mask_all);
        else umma_commit(&empty[s]);
        if (X3) umma_commit(&lo_empty[l]);
      }
      umma_commit(accum_full);
    }
  } else {
    // ===================================================== transform warps, then epilogue
    const int t = threadIdx.x - 64;
    if (need_xform) {
      Drop d = p.drop;
      d.seed = d.key();
      d.layer = p.gd.layer[g];
      d.base = p.gd.base[g];
      const uint8_t* __restrict__ bits = p.drop_bits[g];
      constexpr int A_CH = A_TILE_BYTES / 16 / XFORM_THREADS;                       // 4
      constexpr int B_CH = (C::B_TILE_BYTES / 16 + XFORM_THREADS - 1) / XFORM_THREADS;
      // Packed keep-bits: the byte holding the 4 mask bits of chunk i at k-block kb sits at bit_ptr[i] + kb * bit_step
      // (the element index is affine in the k coordinate and k-blocks start on multiples of 32 elements), so the
      // index arithmetic is done once here.  The bytes of k-block it+1 are requested while k-block it is processed.
      const uint8_t* bit_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
      uint32_t bit_sh[4] = {0, 0, 0, 0};         // first mask bit of the chunk inside its byte; > 4: runs into the next
      int bit_klim[4] = {0, 0, 0, 0};           // chunk i is inside the tensor while k0 < bit_klim[i]
      int64_t bit_step = 0;
      const bool use_bits = p.drop_on && bits != nullptr;
      if (use_bits) {
        bit_step = p.a_mn ? (int64_t)(BK / 8) * p.drop_ld : (int64_t)(BK / 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = t + i * XFORM_THREADS;
          const int pc = ch & 7;
          int64_t e;
          bool ok;
          if (!p.a_mn) {
            const int r = ch >> 3, cc = (pc ^ (r & 7)) << 2;
            e = (int64_t)(m0 + r) * p.drop_ld + cc;
            ok = (m0 + r) < p.drop_rows;
            bit_klim[i] = ok ? (int)min((int64_t)INT32_MAX, p.drop_ld - cc) : 0;
          } else {
            const int j = ch >> 8, r = (ch >> 3) & 31;
            const int col = m0 + j * 32 + (((((pc >> 1) ^ (r & 3)) << 1) | (pc & 1)) << 2);
            e = (int64_t)r * p.drop_ld + col;
            ok = col < p.drop_ld;
            bit_klim[i] = ok ? (int)min((int64_t)INT32_MAX, p.drop_rows - r) : 0;
          }
          bit_ptr[i] = bits + (e >> 3);
          bit_sh[i] = (uint32_t)(e & 7);
        }
      }
      // rows that are not a multiple of 4 elements long put some quads across a byte boundary: those launches fetch
      // the following byte as well (kept apart until use so that no load is waited for inside the fetch)
      const bool wide_bits = use_bits && (p.drop_ld & 3) != 0;
      uint32_t kbyte[4] = {0xFFu, 0xFFu, 0xFFu, 0xFFu}, kbyte_next[4] = {0xFFu, 0xFFu, 0xFFu, 0xFFu};
      uint32_t khigh[4] = {0xFFu, 0xFFu, 0xFFu, 0xFFu}, khigh_next[4] = {0xFFu, 0xFFu, 0xFFu, 0xFFu};
      auto fetch_bits = [&](int it, uint32_t (&dst)[4], uint32_t (&dsth)[4]) {
        const int kb = kb_begin + it;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint8_t* bp = bit_ptr[i] + (int64_t)kb * bit_step;
          const bool in = kb * BK < bit_klim[i];
          dst[i] = in ? (uint32_t)__ldg(bp) : 0xFFu;
          if (wide_bits) dsth[i] = in ? (uint32_t)__ldg(bp + 1) : 0xFFu;
        }
      };
      if (use_bits && nkb > 0) fetch_bits(0, kbyte, khigh);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % NR, l = it % NL;
        const uint32_t ph = (it / NR) & 1;
        const int k0 = (kb_begin + it) * BK;
        if (use_bits && it + 1 < nkb) fetch_bits(it + 1, kbyte_next, khigh_next);
        if (X3) mbar_wait(&lo_empty[l], ((it / NL) & 1) ^ 1);     // lo slot free (MMAs of k-block it-NL done)
        mbar_wait(&full[s], ph);
        if (!(p.debug & 1)) {
          const uint32_t a = smem_u32(stage_a(s));
          const uint32_t alo = smem_u32(stage_alo(l));
          float4 v[A_CH];
#pragma unroll
          for (int i = 0; i < A_CH; ++i) v[i] = lds128(a + (t + i * XFORM_THREADS) * 16);
          if (use_bits) {
#pragma unroll
            for (int i = 0; i < A_CH; ++i) {
              const uint32_t nb = (wide_bits ? (kbyte[i] | (khigh[i] << 8)) : kbyte[i]) >> bit_sh[i];
              v[i].x = (nb & 1u) ? v[i].x * d.scale : 0.0f;
              v[i].y = (nb & 2u) ? v[i].y * d.scale : 0.0f;
              v[i].z = (nb & 4u) ? v[i].z * d.scale : 0.0f;
              v[i].w = (nb & 8u) ? v[i].w * d.scale : 0.0f;
            }
          } else if (p.drop_on) {
            // The 4 lanes of a quartet hold the 4 chunks of one aligned 64-byte run (= 16 consecutive logical
            // elements = one Philox group) in every pass; lane (t & 3) == i computes the group of pass i and the
            // quartet shares it by shuffle: one Philox call per thread per k-block instead of four.
            static_assert(A_CH == 4, "quartet sharing assumes 4 passes");
            uint64_t idx[A_CH];
#pragma unroll
            for (int i = 0; i < A_CH; ++i) {
              const int ch = t + i * XFORM_THREADS;
              const int pc = ch & 7;
              int64_t row, col;
              if (!p.a_mn) {
                const int r = ch >> 3;
                row = m0 + r;
                col = k0 + ((pc ^ (r & 7)) << 2);
              } else {
                // 128B_ATOM_32B swizzle: 32-byte chunk index XOR (row & 3), 16-byte half unchanged
                const int j = ch >> 8, r = (ch >> 3) & 31;
                row = k0 + r;
                col = m0 + j * 32 + (((((pc >> 1) ^ (r & 3)) << 1) | (pc & 1)) << 2);
              }
              idx[i] = d.base + (uint64_t)(row * p.drop_ld + col);
            }
            const int qi = t & 3;
            const uint64_t my_idx = qi == 0 ? idx[0] : (qi == 1 ? idx[1] : (qi == 2 ? idx[2] : idx[3]));
            uint4 mine;
            const bool aligned = (p.drop_ld & 15) == 0 && (d.base & 15) == 0;   // groups never straddle quartets
            if (aligned) mine = philox_group(d.seed, d.layer, my_idx >> 4);
#pragma unroll
            for (int i = 0; i < A_CH; ++i) {
              uint32_t bt;
              if (aligned) {
                const int src = (lane & ~3) | i;
                uint4 r;
                r.x = __shfl_sync(0xffffffffu, mine.x, src);
                r.y = __shfl_sync(0xffffffffu, mine.y, src);
                r.z = __shfl_sync(0xffffffffu, mine.z, src);
                r.w = __shfl_sync(0xffffffffu, mine.w, src);
                bt = pick_word(r, ((uint32_t)idx[i] >> 2) & 3u);
              } else {
                bt = philox_bytes4(d.seed, d.layer, idx[i]);
              }
              v[i].x = (bt & 0xFFu) >= d.thr ? v[i].x * d.scale : 0.0f;
              v[i].y = ((bt >> 8) & 0xFFu) >= d.thr ? v[i].y * d.scale : 0.0f;
              v[i].z = ((bt >> 16) & 0xFFu) >= d.thr ? v[i].z * d.scale : 0.0f;
              v[i].w = (bt >> 24) >= d.thr ? v[i].w * d.scale : 0.0f;
            }
          }
#pragma unroll
          for (int i = 0; i < A_CH; ++i) {
            const int off = (t + i * XFORM_THREADS) * 16;
            if (X3) {
              float4 lo;
              split4(v[i], lo);
              sts128(alo + off, lo);
              if (p.rewrite_hi || p.drop_on) sts128(a + off, v[i]);
            } else {
              sts128(a + off, v[i]);
            }
          }
        }
        if (X3 && !(p.debug & 2)) {
          const uint32_t b = smem_u32(stage_b(s));
          const uint32_t blo = smem_u32(stage_blo(l));
          float4 v[B_CH];
#pragma unroll
          for (int i = 0; i < B_CH; ++i) {
            const int ch = t + i * XFORM_THREADS;
            if (ch < C::B_TILE_BYTES / 16) v[i] = lds128(b + ch * 16);
          }
#pragma unroll
          for (int i = 0; i < B_CH; ++i) {
            const int ch = t + i * XFORM_THREADS;
            if (ch < C::B_TILE_BYTES / 16) {
              float4 lo;
              split4(v[i], lo);
              sts128(blo + ch * 16, lo);
              if (p.rewrite_hi) sts128(b + ch * 16, v[i]);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&ready[s]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { kbyte[i] = kbyte_next[i]; khigh[i] = khigh_next[i]; }
      }
    }
    // ---- epilogue: two warps per TMEM lane quadrant (warp % 4), each takes half of the column blocks
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    constexpr int NCB = BN / 32;
    const int cb0 = half == 0 ? 0 : (NCB + 1) / 2;
    const int cb1 = half == 0 ? (NCB + 1) / 2 : NCB;
    if (nkb > 0) {
      mbar_wait(accum_full, 0);
      tc_fence_after();
      if constexpr (Epi::kStaged) {
        // accumulator rows -> shared (the pipeline stages are free now) -> coalesced row-major epilogue
        constexpr int LDC = BN + 4;
        const uint32_t cs = smem_u32(smem);
#pragma unroll 1
        for (int cb = cb0; cb < cb1; ++cb) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            sts128(cs + (uint32_t)(row * LDC + cb * 32 + c) * 4u, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
        }
        named_bar_sync(1, XFORM_THREADS);
        // A thread keeps ONE 4-column group of the tile and walks down its rows, U rows at a time in two phases:
        // first every global read the functor needs for those rows (old values of a "+=", keep-bits, the
        // multiplicand of the Mutan product) is issued, then the results are combined and stored.  A
        // load->use->store chain per row keeps a single DRAM/L2 round trip in flight per thread and made these
        // epilogues latency-bound.
        constexpr int V4 = BN / 4;
        constexpr int RS = XFORM_THREADS / V4;            // rows covered per pass (BN = 160: 16 threads sit out)
        constexpr int U = OCC == 2 && Epi::kBatch > 4 ? 4 : Epi::kBatch;     // 102 registers per thread at OCC = 2
        const int c4 = t % V4, r0 = t / V4;
        const int n = n0 + c4 * 4;
        if (r0 < RS && n < p.N) {
          typename Epi::Col col;
          p.epi.column(col, g, split, n, p.N);
          const int rows = min(BM, p.M - m0);
#pragma unroll 1
          for (int rb = r0; rb < rows; rb += RS * U) {
            typename Epi::Pre pre[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int r = rb + u * RS;
              if (r < rows) p.epi.preload(pre[u], col, m0 + r);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int r = rb + u * RS;
              if (r < rows) p.epi.row4(col, m0 + r, lds128(cs + (uint32_t)(r * LDC + c4 * 4) * 4u), pre[u]);
            }
          }
        }
      } else {
        // Transposed output (wgrad: the tile row index is the CONTIGUOUS index of the destination): the accumulator
        // is staged column-major, then every thread adds 4 consecutive rows of one column with a single 16-byte
        // reduction.  Scalar reds cost ~1.3 issue cycles per lane on the SM (20480 of them per tile was most of a
        // small wgrad launch); the vector form moves 4 values per lane-op.
        constexpr int LDT = BM + 4;
        static_assert(BN * LDT * 4 <= C::TILE_BYTES, "transposed staging reuses the ring");
        const uint32_t cs = smem_u32(smem);
#pragma unroll 1
        for (int cb = cb0; cb < cb1; ++cb) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
#pragma unroll
          for (int c = 0; c < 32; ++c)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(cs + (uint32_t)((cb * 32 + c) * LDT + row) * 4u), "f"(v[c]) : "memory");
        }
        named_bar_sync(1, XFORM_THREADS);
        const int mq = t & 31;                             // 32 row-quads span the 128 tile rows
        const int m = m0 + mq * 4;
        if (m < p.M) {
          typename Epi::Row rw;
          p.epi.rowquad(rw, g, m, p.M);
#pragma unroll 4
          for (int cc = t >> 5; cc < BN; cc += XFORM_THREADS / 32) {
            const int n = n0 + cc;
            if (n < p.N) p.epi.col4(rw, n, lds128(cs + (uint32_t)(cc * LDT + mq * 4) * 4u));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();          // no CTA leaves while peers may still signal its barriers
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

}  // namespace tc
}  // namespace vqa
