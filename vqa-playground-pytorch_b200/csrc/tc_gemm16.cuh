// tcgen05 / TMEM / TMA GEMM over bf16 OPERAND PLANES for sm_100a:   D[M,N] = A[M,K] . B[N,K]^T,  fp32 accumulate in TMEM.
//
// Operands live in HBM as bf16 planes written by their producers (the input split/mask pass, the weight pack, the
// epilogue of the GEMM that made them): NP = 1 plane = plain bf16 (the reduced-precision mode), NP = 2 planes
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi), |x - hi - lo| <= 2^-16 |x|) = the fp32-parity mode, computed as
// hi.hi + hi.lo + lo.hi: three kind::f16 MMAs per k-step, half the tensor-pipe time and half the shared-memory
// operand reads of 3xTF32 for the same bytes from HBM (4 per element), and NO operand transform in the kernel — the
// main loop is TMA -> tcgen05.mma only (tc_gemm.cuh's transform stage made that kernel shared-memory-bandwidth bound).
//
// Persistent: one CTA per SM walks a static list of work items (group, k-split, m-tile, n-tile; n fastest so that
// concurrently running CTAs share A tiles in L2).  Warp roles (320 threads):
//   warp 0      TMA producer: 3-D boxes (k, rows, plane) / (mn, k, plane), 128-byte swizzle, NS-stage ring
//   warp 1      TMEM allocator + the single thread that issues tcgen05.mma / tcgen05.commit
//   warps 2-9   epilogue: the accumulator is DOUBLE-BUFFERED in TMEM (2 x ACC_COLS columns), so the epilogue of tile t
//               runs under the main loop of tile t+1.  It moves 32 accumulator columns at a time through a
//               double-buffered shared-memory slab (tcgen05.ld -> st.shared -> one named barrier -> functor), with the
//               same functors as tc_gemm.cuh (tc_epilogues.cuh).  (Warp-private slabs without the barrier were
//               measured 5-10 % slower on every epilogue-bound launch: profiles/r2_experiments.md.)
// Operands may be K-major ([rows, K] row-major) or MN-major ([K, rows] row-major, the transposed view used by
// wgrad/dgrad); an MN-major B needs BN % 64 == 0.  Descriptor layouts follow cute/arch/mma_sm100_desc.hpp and
// cute/atom/mma_traits_sm100.hpp (make_umma_desc, 16-bit SWIZZLE_128B canonical layouts).
#pragma once
#include "tc_epilogues.cuh"

namespace vqa {
namespace tc16 {

using tc::MAXG;
using tc::smem_u32;

constexpr int BM = 128;
constexpr int BK = 64;                      // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int A_PLANE_BYTES = BM * 128;     // 16 KB
constexpr int MN_CHUNK_BYTES = BK * 128;    // MN-major: one 64-element MN chunk of one plane = 64 k-rows x 128 B
constexpr int EPI_THREADS = 256;
constexpr int NUM_THREADS = 64 + EPI_THREADS;
constexpr int SLAB_LD = 36;                 // row-major slab: 128 rows x (32 + 4) floats
constexpr int SLAB_LDT = BM + 4;            // transposed slab: 32 columns x (128 + 4) floats
constexpr int SLAB_BYTES = BM * SLAB_LD * 4;   // 18432 >= 32 * SLAB_LDT * 4 = 16896

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

// D[tmem] (+)= A[smem] . B[smem], bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 16 consecutive fp32 columns of the accumulator -> 16 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, 16-bit operands, SWIZZLE_128B, sm_100 version bit set.
//   K-major : rows of 128 B (64 bf16 along K); 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: rows of 128 B (64 bf16 along MN), one row per k; 8-k-row groups 1024 B apart (SBO); the next 64 MN
//             elements lbo_bytes further (LBO).
__device__ __forceinline__ uint64_t make_desc16(uint32_t saddr, bool mn_major, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)(mn_major ? (lbo_bytes >> 4) : 1u) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;      // descriptor version (Blackwell)
  d |= 2ull << 61;      // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = bn
__host__ __device__ constexpr uint32_t make_idesc16(int bn, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <class Epi>
struct Params {
  CUtensorMap tmA[MAXG];   // 3-D: K-major (k, rows, plane) box (64, 128, NP); MN-major (mn, k, plane) box (64, 64, NP)
  CUtensorMap tmB[MAXG];   //      K-major box (64, BN, NP);                    MN-major box (64, 64, NP)
  int M, N, K;             // D is MxN, reduction length K
  int groups, k_splits;
  int a_mn, b_mn;
  Epi epi;
};

template <int BN, int NP>
struct Cfg {
  static constexpr int B_PLANE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = NP * (A_PLANE_BYTES + B_PLANE_BYTES);
  static constexpr int NS_ = (225 * 1024 - 2 * SLAB_BYTES - 1024 - 256) / STAGE_BYTES;
  static constexpr int NS = NS_ > 6 ? 6 : NS_;
  static constexpr int ACC_COLS = BN <= 128 ? 128 : 256;       // column offset of the second accumulator
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = NS * STAGE_BYTES + 2 * SLAB_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(NS >= 2, "tile too large for a two-stage ring");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN, int NP, class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm16_kernel(const __grid_constant__ Params<Epi> p) {
  using C = Cfg<BN, NP>;
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int NS = C::NS;
  uint8_t* slabs = smem + NS * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + 2 * SLAB_BYTES);
  uint64_t* full = bars;                   // [NS] TMA bytes landed
  uint64_t* empty = bars + NS;             // [NS] the MMAs that read the stage have completed
  uint64_t* acc_full = bars + 2 * NS;      // [2]  all MMAs of the tile in accumulator buffer b have completed
  uint64_t* acc_empty = bars + 2 * NS + 2; // [2]  the epilogue has read accumulator buffer b out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
  const int work_total = p.groups * p.k_splits * tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_THREADS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (group, split, m-tile, n-tile), n fastest
  auto decode = [&](int w, int& g, int& split, int& tm, int& tn) {
    tn = w % tiles_n; w /= tiles_n;
    tm = w % tiles_m; w /= tiles_m;
    split = w % p.k_splits; g = w / p.k_splits;
  };
  auto stage_a = [&](int s) { return smem + s * C::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem + s * C::STAGE_BYTES + NP * A_PLANE_BYTES; };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
        int g, split, tm, tn;
        decode(w, g, split, tm, tn);
        const int kb_begin = split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
        const int m0 = tm * BM, n0 = tn * BN;
        if (w == (int)blockIdx.x) { tma_prefetch_desc(&p.tmA[g]); tma_prefetch_desc(&p.tmB[g]); }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % NS;
          mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
          mbar_expect_tx(&full[s], C::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_3d(stage_a(s), &p.tmA[g], &full[s], k0, m0, 0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_3d(stage_a(s) + j * NP * MN_CHUNK_BYTES, &p.tmA[g], &full[s], m0 + j * 64, k0, 0);
          }
          if (!p.b_mn) {
            tma_load_3d(stage_b(s), &p.tmB[g], &full[s], k0, n0, 0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(stage_b(s) + j * NP * MN_CHUNK_BYTES, &p.tmB[g], &full[s], n0 + j * 64, k0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc16(BN, p.a_mn != 0, p.b_mn != 0);
      const uint32_t a_step = p.a_mn ? 2048u : 32u;          // bytes per k-step of 16 bf16
      const uint32_t b_step = p.b_mn ? 2048u : 32u;
      const uint32_t a_lo = p.a_mn ? (uint32_t)MN_CHUNK_BYTES : (uint32_t)A_PLANE_BYTES;       // offset of plane 1
      const uint32_t b_lo = p.b_mn ? (uint32_t)MN_CHUNK_BYTES : (uint32_t)C::B_PLANE_BYTES;
      const uint32_t lbo = (uint32_t)(NP * MN_CHUNK_BYTES);
      int it = 0, ti = 0;
      for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
        int g, split, tm, tn;
        decode(w, g, split, tm, tn);
        const int kb_begin = split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
        const int nkb = kb_end - kb_begin;
        if (nkb <= 0) continue;
        const int buf = ti & 1;
        mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * C::ACC_COLS);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % NS;
          mbar_wait(&full[s], (it / NS) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_a(s)), b_addr = smem_u32(stage_b(s));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_desc16(a_addr + k * a_step, p.a_mn != 0, lbo);
            const uint64_t db = make_desc16(b_addr + k * b_step, p.b_mn != 0, lbo);
            umma_bf16(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
            if (NP == 2) {
              const uint64_t dal = make_desc16(a_addr + a_lo + k * a_step, p.a_mn != 0, lbo);
              const uint64_t dbl = make_desc16(b_addr + b_lo + k * b_step, p.b_mn != 0, lbo);
              umma_bf16(tmem_d, da, dbl, idesc, 1u);
              umma_bf16(tmem_d, dal, db, idesc, 1u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
        ++ti;
      }
    }
  } else {
    // ===================================================== epilogue warps
    const int t = threadIdx.x - 64;
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;             // which 16 of a slab's 32 columns this warp moves
    const int row = q * 32 + lane;
    constexpr int NCB = BN / 32;
    int ti = 0, sc = 0;                           // tiles and slabs done so far (buffer parities)
    for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
      int g, split, tm, tn;
      decode(w, g, split, tm, tn);
      const int kb_begin = split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
      if (kb_end - kb_begin <= 0) continue;
      const int buf = ti & 1;
      mbar_wait(&acc_full[buf], (ti >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * C::ACC_COLS) + ((uint32_t)(q * 32) << 16);
      const int m0 = tm * BM, n0 = tn * BN;
      const int rows = min(BM, p.M - m0);
#pragma unroll 1
      for (int cb = 0; cb < NCB; ++cb, ++sc) {
        const uint32_t sl = smem_u32(slabs + (sc & 1) * SLAB_BYTES);
        {
          float v[16];
          tmem_ld16(tmem_acc + (uint32_t)(cb * 32 + half * 16), v);
          if (cb == NCB - 1) {                    // this thread's last read of the buffer: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
          }
          if constexpr (Epi::kStaged) {
#pragma unroll
            for (int c = 0; c < 16; c += 4)
              sts128(sl + (uint32_t)(row * SLAB_LD + half * 16 + c) * 4u, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(sl + (uint32_t)((half * 16 + c) * SLAB_LDT + row) * 4u), "f"(v[c])
                           : "memory");
          }
        }
        // One barrier per slab is enough with two slab buffers: whoever overwrites buffer b two slabs later has
        // passed the barrier of the slab in between, which every thread reaches only after it finished reading b.
        named_bar_sync(1, EPI_THREADS);
        if constexpr (Epi::kStaged) {
          // a thread keeps ONE 4-column group of the slab and takes rows r0, r0+32, r0+64, r0+96: every global read
          // the functor needs for the four rows is issued before the first dependent instruction
          const int c4 = t & 7, r0 = t >> 3;
          const int n = n0 + cb * 32 + c4 * 4;
          if (n < p.N) {
            typename Epi::Col col;
            p.epi.column(col, g, split, n, p.N);
            typename Epi::Pre pre[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * 32;
              if (r < rows) p.epi.preload(pre[u], col, m0 + r);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * 32;
              if (r < rows) p.epi.row4(col, m0 + r, lds128(sl + (uint32_t)(r * SLAB_LD + c4 * 4) * 4u), pre[u]);
            }
          }
        } else {
          // transposed output (wgrad): the tile ROW index is the contiguous index of the destination; every thread adds
          // 4 consecutive rows of one column with a single 16-byte reduction
          const int mq = t & 31;
          const int m = m0 + mq * 4;
          if (m < p.M) {
            typename Epi::Row rw;
            p.epi.rowquad(rw, g, m, p.M);
#pragma unroll
            for (int cc = t >> 5; cc < 32; cc += EPI_THREADS / 32) {
              const int n = n0 + cb * 32 + cc;
              if (n < p.N) p.epi.col4(rw, n, lds128(sl + (uint32_t)(cc * SLAB_LDT + mq * 4) * 4u));
            }
          }
        }
      }
      ++ti;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

}  // namespace tc16
}  // namespace vqa
