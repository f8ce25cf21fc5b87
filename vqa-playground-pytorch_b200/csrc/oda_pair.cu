// ODA object-difference attention (include/vqacore.h: vqa_oda_pair_attn_{fwd,bwd}).
// Replaces the 36x36 Python pair loop, the stack/transpose/contiguous copy and the 11160-channel
// conv_att of the reference (config/ODA.py:216-226, :192).  The [B,N,N*H] tensor never exists:
//   eval : z[b,i,g] = sum_k ql[b,k] vl[b,i,k] Wsum[g,k] (+ terms constant in i, which the region
//          softmax cancels) — attention.cuh kernels with the FuseOdaEval source;
//   train: every (i,j,k) term is formed in registers with its Philox keep-bit.
#include "attention.cuh"

namespace vqa {

// wsum[g,k] = sum_j W[g, j*H + k].  grid = G
__global__ void oda_wsum_kernel(int64_t N, int64_t H, const float* __restrict__ W, float* __restrict__ wsum) {
  const int g = blockIdx.x;
  for (int64_t k = threadIdx.x; k < H; k += blockDim.x) {
    float s = 0.0f;
    for (int64_t j = 0; j < N; ++j) s += W[(g * N + j) * H + k];
    wsum[g * H + k] = s;
  }
}

// dW[g, j*H+k] (+)= dwsum[g,k] for every j (eval-mode gradient of the factorised form).
__global__ void oda_dw_broadcast_kernel(int64_t N, int64_t H, const float* __restrict__ dwsum, float* __restrict__ dW,
                                        int accumulate) {
  const int64_t total = (int64_t)G * N * H;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t % H, g = t / (N * H);
    const float v = dwsum[g * H + k];
    dW[t] = accumulate ? dW[t] + v : v;
  }
}

// ---- train mode ------------------------------------------------------------------------------
// The [B,N,N*H] tensor of the reference (config/ODA.py:222) is indexed by (b, i, e) with e = j*H + k:
//   z[b,i,g] = bc[g] + scale * sum_e keep(b,i,e) W[g,e] ql[b,k] (vl[b,i,k] - vl[b,j,k]).
// Every kernel below is bound by the instruction issue rate (4 FMAs per element for the G = 4 glimpses, nothing to
// stream: one sample is N*H floats), so the design minimises instructions per (b,i,j,k) element:
//  * the keep flags come from the step's 1-bit-per-element cache (vqa_dropout_bits layout: bit n of the stream =
//    element n), made by ONE Philox pass per step — the forward and the backward both read it instead of drawing
//    the 16-byte Philox groups again, which were two thirds of the old kernels' instructions;
//  * "e-mapping": a thread owns a chunk of CW consecutive k of one region j — W[g,e], vl[b,j,k] stay in registers —
//    and walks over the regions i, whose rows vl[b,i,:] are staged once per CTA in shared memory (one 128-bit
//    broadcast read per 4 elements);
//  * a CTA owns ALL regions j of a k-range, so the sum over j that dvl[b,i,k] needs completes inside the CTA through a
//    double-buffered shared-memory tile (one barrier per row) — the old second backward kernel, which regenerated
//    every mask and re-read W from shared memory four times per element, is gone.
constexpr int PAIR_MAX_THREADS = 512;
constexpr int PF_CW = 16;          // forward: elements per thread and row
constexpr int PF_RB = 8;           // forward: rows per cross-lane reduction batch (PF_RB * G = 32 values = one per lane)
constexpr int PB_CW = 8;           // backward: elements per thread and row (twice the per-element state of the forward)

struct PairGeom { int CC, P, threads; };
// chunk ranges: CC chunks of CW features per CTA, P CTAs per sample, CC * N threads (rounded up to a warp)
static PairGeom pair_geom(int64_t N, int64_t H, int CW) {
  const int NC = (int)cdiv(H, CW);
  int CC = (int)(PAIR_MAX_THREADS / N);
  if (CC < 1) CC = 1;
  if (CC > NC) CC = NC;
  const int P = (NC + CC - 1) / CC;
  CC = (NC + P - 1) / P;
  return PairGeom{CC, P, (int)(((int64_t)CC * N + 31) / 32 * 32)};
}

// ---- keep-bit windows through a thread-private cp.async ring -----------------------------------------
// A thread needs, for every region row i, the <= 32 keep bits that start at bit n0(i) = bit0 + i*N*H of the stream: two
// consecutive 32-bit words (the cache keeps one spare word past its end).  The words are fetched RD - 1 rows ahead
// with cp.async into a shared-memory ring slot that only this thread reads (no barrier, no registers held while the
// load is in flight) — waiting on the global load at its use was 40 % of the first version's stall samples.
__device__ __forceinline__ void cp_async4(uint32_t* smem_dst, const uint32_t* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }

template <int RD>
struct KeepRing {
  uint2* slot;                    // this thread's slot of ring row 0; rows are `stride` uint2 apart
  int stride;
  const uint32_t* words;
  uint64_t n_issue, n_use;        // bit position of the next row to fetch / to hand out
  uint32_t row_bits_lo;           // row length (bits) mod 2^32: enough for the 5-bit shift
  uint64_t row_bits;
  int issued, used;
  __device__ __forceinline__ KeepRing(uint2* slot_, int stride_, const uint32_t* words_, uint64_t bit0, uint64_t row_bits_)
      : slot(slot_), stride(stride_), words(words_), n_issue(bit0), n_use(bit0), row_bits_lo((uint32_t)row_bits_),
        row_bits(row_bits_), issued(0), used(0) {}
  __device__ __forceinline__ void issue() {
    uint32_t* d = reinterpret_cast<uint32_t*>(slot + (issued & (RD - 1)) * stride);
    cp_async4(d, words + (n_issue >> 5));
    cp_async4(d + 1, words + (n_issue >> 5) + 1);
    n_issue += row_bits;
    ++issued;
  }
  // call once per row, in order: prefetches row i + RD - 1 and returns the window of row i
  __device__ __forceinline__ uint32_t next(int nrows) {
    if (issued < nrows) issue();
    cp_async_commit();
    cp_async_wait<RD - 1>();
    const uint2 wv = slot[(used & (RD - 1)) * stride];
    const uint32_t kb = __funnelshift_r(wv.x, wv.y, (uint32_t)n_use & 31u);
    n_use += row_bits;
    ++used;
    return kb;
  }
  __device__ __forceinline__ void prime(int nrows) {
#pragma unroll
    for (int r = 0; r < RD - 1; ++r) {
      if (r < nrows) issue();
      cp_async_commit();
    }
  }
};

// v[0] of lane l <- sum over the 32 lanes of v[l] (butterfly that halves the live values at every step: 31 shuffles)
__device__ __forceinline__ void warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int x = 0; x < s; ++x) {
      const float send = up ? v[x] : v[x + s];
      const float keep = up ? v[x + s] : v[x];
      v[x] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// vl[b, :, kbase .. kbase+KT) -> Vs [N][KT] (zero past H), several loads in flight per thread
__device__ __forceinline__ void stage_rows(float* Vs, const float* __restrict__ vb, int N, int H, int KT, int kbase) {
  const int total = N * KT;
  for (int x0 = threadIdx.x; x0 < total; x0 += 4 * blockDim.x) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = x0 + u * blockDim.x;
      const int i = x / KT, k = kbase + (x - i * KT);
      v[u] = (x < total && k < H) ? __ldg(vb + (int64_t)i * H + k) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = x0 + u * blockDim.x;
      if (x < total) Vs[x] = v[u];
    }
  }
}

// Forward.  grid = (P, B), threads = CC * N rounded up; thread t = (chunk cl = t / N, region j = t % N).
// zpart[b][p][i][g] = this CTA's share of the logit (k-range p); oda_softmax_partials_kernel adds the shares in a
// fixed order (deterministic) with the bias and normalises.
// dynamic smem: Vs [N][KT] (KT = CC * PF_CW) | zred [warps][cdiv(N, 8)][32] | keep ring [PF_RD][threads] uint2
constexpr int PF_RD = 8;
__global__ void __launch_bounds__(PAIR_MAX_THREADS)
oda_pair_fwd_train_kernel(int N, int H, int CC, float scale, const uint32_t* __restrict__ bits,
                          const float* __restrict__ vl, const float* __restrict__ ql, const float* __restrict__ W,
                          float* __restrict__ zpart) {
  extern __shared__ __align__(16) float pair_smem[];
  const int KT = CC * PF_CW;
  const int NB = (N + PF_RB - 1) / PF_RB;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
  float* Vs = pair_smem;
  float* zred = pair_smem + (size_t)N * KT;
  uint2* ring = reinterpret_cast<uint2*>(zred + (size_t)nwarps * NB * 32);
  const int b = blockIdx.y, P = gridDim.x;
  const int64_t NH = (int64_t)N * H;
  const int kbase = blockIdx.x * KT;
  const float* vb = vl + (int64_t)b * NH;
  const int cl = t / N, j = t - cl * N;
  const int k0 = kbase + cl * PF_CW;
  const bool active = cl < CC && k0 < H;
  // idle threads read (and ignore: their weights are zero) the bits and rows of chunk 0 / region 0
  KeepRing<PF_RD> keep(ring + t, (int)blockDim.x, bits, (uint64_t)b * N * NH + (active ? (uint64_t)j * H + k0 : 0),
                       (uint64_t)NH);
  keep.prime(N);
  stage_rows(Vs, vb, N, H, KT, kbase);
  float wq[G][PF_CW], vj[PF_CW];
#pragma unroll
  for (int e = 0; e < PF_CW; ++e) {
    const int k = k0 + e;
    const bool in = active && k < H;
    const float q = in ? __ldg(ql + (int64_t)b * H + k) * scale : 0.0f;
    vj[e] = in ? __ldg(vb + (int64_t)j * H + k) : 0.0f;
#pragma unroll
    for (int g = 0; g < G; ++g) wq[g][e] = in ? __ldg(W + g * NH + (int64_t)j * H + k) * q : 0.0f;
  }
  __syncthreads();
  const float* vrow = Vs + (active ? cl * PF_CW : 0);
  for (int ib = 0; ib < NB; ++ib) {
    float za[PF_RB * G];
#pragma unroll
    for (int x = 0; x < PF_RB * G; ++x) za[x] = 0.0f;
#pragma unroll
    for (int r = 0; r < PF_RB; ++r) {
      const int i = ib * PF_RB + r;
      if (i < N) {
        const uint32_t kb = keep.next(N);
        const float4* vi = reinterpret_cast<const float4*>(vrow + i * KT);
#pragma unroll
        for (int q4 = 0; q4 < PF_CW / 4; ++q4) {
          const float4 x4 = vi[q4];
          const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = q4 * 4 + u;
            const float d = (kb & (1u << e)) ? xv[u] - vj[e] : 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) za[r * G + g] = fmaf(wq[g][e], d, za[r * G + g]);
          }
        }
      }
    }
    warp_transpose_sum32(za, lane);
    zred[(warp * NB + ib) * 32 + lane] = za[0];
  }
  __syncthreads();
  for (int x = t; x < N * G; x += blockDim.x) {       // x = i*G + g = (ib, lane) = (x / 32, x % 32)
    float s = 0.0f;
    for (int wv = 0; wv < nwarps; ++wv) s += zred[(wv * NB + (x >> 5)) * 32 + (x & 31)];
    zpart[((int64_t)b * P + blockIdx.x) * N * G + x] = s;
  }
}

// alpha[b,:,g] = softmax_i(bc[g] + sum_p zpart[b][p][i][g]).  grid = B, dynamic smem N*G floats
__global__ void oda_softmax_partials_kernel(int N, int P, const float* __restrict__ zpart, const float* __restrict__ bc,
                                            float* __restrict__ alpha) {
  extern __shared__ float z_s[];
  const int64_t b = blockIdx.x;
  for (int t = threadIdx.x; t < N * G; t += blockDim.x) {
    float s = bc[t % G];
    for (int p = 0; p < P; ++p) s += zpart[(b * P + p) * N * G + t];
    z_s[t] = s;
  }
  __syncthreads();
  softmax_regions_smem(z_s, N);
  __syncthreads();
  for (int t = threadIdx.x; t < N * G; t += blockDim.x) alpha[b * N * G + t] = z_s[t];
}

// dz = alpha (.) (dalpha - <alpha, dalpha>), dbc += sum dz.  grid = B
__global__ void softmax_regions_bwd_kernel(int64_t N, const float* __restrict__ alpha, const float* __restrict__ dalpha,
                                           float* __restrict__ dz, float* __restrict__ dbc) {
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= G) return;
  float s = 0.0f;
  for (int64_t i = lane; i < N; i += 32) s = fmaf(alpha[(b * N + i) * G + warp], dalpha[(b * N + i) * G + warp], s);
  s = warp_sum(s);
  float tot = 0.0f;
  for (int64_t i = lane; i < N; i += 32) {
    const int64_t o = (b * N + i) * G + warp;
    const float v = alpha[o] * (dalpha[o] - s);
    dz[o] = v;
    tot += v;
  }
  tot = warp_sum(tot);
  if (lane == 0 && dbc) atomicAdd(&dbc[warp], tot);
}

// Backward.  Same mapping as the forward with PB_CW = 8 elements per thread.  With keep = the element's keep flag,
// delta = vl[b,i,k] - vl[b,j,k], D[g,e] = sum_i dz[b,i,g] keep delta and u(i,e) = keep sum_g dz[b,i,g] W[g,e]:
//   dW[g,e]    += scale ql[b,k] D[g,e]                                   (atomic: summed over the samples)
//   dql[b,k]    = scale sum_{j,g} W[g,e] D[g,e]                          (column sum over j inside the CTA)
//   dvl[b,x,k]  = scale ql[b,k] (sum_j u(x,(j,k)) - sum_i u(i,(x,k)))    ("plus": sum over j per row x through the
//                 shared tile; "minus": the thread's own running sum over i)
// The tile is stored TRANSPOSED, tile[column k][region j] with row stride TSC = 16 mod 32 words: the writers (lanes =
// consecutive j) store conflict-free words, the column sums read float4 runs of j (4 lanes per column, two shuffles).
// grid = (P, B).  dynamic smem: Vs [N][KT] | dvs [N][KT] | tile [2][KT][TSC] | dzs [N][G] | keep ring [PB_RD][threads]
constexpr int PB_RD = 4;
// tile row stride: >= roundup(N, 16) (so that TSC/16 float4 reads per lane cover a column) and = 16 mod 32
__host__ __device__ inline int pair_tsc(int N) { return ((N + 3) / 4 * 4 + 15) / 32 * 32 + 16; }

template <int TSC>
__global__ void __launch_bounds__(PAIR_MAX_THREADS)
oda_pair_bwd_train_kernel(int N, int H, int CC, float scale, const uint32_t* __restrict__ bits,
                          const float* __restrict__ vl, const float* __restrict__ ql, const float* __restrict__ W,
                          const float* __restrict__ dz, float* __restrict__ dW, float* __restrict__ dvl,
                          float* __restrict__ dql) {
  extern __shared__ __align__(16) float pair_smem[];
  const int KT = CC * PB_CW;
  const int KTT = (blockDim.x + N - 1) / N * PB_CW;     // tile columns: every thread has a slot (idle chunks write zeros)
  float* Vs = pair_smem;
  float* dvs = Vs + (size_t)N * KT;
  float* tile = dvs + (size_t)N * KT;
  float* dzs = tile + (size_t)2 * KTT * TSC;
  uint2* ring = reinterpret_cast<uint2*>(dzs + (size_t)N * G);
  const int b = blockIdx.y;
  const int64_t NH = (int64_t)N * H;
  const int kbase = blockIdx.x * KT;
  const int t = threadIdx.x;
  const float* vb = vl + (int64_t)b * NH;
  const int cl = t / N, j = t - cl * N;
  const int k0 = kbase + cl * PB_CW;
  const bool owner = cl < CC;
  const bool active = owner && k0 < H;
  KeepRing<PB_RD> keep(ring + t, (int)blockDim.x, bits, (uint64_t)b * N * NH + (active ? (uint64_t)j * H + k0 : 0),
                       (uint64_t)NH);
  keep.prime(N);
  stage_rows(Vs, vb, N, H, KT, kbase);
  for (int x = t; x < N * G; x += blockDim.x) dzs[x] = __ldg(dz + (int64_t)b * N * G + x);
  for (int x = t; x < 2 * KTT * TSC; x += blockDim.x) tile[x] = 0.0f;      // the pad columns j >= N stay zero
  float w[G][PB_CW], D[G][PB_CW], vj[PB_CW], minus[PB_CW];
#pragma unroll
  for (int e = 0; e < PB_CW; ++e) {
    const int k = k0 + e;
    const bool in = active && k < H;
    vj[e] = in ? __ldg(vb + (int64_t)j * H + k) : 0.0f;
    minus[e] = 0.0f;
#pragma unroll
    for (int g = 0; g < G; ++g) { w[g][e] = in ? __ldg(W + g * NH + (int64_t)j * H + k) : 0.0f; D[g][e] = 0.0f; }
  }
  __syncthreads();
  const float* vrow = Vs + (active ? cl * PB_CW : 0);
  float* tcol = tile + (size_t)cl * PB_CW * TSC + j;
  const int toggle = KTT * TSC;
  // column-sum role: 4 consecutive lanes share a column; lane sl takes the float4 runs sl, sl+4, ... (TSC/16 of them:
  // the runs past N are zero padding)
  const int col = t >> 2, sl = t & 3;
  const int cstep = blockDim.x >> 2, colend = (KT + cstep - 1) / cstep * cstep;
  for (int i = 0; i < N; ++i) {
    const uint32_t kb = keep.next(N);
    const float4 z4 = *reinterpret_cast<const float4*>(dzs + i * G);
    const float zg[G] = {z4.x, z4.y, z4.z, z4.w};
    const float4* vi = reinterpret_cast<const float4*>(vrow + i * KT);
    const int boff = (i & 1) * toggle;
    float um[PB_CW];
#pragma unroll
    for (int q4 = 0; q4 < PB_CW / 4; ++q4) {
      const float4 x4 = vi[q4];
      const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int e = q4 * 4 + u4;
        const bool kp = (kb & (1u << e)) != 0;
        const float dm = kp ? xv[u4] - vj[e] : 0.0f;
        float u = zg[0] * w[0][e];
#pragma unroll
        for (int g = 1; g < G; ++g) u = fmaf(zg[g], w[g][e], u);
#pragma unroll
        for (int g = 0; g < G; ++g) D[g][e] = fmaf(zg[g], dm, D[g][e]);
        um[e] = kp ? u : 0.0f;
        minus[e] += um[e];
      }
    }
#pragma unroll
    for (int e = 0; e < PB_CW; ++e) tcol[boff + e * TSC] = um[e];
    __syncthreads();
    // plus term of row i: sums over j of the tile columns.  The other buffer is rewritten only after the next barrier,
    // which every thread reaches after it has finished these reads.
    for (int c = col; c < colend; c += cstep) {       // colend is a multiple of cstep: uniform trip count
      float s = 0.0f;
      if (c < KT) {
        const float4* src = reinterpret_cast<const float4*>(tile + boff + (size_t)c * TSC) + sl;
#pragma unroll
        for (int q = 0; q < TSC / 16; ++q) {
          const float4 a = src[q * 4];
          s += (a.x + a.y) + (a.z + a.w);
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (sl == 0 && c < KT) dvs[i * KT + c] = s;
    }
  }
  __syncthreads();
  // per-sample tail: dW (atomics over the samples), the dql shares into tile buffer 0, minus into dvs
  {
#pragma unroll
    for (int e = 0; e < PB_CW; ++e) {
      const int k = k0 + e;
      float qp = 0.0f;
      if (active && k < H) {
        const float qk = __ldg(ql + (int64_t)b * H + k) * scale;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          atomicAdd(dW + g * NH + (int64_t)j * H + k, qk * D[g][e]);
          qp = fmaf(w[g][e], D[g][e], qp);
        }
      }
      tcol[e * TSC] = qp * scale;
      if (owner) dvs[j * KT + cl * PB_CW + e] -= minus[e];
    }
  }
  __syncthreads();
  for (int c = t; c < KT; c += blockDim.x) {
    const int k = kbase + c;
    if (k < H) {
      float s = 0.0f;
      for (int jj = 0; jj < N; ++jj) s += tile[(size_t)c * TSC + jj];
      dql[(int64_t)b * H + k] = s;
    }
  }
  for (int x = t; x < N * KT; x += blockDim.x) {
    const int i = x / KT, k = kbase + (x - i * KT);
    if (k < H) dvl[((int64_t)b * N + i) * H + k] = scale * __ldg(ql + (int64_t)b * H + k) * dvs[x];
  }
}

}  // namespace vqa

using namespace vqa;

static int oda_check(int64_t B, int64_t N, int64_t H, int64_t D, const char* who) {
  VQA_REQUIRE(B >= 0 && N >= 1 && H >= 1 && D >= 4 && D % 4 == 0, "%s: bad shape B=%lld N=%lld H=%lld D=%lld", who,
              (long long)B, (long long)N, (long long)H, (long long)D);
  return VQA_OK;
}

// train-mode scratch: [keep bits of the dropped [B,N,N*H] tensor | forward's partial logits [B][P][N][G]]
static size_t pair_bits_bytes(int64_t B, int64_t N, int64_t H) {
  return (((size_t)cdiv(B * N * N * H, 16) * 2 + 8) + 255) & ~(size_t)255;
}
extern "C" size_t vqa_oda_pair_attn_workspace_bytes(int64_t B, int64_t N, int64_t H) {
  if (B < 0 || N < 1 || H < 1) return 0;
  const PairGeom gf = pair_geom(N, H, PF_CW);
  return pair_bits_bytes(B, N, H) + (size_t)B * gf.P * N * G * sizeof(float);
}

template <class K>
static int allow_smem(K kern, size_t smem, const char* who) {
  if (smem > 227 * 1024) { set_error("%s: needs %zu bytes of shared memory (N too large)", who, smem); return VQA_EINVAL; }
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cannot reserve %zu bytes of shared memory", who, smem);
    return VQA_ECUDA;
  }
  return VQA_OK;
}

extern "C" int vqa_oda_pair_attn_fwd(const vqa_oda_pair_attn_fwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_oda_pair_attn_fwd: null params");
  VQA_TRY(oda_check(p->B, p->N, p->H, p->D, "vqa_oda_pair_attn_fwd"));
  VQA_REQUIRE(p->vl && p->ql && p->W && p->bc && p->x && p->alpha && p->pooled, "vqa_oda_pair_attn_fwd: null pointer");
  if (p->B == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool train = p->train && p->drop.p > 0.0f;
  if (!train) {
    VQA_REQUIRE(p->wsum != nullptr, "vqa_oda_pair_attn_fwd: wsum workspace required in eval mode");
    oda_wsum_kernel<<<G, 256, 0, st>>>(p->N, p->H, p->W, p->wsum);
    VQA_TRY(check_launch("oda_wsum"));
    FuseOdaEval fs{p->vl, p->ql, p->N, p->H};
    const size_t smem = (size_t)(G * p->H + p->N * G) * sizeof(float);
    att_logits_softmax_kernel<FuseOdaEval><<<dim3((unsigned)p->B, 1), ATT_THREADS, smem, st>>>(fs, p->N, p->H, p->wsum,
                                                                                               p->bc, p->alpha);
    VQA_TRY(check_launch("oda_logits_eval"));
  } else {
    VQA_REQUIRE(p->N <= 144, "vqa_oda_pair_attn_fwd: train mode supports up to 144 regions, got N=%lld", (long long)p->N);
    const size_t need = vqa_oda_pair_attn_workspace_bytes(p->B, p->N, p->H);
    if (!p->workspace || p->workspace_bytes < need) {
      set_error("vqa_oda_pair_attn_fwd: train mode needs a workspace of %zu bytes (vqa_oda_pair_attn_workspace_bytes), got %zu",
                need, p->workspace ? p->workspace_bytes : (size_t)0);
      return VQA_EWORKSPACE;
    }
    const int64_t total = p->B * p->N * p->N * p->H;
    uint8_t* bits = reinterpret_cast<uint8_t*>(p->workspace);
    float* zpart = reinterpret_cast<float*>(bits + pair_bits_bytes(p->B, p->N, p->H));
    if (!p->keep_bits_ready)      // a whole-model plan makes them in its batched keep-bit launch instead
      VQA_TRY(vqa_dropout_bits(p->drop.p, p->drop.seed, p->drop.seed_dev, p->drop.layer, (uint64_t)total, bits, stream));
    const PairGeom gm = pair_geom(p->N, p->H, PF_CW);
    const size_t smem = ((size_t)p->N * gm.CC * PF_CW + (size_t)(gm.threads / 32) * cdiv(p->N, PF_RB) * 32) * sizeof(float) +
                        (size_t)PF_RD * gm.threads * sizeof(uint2);
    VQA_TRY(allow_smem(oda_pair_fwd_train_kernel, smem, "vqa_oda_pair_attn_fwd"));
    {
      KProf kp_(st, "oda_pair_fwd_train", "flop", (double)p->B * p->N * p->N * p->H * (2.0 * G + 2.0));
      oda_pair_fwd_train_kernel<<<dim3((unsigned)gm.P, (unsigned)p->B), gm.threads, smem, st>>>(
          (int)p->N, (int)p->H, gm.CC, 1.0f / (1.0f - p->drop.p), reinterpret_cast<const uint32_t*>(bits), p->vl, p->ql,
          p->W, zpart);
      VQA_TRY(check_launch("oda_pair_fwd_train"));
    }
    oda_softmax_partials_kernel<<<(unsigned)p->B, 128, (size_t)p->N * G * sizeof(float), st>>>((int)p->N, gm.P, zpart,
                                                                                             p->bc, p->alpha);
    VQA_TRY(check_launch("oda_softmax_partials"));
  }
  return launch_pool_fwd(p->B, p->N, p->D, p->x, p->alpha, p->pooled, st);
}

extern "C" int vqa_oda_pair_attn_bwd(const vqa_oda_pair_attn_bwd_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_oda_pair_attn_bwd: null params");
  VQA_TRY(oda_check(p->B, p->N, p->H, p->D, "vqa_oda_pair_attn_bwd"));
  VQA_REQUIRE(p->vl && p->ql && p->W && p->x && p->alpha && p->dpooled && p->dalpha && p->dz && p->dW && p->dvl &&
                  p->dql,
              "vqa_oda_pair_attn_bwd: null pointer");
  if (p->B == 0) return VQA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t NH = p->N * p->H;
  VQA_TRY(launch_pool_bwd(p->B, p->N, p->D, p->x, p->alpha, p->dpooled, nullptr, p->dalpha, nullptr, 0, st));
  if (!p->accumulate_w && p->dbc) cudaMemsetAsync(p->dbc, 0, G * sizeof(float), st);
  const bool train = p->train && p->drop.p > 0.0f;
  if (!train) {
    VQA_REQUIRE(p->wsum && p->dwsum, "vqa_oda_pair_attn_bwd: wsum/dwsum workspaces required in eval mode");
    cudaMemsetAsync(p->dwsum, 0, (size_t)G * p->H * sizeof(float), st);
    FuseOdaEval fs{p->vl, p->ql, p->N, p->H};
    const int64_t groups = p->B < 2 * (int64_t)sm_count() ? p->B : 2 * (int64_t)sm_count();
    dim3 grid((unsigned)groups, (unsigned)cdiv(p->H, ATT_THREADS));
    att_logits_softmax_bwd_kernel<FuseOdaEval, true><<<grid, ATT_THREADS, (size_t)2 * p->N * G * sizeof(float), st>>>(
        fs, p->B, p->N, p->H, p->wsum, p->alpha, p->dalpha, p->dz, p->dwsum, p->dbc, p->dvl, p->dql);
    VQA_TRY(check_launch("oda_logits_eval_bwd"));
    oda_dw_broadcast_kernel<<<(unsigned)cdiv(G * NH, 256), 256, 0, st>>>(p->N, p->H, p->dwsum, p->dW, p->accumulate_w);
    return check_launch("oda_dw_broadcast");
  }
  const size_t need = vqa_oda_pair_attn_workspace_bytes(p->B, p->N, p->H);
  if (!p->workspace || p->workspace_bytes < need) {
    set_error("vqa_oda_pair_attn_bwd: train mode needs the forward's workspace (%zu bytes), got %zu", need,
              p->workspace ? p->workspace_bytes : (size_t)0);
    return VQA_EWORKSPACE;
  }
  softmax_regions_bwd_kernel<<<(unsigned)p->B, 128, 0, st>>>(p->N, p->alpha, p->dalpha, p->dz, p->dbc);
  VQA_TRY(check_launch("softmax_regions_bwd"));
  if (!p->accumulate_w) cudaMemsetAsync(p->dW, 0, (size_t)G * NH * sizeof(float), st);
  const PairGeom gm = pair_geom(p->N, p->H, PB_CW);
  const int KT = gm.CC * PB_CW, TSC = pair_tsc((int)p->N);
  const int KTT = (int)cdiv(gm.threads, p->N) * PB_CW;
  const size_t smem = ((size_t)2 * p->N * KT + (size_t)2 * KTT * TSC + (size_t)p->N * G) * sizeof(float) +
                      (size_t)PB_RD * gm.threads * sizeof(uint2);
  auto kern = TSC == 16 ? oda_pair_bwd_train_kernel<16> : TSC == 48 ? oda_pair_bwd_train_kernel<48>
            : TSC == 80 ? oda_pair_bwd_train_kernel<80> : TSC == 112 ? oda_pair_bwd_train_kernel<112>
            : TSC == 144 ? oda_pair_bwd_train_kernel<144> : nullptr;
  VQA_REQUIRE(kern != nullptr, "vqa_oda_pair_attn_bwd: train mode supports up to 144 regions, got N=%lld", (long long)p->N);
  VQA_TRY(allow_smem(kern, smem, "vqa_oda_pair_attn_bwd"));
  KProf kp_(st, "oda_pair_bwd_train", "flop", (double)p->B * p->N * p->N * p->H * (4.0 * G + 4.0));
  kern<<<dim3((unsigned)gm.P, (unsigned)p->B), gm.threads, smem, st>>>(
      (int)p->N, (int)p->H, gm.CC, 1.0f / (1.0f - p->drop.p), reinterpret_cast<const uint32_t*>(p->workspace), p->vl,
      p->ql, p->W, p->dz, p->dW, p->dvl, p->dql);
  return check_launch("oda_pair_bwd_train");
}
