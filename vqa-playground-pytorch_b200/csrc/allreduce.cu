// Gradient all-reduce over NVLink peer memory (include/vqacore.h: vqa_peer_allreduce_f32).
// Replaces the gradient exchange of the reference's nn.DataParallel (train.py:517): ONE sum over the
// ranks of the flat gradient buffer.  Every rank maps every peer's buffer (symmetric memory: the same allocation made
// on each GPU, exchanged once at start-up by the host side); a bucket [offset, offset+count) is reduced in place:
//   rank r owns the r-th slice of the bucket: it LOADS that slice from every peer over NVLink, adds the world_size
//   values in rank order and STORES the sum back into every peer's buffer (two-shot all-reduce fused into one pass:
//   reduce-scatter and all-gather of a slice happen in the same loop iteration).
// Every element is summed by exactly one rank in a fixed order and then broadcast, so all ranks end up with
// BIT-IDENTICAL gradients (NCCL's ring order depends on the rank), and parameters stay identical across ranks.
// The kernel uses no shared memory and few registers, so its CTAs fit on SMs whose shared memory is taken by the
// persistent tensor-core GEMMs of the backward: the exchange of a finished bucket runs UNDER the rest of the backward
// without taking SMs away from it (NCCL's CTAs do, which made "overlapped" NCCL slower than reduce-after-backward).
// Synchronisation: CTA c of every rank meets CTA c of the other ranks at an entry barrier (all ranks have finished
// producing the bucket) and at an exit barrier (all slices are written everywhere) through 32-bit epoch flags in a
// symmetric signal buffer, release/acquire at system scope.  No intra-grid dependency: CTAs need not be co-resident.
#include "common.cuh"

namespace vqa {

constexpr int AR_MAX_WORLD = VQA_AR_MAX_WORLD;
constexpr int AR_MAX_CTAS = VQA_AR_MAX_CTAS;
constexpr int AR_THREADS = 128;      // default: small CTAs of <= 48 registers, several fit beside a 320-thread GEMM CTA of 128 registers
constexpr int AR_THREADS_WIDE = 256; // cta_threads = 256: twice the loads in flight per CTA (faster alone, heavier beside the GEMMs)
// signal buffer layout (uint32 words): start[c][src], end[c][src], epoch[c], error flag
constexpr int AR_START = 0;
constexpr int AR_END = AR_MAX_CTAS * AR_MAX_WORLD;
constexpr int AR_EPOCH = 2 * AR_MAX_CTAS * AR_MAX_WORLD;
constexpr int AR_ERROR = AR_EPOCH + AR_MAX_CTAS;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data: read once, never reused -> do not keep it in L1
__device__ __forceinline__ float4 ld_peer4(const float4* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_peer4(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// NVLS: one instruction on the MULTICAST mapping of the symmetric buffer makes the NVSwitch fetch the 16 bytes from every
// GPU, add them and return the sum / replicate a store to every GPU — half the NVLink bytes of the peer loads+stores.
__device__ __forceinline__ float4 mc_ld_reduce4(const float4* p) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st4(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct ArArgs {
  float* mc;                         // multicast mapping of the symmetric buffer (NVLS) or nullptr
  float* buf[AR_MAX_WORLD];          // every rank's buffer, as mapped in this process
  uint32_t* sig[AR_MAX_WORLD];       // every rank's signal buffer
  int world, rank;
  int64_t offset, count;             // bucket, in floats (offset and count multiples of 4)
  long long spin_limit;              // clock64 ticks before a barrier gives up (sets the error flag instead of hanging)
};

// meets CTA blockIdx.x of every other rank; `slot` = AR_START or AR_END
__device__ __forceinline__ void ar_barrier(const ArArgs& a, int slot, uint32_t e) {
  const int t = threadIdx.x, c = blockIdx.x;
  if (t < a.world) {
    st_release_sys(a.sig[t] + slot + c * AR_MAX_WORLD + a.rank, e);
    const uint32_t* mine = a.sig[a.rank] + slot + c * AR_MAX_WORLD + t;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > a.spin_limit) { a.sig[a.rank][AR_ERROR] = 1u; break; }
    }
  }
  __syncthreads();
}

template <int W, int THREADS>       // W = world size as a compile-time constant (0: generic)
__global__ void __launch_bounds__(THREADS, THREADS == AR_THREADS ? 10 : 2) peer_allreduce_kernel(ArArgs a) {
  constexpr int AR_THREADS = THREADS;
  __shared__ uint32_t epoch_s;
  const int world = W > 0 ? W : a.world;
  if (threadIdx.x == 0) {
    uint32_t* ep = a.sig[a.rank] + AR_EPOCH + blockIdx.x;
    epoch_s = *ep + 1u;
    *ep = epoch_s;
  }
  __syncthreads();
  const uint32_t e = epoch_s;
  ar_barrier(a, AR_START, e);                       // every rank has produced the bucket
  // this rank's slice, in float4 units
  const int64_t n4 = a.count >> 2;
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = a.rank * per, hi = lo + per < n4 ? lo + per : n4;
  const int64_t base4 = a.offset >> 2;
  const int64_t stride = (int64_t)gridDim.x * AR_THREADS;
  constexpr int U = (W == 2 ? 4 : (W == 4 ? 2 : 1)) * (THREADS == 128 ? 1 : 2);  // independent elements per thread and iteration
  for (int64_t i0 = lo + (int64_t)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi; i0 += U * stride) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < (W > 0 ? W : AR_MAX_WORLD); ++p) {
      if (p >= world) break;
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        v[u] = i < hi ? ld_peer4(reinterpret_cast<const float4*>(a.buf[p]) + base4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
    }
#pragma unroll
    for (int p = 0; p < (W > 0 ? W : AR_MAX_WORLD); ++p) {
      if (p >= world) break;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < hi) st_peer4(reinterpret_cast<float4*>(a.buf[p]) + base4 + i, acc[u]);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  ar_barrier(a, AR_END, e);                         // every slice is written on every rank
}

// NVLS variant: the slice of this rank is reduced in the switch and broadcast by it.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == AR_THREADS ? 10 : 2) peer_allreduce_mc_kernel(ArArgs a) {
  constexpr int AR_THREADS = THREADS;
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) {
    uint32_t* ep = a.sig[a.rank] + AR_EPOCH + blockIdx.x;
    epoch_s = *ep + 1u;
    *ep = epoch_s;
  }
  __syncthreads();
  const uint32_t e = epoch_s;
  ar_barrier(a, AR_START, e);
  const int64_t n4 = a.count >> 2;
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t lo = a.rank * per, hi = lo + per < n4 ? lo + per : n4;
  float4* mc = reinterpret_cast<float4*>(a.mc) + (a.offset >> 2);
  const int64_t stride = (int64_t)gridDim.x * AR_THREADS;
  constexpr int U = THREADS == 128 ? 4 : 8;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi; i0 += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < hi) v[u] = mc_ld_reduce4(mc + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < hi) mc_st4(mc + i, v[u]);
    }
  }
  __threadfence_system();
  __syncthreads();
  ar_barrier(a, AR_END, e);
}

}  // namespace vqa

using namespace vqa;

extern "C" size_t vqa_peer_allreduce_signal_bytes(void) { return (size_t)(AR_ERROR + 1) * sizeof(uint32_t); }

extern "C" int vqa_peer_allreduce_f32(const vqa_peer_allreduce_params* p, void* stream) {
  VQA_REQUIRE(p != nullptr, "vqa_peer_allreduce_f32: null params");
  VQA_REQUIRE(p->world >= 1 && p->world <= AR_MAX_WORLD && p->rank >= 0 && p->rank < p->world,
              "vqa_peer_allreduce_f32: bad world / rank %d / %d", p->world, p->rank);
  VQA_REQUIRE(p->offset >= 0 && p->count >= 0 && p->offset % 4 == 0 && p->count % 4 == 0,
              "vqa_peer_allreduce_f32: offset and count must be multiples of 4 floats (16-byte vectors)");
  if (p->world == 1 || p->count == 0) return VQA_OK;
  ArArgs a = {};
  for (int r = 0; r < p->world; ++r) {
    VQA_REQUIRE(p->buffers[r] && p->signals[r], "vqa_peer_allreduce_f32: null buffer / signal pointer of rank %d", r);
    VQA_REQUIRE((reinterpret_cast<uintptr_t>(p->buffers[r]) & 15) == 0, "vqa_peer_allreduce_f32: buffers must be 16-byte aligned");
    a.buf[r] = reinterpret_cast<float*>(p->buffers[r]);
    a.sig[r] = reinterpret_cast<uint32_t*>(p->signals[r]);
  }
  a.world = p->world; a.rank = p->rank; a.offset = p->offset; a.count = p->count;
  a.mc = reinterpret_cast<float*>(p->multicast);
  VQA_REQUIRE((reinterpret_cast<uintptr_t>(p->multicast) & 15) == 0, "vqa_peer_allreduce_f32: multicast pointer must be 16-byte aligned");
  a.spin_limit = p->spin_limit_ms > 0 ? (long long)p->spin_limit_ms * 2000000ll : (1ll << 62);
  const bool wide = p->cta_threads == AR_THREADS_WIDE;
  VQA_REQUIRE(p->cta_threads == 0 || p->cta_threads == AR_THREADS || wide, "vqa_peer_allreduce_f32: cta_threads must be 0, 128 or 256");
  const int threads = wide ? AR_THREADS_WIDE : AR_THREADS;
  int ctas = p->max_ctas > 0 ? p->max_ctas : (wide ? 128 : 160);
  if (ctas > AR_MAX_CTAS) ctas = AR_MAX_CTAS;
  const int64_t slice4 = (p->count / 4 + p->world - 1) / p->world;
  const int64_t want = cdiv(slice4, threads * (wide ? 8 : 4));
  if (want < ctas) ctas = (int)(want < 1 ? 1 : want);
  // every rank must launch the same grid (CTA c meets CTA c): the grid depends on count, world and the options only
  KProf kp_(stream, "peer_allreduce", "hbm", 4.0 * (double)p->count * 2.0 * (p->world - 1) / p->world);
  cudaStream_t st = (cudaStream_t)stream;
  if (a.mc) {
    if (wide) peer_allreduce_mc_kernel<AR_THREADS_WIDE><<<ctas, threads, 0, st>>>(a);
    else peer_allreduce_mc_kernel<AR_THREADS><<<ctas, threads, 0, st>>>(a);
    return check_launch("peer_allreduce_mc");
  }
#define VQA_AR_LAUNCH(W)                                                                              \
  do {                                                                                                \
    if (wide) peer_allreduce_kernel<W, AR_THREADS_WIDE><<<ctas, threads, 0, st>>>(a);                 \
    else peer_allreduce_kernel<W, AR_THREADS><<<ctas, threads, 0, st>>>(a);                           \
  } while (0)
  switch (p->world) {
    case 2: VQA_AR_LAUNCH(2); break;
    case 4: VQA_AR_LAUNCH(4); break;
    case 8: VQA_AR_LAUNCH(8); break;
    default: VQA_AR_LAUNCH(0); break;
  }
#undef VQA_AR_LAUNCH
  return check_launch("peer_allreduce");
}
