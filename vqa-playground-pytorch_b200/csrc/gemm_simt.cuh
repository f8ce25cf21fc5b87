// fp32 CUDA-core GEMM building block (VQA_MATH_FP32_SIMT): the exact-fp32 parity path and the
// on-device yardstick the tcgen05 kernels (gemm_tc.cu) are validated against.
//
//   acc[m,n] = sum_k A(m,k) * B(n,k)
//
// A and B are FUNCTORS evaluated while the tile is staged into shared memory, which is where
// the reference's elementwise neighbours of each matmul are fused: input dropout (Philox
// regenerated, never stored), act'(y) on the incoming gradient, the Mutan Hadamard factor.
// The epilogue is a functor too (bias, activation, mask, accumulate, transposed store).
//
// Loader concept:   static constexpr bool KC   — true: consecutive k are contiguous in memory
//                   __device__ void select(int z)            — group selection (blockIdx.z)
//                   __device__ float operator()(int64_t r, int64_t k) const   — in-bounds only
#pragma once
#include "common.cuh"

namespace vqa {

constexpr int SIMT_THREADS = 256;
constexpr int SIMT_BK = 16;
constexpr int SIMT_PAD = 4;

template <int BM, int BN, int TM, int TN, class AL, class BL>
__device__ __forceinline__ void simt_mainloop(float (&acc)[TM][TN], int64_t M, int64_t N, int64_t K, int64_t m0,
                                              int64_t n0, const AL& a, const BL& b, float* __restrict__ As,
                                              float* __restrict__ Bs) {
  constexpr int BK = SIMT_BK;
  constexpr int LDA = BM + SIMT_PAD, LDB = BN + SIMT_PAD;
  constexpr int EA = BM * BK / SIMT_THREADS, EB = BN * BK / SIMT_THREADS;
  static_assert((BM / TM) * (BN / TN) == SIMT_THREADS, "thread tiling");
  static_assert(EA >= 1 && EB >= 1, "tile too small");
  const int tid = threadIdx.x;
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);

  float ra[EA], rb[EB];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int e = 0; e < EA; ++e) {
      const int idx = tid + e * SIMT_THREADS;
      const int kk = AL::KC ? idx % BK : idx / BM;
      const int rr = AL::KC ? idx / BK : idx % BM;
      const int64_t m = m0 + rr, k = k0 + kk;
      ra[e] = (m < M && k < K) ? a(m, k) : 0.0f;
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int idx = tid + e * SIMT_THREADS;
      const int kk = BL::KC ? idx % BK : idx / BN;
      const int rr = BL::KC ? idx / BK : idx % BN;
      const int64_t n = n0 + rr, k = k0 + kk;
      rb[e] = (n < N && k < K) ? b(n, k) : 0.0f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int e = 0; e < EA; ++e) {
      const int idx = tid + e * SIMT_THREADS;
      const int kk = AL::KC ? idx % BK : idx / BM;
      const int rr = AL::KC ? idx / BK : idx % BM;
      As[kk * LDA + rr] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int idx = tid + e * SIMT_THREADS;
      const int kk = BL::KC ? idx % BK : idx / BN;
      const int rr = BL::KC ? idx / BK : idx % BN;
      Bs[kk * LDB + rr] = rb[e];
    }
  };

  fetch(0);
  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float fa[TM], fb[TN];
      if constexpr (TM % 4 == 0) {
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&As[kk * LDA + ty * TM + i]);
          fa[i] = t.x; fa[i + 1] = t.y; fa[i + 2] = t.z; fa[i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) fa[i] = As[kk * LDA + ty * TM + i];
      }
      if constexpr (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&Bs[kk * LDB + tx * TN + j]);
          fb[j] = t.x; fb[j + 1] = t.y; fb[j + 2] = t.z; fb[j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) fb[j] = Bs[kk * LDB + tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(fa[i], fb[j], acc[i][j]);
    }
    __syncthreads();
  }
}

template <int BM, int BN, int TM, int TN, class AL, class BL, class EP>
__global__ void __launch_bounds__(SIMT_THREADS) gemm_simt_kernel(int64_t M, int64_t N, int64_t K, AL a, BL b, EP ep) {
  __shared__ __align__(16) float As[SIMT_BK * (BM + SIMT_PAD)];
  __shared__ __align__(16) float Bs[SIMT_BK * (BN + SIMT_PAD)];
  a.select(blockIdx.z);
  b.select(blockIdx.z);
  ep.select(blockIdx.z);
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
  simt_mainloop<BM, BN, TM, TN>(acc, M, N, K, m0, n0, a, b, As, Bs);
  const int ty = threadIdx.x / (BN / TN), tx = threadIdx.x % (BN / TN);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n < N) ep(m, n, acc[i][j]);
    }
  }
}

// Tile choice: 128x64 for tall problems, 64x64 otherwise, 32x64 (more CTAs) when M is small.
template <class AL, class BL, class EP>
int launch_gemm_simt(int64_t M, int64_t N, int64_t K, int groups, const AL& a, const BL& b, const EP& ep,
                     cudaStream_t st, const char* what) {
  if (M <= 0 || N <= 0 || groups <= 0) return VQA_OK;
  const int64_t tiles64 = cdiv(M, 64) * cdiv(N, 64) * groups;
  if (M >= 2048 && tiles64 >= 4 * (int64_t)sm_count()) {
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 128), (unsigned)groups);
    gemm_simt_kernel<128, 64, 8, 4><<<grid, SIMT_THREADS, 0, st>>>(M, N, K, a, b, ep);
  } else if (tiles64 >= (int64_t)sm_count()) {
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64), (unsigned)groups);
    gemm_simt_kernel<64, 64, 4, 4><<<grid, SIMT_THREADS, 0, st>>>(M, N, K, a, b, ep);
  } else {
    dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(M, 32), (unsigned)groups);
    gemm_simt_kernel<32, 32, 4, 1><<<grid, SIMT_THREADS, 0, st>>>(M, N, K, a, b, ep);
  }
  return check_launch(what);
}

}  // namespace vqa
