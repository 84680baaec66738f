// CUDA-core GEMM  D[M,N] = A[M,K] * W[N,K]^T  with fp32 accumulation, for the small decision-critical
// fp32 layers of the path (dim_reduction, selection MLP: < 0.1 % of the FLOPs, kept in fp32 so that the
// `logit > -1` decisions see reference-grade arithmetic) and as the independent cross-check the unit tests
// use against the tcgen05 kernel.  Same epilogue functors as gemm_tc.cuh (apply<4>).
#pragma once
#include "common.cuh"

namespace rgrg {
namespace simt {

__device__ __forceinline__ float ld(const float* p) { return *p; }
__device__ __forceinline__ float ld(const bf16* p) { return bf2f(*p); }

template <class TA, class TW, class Epi>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TW* __restrict__ Wt, int M, int N,
                                                       int K, int lda, int ldw, const Epi epi) {
  constexpr int T = 64, KB = 16;
  __shared__ float sA[KB][T + 1];
  __shared__ float sW[KB][T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += KB) {
    for (int i = threadIdx.x; i < T * KB; i += 256) {
      const int r = i / KB, k = i % KB;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + k;
      sA[k][r] = (gm < M && gk < K) ? ld(A + static_cast<size_t>(gm) * lda + gk) : 0.0f;
      sW[k][r] = (gn < N && gk < K) ? ld(Wt + static_cast<size_t>(gn) * ldw + gk) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = sW[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int col0 = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row < M && col0 < N) {
      typename Epi::State st;
      epi.init(st, 0);
      epi.template apply<4>(st, row, col0, acc[i], N);
    }
  }
}

template <class TA, class TW, class Epi>
inline void launch(const TA* A, const TW* Wt, int M, int N, int K, const Epi& epi, cudaStream_t stream) {
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  gemm_simt_kernel<TA, TW, Epi><<<grid, 256, 0, stream>>>(A, Wt, M, N, K, K, K, epi);
  KERNEL_CHECK();
}

}  // namespace simt
}  // namespace rgrg
