// Generic strided fp32 FMA GEMM used by the loss and the backward kernels where the tensor-core
// path does not apply:  C[i*ldc + j] (+)= sum_k A(i,k) * B(k,j),
//   A(i,k) = A[i*sai + k*sak],  B(k,j) = B[k*sbk + j*sbj].
// 64x64x16 tiles, 256 threads, 4x4 micro-tile.
#pragma once
#include "common.cuh"

namespace sga {

namespace gemm_detail {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

static __global__ void __launch_bounds__(NT)
gemm_kernel(const float* __restrict__ A, int64_t sai, int64_t sak, const float* __restrict__ B, int64_t sbk,
            int64_t sbj, float* __restrict__ C, int64_t ldc, int M, int N, int K, int accumulate, int kchunk) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  // split-K: slice blockIdx.z covers [kbeg, kend) and adds its partial product atomically
  const int kbeg = blockIdx.z * kchunk;
  if (gridDim.z > 1) K = min(K, kbeg + kchunk);
  for (int k0 = kbeg; k0 < K; k0 += BK) {
    __syncthreads();
    for (int t = tid; t < BM * BK; t += NT) {
      int i, k;
      if (a_kfast) { i = t / BK; k = t % BK; } else { k = t / BM; i = t % BM; }
      As[k][i] = (i0 + i < M && k0 + k < K) ? A[(int64_t)(i0 + i) * sai + (int64_t)(k0 + k) * sak] : 0.f;
    }
    for (int t = tid; t < BN * BK; t += NT) {
      int j, k;
      if (b_kfast) { j = t / BK; k = t % BK; } else { k = t / BN; j = t % BN; }
      Bs[k][j] = (j0 + j < N && k0 + k < K) ? B[(int64_t)(k0 + k) * sbk + (int64_t)(j0 + j) * sbj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = i0 + ty + 16 * i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = j0 + tx + 16 * j;
      if (c < N) {
        float* p = C + (int64_t)r * ldc + c;
        if (gridDim.z > 1) atomicAdd(p, acc[i][j]);
        else *p = accumulate ? *p + acc[i][j] : acc[i][j];
      }
    }
  }
}
}  // namespace gemm_detail

// splitk > 1 requires accumulate semantics (C already holds the value to add to, e.g. zeros).
static inline cudaError_t launch_gemm(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj,
                               float* C, int64_t ldc, int M, int N, int K, int accumulate, cudaStream_t st,
                               int splitk = 1) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (!accumulate || splitk < 1) splitk = 1;
  int kchunk = ((K + splitk - 1) / splitk + gemm_detail::BK - 1) / gemm_detail::BK * gemm_detail::BK;
  if (kchunk < gemm_detail::BK) kchunk = gemm_detail::BK;
  splitk = (K + kchunk - 1) / kchunk;
  if (splitk < 1) splitk = 1;
  dim3 grid((N + gemm_detail::BN - 1) / gemm_detail::BN, (M + gemm_detail::BM - 1) / gemm_detail::BM, splitk);
  gemm_detail::gemm_kernel<<<grid, gemm_detail::NT, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, accumulate, kchunk);
  return cudaGetLastError();
}

// number of K slices that fills the machine for a small-output / long-K product
static inline int splitk_for(int M, int N, int K) {
  int tiles = ((M + gemm_detail::BM - 1) / gemm_detail::BM) * ((N + gemm_detail::BN - 1) / gemm_detail::BN);
  int want = (2 * sm_count() + tiles - 1) / tiles;
  int maxs = (K + 4 * gemm_detail::BK - 1) / (4 * gemm_detail::BK);
  if (want > maxs) want = maxs;
  return want < 1 ? 1 : want;
}

}  // namespace sga
