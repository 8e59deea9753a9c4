// Small dense layer  Y[n, o] = sum_k X[n, k] W[o, k]  on the fp32 FMA pipe, shared by the GAT linear
// stage (gat.cu) and the modality projections (project_fuse.cu).  These layers are a few hundred MMAC
// on [N x 256] activations: launch- and latency-bound, not tensor-bound, so the design goal is one
// wave of CTAs with the global loads of chunk c+1 in flight while chunk c is multiplied.
//
// CTA tile: 32 rows x 128 output columns, K chunks of 16, 256 threads, 2 x 8 micro-tile per thread
// (rows ty*2+i, columns col_of(tx, j): two groups of 4 consecutive columns 64 apart, so that the 16
// threads of a row read 256 contiguous bytes per LDS.128 -- conflict-free), operands transposed in
// shared memory so that a thread reads its 2 activations with one LDS.64 and its 8 weights with two
// LDS.128; register-staged double buffering.
#pragma once
#include "common.cuh"

namespace sga {
namespace linear {

constexpr int BM = 32, BN = 128, BK = 16, NT = 256;
constexpr int XS_LD = BM + 2;    // floats per k row of the transposed X tile (even: LDS.64 alignment)
constexpr int WS_LD = BN + 4;    // multiple of 4: LDS.128 alignment

__device__ __forceinline__ int col_of(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

struct Smem {
  float xs[2][BK][XS_LD];
  float ws[2][BK][WS_LD];
};

// acc[i][j] = sum_k X[n0 + ty*2 + i, k] * W[o0 + col_of(tx, j), k];  rows >= N / columns >= O read as 0.
// X is fp32 or fp64 (x_is_f64; the reference's `.float()` on the dataloader tensors), row stride in_dim.
__device__ __forceinline__ void tile_mma(Smem& sm, const void* __restrict__ X, int x_is_f64, int64_t N, int in_dim,
                                         const float* __restrict__ W, int O, int64_t n0, int o0, float (&acc)[2][8]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // loader mapping: X tile 32 x 16 = 512 elements -> 2 per thread (row lr, k pair lk..lk+1);
  //                 W tile 128 x 16 = 2048 elements -> 8 per thread (row wr, k wk..wk+7)
  const int lr = tid >> 3, lk = (tid & 7) * 2;
  const int wr = tid >> 1, wk = (tid & 1) * 8;
  const bool x_vec = !x_is_f64 && (in_dim % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 7) == 0);
  const bool w_vec = (in_dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int64_t xrow = n0 + lr;
  const int wrow = o0 + wr;
  float xr[2], wv[8];
  auto fetch = [&](int k0) {
    xr[0] = xr[1] = 0.f;
    if (xrow < N) {
      const int k = k0 + lk;
      if (x_vec && k + 1 < in_dim) {
        const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(X) + xrow * in_dim + k);
        xr[0] = v.x; xr[1] = v.y;
      } else {
        if (k < in_dim) xr[0] = load_as_float<float>(X, xrow * in_dim + k, x_is_f64);
        if (k + 1 < in_dim) xr[1] = load_as_float<float>(X, xrow * in_dim + k + 1, x_is_f64);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) wv[e] = 0.f;
    if (wrow < O) {
      const int k = k0 + wk;
      const float* p = W + (int64_t)wrow * in_dim + k;
      if (w_vec && k + 7 < in_dim) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (k + e < in_dim) wv[e] = p[e];
      }
    }
  };
  auto stash = [&](int buf) {
    sm.xs[buf][lk][lr] = xr[0];
    sm.xs[buf][lk + 1][lr] = xr[1];
#pragma unroll
    for (int e = 0; e < 8; ++e) sm.ws[buf][wk + e][wr] = wv[e];
  };
  const int nchunk = (in_dim + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) fetch((c + 1) * BK);          // global loads in flight during the FMAs below
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float2 a = *reinterpret_cast<const float2*>(&sm.xs[buf][k][ty * 2]);
      const float4 w0 = *reinterpret_cast<const float4*>(&sm.ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sm.ws[buf][k][64 + tx * 4]);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] = fmaf(a.x, w[j], acc[0][j]);
        acc[1][j] = fmaf(a.y, w[j], acc[1][j]);
      }
    }
    if (c + 1 < nchunk) {
      stash(buf ^ 1);          // the other buffer was last read in iteration c-1, fenced by the barrier below
      __syncthreads();
    }
  }
}

// sum over the 16 threads (tx) that share a row
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace linear
}  // namespace sga
