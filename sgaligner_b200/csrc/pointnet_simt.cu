// PointNet feature encoder, fp32 FMA path (SGA_POINTNET_SIMT): the on-GPU cross-check for the
// tensor-core kernel and the path for shapes the tensor-core kernel does not take.
// Reference: src/aligner/networks/pointnet.py:140-163 (conv1/2/3 + ReLU, max over points; the
// BatchNorm outputs are discarded there, so they do not appear).
//
// One persistent CTA per SM walks objects; W2^T and a 256-channel block of W3^T stay in shared
// memory; a tile of 32 points flows conv1 -> conv2 -> conv3 through shared memory and the running
// (max, first-argmax) per channel is kept as a packed 64-bit key.
#include "common.cuh"

namespace sga {
namespace {

constexpr int TP = 32;    // points per tile
constexpr int NT = 256;   // threads per CTA
constexpr int CB = 256;   // conv3 output channels per CTA (grid.y walks channel blocks)
constexpr int H1LD = 65;
constexpr int H2LD = 129;

__device__ __forceinline__ unsigned long long pack_key(float v, int p) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)p);
}

// kMoments: instead of the pooled feature, accumulate per-channel sum / sum of squares of the three
// PRE-ReLU conv outputs over all points (the batch statistics the reference's discarded
// BatchNorm1d calls fold into running_mean / running_var in train mode).
template <bool kMoments>
__global__ void __launch_bounds__(NT, 1)
pointnet_fwd_simt_kernel(const float* __restrict__ pts, int64_t N, int P,
                         const float* __restrict__ W1, const float* __restrict__ b1,
                         const float* __restrict__ W2, const float* __restrict__ b2,
                         const float* __restrict__ W3, const float* __restrict__ b3, int C3,
                         float* __restrict__ out, int32_t* __restrict__ argmax, double* __restrict__ moments) {
  extern __shared__ __align__(16) float smem[];
  float* W2t = smem;                    // [64][128]
  float* W3t = W2t + 64 * 128;          // [128][CB]
  float* xs = W3t + 128 * CB;           // [TP][4]
  float* h1 = xs + TP * 4;              // [TP][H1LD]
  float* h2 = h1 + TP * H1LD;           // [TP][H2LD]
  unsigned long long* best = reinterpret_cast<unsigned long long*>(h2 + TP * H2LD + 2);  // [CB]

  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int cb0 = blockIdx.y * CB;
  const int ncb = min(CB, C3 - cb0);

  for (int i = tid; i < 64 * 128; i += NT) {
    int k = i >> 7, c = i & 127;
    W2t[i] = W2[c * 64 + k];
  }
  for (int i = tid; i < 128 * CB; i += NT) {
    int k = i / CB, c = i % CB;
    W3t[i] = (c < ncb) ? W3[(int64_t)(cb0 + c) * 128 + k] : 0.f;
  }
  float bias2[4], bias3[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias2[j] = b2[tx + 32 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias3[j] = (tx + 32 * j < ncb) ? b3[cb0 + tx + 32 * j] : 0.f;
  __syncthreads();
  double ms1 = 0, mq1 = 0, ms2[4] = {0, 0, 0, 0}, mq2[4] = {0, 0, 0, 0}, ms3[8], mq3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ms3[j] = mq3[j] = 0;

  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    for (int c = tid; c < CB; c += NT) best[c] = 0ull;
    const float* po = pts + n * (int64_t)P * 3;
    for (int t0 = 0; t0 < P; t0 += TP) {
      __syncthreads();
      if (tid < TP * 3) {
        int p = tid / 3, d = tid % 3;
        xs[p * 4 + d] = (t0 + p < P) ? po[(int64_t)(t0 + p) * 3 + d] : 0.f;
      }
      __syncthreads();
      // conv1 + ReLU
      {
        float ts = 0.f, tq = 0.f;
        for (int i = tid; i < TP * 64; i += NT) {
          int p = i >> 6, c = i & 63;
          float v = b1[c];
          v = fmaf(W1[c * 3 + 0], xs[p * 4 + 0], v);
          v = fmaf(W1[c * 3 + 1], xs[p * 4 + 1], v);
          v = fmaf(W1[c * 3 + 2], xs[p * 4 + 2], v);
          h1[p * H1LD + c] = v > 0.f ? v : 0.f;
          if (kMoments && blockIdx.y == 0 && t0 + p < P) { ts += v; tq = fmaf(v, v, tq); }
        }
        if (kMoments) { ms1 += ts; mq1 += tq; }
      }
      __syncthreads();
      // conv2 + ReLU : thread -> points ty*4+i, channels tx+32j
      {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = bias2[j];
        for (int k = 0; k < 64; ++k) {
          float a[4], w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = h1[(ty * 4 + i) * H1LD + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = W2t[k * 128 + tx + 32 * j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) h2[(ty * 4 + i) * H2LD + tx + 32 * j] = acc[i][j] > 0.f ? acc[i][j] : 0.f;
        if (kMoments && blockIdx.y == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (t0 + ty * 4 + i < P) { ts += acc[i][j]; tq = fmaf(acc[i][j], acc[i][j], tq); }
            ms2[j] += ts; mq2[j] += tq;
          }
        }
      }
      __syncthreads();
      // conv3 + ReLU + running max
      {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = bias3[j];
        for (int k = 0; k < 128; ++k) {
          float a[4], w[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = h2[(ty * 4 + i) * H2LD + k];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = W3t[k * CB + tx + 32 * j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kMoments) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (t0 + ty * 4 + i < P) { ts += acc[i][j]; tq = fmaf(acc[i][j], acc[i][j], tq); }
            ms3[j] += ts; mq3[j] += tq;
          }
        }
#pragma unroll
        for (int j = 0; j < 8 && !kMoments; ++j) {
          unsigned long long key = 0ull;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int p = t0 + ty * 4 + i;
            if (p < P) {
              float v = acc[i][j] > 0.f ? acc[i][j] : 0.f;
              unsigned long long kk = pack_key(v, p);
              key = kk > key ? kk : key;
            }
          }
          if (key) atomicMax(&best[tx + 32 * j], key);
        }
      }
    }
    __syncthreads();
    if (!kMoments) {
      for (int c = tid; c < ncb; c += NT) {
        unsigned long long key = best[c];
        out[n * C3 + cb0 + c] = __uint_as_float((unsigned)(key >> 32));
        if (argmax) argmax[n * C3 + cb0 + c] = (int32_t)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
      }
    }
    __syncthreads();
  }
  if (kMoments) {
    // moments = {sum1[64], sq1[64], sum2[128], sq2[128], sum3[C3], sq3[C3]}
    if (blockIdx.y == 0) {
      atomicAdd(&moments[tid & 63], ms1);
      atomicAdd(&moments[64 + (tid & 63)], mq1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&moments[128 + tx + 32 * j], ms2[j]);
        atomicAdd(&moments[256 + tx + 32 * j], mq2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = tx + 32 * j;
      if (c < ncb) {
        atomicAdd(&moments[384 + cb0 + c], ms3[j]);
        atomicAdd(&moments[384 + C3 + cb0 + c], mq3[j]);
      }
    }
  }
}

}  // namespace

int pointnet_fwd_simt(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                      const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                      float* out, int32_t* argmax, double* moments, cudaStream_t st) {
  size_t smem = (64 * 128 + 128 * CB + TP * 4 + TP * H1LD + TP * H2LD + 2) * sizeof(float) + CB * 8;
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int nblk_y = (C3 + CB - 1) / CB;
  int gx = (int)((N < (int64_t)sm_count()) ? N : sm_count());
  dim3 grid(gx, nblk_y);
  if (moments)
    pointnet_fwd_simt_kernel<true><<<grid, NT, smem, st>>>(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, moments);
  else
    pointnet_fwd_simt_kernel<false><<<grid, NT, smem, st>>>(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, nullptr);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
