// NaivePCT: the small kernels around the tensor-core stages (reference: src/aligner/networks/pct.py).
//   * BatchNorm bookkeeping: batch statistics -> the folded per-channel affine pair (a, b) the NEXT kernel's prologue
//     applies (BN(z) = a z + b, a = gamma / sqrt(var + eps), b = beta - a mean), plus torch's train-mode side effect
//     on running_mean / running_var / num_batches_tracked (momentum update with the UNBIASED variance);
//   * the statistics of embedding.bn1 analytically from the first and second moments of the points (conv1 has no
//     bias and is linear in the point: sum z = w . sum p, sum z^2 = w^T (sum p p^T) w);
//   * the head (pct.py:311-316): per-column statistics over the objects and the fused BN + ReLU + dropout-mask pass
//     between the two small GEMMs.
#include "common.cuh"

namespace sga {
namespace pct {
namespace {

// sum p (3), sum p p^T (6: xx xy xz yy yz zz) over all NP points -> mom[9] (f64, zeroed by the caller)
__global__ void __launch_bounds__(256) point_moments_kernel(const float* __restrict__ pts, int64_t NP, double* __restrict__ mom) {
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < NP; i += (int64_t)gridDim.x * 256) {
    const float x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
    s[0] += x; s[1] += y; s[2] += z;
    s[3] += (double)x * x; s[4] += (double)x * y; s[5] += (double)x * z;
    s[6] += (double)y * y; s[7] += (double)y * z; s[8] += (double)z * z;
  }
  __shared__ double red[8][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const double v = warp_sum_d(s[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    atomicAdd(&mom[threadIdx.x], v);
  }
}

// per-channel {sum z, sum z^2} of z = W p from the point moments: stats[2C]
__global__ void affine_stats_kernel(const double* __restrict__ mom, const float* __restrict__ W, int C, double* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double wx = W[c * 3], wy = W[c * 3 + 1], wz = W[c * 3 + 2];
  stats[c] = wx * mom[0] + wy * mom[1] + wz * mom[2];
  stats[C + c] = wx * wx * mom[3] + wy * wy * mom[6] + wz * wz * mom[8] + 2.0 * (wx * wy * mom[4] + wx * wz * mom[5] + wy * wz * mom[7]);
}

// stats: {sum z [C], sum z^2 [C]} of the tensor the BatchNorm sees MINUS lin_bias (when the producing kernel stored
// z without its bias); training: batch statistics + running-stat update, else running statistics.  Writes the folded
// affine pair that applies to the STORED tensor: BN(z + bias) = a z + b.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double cnt, const float* __restrict__ lin_bias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, int64_t* __restrict__ nbt, int training, float momentum, float eps,
                                   int C, float* __restrict__ a_out, float* __restrict__ b_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  const double lb = lin_bias ? (double)lin_bias[c] : 0.0;
  double mean_z, var;
  if (training) {
    mean_z = stats[c] / cnt;
    var = stats[C + c] / cnt - mean_z * mean_z;               // biased, what normalises the batch
    if (var < 0) var = 0;
    const double mean_y = mean_z + lb;
    const double var_unb = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
    rmean[c] = rmean[c] * (1.f - momentum) + momentum * (float)mean_y;
    rvar[c] = rvar[c] * (1.f - momentum) + momentum * (float)var_unb;
  } else {
    mean_z = (double)rmean[c] - lb;
    var = (double)rvar[c];
  }
  const double a = (double)gamma[c] / sqrt(var + (double)eps);
  a_out[c] = (float)a;
  b_out[c] = (float)((double)beta[c] - a * mean_z);
}

// column sums / sums of squares of x [N, C] over the N rows -> stats[2C] (f64, zeroed by the caller)
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ x, int64_t N, int C, double* __restrict__ stats) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  double s = 0, q = 0;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + ry; r < N; r += (int64_t)gridDim.y * 8) {
      const double v = x[r * C + c];
      s += v;
      q += v * v;
    }
  __shared__ double rs[8][33], rq[8][33];
  rs[ry][threadIdx.x & 31] = s;
  rq[ry][threadIdx.x & 31] = q;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int k = 1; k < 8; ++k) {
      s += rs[k][threadIdx.x];
      q += rq[k][threadIdx.x];
    }
    atomicAdd(&stats[c], s);
    atomicAdd(&stats[C + c], q);
  }
}

// the same for long tensors ([N*P, 128] activations / gradients of the NaivePCT backward): float4 rows, fp32 partial sums
// over at most a few hundred rows per thread, fp64 across threads -- HBM bound instead of FP64-latency bound
__global__ void __launch_bounds__(256) col_stats_long_kernel(const float* __restrict__ x, int64_t N, int C, double* __restrict__ stats) {
  const int tpr = C >> 2, rpi = 256 / tpr;
  const int col4 = threadIdx.x % tpr, rslot = threadIdx.x / tpr;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
  int run = 0;
  for (int64_t r = (int64_t)blockIdx.x * rpi + rslot; r < N; r += (int64_t)gridDim.x * rpi) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C + col4 * 4));
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    if (++run == 64) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ds[i] += (double)s[i]; dq[i] += (double)q[i];
        s[i] = 0.f; q[i] = 0.f;
      }
      run = 0;
    }
  }
  __shared__ double red[256][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[threadIdx.x][i] = ds[i] + (double)s[i];
    red[threadIdx.x][4 + i] = dq[i] + (double)q[i];
  }
  __syncthreads();
  if (rslot == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double t1 = 0, t2 = 0;
      for (int k = 0; k < rpi; ++k) {
        t1 += red[k * tpr + col4][i];
        t2 += red[k * tpr + col4][4 + i];
      }
      atomicAdd(&stats[col4 * 4 + i], t1);
      atomicAdd(&stats[C + col4 * 4 + i], t2);
    }
  }
}

// out = relu(a_c x + b_c) * (mask ? mask * scale : 1)        (BatchNorm + ReLU + nn.Dropout, pct.py:312-316)
__global__ void bn_act_rows_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                   const float* __restrict__ mask, float scale, int64_t total, int C, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  float y = fmaf(a[c], x[i], b[c]);
  y = y > 0.f ? y : 0.f;
  if (mask) y *= mask[i] * scale;
  out[i] = y;
}

}  // namespace
}  // namespace pct
}  // namespace sga

extern "C" int sga_pct_point_moments(const float* pts, int64_t NP, double* mom9, void* stream) {
  if (NP <= 0) return SGA_OK;
  SGA_REQUIRE(pts && mom9, "sga_pct_point_moments: null pointer");
  int64_t blocks = (NP + 255) / 256;
  const int64_t cap = (int64_t)sga::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sga::pct::point_moments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pts, NP, mom9);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_affine_stats(const double* mom9, const float* W, int C, double* stats, void* stream) {
  SGA_REQUIRE(mom9 && W && stats && C >= 1, "sga_pct_affine_stats: bad arguments");
  sga::pct::affine_stats_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mom9, W, C, stats);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_bn_fold(const double* stats, double cnt, const float* lin_bias, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, int training,
                           float momentum, float eps, int C, float* a_out, float* b_out, void* stream) {
  SGA_REQUIRE(gamma && beta && running_mean && running_var && a_out && b_out && C >= 1, "sga_bn_fold: null pointer");
  SGA_REQUIRE(!training || (stats && cnt >= 1.0), "sga_bn_fold: training needs batch statistics");
  sga::pct::bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, cnt, lin_bias, gamma, beta, running_mean, running_var,
                                                                                num_batches_tracked, training, momentum, eps, C, a_out, b_out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_col_stats(const float* x, int64_t N, int C, double* stats, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && stats && C >= 1, "sga_col_stats: bad arguments");
  if (N >= 8192 && (C == 128 || C == 256 || C == 512 || C == 1024) && ((uintptr_t)x & 15) == 0) {
    const int rpi = 256 / (C / 4);
    int64_t blocks = (N + (int64_t)rpi * 16 - 1) / ((int64_t)rpi * 16);
    const int64_t cap = (int64_t)sga::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    sga::pct::col_stats_long_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, N, C, stats);
    SGA_LAUNCH_CHECK();
    return SGA_OK;
  }
  int gy = (int)((N + 63) / 64);
  if (gy > 64) gy = 64;
  sga::pct::col_stats_kernel<<<dim3((C + 31) / 32, gy), 256, 0, (cudaStream_t)stream>>>(x, N, C, stats);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_bn_act_rows(const float* x, const float* a, const float* b, const float* mask, float scale, int64_t N, int C,
                               float* out, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && a && b && out && C >= 1, "sga_bn_act_rows: bad arguments");
  const int64_t total = N * C;
  sga::pct::bn_act_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, a, b, mask, scale, total, C, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
