// NaivePCT backward: the HBM-bound element-wise / reduction kernels around the tensor-core products (reference: autograd
// of src/aligner/networks/pct.py:101-125,187-232,275-317; orchestration in sgaligner_b200/pct.py).
//
// BatchNorm in closed form.  For y -> BN -> activation with upstream gradient g:  gy = g * act'(a y + b) (* dropout mask),
//     S1 = sum gy,  S2 = sum gy y  over the batch          d beta = S1,  d gamma = (S2 - mean S1) / sigma
//     train():  dy = a gy - e - f (y - mean),   f = a (S2 - mean S1) / (sigma^2 cnt),   e = a S1 / cnt
//     eval():   dy = a gy                                      (running statistics are constants)
// (y - mean) is formed first: with |mean| >> sigma the two halves of  -f y + f mean  would cancel digits.
// so every BatchNorm costs one reduction pass (bn_bwd_stats) and one element-wise pass (bn_bwd_apply) over tensors the
// forward stored anyway -- and for the 512 -> 1024 convolution, whose [N, P, 1024] output is never stored, the same
// algebra turns the dense part of the gradient into a 512 x 512 pointwise map of the inputs (pct_cat.cu, kMode 2).
#include "common.cuh"

namespace sga {
namespace pct {
namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------- BatchNorm backward
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ mask, float scale, float slope, int64_t rows,
                                                           int C, double* __restrict__ sums) {
  const int tpr = C >> 2, rpi = 256 / tpr;
  const int col4 = threadIdx.x % tpr, rslot = threadIdx.x / tpr;
  const float4 av = ld4(a + col4 * 4), bv = ld4(b + col4 * 4);
  const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = (int64_t)blockIdx.x * rpi + rslot; r < rows; r += (int64_t)gridDim.x * rpi) {
    const float4 gv = ld4(g + r * C + col4 * 4), yv = ld4(y + r * C + col4 * 4);
    float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    if (mask) {
      const float4 mv = ld4(mask + r * C + col4 * 4);
      gg[0] *= mv.x * scale; gg[1] *= mv.y * scale; gg[2] *= mv.z * scale; gg[3] *= mv.w * scale;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float gy = gg[i] * (fmaf(aa[i], yy[i], bb[i]) > 0.f ? 1.f : slope);
      s1[i] += gy;
      s2[i] = fmaf(gy, yy[i], s2[i]);
    }
  }
  __shared__ float red[256][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[threadIdx.x][i] = s1[i];
    red[threadIdx.x][4 + i] = s2[i];
  }
  __syncthreads();
  if (rslot == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double t1 = 0, t2 = 0;
      for (int k = 0; k < rpi; ++k) {
        t1 += (double)red[k * tpr + col4][i];
        t2 += (double)red[k * tpr + col4][4 + i];
      }
      atomicAdd(&sums[col4 * 4 + i], t1);
      atomicAdd(&sums[C + col4 * 4 + i], t2);
    }
  }
}

__global__ void bn_bwd_coef_kernel(const double* __restrict__ sums, const double* __restrict__ stats, double cnt,
                                   const float* __restrict__ lin_bias, const float* __restrict__ gamma, const float* __restrict__ rmean,
                                   const float* __restrict__ rvar, int training, float eps, int C, float* __restrict__ e_out,
                                   float* __restrict__ f_out, float* __restrict__ mean_out, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double S1 = sums[c], S2 = sums[C + c];
  double mean, var;
  if (training) {
    mean = stats[c] / cnt;
    var = stats[C + c] / cnt - mean * mean;
    if (var < 0) var = 0;
  } else {
    mean = (double)rmean[c] - (lin_bias ? (double)lin_bias[c] : 0.0);
    var = (double)rvar[c];
  }
  const double s2 = var + (double)eps, inv = 1.0 / sqrt(s2);
  const double a = (double)gamma[c] * inv;
  const double proj = S2 - mean * S1;                 // sum gy (y - mean)
  if (dgamma) dgamma[c] = (float)(proj * inv);
  if (dbeta) dbeta[c] = (float)S1;
  double e = 0, f = 0;
  if (training) {
    f = a * proj / (s2 * cnt);
    e = a * S1 / cnt;
  }
  e_out[c] = (float)e;
  f_out[c] = (float)f;
  mean_out[c] = (float)mean;
}

constexpr int kApplyPer = 4;       // float4 per thread: a CTA covers 1024 float4 (4096 elements)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ mask, float scale, float slope,
                                                           const float* __restrict__ e, const float* __restrict__ f,
                                                           const float* __restrict__ mean, int64_t total4, int C,
                                                           float* __restrict__ out, int64_t per_obj4, float* __restrict__ absmax) {
  const int64_t i0 = (int64_t)blockIdx.x * (256 * kApplyPer);
  const int64_t i1 = min(i0 + 256 * kApplyPer - 1, total4 - 1);
  // max |out| per object (the operand scale of the next tensor-core product): one atomic per CTA when its 4096 elements
  // belong to one object (always, when P * C is a multiple of 4096), one per thread and element group in a CTA that straddles two
  const bool one_obj = absmax != nullptr && (i0 / per_obj4 == i1 / per_obj4);
  float amax = 0.f;
  float4 gv[kApplyPer], yv[kApplyPer];
#pragma unroll
  for (int u = 0; u < kApplyPer; ++u) {
    const int64_t i = i0 + u * 256 + threadIdx.x;
    if (i < total4) {
      gv[u] = ld4(g + i * 4);
      yv[u] = ld4(y + i * 4);
    }
  }
#pragma unroll
  for (int u = 0; u < kApplyPer; ++u) {
    const int64_t i = i0 + u * 256 + threadIdx.x;
    if (i >= total4) continue;
    const int c = (int)((i * 4) % C);
    const float4 av = ld4(a + c), bv = ld4(b + c);
    float gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
    const float yy[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w}, aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
    float ee[4] = {0.f, 0.f, 0.f, 0.f}, ff[4] = {0.f, 0.f, 0.f, 0.f}, mm[4] = {0.f, 0.f, 0.f, 0.f};
    if (e) {
      const float4 ev = ld4(e + c), fv = ld4(f + c), mv = ld4(mean + c);
      ee[0] = ev.x; ee[1] = ev.y; ee[2] = ev.z; ee[3] = ev.w;
      ff[0] = fv.x; ff[1] = fv.y; ff[2] = fv.z; ff[3] = fv.w;
      mm[0] = mv.x; mm[1] = mv.y; mm[2] = mv.z; mm[3] = mv.w;
    }
    if (mask) {
      const float4 mv = ld4(mask + i * 4);
      gg[0] *= mv.x * scale; gg[1] *= mv.y * scale; gg[2] *= mv.z * scale; gg[3] *= mv.w * scale;
    }
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gy = gg[k] * (fmaf(aa[k], yy[k], bb[k]) > 0.f ? 1.f : slope);
      o[k] = fmaf(aa[k], gy, -ee[k]) - ff[k] * (yy[k] - mm[k]);
    }
    *reinterpret_cast<float4*>(out + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
    const float m = fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3])));
    if (one_obj) amax = fmaxf(amax, m);
    else if (absmax) atomicMax(reinterpret_cast<unsigned int*>(absmax + i / per_obj4), __float_as_uint(m));
  }
  if (!one_obj) return;
  __shared__ float red[8];
  amax = warp_max(amax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    atomicMax(reinterpret_cast<unsigned int*>(absmax + i0 / per_obj4), __float_as_uint(m));
  }
}

// ---------------------------------------------------------------------------------------------- small row kernels
// delta[n, p] = scale[n][0] * sum_c v[n,p,c] dv[n,p,c]   (= sum_j A[i,j] dA[i,j] of the softmax backward, in scaled units)
__global__ void __launch_bounds__(256) rowdot_scaled_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ scale, int64_t rows, int P, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4 a = ld4(x + r * 128 + lane * 4), b = ld4(y + r * 128 + lane * 4);
  const float s = warp_sum(a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w);
  if (lane == 0) out[r] = s * __ldg(scale + 2 * (r / P));
}

// out = gx (+ gcat) + dxv + (dk1 + dk2) Wk; dk1 <- dk1 + dk2.  One warp per group of four rows, lane = 4 output channels; the loads of
// all rows are issued before the first use (the 32-step shuffle / FMA chain of one row runs under the other's loads).
__global__ void __launch_bounds__(256) sa_input_grad_kernel(const float* __restrict__ gx, const float* __restrict__ gcat,
                                                            const float* __restrict__ dxv, float* __restrict__ dk1,
                                                            const float* __restrict__ dk2, const float* __restrict__ Wk, int64_t rows,
                                                            float* __restrict__ out) {
  __shared__ float4 ws[32][32];
  for (int i = threadIdx.x; i < 32 * 32; i += 256) ws[i >> 5][i & 31] = ld4(Wk + (int64_t)i * 4);
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  constexpr int R = 4;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + wrp) * R; r0 < rows; r0 += (int64_t)gridDim.x * 8 * R) {
    float dks[R];
    float4 acc[R], dv[R], gc[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t r = min(r0 + u, rows - 1);
      dks[u] = dk1[r * 32 + lane] + dk2[r * 32 + lane];
      acc[u] = ld4(gx + r * 128 + lane * 4);
      dv[u] = ld4(dxv + r * 128 + lane * 4);
      gc[u] = gcat ? ld4(gcat + r * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      acc[u].x += dv[u].x; acc[u].y += dv[u].y; acc[u].z += dv[u].z; acc[u].w += dv[u].w;
      if (gcat) { acc[u].x += gc[u].x; acc[u].y += gc[u].y; acc[u].z += gc[u].z; acc[u].w += gc[u].w; }
    }
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float4 w = ws[kk][lane];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const float d = __shfl_sync(0xffffffffu, dks[u], kk);
        acc[u].x = fmaf(d, w.x, acc[u].x); acc[u].y = fmaf(d, w.y, acc[u].y); acc[u].z = fmaf(d, w.z, acc[u].z); acc[u].w = fmaf(d, w.w, acc[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t r = r0 + u;
      if (r < rows) {
        dk1[r * 32 + lane] = dks[u];
        *reinterpret_cast<float4*>(out + r * 128 + lane * 4) = acc[u];
      }
    }
  }
}

__global__ void __launch_bounds__(256) residual_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                       const float* __restrict__ a, const float* __restrict__ b, int64_t total4,
                                                       float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total4) return;
  const int c = (int)((i * 4) & 127);
  const float4 xv = ld4(x + i * 4), tv = ld4(t + i * 4), av = ld4(a + c), bv = ld4(b + c);
  float4 o;
  o.x = xv.x + fmaxf(fmaf(av.x, tv.x, bv.x), 0.f);
  o.y = xv.y + fmaxf(fmaf(av.y, tv.y, bv.y), 0.f);
  o.z = xv.z + fmaxf(fmaf(av.z, tv.z, bv.z), 0.f);
  o.w = xv.w + fmaxf(fmaf(av.w, tv.w, bv.w), 0.f);
  *reinterpret_cast<float4*>(out + i * 4) = o;
}

__global__ void __launch_bounds__(256) axpby_rows_kernel(float* __restrict__ dst, float alpha, const float* __restrict__ src, float beta,
                                                         const float* __restrict__ rs, float gamma, const float* __restrict__ rs2,
                                                         const double* __restrict__ colvec, int64_t R, int C) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= R * C) return;
  const int64_t r = i / C;
  const int c = (int)(i - r * C);
  float v = (alpha != 0.f) ? alpha * dst[i] : 0.f;
  if (src) v = fmaf(beta * (rs ? rs[r] : 1.f), src[i], v);
  if (colvec) v = fmaf(gamma * (rs2 ? rs2[r] : 1.f), (float)colvec[c], v);
  dst[i] = v;
}

// ---------------------------------------------------------------------------------------------- Embedding.conv1 / bn1
__global__ void __launch_bounds__(256) embed_a1_kernel(const float* __restrict__ pts, const float* __restrict__ W1,
                                                       const float* __restrict__ a, const float* __restrict__ b, int64_t rows,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  float wx[4], wy[4], wz[4], aa[4], bb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane * 4 + i;
    wx[i] = W1[c * 3]; wy[i] = W1[c * 3 + 1]; wz[i] = W1[c * 3 + 2];
    aa[i] = a[c]; bb[i] = b[c];
  }
  for (int64_t r = (int64_t)blockIdx.x * 8 + wrp; r < rows; r += (int64_t)gridDim.x * 8) {
    const float px = __ldg(pts + r * 3), py = __ldg(pts + r * 3 + 1), pz = __ldg(pts + r * 3 + 2);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float z = fmaf(wx[i], px, fmaf(wy[i], py, wz[i] * pz));
      o[i] = fmaxf(fmaf(aa[i], z, bb[i]), 0.f);
    }
    *reinterpret_cast<float4*>(out + r * 128 + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256) embed1_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ pts,
                                                               const float* __restrict__ W1, const float* __restrict__ a,
                                                               const float* __restrict__ b, int64_t rows, double* __restrict__ sums5) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  float wx[4], wy[4], wz[4], aa[4], bb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane * 4 + i;
    wx[i] = W1[c * 3]; wy[i] = W1[c * 3 + 1]; wz[i] = W1[c * 3 + 2];
    aa[i] = a[c]; bb[i] = b[c];
  }
  float s[4][5];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 5; ++k) s[i][k] = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * 8 + wrp; r < rows; r += (int64_t)gridDim.x * 8) {
    const float px = __ldg(pts + r * 3), py = __ldg(pts + r * 3 + 1), pz = __ldg(pts + r * 3 + 2);
    const float4 gv = ld4(g + r * 128 + lane * 4);
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float z = fmaf(wx[i], px, fmaf(wy[i], py, wz[i] * pz));
      const float gy = fmaf(aa[i], z, bb[i]) > 0.f ? gg[i] : 0.f;
      s[i][0] += gy;
      s[i][1] = fmaf(gy, z, s[i][1]);
      s[i][2] = fmaf(gy, px, s[i][2]);
      s[i][3] = fmaf(gy, py, s[i][3]);
      s[i][4] = fmaf(gy, pz, s[i][4]);
    }
  }
  __shared__ float red[8][128][5];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 5; ++k) red[wrp][lane * 4 + i][k] = s[i][k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 128 * 5; idx += 256) {
    const int c = idx / 5, k = idx - c * 5;
    double t = 0;
    for (int w = 0; w < 8; ++w) t += (double)red[w][c][k];
    atomicAdd(&sums5[k * 128 + c], t);
  }
}

// dz1 = a gy - e - f (z1 - mean), z1 = W1 p  ->  dW1[c,:] = a T_c - e m1 - f (W1_c M2 - mean_c m1),  mean_c = W1_c . m1 / cnt
__global__ void embed1_wgrad_kernel(const double* __restrict__ sums5, const double* __restrict__ mom, double cnt, const float* __restrict__ W1,
                                    const float* __restrict__ a, const float* __restrict__ e, const float* __restrict__ f,
                                    float* __restrict__ dW1) {
  const int c = threadIdx.x;
  if (c >= 128) return;
  const double wx = W1[c * 3], wy = W1[c * 3 + 1], wz = W1[c * 3 + 2];
  const double mean = (wx * mom[0] + wy * mom[1] + wz * mom[2]) / cnt;
  // M2 = sum p p^T: mom[3..8] = xx xy xz yy yz zz; centred: M2 - m1 m1^T / cnt, contracted with W1_c
  const double m2x = wx * mom[3] + wy * mom[4] + wz * mom[5] - mean * mom[0];
  const double m2y = wx * mom[4] + wy * mom[6] + wz * mom[7] - mean * mom[1];
  const double m2z = wx * mom[5] + wy * mom[7] + wz * mom[8] - mean * mom[2];
  const double ac = a[c], ec = e[c], fc = f[c];
  dW1[c * 3] = (float)(ac * sums5[2 * 128 + c] - ec * mom[0] - fc * m2x);
  dW1[c * 3 + 1] = (float)(ac * sums5[3 * 128 + c] - ec * mom[1] - fc * m2y);
  dW1[c * 3 + 2] = (float)(ac * sums5[4 * 128 + c] - ec * mom[2] - fc * m2z);
}

// ---------------------------------------------------------------------------------------------- concat-conv, sparse half
// One CTA per object: bucket the 1024 channels by their arg-max point (counting sort in shared memory), then a warp per
// point sums coef_c WL[c, :] over the point's channels and adds the 512 values to the four g_x rows of that point.
__global__ void __launch_bounds__(256) cat_sparse_bwd_x_kernel(const float* __restrict__ coef, const int32_t* __restrict__ pstar,
                                                               const float* __restrict__ WL, int P, float* __restrict__ g1,
                                                               float* __restrict__ g2, float* __restrict__ g3, float* __restrict__ g4) {
  __shared__ int cnt[512], off[513], cur[512];
  __shared__ int order[1024], ps[1024];
  __shared__ float cf[1024];
  const int64_t n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  for (int i = tid; i < 512; i += 256) cnt[i] = 0;
  __syncthreads();
  for (int c = tid; c < 1024; c += 256) {
    const int p = pstar[n * 1024 + c];
    ps[c] = p;
    cf[c] = coef[n * 1024 + c];
    atomicAdd(&cnt[p], 1);
  }
  __syncthreads();
  if (wrp == 0) {
    int loc[16], tot = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      loc[i] = tot;
      tot += cnt[lane * 16 + i];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int base = incl - tot;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      off[lane * 16 + i] = base + loc[i];
      cur[lane * 16 + i] = base + loc[i];
    }
    if (lane == 31) off[512] = incl;
  }
  __syncthreads();
  for (int c = tid; c < 1024; c += 256) order[atomicAdd(&cur[ps[c]], 1)] = c;
  __syncthreads();
  float* const gs[4] = {g1, g2, g3, g4};
  for (int p = wrp; p < P; p += 8) {
    const int beg = off[p], end = off[p + 1];
    if (beg == end) continue;
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = beg; k < end; ++k) {
      const int c = order[k];
      const float w = cf[c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 wv = ld4(WL + (int64_t)c * 512 + i * 128 + lane * 4);
        acc[i].x = fmaf(w, wv.x, acc[i].x); acc[i].y = fmaf(w, wv.y, acc[i].y);
        acc[i].z = fmaf(w, wv.z, acc[i].z); acc[i].w = fmaf(w, wv.w, acc[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4* dst = reinterpret_cast<float4*>(gs[i] + (n * (int64_t)P + p) * 128 + lane * 4);
      float4 v = *dst;
      v.x += acc[i].x; v.y += acc[i].y; v.z += acc[i].z; v.w += acc[i].w;
      *dst = v;
    }
  }
}

// dWL[c, :] += sum_n coef[n, c] xcat[n, pstar[n, c], :]: a warp per channel, the objects cut into gridDim.y slices
__global__ void __launch_bounds__(256) cat_sparse_bwd_w_kernel(const float* __restrict__ coef, const int32_t* __restrict__ pstar,
                                                               const float* __restrict__ x1, const float* __restrict__ x2,
                                                               const float* __restrict__ x3, const float* __restrict__ x4, int64_t N,
                                                               int P, float* __restrict__ dWL) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + wrp;
  const int64_t per = (N + gridDim.y - 1) / gridDim.y;
  const int64_t n0 = (int64_t)blockIdx.y * per, n1 = min(N, n0 + per);
  const float* const xs[4] = {x1, x2, x3, x4};
  float4 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int64_t n = n0; n < n1; ++n) {
    const float w = __ldg(coef + n * 1024 + c);
    const int64_t row = n * (int64_t)P + __ldg(pstar + n * 1024 + c);
    if (w == 0.f) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 xv = ld4(xs[i] + row * 128 + lane * 4);
      acc[i].x = fmaf(w, xv.x, acc[i].x); acc[i].y = fmaf(w, xv.y, acc[i].y);
      acc[i].z = fmaf(w, xv.z, acc[i].z); acc[i].w = fmaf(w, xv.w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* dst = dWL + (int64_t)c * 512 + i * 128 + lane * 4;
    atomicAdd(dst, acc[i].x);
    atomicAdd(dst + 1, acc[i].y);
    atomicAdd(dst + 2, acc[i].z);
    atomicAdd(dst + 3, acc[i].w);
  }
}

// Per-object power-of-two operand scale of the backward's tensor-core products (pct_common.cuh): one CTA per object,
// scale[n] = {s, 1/s}, s = 2^floor(log2(target / (max|x_n| * (y ? max|y_n| : 1)))), 1 when the object's x is all zero.
__global__ void __launch_bounds__(256) pow2_scale_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t per,
                                                         float target, float* __restrict__ scale) {
  const int64_t n = blockIdx.x;
  const float4* xp = reinterpret_cast<const float4*>(x + n * per);
  const float4* yp = y ? reinterpret_cast<const float4*>(y + n * per) : nullptr;
  float mx = 0.f, my = 0.f;
  for (int64_t i = threadIdx.x; i < per / 4; i += 256) {
    const float4 v = __ldg(xp + i);
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    if (yp) {
      const float4 w = __ldg(yp + i);
      my = fmaxf(my, fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fmaxf(fabsf(w.z), fabsf(w.w))));
    }
  }
  __shared__ float rx[8], ry[8];
  mx = warp_max(mx);
  my = warp_max(my);
  if ((threadIdx.x & 31) == 0) {
    rx[threadIdx.x >> 5] = mx;
    ry[threadIdx.x >> 5] = my;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      mx = fmaxf(mx, rx[w]);
      my = fmaxf(my, ry[w]);
    }
    float den = mx * (yp ? fmaxf(my, 1e-30f) : 1.f);
    int ex = 0;
    if (den > 0.f && isfinite(den)) ex = (int)floorf(log2f(target / den));
    ex = max(-100, min(100, ex));
    scale[2 * n] = exp2f((float)ex);
    scale[2 * n + 1] = exp2f((float)-ex);
  }
}

// scale[n] = {s, 1/s}, s = 2^floor(log2(target / absmax[n])) -- the per-object scale from a maximum a producing kernel recorded
__global__ void scale_from_absmax_kernel(const float* __restrict__ absmax, const float* __restrict__ absmax2, int64_t N, float target,
                                         float* __restrict__ scale) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float den = absmax[n] * (absmax2 ? fmaxf(absmax2[n], 1e-30f) : 1.f);      // as pow2_scale_kernel with (x, y)
  int ex = 0;
  if (den > 0.f && isfinite(den)) ex = (int)floorf(log2f(target / den));
  ex = max(-100, min(100, ex));
  scale[2 * n] = exp2f((float)ex);
  scale[2 * n + 1] = exp2f((float)-ex);
}

inline unsigned grid_for(int64_t work_rows, int per_block) {
  int64_t blocks = (work_rows + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace
}  // namespace pct
}  // namespace sga

using namespace sga::pct;

static bool bn_c_ok(int C) { return C == 128 || C == 256 || C == 512 || C == 1024; }

extern "C" int sga_bn_bwd_stats(const float* g, const float* y, const float* a, const float* b, const float* mask, float scale,
                                float slope, int64_t rows, int C, double* sums, void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(g && y && a && b && sums && bn_c_ok(C), "sga_bn_bwd_stats: bad arguments (C=%d)", C);
  SGA_REQUIRE((((uintptr_t)g | (uintptr_t)y | (uintptr_t)a | (uintptr_t)b | (uintptr_t)mask) & 15) == 0, "sga_bn_bwd_stats: 16-byte alignment");
  const int rpi = 256 / (C / 4);
  bn_bwd_stats_kernel<<<grid_for(rows, rpi * 4), 256, 0, (cudaStream_t)stream>>>(g, y, a, b, mask, scale, slope, rows, C, sums);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_bn_bwd_coef(const double* sums, const double* stats, double cnt, const float* lin_bias, const float* gamma,
                               const float* running_mean, const float* running_var, int training, float eps, int C, float* e,
                               float* f, float* mean, float* dgamma, float* dbeta, void* stream) {
  SGA_REQUIRE(sums && gamma && e && f && mean && C >= 1, "sga_bn_bwd_coef: null pointer");
  SGA_REQUIRE(training ? (stats && cnt >= 1.0) : (running_mean && running_var), "sga_bn_bwd_coef: statistics missing");
  bn_bwd_coef_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, stats, cnt, lin_bias, gamma, running_mean, running_var,
                                                                        training, eps, C, e, f, mean, dgamma, dbeta);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_bn_bwd_apply(const float* g, const float* y, const float* a, const float* b, const float* mask, float scale,
                                float slope, const float* e, const float* f, const float* mean, int64_t rows, int C, float* out,
                                void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(g && y && a && b && out && C >= 4 && C % 4 == 0 && (e == nullptr) == (f == nullptr) && (e == nullptr) == (mean == nullptr),
              "sga_bn_bwd_apply: bad arguments");
  SGA_REQUIRE((((uintptr_t)g | (uintptr_t)y | (uintptr_t)a | (uintptr_t)b | (uintptr_t)mask | (uintptr_t)e | (uintptr_t)f | (uintptr_t)out) & 15) == 0,
              "sga_bn_bwd_apply: 16-byte alignment");
  const int64_t total4 = rows * C / 4;
  bn_bwd_apply_kernel<<<(unsigned)((total4 + 256 * kApplyPer - 1) / (256 * kApplyPer)), 256, 0, (cudaStream_t)stream>>>(g, y, a, b, mask, scale, slope, e, f, mean, total4, C, out,
                                                                                          0, nullptr);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

// The same pass that also records absmax [rows / rows_per_object] (zeroed by the caller) = max |out| per object.
extern "C" int sga_bn_bwd_apply_absmax(const float* g, const float* y, const float* a, const float* b, const float* mask, float scale,
                                       float slope, const float* e, const float* f, const float* mean, int64_t rows, int C, float* out,
                                       int64_t rows_per_object, float* absmax, void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(g && y && a && b && out && absmax && C >= 4 && C % 4 == 0 && (e == nullptr) == (f == nullptr) && (e == nullptr) == (mean == nullptr),
              "sga_bn_bwd_apply_absmax: bad arguments");
  SGA_REQUIRE(rows_per_object >= 1 && rows % rows_per_object == 0, "sga_bn_bwd_apply_absmax: rows=%lld rows_per_object=%lld", (long long)rows,
              (long long)rows_per_object);
  SGA_REQUIRE((((uintptr_t)g | (uintptr_t)y | (uintptr_t)a | (uintptr_t)b | (uintptr_t)mask | (uintptr_t)e | (uintptr_t)f | (uintptr_t)out) & 15) == 0,
              "sga_bn_bwd_apply_absmax: 16-byte alignment");
  const int64_t total4 = rows * C / 4;
  bn_bwd_apply_kernel<<<(unsigned)((total4 + 256 * kApplyPer - 1) / (256 * kApplyPer)), 256, 0, (cudaStream_t)stream>>>(g, y, a, b, mask, scale, slope, e, f, mean, total4, C, out,
                                                                                          rows_per_object * C / 4, absmax);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_rowdot_scaled(const float* x, const float* y, const float* scale, int64_t N, int P, float* out, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && y && scale && out && P >= 1, "sga_pct_rowdot_scaled: bad arguments");
  const int64_t rows = N * P;
  rowdot_scaled_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, y, scale, rows, P, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_sa_input_grad(const float* gx, const float* gcat, const float* dxv, float* dk1, const float* dk2,
                                     const float* Wk, int64_t rows, float* out, void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(gx && dxv && dk1 && dk2 && Wk && out, "sga_pct_sa_input_grad: null pointer");
  sa_input_grad_kernel<<<grid_for(rows, 128), 256, 0, (cudaStream_t)stream>>>(gx, gcat, dxv, dk1, dk2, Wk, rows, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_residual(const float* x, const float* t, const float* a, const float* b, int64_t rows, float* out, void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(x && t && a && b && out, "sga_pct_residual: null pointer");
  const int64_t total4 = rows * 32;
  residual_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, t, a, b, total4, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_axpby_rows(float* dst, float alpha, const float* src, float beta, const float* rowscale, float gamma,
                              const float* rowscale2, const double* colvec, int64_t R, int C, void* stream) {
  if (R <= 0 || C <= 0) return SGA_OK;
  SGA_REQUIRE(dst, "sga_axpby_rows: null pointer");
  axpby_rows_kernel<<<(unsigned)((R * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, alpha, src, beta, rowscale, gamma, rowscale2, colvec, R, C);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_embed_a1(const float* pts, const float* W1, const float* a, const float* b, int64_t rows, float* out, void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(pts && W1 && a && b && out, "sga_pct_embed_a1: null pointer");
  embed_a1_kernel<<<grid_for(rows, 32), 256, 0, (cudaStream_t)stream>>>(pts, W1, a, b, rows, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_embed1_bwd_stats(const float* g, const float* pts, const float* W1, const float* a, const float* b,
                                        int64_t rows, double* sums5, void* stream) {
  if (rows <= 0) return SGA_OK;
  SGA_REQUIRE(g && pts && W1 && a && b && sums5, "sga_pct_embed1_bwd_stats: null pointer");
  embed1_bwd_stats_kernel<<<grid_for(rows, 64), 256, 0, (cudaStream_t)stream>>>(g, pts, W1, a, b, rows, sums5);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_embed1_wgrad(const double* sums5, const double* mom9, double cnt, const float* W1, const float* a, const float* e,
                                    const float* f, float* dW1, void* stream) {
  SGA_REQUIRE(sums5 && mom9 && W1 && a && e && f && dW1, "sga_pct_embed1_wgrad: null pointer");
  embed1_wgrad_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sums5, mom9, cnt, W1, a, e, f, dW1);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_cat_sparse_bwd_x(const float* coef, const int32_t* pstar, const float* WL, int64_t N, int P, float* g1,
                                        float* g2, float* g3, float* g4, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(coef && pstar && WL && g1 && g2 && g3 && g4 && P >= 1 && P <= 512, "sga_pct_cat_sparse_bwd_x: bad arguments (P=%d)", P);
  cat_sparse_bwd_x_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(coef, pstar, WL, P, g1, g2, g3, g4);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_cat_sparse_bwd_w(const float* coef, const int32_t* pstar, const float* x1, const float* x2, const float* x3,
                                        const float* x4, int64_t N, int P, float* dWL, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(coef && pstar && x1 && x2 && x3 && x4 && dWL && P >= 1, "sga_pct_cat_sparse_bwd_w: bad arguments");
  int gy = (int)((N + 63) / 64);
  if (gy > 16) gy = 16;
  cat_sparse_bwd_w_kernel<<<dim3(128, gy), 256, 0, (cudaStream_t)stream>>>(coef, pstar, x1, x2, x3, x4, N, P, dWL);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_pow2_scale(const float* x, const float* y, int64_t N, int64_t per, float target, float* scale, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && scale && per >= 4 && per % 4 == 0 && target > 0.f, "sga_pct_pow2_scale: bad arguments");
  SGA_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "sga_pct_pow2_scale: 16-byte alignment");
  pow2_scale_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(x, y, per, target, scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_scale_from_absmax(const float* absmax, int64_t N, float target, float* scale, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(absmax && scale && target > 0.f, "sga_pct_scale_from_absmax: bad arguments");
  scale_from_absmax_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(absmax, nullptr, N, target, scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_scale_from_absmax_pair(const float* absmax_x, const float* absmax_y, int64_t N, float target, float* scale,
                                              void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(absmax_x && absmax_y && scale && target > 0.f, "sga_pct_scale_from_absmax_pair: bad arguments");
  scale_from_absmax_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(absmax_x, absmax_y, N, target, scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
