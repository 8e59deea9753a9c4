// Fused Adam step over one flat parameter buffer (torch.optim.Adam semantics as configured at
// src/trainers/trainval_sgaligner.py:53: L2 weight decay added to the gradient, bias-corrected
// first/second moments, eps added after the square root of the corrected second moment).
#include "common.cuh"

namespace sga {
namespace {
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float gi = g[i] * gscale;
    float pi = p[i];
    gi = fmaf(wd, pi, gi);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// ---- torch.optim.Adam skips a parameter whose .grad is None (never produced in the step: a module that is not
// selected, BatchNorm affine parameters whose outputs the reference discards).  Over a flat buffer every gradient is
// allocated, so "not produced" is recognised as a segment that is exactly zero: such a segment is left untouched
// (no weight-decay drift, no moment update), as torch leaves a grad-None parameter.
__global__ void __launch_bounds__(1024)
seg_nonzero_kernel(const float* __restrict__ g, const int64_t* __restrict__ seg_off, int32_t* __restrict__ active) {
  const int s = blockIdx.x;
  const int64_t a = seg_off[s], b = seg_off[s + 1];     // segments start 256-byte aligned and are padded to 64 floats
  const float4* g4 = reinterpret_cast<const float4*>(g + a);
  const int64_t n4 = (b - a) >> 2;
  int any = 0;
  for (int64_t i = threadIdx.x; i < n4; i += 1024) {
    const float4 v = g4[i];
    any |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) active[s] = any;
}

__global__ void __launch_bounds__(256)
adam_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                int64_t n, const int64_t* __restrict__ seg_off, const int32_t* __restrict__ active, int nseg,
                float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  extern __shared__ int64_t soff[];
  for (int i = threadIdx.x; i <= nseg; i += 256) soff[i] = seg_off[i];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    int lo = 0, hi = nseg;                 // segment with soff[lo] <= i < soff[lo + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (soff[mid] <= i) lo = mid; else hi = mid;
    }
    if (!active[lo]) continue;
    float gi = g[i] * gscale;
    float pi = p[i];
    gi = fmaf(wd, pi, gi);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}
}  // namespace
}  // namespace sga

extern "C" int sga_adam_step_segments(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                      const int64_t* seg_off, int nseg, int32_t* seg_active, float lr, float beta1,
                                      float beta2, float eps, float weight_decay, int step, float grad_scale,
                                      void* stream) {
  if (n <= 0) return SGA_OK;
  SGA_REQUIRE(step >= 1 && nseg >= 1 && nseg <= 4096 && seg_off && seg_active, "sga_adam_step_segments: step=%d nseg=%d", step, nseg);
  SGA_REQUIRE(((uintptr_t)grad & 15) == 0, "sga_adam_step_segments: grad must be 16-byte aligned (segments start at multiples of 4 floats)");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  sga::seg_nonzero_kernel<<<nseg, 1024, 0, (cudaStream_t)stream>>>(grad, seg_off, seg_active);
  SGA_LAUNCH_CHECK();
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)sga::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sga::adam_seg_kernel<<<(unsigned)blocks, 256, (nseg + 1) * sizeof(int64_t), (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, seg_off, seg_active, nseg, lr, beta1, beta2, eps, weight_decay, (float)bc1,
      (float)sqrt(bc2), grad_scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             float grad_scale, void* stream) {
  if (n <= 0) return SGA_OK;
  SGA_REQUIRE(step >= 1, "sga_adam_step: step=%d must be >= 1", step);
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)sga::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sga::adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                      weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
