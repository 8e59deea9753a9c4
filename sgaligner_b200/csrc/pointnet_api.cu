// C-ABI entry for the PointNet feature encoder: argument checks + kernel selection.
#include "common.cuh"

namespace sga {
int pointnet_fwd_simt(const float*, int64_t, int, const float*, const float*, const float*, const float*, const float*,
                      const float*, int, float*, int32_t*, double*, cudaStream_t);
int pointnet_fwd_tc(const float*, int64_t, int, const float*, const float*, const float*, const float*, const float*,
                    const float*, int, float*, int32_t*, double*, double*, cudaStream_t);
size_t pointnet_tc_raw_doubles(int C3);
void pointnet_tc_set_max_ctas(int n);
int debug_set_trace(long long* ptr);
int debug_tie_stats(unsigned long long* host_out2, int reset);
}  // namespace sga

extern "C" int sga_pointnet_fwd(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                float* out, int32_t* argmax, int mode, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(P >= 1 && C3 >= 1, "sga_pointnet_fwd: P=%d C3=%d", P, C3);
  SGA_REQUIRE(pts && W1 && b1 && W2 && b2 && W3 && b3 && out, "sga_pointnet_fwd: null pointer");
  if (mode == SGA_POINTNET_SIMT) return sga::pointnet_fwd_simt(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, nullptr, (cudaStream_t)stream);
  if (mode == SGA_POINTNET_TC) {
    SGA_REQUIRE(((uintptr_t)W2 & 15) == 0 && ((uintptr_t)W3 & 15) == 0, "sga_pointnet_fwd(TC): W2/W3 must be 16-byte aligned");
    return sga::pointnet_fwd_tc(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, nullptr, nullptr, (cudaStream_t)stream);
  }
  sga::set_error("sga_pointnet_fwd: unknown mode %d", mode);
  return SGA_EINVAL;
}

extern "C" int sga_pointnet_set_max_ctas(int n) {
  sga::pointnet_tc_set_max_ctas(n);
  return SGA_OK;
}

extern "C" size_t sga_pointnet_stats_scratch_bytes(int C3) { return sga::pointnet_tc_raw_doubles(C3) * sizeof(double); }

extern "C" int sga_pointnet_fwd_stats(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                      const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                      float* out, int32_t* argmax, double* moments, void* scratch, size_t scratch_bytes,
                                      void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(P >= 1 && C3 >= 1, "sga_pointnet_fwd_stats: P=%d C3=%d", P, C3);
  SGA_REQUIRE(pts && W1 && b1 && W2 && b2 && W3 && b3 && out && moments && scratch, "sga_pointnet_fwd_stats: null pointer");
  SGA_REQUIRE(((uintptr_t)W2 & 15) == 0 && ((uintptr_t)W3 & 15) == 0, "sga_pointnet_fwd_stats: W2/W3 must be 16-byte aligned");
  SGA_REQUIRE(((uintptr_t)scratch & 7) == 0, "sga_pointnet_fwd_stats: scratch must be 8-byte aligned");
  if (scratch_bytes < sga_pointnet_stats_scratch_bytes(C3)) {
    sga::set_error("sga_pointnet_fwd_stats: scratch of %zu bytes, need %zu", scratch_bytes, sga_pointnet_stats_scratch_bytes(C3));
    return SGA_EWORKSPACE;
  }
  return sga::pointnet_fwd_tc(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, (double*)scratch, moments, (cudaStream_t)stream);
}

extern "C" int sga_pointnet_bn_moments(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                       const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                       double* moments, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(P >= 1 && C3 >= 1 && moments, "sga_pointnet_bn_moments: bad arguments");
  return sga::pointnet_fwd_simt(pts, N, P, W1, b1, W2, b2, W3, b3, C3, nullptr, nullptr, moments, (cudaStream_t)stream);
}

/* diagnostics only: device buffer of >= 2048 int64 that CTA 0 of the tensor-core PointNet kernel
 * fills with clock64() stamps of its pipeline events (NULL switches tracing off) */
extern "C" int sga_debug_set_trace(long long* trace) { return sga::debug_set_trace(trace); }
extern "C" int sga_debug_tie_stats(unsigned long long* counts_host, int reset) { return sga::debug_tie_stats(counts_host, reset); }

// Train-mode side effect of the discarded BatchNorm1d calls (pointnet.py:141-142,154-155,158-159; torch BatchNorm:
// running <- (1 - momentum) running + momentum batch_stat, unbiased variance, num_batches_tracked += 1) for the three
// layers in ONE launch.  moments: f64 {sum1[64], sq1[64], sum2[128], sq2[128], sum3[C3], sq3[C3]} of the pre-ReLU conv
// outputs over cnt = N*P points.
namespace sga {
namespace {
__global__ void bn_running_update_kernel(const double* __restrict__ mom, double cnt, float momentum, int C3,
                                         float* rm1, float* rv1, float* rm2, float* rv2, float* rm3, float* rv3,
                                         int64_t* nbt1, int64_t* nbt2, int64_t* nbt3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 64 + 128 + C3;
  if (i == 0) {
    *nbt1 += 1;
    *nbt2 += 1;
    *nbt3 += 1;
  }
  if (i >= total) return;
  int c, C, o;
  float *rm, *rv;
  if (i < 64) { c = i; C = 64; o = 0; rm = rm1; rv = rv1; }
  else if (i < 192) { c = i - 64; C = 128; o = 128; rm = rm2; rv = rv2; }
  else { c = i - 192; C = C3; o = 384; rm = rm3; rv = rv3; }
  const double s = mom[o + c], sq = mom[o + C + c];
  const double mean = s / cnt;
  const double var_unbiased = (sq - s * mean) / (cnt - 1.0 > 1.0 ? cnt - 1.0 : 1.0);
  // same rounding sequence as the torch ops this replaces: running.mul_(1 - m).add_(float(stat), alpha = m)
  rm[c] = rm[c] * (1.f - momentum) + momentum * (float)mean;
  rv[c] = rv[c] * (1.f - momentum) + momentum * (float)var_unbiased;
}
}  // namespace
}  // namespace sga

extern "C" int sga_bn_running_update(const double* moments, double cnt, float momentum, int C3, float* rm1, float* rv1,
                                     float* rm2, float* rv2, float* rm3, float* rv3, int64_t* nbt1, int64_t* nbt2,
                                     int64_t* nbt3, void* stream) {
  SGA_REQUIRE(moments && rm1 && rv1 && rm2 && rv2 && rm3 && rv3 && nbt1 && nbt2 && nbt3 && C3 >= 1 && cnt >= 1.0,
              "sga_bn_running_update: bad arguments");
  const int total = 64 + 128 + C3;
  sga::bn_running_update_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(moments, cnt, momentum, C3, rm1, rv1, rm2, rv2,
                                                                                        rm3, rv3, nbt1, nbt2, nbt3);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
