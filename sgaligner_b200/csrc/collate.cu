// Device side of the batch collation: the per-pair centring of the object points
// (src/datasets/scan3r.py:99-100: obj_points - pcl_center, a float32 subtraction) applied in place to the raw points
// after the H2D copy, so the host never touches the 12*P bytes/object it ships.  HBM bound: 24*P bytes per object.
#include "common.cuh"

namespace sga {
namespace {

__global__ void __launch_bounds__(256)
center_points_kernel(float* __restrict__ pts, int64_t N, int P, const float* __restrict__ center,
                     const int32_t* __restrict__ node_pair) {
  // one object = 3*P consecutive floats; a block strides over objects, threads over the floats of one object
  for (int64_t o = blockIdx.x; o < N; o += gridDim.x) {
    const float* c = center + 3 * (int64_t)node_pair[o];
    const float c0 = c[0], c1 = c[1], c2 = c[2];
    float* p = pts + o * 3 * (int64_t)P;
    const int n = 3 * P;
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      // 12 floats = 4 points = 3 float4; thread t handles float4 index t, whose coordinate phase is (4t) % 3
      for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
        float4 v = reinterpret_cast<float4*>(p)[i];
        const int ph = (4 * i) % 3;
        const float a = ph == 0 ? c0 : (ph == 1 ? c1 : c2);
        const float b = ph == 0 ? c1 : (ph == 1 ? c2 : c0);
        const float d = ph == 0 ? c2 : (ph == 1 ? c0 : c1);
        v.x -= a; v.y -= b; v.z -= d; v.w -= a;
        reinterpret_cast<float4*>(p)[i] = v;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ph = i % 3;
        p[i] -= ph == 0 ? c0 : (ph == 1 ? c1 : c2);
      }
    }
  }
}

}  // namespace
}  // namespace sga

extern "C" int sga_center_points(float* pts, int64_t N, int P, const float* center, const int32_t* node_pair, void* stream) {
  if (N <= 0 || P <= 0) return SGA_OK;
  const int64_t cap = (int64_t)sga::sm_count() * 8;
  sga::center_points_kernel<<<(unsigned)(N < cap ? N : cap), 256, 0, (cudaStream_t)stream>>>(pts, N, P, center, node_pair);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
