// Backward of the PointNet feature encoder through the max-pool (autograd of
// src/aligner/networks/pointnet.py:140-163; BatchNorm layers get no gradient there because their
// outputs are discarded).
//
// Only the point that attains the max of channel c of object n receives gradient for that channel,
// so the backward works on "instances" (n, c, p = argmax[n][c]) -- at most C3 per object instead
// of P points: conv1/conv2 activations are recomputed for those points only, never stored.
//   dW3[c,:] += g_c h2[p_c,:]            db3[c] += g_c             (g_c = grad_out if out > 0)
//   dz2_c     = (h2[p_c,:] > 0) . g_c W3[c,:]                       (one row per instance; linear, so
//   dW2      += dz2_c^T h1[p_c,:]         db2 += dz2_c                instances sharing a point just add)
//   dz1_c     = (h1[p_c,:] > 0) . (dz2_c W2)
//   dW1      += dz1_c^T x[p_c]            db1 += dz1_c
// One persistent CTA per SM and 128-channel block; dW3 block (64 KiB) lives in shared memory, dW2
// as a 4x8 register tile per thread, both flushed with one round of atomics at the end.
#include "common.cuh"

namespace sga {
namespace {

constexpr int NT = 256;
constexpr int CBW = 128;   // conv3 channels per CTA
constexpr int IC = 32;     // instances per chunk

struct Smem {
  float dW3[CBW][128];     // 64 KiB
  float W2s[128][64];      // natural layout  (dh1 = dz2 W2)
  float W2t[64][128];      // transposed      (h2 = h1 W2^T)
  float h1[IC][64 + 4];
  float h2[IC][128 + 4];
  float dz2[IC][128 + 4];
  float dz1[IC][64 + 4];
  float x[IC][4];
  float g[IC];
  float db3[CBW];
  float W1s[64][4];        // {w0,w1,w2,b1}
  float b2s[128];
};

__global__ void __launch_bounds__(NT, 1)
pointnet_bwd_kernel(const float* __restrict__ pts, int64_t N, int P,
                    const float* __restrict__ W1, const float* __restrict__ b1,
                    const float* __restrict__ W2, const float* __restrict__ b2,
                    const float* __restrict__ W3, int C3,
                    const float* __restrict__ out, const int32_t* __restrict__ argmax,
                    const float* __restrict__ gout,
                    float* __restrict__ gW1, float* __restrict__ gb1, float* __restrict__ gW2,
                    float* __restrict__ gb2, float* __restrict__ gW3, float* __restrict__ gb3) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int cb0 = blockIdx.y * CBW;
  const int ncb = min(CBW, C3 - cb0);

  for (int i = tid; i < CBW * 128; i += NT) (&s.dW3[0][0])[i] = 0.f;
  for (int i = tid; i < 128 * 64; i += NT) {
    float w = W2[i];
    s.W2s[i >> 6][i & 63] = w;
    s.W2t[i & 63][i >> 6] = w;
  }
  for (int i = tid; i < CBW; i += NT) s.db3[i] = 0.f;
  for (int i = tid; i < 64; i += NT) {
    s.W1s[i][0] = W1[i * 3]; s.W1s[i][1] = W1[i * 3 + 1]; s.W1s[i][2] = W1[i * 3 + 2]; s.W1s[i][3] = b1[i];
  }
  for (int i = tid; i < 128; i += NT) s.b2s[i] = b2[i];

  // persistent register accumulators
  // Column ownership is float4-interleaved everywhere (thread t7 = tid%8 owns columns t7*4 + 32*h + e):
  // 8 consecutive threads then read/write 128 contiguous bytes -> no shared-memory bank conflicts.
  // dW2 tile: rows k2 = 4*(tid/8) .. +3, cols k1 = 4*(tid%8) + 32*h + e  (h < 2, e < 4)
  const int r2 = 4 * (tid >> 3), t7 = tid & 7;
  float aW2[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) aW2[a][b] = 0.f;
  float aW1 = 0.f;   // threads < 192: dW1[k1 = tid/3][d = tid%3]
  float ab1 = 0.f;   // threads < 64
  float ab2 = 0.f;   // threads < 128
  __syncthreads();

  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    for (int ch0 = 0; ch0 < ncb; ch0 += IC) {
      // ---- instance setup
      if (tid < IC) {
        int c = ch0 + tid;
        float g = 0.f;
        int p = 0;
        if (c < ncb) {
          int64_t o = n * C3 + cb0 + c;
          g = out[o] > 0.f ? gout[o] : 0.f;
          p = argmax[o];
          p = min(max(p, 0), P - 1);
        }
        s.g[tid] = g;
        const float* pp = pts + (n * P + p) * 3;
        s.x[tid][0] = pp[0]; s.x[tid][1] = pp[1]; s.x[tid][2] = pp[2];
      }
      __syncthreads();
      // ---- h1 = relu(W1 x + b1): 32 x 64
      for (int i = tid; i < IC * 64; i += NT) {
        int r = i >> 6, k = i & 63;
        float v = fmaf(s.W1s[k][0], s.x[r][0], fmaf(s.W1s[k][1], s.x[r][1], fmaf(s.W1s[k][2], s.x[r][2], s.W1s[k][3])));
        s.h1[r][k] = v > 0.f ? v : 0.f;
      }
      __syncthreads();
      // ---- h2 = relu(h1 W2^T + b2): thread -> instance r = tid/8, k2 = 4*t7 + 32*j4 + e
      {
        const int r = tid >> 3;
        float acc[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float4 bv = *reinterpret_cast<const float4*>(&s.b2s[4 * t7 + 32 * j4]);
          acc[4 * j4 + 0] = bv.x; acc[4 * j4 + 1] = bv.y; acc[4 * j4 + 2] = bv.z; acc[4 * j4 + 3] = bv.w;
        }
        for (int k1 = 0; k1 < 64; ++k1) {
          float a = s.h1[r][k1];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 wv = *reinterpret_cast<const float4*>(&s.W2t[k1][4 * t7 + 32 * j4]);
            acc[4 * j4 + 0] = fmaf(a, wv.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(a, wv.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(a, wv.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(a, wv.w, acc[4 * j4 + 3]);
          }
        }
        const float g = s.g[r];
        const int c = ch0 + r;
        const bool live = c < ncb;
        const float* w3 = W3 + (int64_t)(cb0 + (live ? c : 0)) * 128;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int k2 = 4 * t7 + 32 * j4;
          float4 wv = *reinterpret_cast<const float4*>(w3 + k2);
          float h[4] = {acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]};
          const float wq[4] = {wv.x, wv.y, wv.z, wv.w};
          float dz[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = h[e] > 0.f ? h[e] : 0.f;
            dz[e] = (h[e] > 0.f && live) ? g * wq[e] : 0.f;
          }
          *reinterpret_cast<float4*>(&s.h2[r][k2]) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(&s.dz2[r][k2]) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        }
      }
      __syncthreads();
      // ---- dW3 block rows, db3
      for (int i = tid; i < IC * 128; i += NT) {
        int r = i >> 7, k = i & 127;
        if (ch0 + r < ncb) s.dW3[ch0 + r][k] = fmaf(s.g[r], s.h2[r][k], s.dW3[ch0 + r][k]);
      }
      if (tid < IC && ch0 + tid < ncb) s.db3[ch0 + tid] += s.g[tid];
      // ---- dW2 += dz2^T h1 (register tile), db2
#pragma unroll 4
      for (int r = 0; r < IC; ++r) {
        float4 a = *reinterpret_cast<const float4*>(&s.dz2[r][r2]);
        float4 b0 = *reinterpret_cast<const float4*>(&s.h1[r][4 * t7]);
        float4 b1v = *reinterpret_cast<const float4*>(&s.h1[r][4 * t7 + 32]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) aW2[i][j] = fmaf(av[i], bv[j], aW2[i][j]);
      }
      if (tid < 128) {
        float t = 0.f;
        for (int r = 0; r < IC; ++r) t += s.dz2[r][tid];
        ab2 += t;
      }
      // ---- dz1 = (h1 > 0) . (dz2 W2): thread -> instance r = tid/8, k1 = 4*t7 + 32*h + e
      {
        const int r = tid >> 3;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k2 = 0; k2 < 128; ++k2) {
          float a = s.dz2[r][k2];
          float4 w0 = *reinterpret_cast<const float4*>(&s.W2s[k2][4 * t7]);
          float4 w1 = *reinterpret_cast<const float4*>(&s.W2s[k2][4 * t7 + 32]);
          acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
          acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
          acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
          acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 hv = *reinterpret_cast<const float4*>(&s.h1[r][4 * t7 + 32 * h]);
          *reinterpret_cast<float4*>(&s.dz1[r][4 * t7 + 32 * h]) =
              make_float4(hv.x > 0.f ? acc[4 * h] : 0.f, hv.y > 0.f ? acc[4 * h + 1] : 0.f, hv.z > 0.f ? acc[4 * h + 2] : 0.f,
                          hv.w > 0.f ? acc[4 * h + 3] : 0.f);
        }
      }
      __syncthreads();
      // ---- dW1, db1
      if (tid < 192) {
        const int k1 = tid / 3, d = tid % 3;
        float t = 0.f;
        for (int r = 0; r < IC; ++r) t = fmaf(s.dz1[r][k1], s.x[r][d], t);
        aW1 += t;
      }
      if (tid < 64) {
        float t = 0.f;
        for (int r = 0; r < IC; ++r) t += s.dz1[r][tid];
        ab1 += t;
      }
      __syncthreads();
    }
  }
  // ---- flush
  for (int i = tid; i < ncb * 128; i += NT) {
    float v = (&s.dW3[0][0])[i];
    if (v != 0.f) atomicAdd(&gW3[(int64_t)cb0 * 128 + i], v);
  }
  for (int i = tid; i < ncb; i += NT) atomicAdd(&gb3[cb0 + i], s.db3[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&gW2[(r2 + i) * 64 + 4 * t7 + 32 * (j >> 2) + (j & 3)], aW2[i][j]);
  if (tid < 128) atomicAdd(&gb2[tid], ab2);
  if (tid < 192) atomicAdd(&gW1[tid], aW1);
  if (tid < 64) atomicAdd(&gb1[tid], ab1);
}

}  // namespace
}  // namespace sga

namespace sga {
int pointnet_bwd_tc(const float* pts, int64_t N, int P, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* W3, int C3, const float* out, const int32_t* argmax, const float* gout, float* gW1, float* gb1,
                    float* gW2, float* gb2, float* gW3, float* gb3, cudaStream_t st);
}

// mode: SGA_POINTNET_TC (tensor cores; C3 % 128 == 0) or SGA_POINTNET_SIMT (fp32 FMA, any shape)
extern "C" int sga_pointnet_bwd_mode(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                     const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                     const float* out, const int32_t* argmax, const float* grad_out, float* gW1,
                                     float* gb1, float* gW2, float* gb2, float* gW3, float* gb3, int mode, void* stream) {
  if (N <= 0) return SGA_OK;
  if (mode == SGA_POINTNET_TC) {
    SGA_REQUIRE(P >= 1 && C3 >= 128 && C3 % 128 == 0, "sga_pointnet_bwd(TC): P=%d C3=%d (C3 must be a multiple of 128)", P, C3);
    SGA_REQUIRE(out && argmax && grad_out, "sga_pointnet_bwd: forward results / grad_out missing");
    SGA_REQUIRE(((uintptr_t)W2 & 15) == 0 && ((uintptr_t)W3 & 15) == 0, "sga_pointnet_bwd(TC): W2/W3 must be 16-byte aligned");
    return sga::pointnet_bwd_tc(pts, N, P, W1, b1, W2, b2, W3, C3, out, argmax, grad_out, gW1, gb1, gW2, gb2, gW3, gb3,
                                (cudaStream_t)stream);
  }
  return sga_pointnet_bwd(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, grad_out, gW1, gb1, gW2, gb2, gW3, gb3, stream);
}

extern "C" int sga_pointnet_bwd(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                const float* out, const int32_t* argmax, const float* grad_out, float* gW1,
                                float* gb1, float* gW2, float* gb2, float* gW3, float* gb3, void* stream) {
  (void)b3;
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(P >= 1 && C3 >= 1, "sga_pointnet_bwd: P=%d C3=%d", P, C3);
  SGA_REQUIRE(out && argmax && grad_out, "sga_pointnet_bwd: forward results / grad_out missing");
  size_t smem = sizeof(sga::Smem);
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(sga::pointnet_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int nby = (C3 + sga::CBW - 1) / sga::CBW;
  int gx = sga::sm_count();   // one CTA per SM per channel block keeps the flush atomics bounded
  if ((int64_t)gx > N) gx = (int)N;
  dim3 grid(gx, nby);
  sga::pointnet_bwd_kernel<<<grid, sga::NT, smem, (cudaStream_t)stream>>>(pts, N, P, W1, b1, W2, b2, W3, C3, out, argmax, grad_out, gW1, gb1,
                                                                        gW2, gb2, gW3, gb3);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
