// BatchNorm batch statistics of the PointNet feature encoder from GRAM MATRICES on the tensor cores.
//
// The reference's BatchNorm1d layers (pointnet.py:141-142,154-155,158-159) have their outputs discarded but, in
// train(), still fold the batch statistics of the three pre-ReLU conv outputs z = W h + b into running_mean/var.
// Both moments of z follow from the moments of its INPUT h:
//     sum_p z_c  = w_c . (sum_p h) + n b_c
//     sum_p z_c^2 = w_c^T (sum_p h h^T) w_c + 2 b_c w_c . (sum_p h) + n b_c^2
// so instead of summing every conv2 / conv3 output element in the forward epilogues (ALU bound: it doubled the
// forward), this kernel accumulates G1 = sum_p h1 h1^T [64x64] and G2 = sum_p h2 h2^T [128x128] (plus the first
// moments, through a ones column) with tcgen05 MMAs that read the activation tiles MN-major -- the same bytes the
// conv2 MMA reads K-major -- and a small finalize kernel applies the weights in fp64.  conv1 is affine in the
// point, so its statistics come from the 9 point moments as before.
//
// Precision: h1 / h2 enter the Grams as single bf16 (round-to-nearest): the rounding errors are independent per
// point and average out over the N*P >= 10^5 points of a batch (relative error of a Gram entry ~ 2^-9 / sqrt(n));
// W2 is a FIXED operand, so conv2 keeps its hi/lo split (two passes) -- a rounded W2 would bias every h2.
// Tensor work per 128-point tile: 256 (G1) + 512 (conv2) + 768 (G2) cycles vs 3840 of the forward.
#include "common.cuh"
#include "ptx.cuh"

namespace sga {
namespace {

constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;
constexpr uint32_t kBlk = 16384;
constexpr int kTile = 128;

// shared-memory map (A1 / H2 / the point table are double-buffered: conv1 of tile t+1 runs under the MMAs of tile t)
constexpr uint32_t W2HI = 0;                 // [128 k2 rows][64 k1] K-major SW128: B of conv2
constexpr uint32_t W2LO = W2HI + kBlk;
constexpr uint32_t A1 = W2LO + kBlk;         // 2 x { [128 pts][64 ch] bf16: A of conv2 (K-major), both operands of G1 (MN-major);
constexpr uint32_t A1_STRIDE = 2 * kBlk;     //       E1: second MN atom of G1's A operand, channel 0 = 1 (first moments) }
constexpr uint32_t H2 = A1 + 2 * A1_STRIDE;  // 2 x { 2 blocks [128 pts][64 ch] bf16: both operands of G2 (MN-major);
constexpr uint32_t H2_STRIDE = 3 * kBlk;     //       E2: third MN atom of G2's B operand, the ones column }
constexpr uint32_t SMALL = H2 + 2 * H2_STRIDE;   // 196608
constexpr uint32_t W1B1 = SMALL;             // float4[64] = {w0,w1,w2,b}
constexpr uint32_t B2 = W1B1 + 1024;         // float[128]
constexpr uint32_t XS = B2 + 512;            // 2 x float4[128] = {x,y,z,valid}
constexpr uint32_t BARS = XS + 4096;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;

constexpr uint32_t D1_COL = 0;     // [128 x 64]  rows 0..63 = G1, row 64 = sum h1
constexpr uint32_t D2_COL = 64;    // 2 x [128 pts x 128 ch] conv2 accumulators
constexpr uint32_t D3_COL = 320;   // [128 x 192] cols 0..127 = G2, col 128 = sum h2
constexpr int kTmemCols = 512;
constexpr int kPartial = 128 * 64 + 128 * 192;      // floats one CTA writes

enum { BAR_A1_FULL = 0, BAR_D2_FULL = 2, BAR_H2_FULL = 4, BAR_G2_DONE = 6 };     // two of each (buffer t & 1)

__device__ __forceinline__ uint32_t pack2(float a, float b) {   // a -> low half (lower channel)
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_chunk(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// MN-major view of [rows = K index][64 x bf16 = 128 B] SWIZZLE_128B tiles (pointnet_bwd_tc.cu): 64-element MN atoms
// `lbo` bytes apart, groups of 8 K rows 1024 B apart; one K = 16 step = 2048 B.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
pointnet_gram_kernel(const float* __restrict__ pts, int64_t N, int P, const float* __restrict__ W1,
                     const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                     float* __restrict__ partial, double* __restrict__ raw_pts) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  const float4* w1b1 = reinterpret_cast<const float4*>(sm + W1B1);
  const float* b2s = reinterpret_cast<const float*>(sm + B2);
  float4* xs = reinterpret_cast<float4*>(sm + XS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---------------- one-time setup
  for (int i = tid; i < 128 * 8; i += kThreads) {          // W2 [128 k2][64 k1] hi/lo
    const int r = i >> 3, j = i & 7;
    const float4* src = reinterpret_cast<const float4*>(W2 + r * 64 + j * 8);
    const float4 a = src[0], b = src[1];
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = pack2(f[2 * e], f[2 * e + 1]);
      l[e] = pack2(f[2 * e] - __uint_as_float(h[e] << 16), f[2 * e + 1] - __uint_as_float(h[e] & 0xFFFF0000u));
    }
    const uint32_t off = ptx::sw128_offset(r, j);
    st_chunk(sm_base + W2HI + off, make_uint4(h[0], h[1], h[2], h[3]));
    st_chunk(sm_base + W2LO + off, make_uint4(l[0], l[1], l[2], l[3]));
  }
  for (int i = tid; i < 128 * 8; i += kThreads) {          // ones columns: channel 0 of every row = bf16 1.0
    const int r = i >> 3, j = i & 7;
    const uint4 v = make_uint4(j == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);
    const uint32_t off = ptx::sw128_offset(r, j);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      st_chunk(sm_base + A1 + b * A1_STRIDE + kBlk + off, v);
      st_chunk(sm_base + H2 + b * H2_STRIDE + 2 * kBlk + off, v);
    }
  }
  for (int i = tid; i < 64; i += kThreads)
    reinterpret_cast<float4*>(sm + W1B1)[i] = make_float4(W1[i * 3], W1[i * 3 + 1], W1[i * 3 + 2], b1[i]);
  for (int i = tid; i < 128; i += kThreads) reinterpret_cast<float*>(sm + B2)[i] = b2[i];
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars[BAR_A1_FULL + b], kComputeThreads);
      ptx::mbar_init(&bars[BAR_D2_FULL + b], 1);
      ptx::mbar_init(&bars[BAR_H2_FULL + b], kComputeThreads);
      ptx::mbar_init(&bars[BAR_G2_DONE + b], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int ntile = (P + kTile - 1) / kTile;
  const int64_t nobj = (N > (int64_t)blockIdx.x) ? (N - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t G = nobj * ntile;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    // issue order: G1 + conv2 of tile t+1 go out BEFORE the wait for H2 of tile t, so they run under its epilogue
    const uint32_t idesc2 = ptx::make_idesc(1, 128, 128);
    const uint32_t idesc_g1 = ptx::make_idesc(1, 128, 64) | (1u << 15) | (1u << 16);    // both operands MN-major
    const uint32_t idesc_g2 = ptx::make_idesc(1, 128, 192) | (1u << 15) | (1u << 16);
    const uint64_t dW2hi = ptx::smem_desc_sw128(sm_base + W2HI), dW2lo = ptx::smem_desc_sw128(sm_base + W2LO);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    auto issue_front = [&](int64_t t) {        // G1 += A1^T [A1 | 1];  D2[b] = A1 W2^T (W2 hi, lo)
      const int b = (int)(t & 1);
      ptx::mbar_wait(&bars[BAR_A1_FULL + b], (uint32_t)((t >> 1) & 1));
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t a1 = sm_base + A1 + b * A1_STRIDE;
        const uint64_t dA1 = ptx::smem_desc_sw128(a1);
        const uint64_t mA1 = desc_mn_sw128(a1, kBlk);          // atoms: A1, E1 (A of G1) / A1 alone (B of G1, N = 64)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::umma_bf16(tmem_u + D1_COL, mA1 + (uint64_t)(ks * 128), mA1 + (uint64_t)(ks * 128), idesc_g1, (t | ks) != 0);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const uint64_t bb = pass ? dW2lo : dW2hi;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16(tmem_u + D2_COL + b * 128, dA1 + (uint64_t)(ks * 2), bb + (uint64_t)(ks * 2), idesc2, (pass | ks) != 0);
        }
        ptx::umma_commit(&bars[BAR_D2_FULL + b]);
      }
      __syncwarp();
    };
    if (G > 0) issue_front(0);
    for (int64_t t = 0; t < G; ++t) {
      const int b = (int)(t & 1);
      if (t + 1 < G) issue_front(t + 1);
      ptx::mbar_wait(&bars[BAR_H2_FULL + b], (uint32_t)((t >> 1) & 1));
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        // G2[ch x (ch | 1)] += H2^T [H2 | 1]
        const uint64_t mH2 = desc_mn_sw128(sm_base + H2 + b * H2_STRIDE, kBlk);   // atoms: H2 block 0, block 1 (A), + E2 (B, N = 192)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::umma_bf16(tmem_u + D3_COL, mH2 + (uint64_t)(ks * 128), mH2 + (uint64_t)(ks * 128), idesc_g2, (t | ks) != 0);
        ptx::umma_commit(&bars[BAR_G2_DONE + b]);
      }
      __syncwarp();
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, wh = warp >> 2;
    const int row = 32 * q + lane;                       // tile row = TMEM lane
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int pg = tid & 31, cg = tid >> 5;              // conv1 mapping: 8 channels (cg) x 4 rows (pg + 32 i)
    double pm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};          // threads < 128: sum x,y,z, xx,xy,xz,yy,yz,zz
    // point table + conv1 of tile t into buffer t & 1 (A1[b] is free: this thread has waited for D2_FULL of tile t-2)
    auto front = [&](int64_t t) -> bool {
      const int b = (int)(t & 1);
      float4* xb = xs + 128 * b;
      const int64_t n = blockIdx.x + (t / ntile) * (int64_t)gridDim.x;
      const int p0 = (int)(t % ntile) * kTile;
      if (tid < 128) {
        const int p = p0 + tid;
        const bool ok = p < P;
        float x = 0.f, y = 0.f, z = 0.f;
        if (ok) {
          const float* src = pts + (n * P + p) * 3;
          x = src[0]; y = src[1]; z = src[2];
          pm[0] += x; pm[1] += y; pm[2] += z;
          pm[3] += (double)x * x; pm[4] += (double)x * y; pm[5] += (double)x * z;
          pm[6] += (double)y * y; pm[7] += (double)y * z; pm[8] += (double)z * z;
        }
        xb[tid] = make_float4(x, y, z, ok ? 1.f : 0.f);
      }
      compute_barrier();
      // conv1 -> A1 (single bf16; rows beyond P are zero so that they drop out of every sum)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 xr = xb[pg + 32 * i];
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 w = w1b1[8 * cg + e];
          const float v = fmaf(w.x, xr.x, fmaf(w.y, xr.y, fmaf(w.z, xr.z, w.w)));
          f[e] = (v > 0.f && xr.w != 0.f) ? v : 0.f;
        }
        st_chunk(sm_base + A1 + b * A1_STRIDE + ptx::sw128_offset(pg + 32 * i, cg),
                 make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7])));
      }
      const bool ok_row = xb[row].w != 0.f;
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_A1_FULL + b]);
      return ok_row;
    };
    bool row_ok = G > 0 ? front(0) : false;
    for (int64_t t = 0; t < G; ++t) {
      const int b = (int)(t & 1);
      const uint32_t ph = (uint32_t)((t >> 1) & 1);
      const bool row_ok_next = (t + 1 < G) ? front(t + 1) : false;     // runs under the MMAs of tile t

      // ---- E2: h2 = relu(D2[b] + b2) -> H2[b] (single bf16), this thread's 64 channels [64 wh, +64)
      ptx::mbar_wait(&bars[BAR_D2_FULL + b], ph);
      ptx::tc_fence_after();
      uint32_t v[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld16(tmem + lane_addr + D2_COL + b * 128 + 64 * wh + 16 * c, v[c]);
      ptx::tmem_ld_wait();
      if (t >= 2) {                                      // G2 of tile t-2 has finished reading H2[b]
        ptx::mbar_wait(&bars[BAR_G2_DONE + b], ph ^ 1u);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float z = __uint_as_float(v[c][e]) + b2s[64 * wh + 16 * c + e];
          f[e] = (z > 0.f && row_ok) ? z : 0.f;
        }
        const uint32_t base = sm_base + H2 + b * H2_STRIDE + (uint32_t)wh * kBlk;
        st_chunk(base + ptx::sw128_offset(row, 2 * c),
                 make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7])));
        st_chunk(base + ptx::sw128_offset(row, 2 * c + 1),
                 make_uint4(pack2(f[8], f[9]), pack2(f[10], f[11]), pack2(f[12], f[13]), pack2(f[14], f[15])));
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_H2_FULL + b]);
      row_ok = row_ok_next;
    }
    // ---- drain: the accumulators of this CTA -> its slot of `partial`
    if (G > 0) {
      // the commit of the last tile covers every MMA issued before it (G1 included)
      ptx::mbar_wait(&bars[BAR_G2_DONE + (int)((G - 1) & 1)], (uint32_t)(((G - 1) >> 1) & 1));
      ptx::tc_fence_after();
      float* out = partial + (size_t)blockIdx.x * kPartial;
      if (wh == 0) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t u[16];
          ptx::tmem_ld16(tmem + lane_addr + D1_COL + 16 * c, u);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) out[row * 64 + 16 * c + e] = __uint_as_float(u[e]);
        }
      }
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {                      // D3: 192 columns, 96 per half
        uint32_t u[16];
        ptx::tmem_ld16(tmem + lane_addr + D3_COL + 96 * wh + 16 * c, u);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) out[128 * 64 + row * 192 + 96 * wh + 16 * c + e] = __uint_as_float(u[e]);
      }
    } else {
      float* out = partial + (size_t)blockIdx.x * kPartial;
      for (int i = tid; i < kPartial; i += kComputeThreads) out[i] = 0.f;
    }
    if (tid < 128) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double s = warp_sum_d(pm[k]);
        if (lane == 0 && s != 0.0) atomicAdd(&raw_pts[k], s);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<kTmemCols>(tmem);
}

// partial[nCTA][kPartial] -> red[kPartial] (fp64)
__global__ void gram_reduce_kernel(const float* __restrict__ partial, int ncta, double* __restrict__ red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kPartial) return;
  double s = 0.0;
  for (int c = 0; c < ncta; ++c) s += (double)partial[(size_t)c * kPartial + i];
  red[i] = s;
}

// one block per output channel: raw sums S = w.m, Q = w^T G w, then the same bias algebra as
// pointnet_moments_finalize_kernel; blocks [0,64): conv1 (point moments), [64,192): conv2, [192,192+C3): conv3
__global__ void __launch_bounds__(128)
gram_moments_kernel(const double* __restrict__ red, const double* __restrict__ raw_pts, const float* __restrict__ W1,
                    const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                    const float* __restrict__ W3, const float* __restrict__ b3, int C3, double n,
                    double* __restrict__ moments) {
  __shared__ double sh[128];
  const int blk = blockIdx.x, t = threadIdx.x;
  if (blk < 64) {
    if (t != 0) return;
    const int i = blk;
    const double wx = W1[i * 3], wy = W1[i * 3 + 1], wz = W1[i * 3 + 2], b = b1[i];
    const double* m = raw_pts;
    const double S = wx * m[0] + wy * m[1] + wz * m[2];
    const double Q = wx * wx * m[3] + wy * wy * m[6] + wz * wz * m[8] + 2.0 * (wx * wy * m[4] + wx * wz * m[5] + wy * wz * m[7]);
    moments[i] += S + n * b;
    moments[64 + i] += Q + 2.0 * b * S + n * b * b;
    return;
  }
  const bool l2 = blk < 192;
  const int c = l2 ? blk - 64 : blk - 192;
  const int K = l2 ? 64 : 128;
  const float* w = l2 ? W2 + c * 64 : W3 + (int64_t)c * 128;
  const double* Gm = l2 ? red : red + 128 * 64;            // row stride: 64 (D1) / 192 (D3)
  const int ld = l2 ? 64 : 192;
  double sp = 0.0, qp = 0.0;
  if (t < K) {
    const double wi = w[t];
    // first moments: D1 row 64 (ones row) / D3 column 128 (ones column)
    const double mi = l2 ? red[64 * 64 + t] : red[128 * 64 + t * 192 + 128];
    sp = wi * mi;
    double acc = 0.0;
    for (int j = 0; j < K; ++j) acc += Gm[t * ld + j] * (double)w[j];
    qp = wi * acc;
  }
  sh[t] = sp;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  const double S = sh[0];
  __syncthreads();
  sh[t] = qp;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  const double Q = sh[0];
  if (t == 0) {
    const double b = l2 ? b2[c] : b3[c];
    const int o1 = l2 ? 128 : 384, o2 = l2 ? 256 : 384 + C3;
    moments[o1 + c] += S + n * b;
    moments[o2 + c] += Q + 2.0 * b * S + n * b * b;
  }
}

}  // namespace

size_t pointnet_gram_scratch_bytes() {
  // per-CTA partial accumulators (fp32) + their fp64 reduction + the 9 point moments
  return (size_t)sm_count() * kPartial * sizeof(float) + (size_t)kPartial * sizeof(double) + 16 * sizeof(double);
}

int pointnet_gram_moments(const float* pts, int64_t N, int P, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* W3, const float* b3, int C3, double* moments, void* scratch,
                          cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pointnet_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_done = true;
  }
  int grid = persistent_ctas();
  if ((int64_t)grid > N) grid = (int)N;
  unsigned char* ws = (unsigned char*)scratch;
  float* partial = (float*)ws;
  double* red = (double*)(ws + (size_t)sm_count() * kPartial * sizeof(float));
  double* raw_pts = red + kPartial;
  SGA_CUDA(cudaMemsetAsync(raw_pts, 0, 16 * sizeof(double), st));
  pointnet_gram_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(pts, N, P, W1, b1, W2, b2, partial, raw_pts);
  SGA_LAUNCH_CHECK();
  gram_reduce_kernel<<<(kPartial + 255) / 256, 256, 0, st>>>(partial, grid, red);
  SGA_LAUNCH_CHECK();
  gram_moments_kernel<<<192 + C3, 128, 0, st>>>(red, raw_pts, W1, b1, W2, b2, W3, b3, C3, (double)N * (double)P, moments);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga

extern "C" size_t sga_pointnet_gram_scratch_bytes(void) { return sga::pointnet_gram_scratch_bytes(); }

extern "C" int sga_pointnet_bn_moments_gram(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                            const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                            double* moments, void* scratch, size_t scratch_bytes, void* stream) {
  SGA_REQUIRE(pts && W1 && b1 && W2 && b2 && W3 && b3 && moments && scratch, "sga_pointnet_bn_moments_gram: null pointer");
  SGA_REQUIRE(N >= 1 && P >= 1 && C3 >= 1, "sga_pointnet_bn_moments_gram: bad sizes");
  SGA_REQUIRE(((uintptr_t)W2 & 15) == 0 && ((uintptr_t)scratch & 15) == 0, "sga_pointnet_bn_moments_gram: W2 / scratch must be 16-byte aligned");
  if (scratch_bytes < sga::pointnet_gram_scratch_bytes()) {
    sga::set_error("sga_pointnet_bn_moments_gram: scratch of %zu bytes, need %zu", scratch_bytes, sga::pointnet_gram_scratch_bytes());
    return SGA_EWORKSPACE;
  }
  return sga::pointnet_gram_moments(pts, N, P, W1, b1, W2, b2, W3, b3, C3, moments, scratch, (cudaStream_t)stream);
}
