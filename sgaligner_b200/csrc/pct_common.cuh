// Shared pieces of the NaivePCT kernels (pct_pw.cu, pct_attn.cu, pct_cat.cu): bf16 hi/lo operand splitting, the
// shared-memory tile images tcgen05 reads, the coalescing output stage.
//
// Activations live in HBM as fp32 [N, P, C] ("point-major": one row of C channels per point).  A tile of 128 points
// x 64 channels becomes one 16 KiB block of 128-byte rows (64 bf16) with the 128B swizzle; the SAME bytes are read
//   K-major  (rows = M or N index, the 64 channels = K)          by the convolutions            and
//   MN-major (rows = K index, the 64 channels = M/N, lbo = 16 KiB between the channel blocks) by x_v * attention.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace sga {
namespace pct {

constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;       // + the MMA-issuing warp
constexpr int kTile = 128;                           // points per tile
constexpr uint32_t kBlk = 16384;                     // [128 rows x 128 B]

// Split-operand format of every NaivePCT contraction: FP16 pairs.  x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// carries 22 significant bits (|x| >= 2^-3; below that the lo part is subnormal: absolute error <= 2^-25), against 16
// for a bf16 pair -- and the self-attention needs them: the energies k_i . k_j / sqrt(32) sit in an exponent, so an
// operand error of 2^-17 (bf16 pair) on energies of 10^3..10^4 is an O(1) error of attention weights (measured on the
// golden input: 3.3e-4 on the encoder output with bf16 pairs, tools/pct_numerics_emul.py; 1.4e-6 emulated with fp16
// pairs at the same three tensor passes).  Price: the fp16 range -- an activation beyond 65504 becomes inf and the
// output NaN (loud), where the fp32 reference carries on; BatchNorm after every convolution keeps real activations
// orders of magnitude below that.
constexpr int kFmt = 0;                               // tcgen05 instruction descriptor a_format / b_format: 0 = F16, 1 = BF16

__device__ __forceinline__ uint32_t pack2(float a, float b) {   // a -> low half (lower k)
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }

// 8 fp32 -> fp16 hi / lo parts as two 16-byte chunks (x = hi + lo to ~2^-22 relative)
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack2(f[2 * i], f[2 * i + 1]);
    l[i] = pack2(f[2 * i] - bf_lo(h[i]), f[2 * i + 1] - bf_hi(h[i]));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// ---- format-templated variants (kF = 0: fp16 pairs as above; kF = 1: bf16 pairs, kept for experiments).
// The BACKWARD kernels also use FP16 pairs: the chain through a peaked softmax cancels three digits (sum_j dS[i,j] = 0
// against |k| ~ 10^2), so an operand error of 2^-17 (bf16 pair) shows up as 1e-2 on d k where fp32 autograd has 5e-5.
// Gradients live far below the fp16 range, so every gradient operand is multiplied by a per-object power of two on the
// way in (sga_pct_pow2_scale: the object's largest element lands near 2^12) and the result divided by it on the way out
// -- exact operations; what is left below 2^-3 after scaling has an absolute error of 2^-25, i.e. 1e-11 of the largest
// element of the same object.
template <int kF>
__device__ __forceinline__ uint32_t pack2f(float a, float b) {
  if (kF == 0) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
template <int kF>
__device__ __forceinline__ float unpack_lo(uint32_t w) {
  if (kF == 0) return __low2float(*reinterpret_cast<const __half2*>(&w));
  return __uint_as_float(w << 16);
}
template <int kF>
__device__ __forceinline__ float unpack_hi(uint32_t w) {
  if (kF == 0) return __high2float(*reinterpret_cast<const __half2*>(&w));
  return __uint_as_float(w & 0xFFFF0000u);
}
template <int kF>
__device__ __forceinline__ void split2f(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack2f<kF>(a, b);
  lo = pack2f<kF>(a - unpack_lo<kF>(hi), b - unpack_hi<kF>(hi));
}
template <int kF>
__device__ __forceinline__ void split8f(const float (&f)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2f<kF>(f[2 * i], f[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_chunk(uint32_t smem_addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// MN-major view of [rows = K index][64 x bf16] SWIZZLE_128B blocks: 64-element MN atoms `lbo` bytes apart, groups of
// 8 K rows 1024 B apart; one K = 16 step advances the start address by 2048 B (128 descriptor units).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- coalescing output stage.  After tcgen05.ld a thread holds 16 consecutive columns of ITS row (32 rows per warp):
// written straight to a row-major tensor that is 32 scattered 64-byte pieces per instruction.  Through a private
// [32][17] shared-memory patch the warp writes two rows x 64 contiguous bytes per instruction instead, and a lane
// sees all 32 rows of one column on the way (per-channel BatchNorm sums for free).
constexpr int kStageLd = 17;
constexpr int kStageFloats = 32 * kStageLd;       // per warp

// v: the thread's 16 values of row (row0 + lane); dst: row-major output, ld floats per row, column c0;
// nvalid: rows row0 .. row0+nvalid-1 are stored / counted.  s, q: this lane's running sum / sum of squares of
// column (lane & 15) over the rows 2*it + (lane >> 4).
__device__ __forceinline__ void stage_store16(float* stage, const float (&v)[16], float* __restrict__ dst, int64_t ld,
                                              int nvalid, int lane, float& s, float& q, bool want_stats) {
#pragma unroll
  for (int c = 0; c < 16; ++c) stage[lane * kStageLd + c] = v[c];
  __syncwarp();
  const int col = lane & 15, rh = lane >> 4;
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int r = 2 * it + rh;
    const float x = stage[r * kStageLd + col];
    if (r < nvalid) {
      dst[(int64_t)r * ld + col] = x;
      if (want_stats) {
        s += x;
        q = fmaf(x, x, q);
      }
    }
  }
  __syncwarp();
}

}  // namespace pct
}  // namespace sga
