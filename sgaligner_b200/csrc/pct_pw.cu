// NaivePCT: the 128-input-channel pointwise convolutions on the tensor cores (reference:
// src/aligner/networks/pct.py -- Embedding.conv2 :107, SA.k_conv/v_conv :197-200, SA.trans_conv :202), one launch
// per BatchNorm segment.  Y[pts, Cout] = X[pts, 128] W^T (+ bias), where X is produced on the fly by a PROLOGUE:
//     X = g1(src1) + g2(src2),   g(s) = s  or  relu(a_c s + b_c)          (BatchNorm folded into a_c, b_c)
// which covers  x0 = relu(bn2(z2)),  x_l = x_{l-1} + relu(after_norm(t_l))  (the SA residual, pct.py:228-230) and the
// plain x_s -> trans_conv input; the EMBED variant computes X = relu(bn1(conv1(point))) from the raw points
// (pct.py:122).  X itself can be written out (x_l is needed again by the next layer and by the concat).
// EPILOGUE: + bias, fp32 rows to one or two row-major tensors (k | v share one launch: Cout = 32 + 128), per-channel
// sum / sum of squares of what was stored (the batch statistics of the NEXT BatchNorm) as fp64 atomics.
//
// bf16x3 split operands (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM): ~1e-5, where one bf16 / tf32 pass is not
// inside the 1e-4 gate.  Per tile of 128 points: 8 compute warps load + prologue + split -> A tile (128B-swizzled
// K-major smem image), one elected thread issues 24 tcgen05.mma (M = 128 points, N = Cout, K = 16), the same 8 warps
// drain the accumulator through a coalescing stage.  These convolutions are HBM bound (0.13-0.27 MB in+out per tile
// against 0.5 us of MMAs), so the tile pipeline is kept simple: what matters is bytes, and those are minimal.
#include "pct_common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace sga {
namespace pct {
namespace {

constexpr int kMaxCout = 160;
constexpr uint32_t WBLK = kMaxCout * 128;                    // one 64-channel K block of the weights (<= 160 rows)
constexpr uint32_t WHI = 0;
constexpr uint32_t WLO = WHI + 2 * WBLK;
constexpr uint32_t AHI = WLO + 2 * WBLK;                     // 81920
constexpr uint32_t ALO = AHI + 2 * kBlk;
constexpr uint32_t STAGE = ALO + 2 * kBlk;                   // 147456
constexpr uint32_t BIAS = STAGE + 8 * kStageFloats * 4;      // + 17408
constexpr uint32_t BARS = BIAS + kMaxCout * 4;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_USED = TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;
constexpr int kTmemCols = 256;

enum { BAR_A_FULL = 0, BAR_D_FULL = 1 };

struct PwArgs {
  const float* src1; const float* a1; const float* b1; int mode1;    // 0 absent, 1 identity, 2 relu(a s + b)
  const float* src2; const float* a2; const float* b2; int mode2;
  const float* pts; const float* w1;                                  // EMBED: points [N,P,3], conv1 weight [128,3] (a1, b1 = folded bn1)
  int64_t N; int P;
  const float* W; const float* bias; int Cout; int c0;
  float* out_x; float* out0; float* out1; double* stats;
  const float* scale;                                                 // backward: [N,2] per-object {s, 1/s}: X is multiplied by s on the way in, Y by 1/s on the way out
  float* absmax;                                                      // or NULL: [N] (zeroed) per-object max |main output| (second-generation kernel only)
};

template <bool kEmbed, bool kScaled>
__global__ void __launch_bounds__(kThreads, 1) pct_pw_kernel(const PwArgs A) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  float* bias_s = reinterpret_cast<float*>(sm + BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cout = A.Cout;

  // ---- one-time setup: weights -> bf16 hi/lo K-major images, bias, barriers, TMEM
  for (int i = tid; i < Cout * 16; i += kThreads) {
    const int r = i >> 4, j = i & 15;
    const float4* src = reinterpret_cast<const float4*>(A.W + (int64_t)r * 128 + j * 8);
    const float4 x = src[0], y = src[1];
    const float f[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    uint4 hi, lo;
    split8(f, hi, lo);
    const uint32_t off = (uint32_t)(j >> 3) * WBLK + ptx::sw128_offset(r, j & 7);
    st_chunk(sm_base + WHI + off, hi);
    st_chunk(sm_base + WLO + off, lo);
  }
  for (int i = tid; i < Cout; i += kThreads) bias_s[i] = A.bias ? A.bias[i] : 0.f;
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_A_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_D_FULL], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int T = (A.P + kTile - 1) / kTile;
  const int64_t G = A.N * T;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = ptx::make_idesc(kFmt, 128, Cout);
    const uint64_t dAhi = ptx::smem_desc_sw128(sm_base + AHI), dAlo = ptx::smem_desc_sw128(sm_base + ALO);
    const uint64_t dWhi = ptx::smem_desc_sw128(sm_base + WHI), dWlo = ptx::smem_desc_sw128(sm_base + WLO);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t it = 0;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x, ++it) {
      ptx::mbar_wait(&bars[BAR_A_FULL], it & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        // all FOUR partial products (lo*lo included): these convolutions produce k, and at attention energies of 10^3..10^4
        // an error of 4e-7 |k| (three products) is a 0.4 % error of attention weights -- measured on the parameter
        // gradients at 512 x 512 (tests/test_gpu_pct_backward.py); the tensor pipe is 10 % busy here, the pass is free
        for (int pass = 0; pass < 4; ++pass) {
          const uint64_t ab = (pass & 1) ? dAlo : dAhi;
          const uint64_t bb = (pass & 2) ? dWlo : dWhi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ao = (uint64_t)((ks >> 2) * (kBlk >> 4) + (ks & 3) * 2);
            const uint64_t bo = (uint64_t)((ks >> 2) * (WBLK >> 4) + (ks & 3) * 2);
            ptx::umma_bf16(tmem_u, ab + ao, bb + bo, idesc, (pass | ks) != 0);
          }
        }
        ptx::umma_commit(&bars[BAR_D_FULL]);
      }
      __syncwarp();
    }
  } else {
    // =============================== compute warps ===============================
    const int cc = tid & 15, r0 = tid >> 4;          // loader: 8-channel chunk cc of rows r0 + 16 q
    const int ch0 = cc * 8;
    // per-channel prologue parameters of this thread's 8 channels, resident in registers for the whole kernel
    // (EMBED: the folded conv1 rows; otherwise the two BatchNorm affine pairs)
    float pa1[8], pb1[8], pa2[8], pb2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (kEmbed) {      // bn1(conv1(p)) = (a w) . p + b: the BatchNorm scale folded into the three conv1 weights of the channel
        const float a = A.a1[ch0 + e];
        pa1[e] = a * A.w1[(ch0 + e) * 3];
        pb1[e] = a * A.w1[(ch0 + e) * 3 + 1];
        pa2[e] = a * A.w1[(ch0 + e) * 3 + 2];
        pb2[e] = A.b1[ch0 + e];
      } else {
        pa1[e] = (A.mode1 == 2) ? A.a1[ch0 + e] : 1.f;
        pb1[e] = (A.mode1 == 2) ? A.b1[ch0 + e] : 0.f;
        pa2[e] = (A.mode2 == 2) ? A.a2[ch0 + e] : 1.f;
        pb2[e] = (A.mode2 == 2) ? A.b2[ch0 + e] : 0.f;
      }
    }
    const uint32_t a_off_blk = (uint32_t)(cc >> 3) * kBlk;
    const int q = warp & 3, hc = warp >> 2;           // epilogue: TMEM lane quarter, column half
    const int ncol_half = Cout >> 1, nch = ncol_half >> 4;
    float* stage = reinterpret_cast<float*>(sm + STAGE) + warp * kStageFloats;
    double ds[5] = {0, 0, 0, 0, 0}, dq[5] = {0, 0, 0, 0, 0};
    const bool want_stats = A.stats != nullptr;

    uint32_t it = 0;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x, ++it) {
      const int64_t n = g / T;
      const int t = (int)(g - n * T);
      const int64_t rowbase = n * A.P + (int64_t)t * kTile;
      const int valid = min(kTile, A.P - t * kTile);
      // ---- load + prologue + split -> A tile (the previous tile's MMAs have completed: D_FULL was waited for)
      // The asm shared-memory stores are ordering barriers for the compiler, so the loads of a batch are issued before
      // its first store: a batch = 8 rows of one source (16 loads of 16 bytes in flight per thread), or 4 rows of two.
      auto batch = [&](auto rows_tag, int bt) {
        constexpr int kRows = decltype(rows_tag)::value;
        float4 u[kRows][2], w[kRows][2];
        float3 p3[kRows];
#pragma unroll
        for (int qq = 0; qq < kRows; ++qq) {
          const int row = r0 + 16 * (bt * kRows + qq);
          const bool ok = row < valid;
          if (kEmbed) {
            const float* pp = A.pts + (rowbase + row) * 3;
            p3[qq] = ok ? make_float3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)) : make_float3(0.f, 0.f, 0.f);
          } else {
            const float4* s1 = reinterpret_cast<const float4*>(A.src1 + (rowbase + row) * 128 + ch0);
            u[qq][0] = ok ? __ldg(s1) : make_float4(0.f, 0.f, 0.f, 0.f);
            u[qq][1] = ok ? __ldg(s1 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (A.mode2) {
              const float4* s2 = reinterpret_cast<const float4*>(A.src2 + (rowbase + row) * 128 + ch0);
              w[qq][0] = ok ? __ldg(s2) : make_float4(0.f, 0.f, 0.f, 0.f);
              w[qq][1] = ok ? __ldg(s2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
#pragma unroll
        for (int qq = 0; qq < kRows; ++qq) {
          const int row = r0 + 16 * (bt * kRows + qq);
          const bool ok = row < valid;
          float f[8];
          if (kEmbed) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float v = fmaf(pa1[e], p3[qq].x, fmaf(pb1[e], p3[qq].y, fmaf(pa2[e], p3[qq].z, pb2[e])));
              f[e] = v > 0.f ? v : 0.f;
            }
          } else {
            const float s1[8] = {u[qq][0].x, u[qq][0].y, u[qq][0].z, u[qq][0].w, u[qq][1].x, u[qq][1].y, u[qq][1].z, u[qq][1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = s1[e];
              if (A.mode1 == 2) {
                v = fmaf(pa1[e], v, pb1[e]);
                v = v > 0.f ? v : 0.f;
              }
              f[e] = v;
            }
            if (A.mode2) {
              const float s2[8] = {w[qq][0].x, w[qq][0].y, w[qq][0].z, w[qq][0].w, w[qq][1].x, w[qq][1].y, w[qq][1].z, w[qq][1].w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float v = s2[e];
                if (A.mode2 == 2) {
                  v = fmaf(pa2[e], v, pb2[e]);
                  v = v > 0.f ? v : 0.f;
                }
                f[e] += v;
              }
            }
          }
          if (!ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = 0.f;
          }
          if (kScaled) {
            const float sc = __ldg(A.scale + 2 * n);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= sc;
          }
          if (!kEmbed && A.out_x && ok) {
            float4* ox = reinterpret_cast<float4*>(A.out_x + (rowbase + row) * 128 + ch0);
            ox[0] = make_float4(f[0], f[1], f[2], f[3]);
            ox[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
          uint4 hi, lo;
          split8(f, hi, lo);
          const uint32_t off = a_off_blk + ptx::sw128_offset(row, cc & 7);
          st_chunk(sm_base + AHI + off, hi);
          st_chunk(sm_base + ALO + off, lo);
        }
      };
      if (!kEmbed && A.mode2) {
        batch(std::integral_constant<int, 4>{}, 0);
        batch(std::integral_constant<int, 4>{}, 1);
      } else {
        batch(std::integral_constant<int, 8>{}, 0);
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_A_FULL]);

      // ---- epilogue: accumulator -> (+ bias) -> coalesced rows, BatchNorm sums
      ptx::mbar_wait(&bars[BAR_D_FULL], it & 1);
      ptx::tc_fence_after();
      const int nvalid = max(0, min(32, valid - 32 * q));
#pragma unroll
      for (int ch = 0; ch < 5; ++ch) {
        if (ch >= nch) break;
        const int col0 = hc * ncol_half + ch * 16;
        uint32_t v[16];
        ptx::tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
        ptx::tmem_ld_wait();
        float f[16];
        const float osc = kScaled ? __ldg(A.scale + 2 * n + 1) : 1.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = kScaled ? __uint_as_float(v[e]) * osc : __uint_as_float(v[e]) + bias_s[col0 + e];
        float* dst;
        int64_t ld;
        if (col0 < A.c0) {
          ld = A.c0;
          dst = A.out0 + (rowbase + 32 * q) * ld + col0;
        } else {
          ld = Cout - A.c0;
          dst = A.out1 + (rowbase + 32 * q) * ld + (col0 - A.c0);
        }
        float s = 0.f, qv = 0.f;
        stage_store16(stage, f, dst, ld, nvalid, lane, s, qv, want_stats);
        if (want_stats) {
          ds[ch] += (double)s;
          dq[ch] += (double)qv;
        }
      }
      ptx::tc_fence_before();     // the next A_FULL arrival orders the next tile's MMAs after these TMEM reads
    }
    if (want_stats) {
#pragma unroll
      for (int ch = 0; ch < 5; ++ch) {
        if (ch >= nch) break;
        const double s = ds[ch] + __shfl_xor_sync(0xffffffffu, ds[ch], 16);
        const double qv = dq[ch] + __shfl_xor_sync(0xffffffffu, dq[ch], 16);
        if (lane < 16) {
          const int col = hc * ncol_half + ch * 16 + lane;
          atomicAdd(&A.stats[col], s);
          atomicAdd(&A.stats[Cout + col], qv);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<kTmemCols>(tmem);
}


// =====================================================================================================================
// Second generation for Cout = 128 (the trans convolutions of the forward, every input-gradient product of the backward):
// CHANNELS ON THE TMEM LANES and two co-resident CTAs per SM.
//     D[ch, pt] = sum_k W[ch, k] X[pt, k]          A operand = W from TENSOR MEMORY (hi 64 + lo 64 columns, staged once per
//                                                  CTA), B operand = the X tile's K-major image, N = 128 points
// The first kernel keeps 64-80 KiB of weight images next to the 64 KiB tile and one serial chain load -> product -> store
// per SM (ncu: issue slots 28 %, DRAM 26 %, long scoreboard on top).  With W in tensor memory a CTA needs the tile only
// (66 KiB, 256 TMEM columns), so a second CTA's chain fills the gaps of the first.  The epilogue thread owns an output
// channel: BatchNorm sums are thread-local, and a warp stores 32 consecutive channels of one point per instruction.
namespace p2 {
constexpr uint32_t AHI = 0;
constexpr uint32_t ALO = AHI + 2 * kBlk;
constexpr uint32_t BARS = ALO + 2 * kBlk;                  // 65536
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;
constexpr uint32_t D_COL = 0, WHI_COL = 128, WLO_COL = 192;
// fused k | v convolution (Cout = 32 + 128): W_v.lo and both halves of W_k are shared-memory images, D_k takes W_v.lo's columns
constexpr uint32_t KV_WLO = ALO + 2 * kBlk;                // W_v.lo  [128 x 128] K-major, 2 blocks
constexpr uint32_t KV_WKH = KV_WLO + 2 * kBlk;             // W_k.hi  [32 x 128]: 2 blocks of 4 KiB
constexpr uint32_t KV_WKL = KV_WKH + 8192;
constexpr uint32_t KV_BARS = KV_WKL + 8192;                // 114688
constexpr uint32_t KV_TMEMPTR = KV_BARS + 64;
constexpr uint32_t KV_SMEM_BYTES = KV_TMEMPTR + 16;        // 114768: no alignment slack (two CTAs per SM), the base is checked
constexpr uint32_t DK_COL = 192;
enum { BAR_A_FULL = 0, BAR_D_FULL = 1 };
}  // namespace p2

__device__ __forceinline__ void ld8(const float* __restrict__ p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

template <bool kScaled, bool kKV>
__global__ void __launch_bounds__(kThreads, 2) pct_pw2_kernel(const PwArgs A) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (kKV && sm != smem_raw) __trap();       // KV_SMEM_BYTES has no slack: the dynamic shared memory window must start 1 KiB aligned
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (kKV ? p2::KV_BARS : p2::BARS));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (kKV ? p2::KV_TMEMPTR : p2::TMEMPTR));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(&bars[p2::BAR_A_FULL], kComputeThreads);
    ptx::mbar_init(&bars[p2::BAR_D_FULL], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float* Wmain = kKV ? A.W + 32 * 128 : A.W;       // k | v: rows 0..31 are W_k, 32..159 W_v
  if (warp < 4) {       // W: output channel on the TMEM lane, 128 k packed two per 32-bit column, hi and lo
    const int r = 32 * warp + lane;
    const float4* src = reinterpret_cast<const float4*>(Wmain + (int64_t)r * 128);
#pragma unroll 1
    for (int grp = 0; grp < 4; ++grp) {
      uint32_t wh[16], wl[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = __ldg(src + grp * 8 + j);
        split2f<0>(a.x, a.y, wh[2 * j], wl[2 * j]);
        split2f<0>(a.z, a.w, wh[2 * j + 1], wl[2 * j + 1]);
      }
      ptx::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + p2::WHI_COL + grp * 16, wh);
      if (!kKV) {
        ptx::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + p2::WLO_COL + grp * 16, wl);
      } else {          // 32 k = four 16-byte chunks of the K-major image
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int chunk = grp * 4 + c4;
          st_chunk(sm_base + p2::KV_WLO + (uint32_t)(chunk >> 3) * kBlk + ptx::sw128_offset(r, chunk & 7),
                   make_uint4(wl[4 * c4], wl[4 * c4 + 1], wl[4 * c4 + 2], wl[4 * c4 + 3]));
        }
      }
    }
    ptx::tmem_st_wait();
  } else if (kKV && warp < 8) {       // W_k: 32 rows x 16 chunks, four per thread
#pragma unroll 1
    for (int i = tid - 128; i < 32 * 16; i += 128) {
      const int r = i >> 4, chunk = i & 15;
      const float4* src = reinterpret_cast<const float4*>(A.W + (int64_t)r * 128 + chunk * 8);
      const float4 x = __ldg(src), y = __ldg(src + 1);
      const float f[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
      uint4 hi, lo;
      split8(f, hi, lo);
      const uint32_t off = (uint32_t)(chunk >> 3) * 4096 + ptx::sw128_offset(r, chunk & 7);
      st_chunk(sm_base + p2::KV_WKH + off, hi);
      st_chunk(sm_base + p2::KV_WKL + off, lo);
    }
  }
  if (kKV) ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  const int T = (A.P + kTile - 1) / kTile;
  const int64_t G = A.N * T;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = ptx::make_idesc(kFmt, 128, 128);
    const uint64_t dAhi = ptx::smem_desc_sw128(sm_base + p2::AHI), dAlo = ptx::smem_desc_sw128(sm_base + p2::ALO);
    const uint32_t idesc_k = ptx::make_idesc(kFmt, 128, 32);
    const uint64_t dWlo = ptx::smem_desc_sw128(sm_base + p2::KV_WLO);
    const uint64_t dWkh = ptx::smem_desc_sw128(sm_base + p2::KV_WKH), dWkl = ptx::smem_desc_sw128(sm_base + p2::KV_WKL);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t it = 0;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x, ++it) {
      ptx::mbar_wait(&bars[p2::BAR_A_FULL], it & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {          // (W.hi, X.hi) (W.hi, X.lo) (W.lo, X.hi) (W.lo, X.lo)
          const uint32_t wcol = (pass & 2) ? p2::WLO_COL : p2::WHI_COL;
          const uint64_t xd = (pass & 1) ? dAlo : dAhi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t koff = (uint64_t)((ks >> 2) * (kBlk >> 4) + (ks & 3) * 2);
            if (kKV && (pass & 2)) ptx::umma_bf16(tmem_u + p2::D_COL, dWlo + koff, xd + koff, idesc, 1);
            else ptx::umma_bf16_ts(tmem_u + p2::D_COL, tmem_u + wcol + (uint32_t)(ks * 8), xd + koff, idesc, (pass | ks) != 0);
          }
        }
        if (kKV) {        // k = X W_k^T with the POINTS on the lanes: the X tile is the A operand of this one
#pragma unroll
          for (int pass = 0; pass < 4; ++pass) {
            const uint64_t xd = (pass & 1) ? dAlo : dAhi;
            const uint64_t wd = (pass & 2) ? dWkl : dWkh;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::umma_bf16(tmem_u + p2::DK_COL, xd + (uint64_t)((ks >> 2) * (kBlk >> 4) + (ks & 3) * 2),
                             wd + (uint64_t)((ks >> 2) * (4096 >> 4) + (ks & 3) * 2), idesc_k, (pass | ks) != 0);
          }
        }
        ptx::umma_commit(&bars[p2::BAR_D_FULL]);
      }
      __syncwarp();
    }
  } else {
    // =============================== compute warps ===============================
    const int cc = tid & 15, r0 = tid >> 4;          // loader: 8-channel chunk cc of rows r0 + 16 q
    const int ch0 = cc * 8;
    const uint32_t a_off_blk = (uint32_t)(cc >> 3) * kBlk;
    const int q = warp & 3, hc = warp >> 2;           // epilogue: channel = 32 q + lane, point half hc
    const int och = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const float bias = A.bias ? A.bias[(kKV ? 32 : 0) + och] : 0.f;
    float* const out_main = kKV ? A.out1 : A.out0;
    const bool want_stats = A.stats != nullptr;
    double ds = 0, dq = 0;

    uint32_t it = 0;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x, ++it) {
      const int64_t n = g / T;
      const int t = (int)(g - n * T);
      const int64_t rowbase = n * A.P + (int64_t)t * kTile;
      const int valid = min(kTile, A.P - t * kTile);
      const float sc = kScaled ? __ldg(A.scale + 2 * n) : 1.f;
      // ---- load + prologue + split -> X tile (the previous tile's products have completed: D_FULL was waited for).
      auto issue = [&](auto rows_tag, int bt, int64_t rb, int vld, float4 (&u)[4][2], float4 (&w)[4][2]) {
        constexpr int kRows = decltype(rows_tag)::value;
#pragma unroll
        for (int qq = 0; qq < kRows; ++qq) {
          const int row = r0 + 16 * (bt * kRows + qq);
          const bool ok = row < vld;
          const float4* s1 = reinterpret_cast<const float4*>(A.src1 + (rb + row) * 128 + ch0);
          u[qq][0] = ok ? __ldg(s1) : make_float4(0.f, 0.f, 0.f, 0.f);
          u[qq][1] = ok ? __ldg(s1 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (A.mode2) {
            const float4* s2 = reinterpret_cast<const float4*>(A.src2 + (rb + row) * 128 + ch0);
            w[qq][0] = ok ? __ldg(s2) : make_float4(0.f, 0.f, 0.f, 0.f);
            w[qq][1] = ok ? __ldg(s2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      auto finish = [&](auto rows_tag, int bt, const float4 (&u)[4][2], const float4 (&w)[4][2]) {
        constexpr int kRows = decltype(rows_tag)::value;
        float pa1[8], pb1[8], pa2[8], pb2[8];        // fetched per batch (L1 hits): 32 registers not held across the epilogue
        if (A.mode1 == 2) {
          ld8(A.a1 + ch0, pa1);
          ld8(A.b1 + ch0, pb1);
        }
        if (A.mode2 == 2) {
          ld8(A.a2 + ch0, pa2);
          ld8(A.b2 + ch0, pb2);
        }
#pragma unroll
        for (int qq = 0; qq < kRows; ++qq) {
          const int row = r0 + 16 * (bt * kRows + qq);
          const bool ok = row < valid;
          float f[8];
          const float s1[8] = {u[qq][0].x, u[qq][0].y, u[qq][0].z, u[qq][0].w, u[qq][1].x, u[qq][1].y, u[qq][1].z, u[qq][1].w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = s1[e];
            if (A.mode1 == 2) {
              v = fmaf(pa1[e], v, pb1[e]);
              v = v > 0.f ? v : 0.f;
            }
            f[e] = v;
          }
          if (A.mode2) {
            const float s2[8] = {w[qq][0].x, w[qq][0].y, w[qq][0].z, w[qq][0].w,
                                 w[qq][1].x, w[qq][1].y, w[qq][1].z, w[qq][1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = s2[e];
              if (A.mode2 == 2) {
                v = fmaf(pa2[e], v, pb2[e]);
                v = v > 0.f ? v : 0.f;
              }
              f[e] += v;
            }
          }
          if (!ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = 0.f;
          }
          if (A.out_x && ok) {
            float4* ox = reinterpret_cast<float4*>(A.out_x + (rowbase + row) * 128 + ch0);
            ox[0] = make_float4(f[0], f[1], f[2], f[3]);
            ox[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
          if (kScaled) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= sc;
          }
          uint4 hi, lo;
          split8(f, hi, lo);
          const uint32_t off = a_off_blk + ptx::sw128_offset(row, cc & 7);
          st_chunk(sm_base + p2::AHI + off, hi);
          st_chunk(sm_base + p2::ALO + off, lo);
        }
      };
      using I2 = std::integral_constant<int, 2>;
      using I4 = std::integral_constant<int, 4>;
      if (A.mode2) {          // 96 registers per thread (two CTAs per SM): smaller batches than the first kernel's
        float4 u[4][2], w[4][2];
#pragma unroll 1
        for (int bt = 0; bt < 4; ++bt) {
          issue(I2{}, bt, rowbase, valid, u, w);
          finish(I2{}, bt, u, w);
        }
      } else {
        float4 u[4][2];
#pragma unroll 1
        for (int bt = 0; bt < 2; ++bt) {
          issue(I4{}, bt, rowbase, valid, u, u);
          finish(I4{}, bt, u, u);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[p2::BAR_A_FULL]);
      if (g + gridDim.x < G) {          // next tile -> L2 while this one's products and epilogue run (2 lines per thread and source):
                                        // the loader then pays an L2 hit instead of a DRAM round trip per batch
        const int64_t gn = g + gridDim.x;
        const int64_t nn = gn / T;
        const int64_t rb = nn * A.P + (gn - nn * T) * kTile;
        const int vld = min(kTile, A.P - (int)(gn - nn * T) * kTile);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int line = tid + 256 * j;           // 512 lines of 128 B: row = line / 4
          if ((line >> 2) < vld) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.src1 + (rb + (line >> 2)) * 128 + (line & 3) * 32));
            if (A.mode2) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.src2 + (rb + (line >> 2)) * 128 + (line & 3) * 32));
          }
        }
      }

      // ---- epilogue: this thread's channel, the 64 points of its half
      ptx::mbar_wait(&bars[p2::BAR_D_FULL], it & 1);
      ptx::tc_fence_after();
      const float osc = kScaled ? __ldg(A.scale + 2 * n + 1) : 1.f;
      float s = 0.f, sq = 0.f, amax = 0.f;
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {           // 16 points at a time: the prefetched rows of the next tile stay in registers
        uint32_t v[16];
        ptx::tmem_ld16(tmem + lane_addr + p2::D_COL + (uint32_t)(hc * 64 + h * 16), v);
        ptx::tmem_ld_wait();
        const int p0 = hc * 64 + h * 16;
        float* dst = out_main + (rowbase + p0) * 128 + och;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (p0 + e < valid) {
            const float y = kScaled ? __uint_as_float(v[e]) * osc : __uint_as_float(v[e]) + bias;
            dst[(int64_t)e * 128] = y;
            amax = fmaxf(amax, fabsf(y));
            if (want_stats) {
              s += y;
              sq = fmaf(y, y, sq);
            }
          }
        }
      }
      if (A.absmax) {             // the object's largest |output|: the operand scale of a later product comes from it
        amax = warp_max(amax);
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(A.absmax + n), __float_as_uint(amax));
      }
      if (kKV) {                  // k: this thread's point (lane 32 q + lane), 16 of the 32 channels
        uint32_t v[16];
        ptx::tmem_ld16(tmem + lane_addr + p2::DK_COL + (uint32_t)(hc * 16), v);
        ptx::tmem_ld_wait();
        const int pt = 32 * q + lane;
        if (pt < valid) {
          float4* dk = reinterpret_cast<float4*>(A.out0 + (rowbase + pt) * 32 + hc * 16);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float4 o = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
            if (A.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(A.bias + hc * 16 + 4 * e));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            dk[e] = o;
          }
        }
      }
      ptx::tc_fence_before();     // the next A_FULL arrival orders the next tile's products after these TMEM reads
      if (want_stats) {
        ds += (double)s;
        dq += (double)sq;
      }
    }
    if (want_stats) {             // the two point halves of a channel
      atomicAdd(&A.stats[och], ds);
      atomicAdd(&A.stats[128 + och], dq);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<256>(tmem);
}

}  // namespace

int pw_launch(const PwArgs& a, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_pw_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_pw_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_pw_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_done = true;
  }
  const int T = (a.P + kTile - 1) / kTile;
  int64_t G = a.N * T;
  int grid = sm_count();
  if ((int64_t)grid > G) grid = (int)G;
  static const bool v1 = [] { const char* e = getenv("SGA_PCT_PW"); return e && e[0] == 'v' && e[1] == '1'; }();
  const bool plain = !a.pts && a.Cout == 128 && a.c0 == 128;
  const bool kv = !a.pts && a.Cout == 160 && a.c0 == 32 && !a.scale && !a.stats;
  SGA_REQUIRE(!a.absmax || ((plain || kv) && !v1), "sga_pct_pointwise: the per-object maximum is recorded by the second-generation kernel only");
  if ((plain || kv) && !v1) {       // second generation: channels on lanes, two CTAs per SM
    static bool attr2 = false;
    if (!attr2) {
      SGA_CUDA(cudaFuncSetAttribute(pct_pw2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2::SMEM_BYTES));
      SGA_CUDA(cudaFuncSetAttribute(pct_pw2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2::SMEM_BYTES));
      SGA_CUDA(cudaFuncSetAttribute(pct_pw2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2::KV_SMEM_BYTES));
      attr2 = true;
    }
    int64_t g2 = 2 * (int64_t)sm_count();
    if (g2 > G) g2 = G;
    if (kv) pct_pw2_kernel<false, true><<<(unsigned)g2, kThreads, p2::KV_SMEM_BYTES, st>>>(a);
    else if (a.scale) pct_pw2_kernel<true, false><<<(unsigned)g2, kThreads, p2::SMEM_BYTES, st>>>(a);
    else pct_pw2_kernel<false, false><<<(unsigned)g2, kThreads, p2::SMEM_BYTES, st>>>(a);
    SGA_LAUNCH_CHECK();
    return SGA_OK;
  }
  if (a.pts) pct_pw_kernel<true, false><<<grid, kThreads, SMEM_BYTES, st>>>(a);
  else if (a.scale) pct_pw_kernel<false, true><<<grid, kThreads, SMEM_BYTES, st>>>(a);
  else pct_pw_kernel<false, false><<<grid, kThreads, SMEM_BYTES, st>>>(a);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace pct
}  // namespace sga

using sga::pct::PwArgs;

static int check_pw_common(const char* who, int64_t N, int P, const float* W, int Cout) {
  SGA_REQUIRE(N >= 1 && P >= 1, "%s: N=%lld P=%d", who, (long long)N, P);
  SGA_REQUIRE(W && ((uintptr_t)W & 15) == 0, "%s: W must be a 16-byte aligned device pointer", who);
  SGA_REQUIRE(Cout == 128 || Cout == 160, "%s: Cout=%d (128, or 160 = 32 + 128 for the fused k | v convolution)", who, Cout);
  return SGA_OK;
}

static int pct_pointwise_impl(const float* src1, const float* a1, const float* b1, const float* src2, const float* a2,
                              const float* b2, int64_t N, int P, const float* W, const float* bias, int Cout, int c0,
                              float* out_x, float* out0, float* out1, double* stats, const float* scale, float* absmax, void* stream) {
  if (N <= 0) return SGA_OK;
  int rc = check_pw_common("sga_pct_pointwise", N, P, W, Cout);
  if (rc) return rc;
  SGA_REQUIRE(src1 && out0, "sga_pct_pointwise: null pointer");
  SGA_REQUIRE((a1 == nullptr) == (b1 == nullptr) && (a2 == nullptr) == (b2 == nullptr), "sga_pct_pointwise: a/b must come in pairs");
  SGA_REQUIRE(src2 || (!a2 && !b2), "sga_pct_pointwise: a2/b2 without src2");
  SGA_REQUIRE(c0 == Cout || (c0 == 32 && Cout == 160 && out1), "sga_pct_pointwise: c0=%d Cout=%d", c0, Cout);
  SGA_REQUIRE((((uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)out_x) & 15) == 0, "sga_pct_pointwise: activations must be 16-byte aligned");
  PwArgs a{};
  a.src1 = src1; a.a1 = a1; a.b1 = b1; a.mode1 = a1 ? 2 : 1;
  a.src2 = src2; a.a2 = a2; a.b2 = b2; a.mode2 = src2 ? (a2 ? 2 : 1) : 0;
  a.pts = nullptr; a.w1 = nullptr;
  a.N = N; a.P = P; a.W = W; a.bias = bias; a.Cout = Cout; a.c0 = c0;
  a.out_x = out_x; a.out0 = out0; a.out1 = out1; a.stats = stats; a.scale = scale; a.absmax = absmax;
  return sga::pct::pw_launch(a, (cudaStream_t)stream);
}

extern "C" int sga_pct_pointwise(const float* src1, const float* a1, const float* b1, const float* src2, const float* a2,
                                 const float* b2, int64_t N, int P, const float* W, const float* bias, int Cout, int c0,
                                 float* out_x, float* out0, float* out1, double* stats, void* stream) {
  return pct_pointwise_impl(src1, a1, b1, src2, a2, b2, N, P, W, bias, Cout, c0, out_x, out0, out1, stats, nullptr, nullptr, stream);
}

// The fused k | v convolution of an SA layer (Cout = 32 + 128, no statistics) that also records max |v| per object
// (v_absmax [N], zeroed by the caller): the backward's operand scale needs it and would otherwise re-read v.
extern "C" int sga_pct_pointwise_kv(const float* src1, const float* a1, const float* b1, const float* src2, const float* a2,
                                    const float* b2, int64_t N, int P, const float* W, const float* bias, float* out_x, float* k,
                                    float* v, float* v_absmax, void* stream) {
  return pct_pointwise_impl(src1, a1, b1, src2, a2, b2, N, P, W, bias, 160, 32, out_x, k, v, nullptr, nullptr, v_absmax, stream);
}

// The same kernel for the input-gradient products of the backward (dX = dY W: pass W^T as the weight): the operand is a
// gradient, so it is multiplied by the object's power-of-two scale[n][0] before the fp16 split and the result by
// scale[n][1] = 1 / scale[n][0] (sga_pct_pow2_scale).
extern "C" int sga_pct_pointwise_scaled(const float* src, const float* scale, int64_t N, int P, const float* Wt, float* out, void* stream) {
  SGA_REQUIRE(scale, "sga_pct_pointwise_scaled: null scale");
  return pct_pointwise_impl(src, nullptr, nullptr, nullptr, nullptr, nullptr, N, P, Wt, nullptr, 128, 128, nullptr, out, nullptr,
                            nullptr, scale, nullptr, stream);
}

// ... and with max |out| per object recorded on the way out (out_absmax [N], zeroed by the caller)
extern "C" int sga_pct_pointwise_scaled_absmax(const float* src, const float* scale, int64_t N, int P, const float* Wt, float* out,
                                               float* out_absmax, void* stream) {
  SGA_REQUIRE(scale && out_absmax, "sga_pct_pointwise_scaled_absmax: null pointer");
  return pct_pointwise_impl(src, nullptr, nullptr, nullptr, nullptr, nullptr, N, P, Wt, nullptr, 128, 128, nullptr, out, nullptr,
                            nullptr, scale, out_absmax, stream);
}

extern "C" int sga_pct_embed(const float* pts, int64_t N, int P, const float* W1, const float* a1, const float* b1,
                             const float* W2, float* z2, double* stats, void* stream) {
  if (N <= 0) return SGA_OK;
  int rc = check_pw_common("sga_pct_embed", N, P, W2, 128);
  if (rc) return rc;
  SGA_REQUIRE(pts && W1 && a1 && b1 && z2, "sga_pct_embed: null pointer");
  PwArgs a{};
  a.mode1 = 0; a.mode2 = 0;
  a.a1 = a1; a.b1 = b1;
  a.pts = pts; a.w1 = W1;
  a.N = N; a.P = P; a.W = W2; a.bias = nullptr; a.Cout = 128; a.c0 = 128;
  a.out_x = nullptr; a.out0 = z2; a.out1 = nullptr; a.stats = stats;
  return sga::pct::pw_launch(a, (cudaStream_t)stream);
}
