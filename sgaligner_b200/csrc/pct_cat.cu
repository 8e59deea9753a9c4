// NaivePCT: concat(x1..x4) -> Conv1d(512, 1024, bias=False) -> BatchNorm -> LeakyReLU(0.2) -> max over the points
// (reference: src/aligner/networks/pct.py:285-289 `self.linear`, :308-313) -- half of the encoder's FLOPs.
// The [N, 1024, P] activation is never formed: BatchNorm is a per-channel affine map and LeakyReLU is monotone, so
//     max_p lrelu(a z_p + b) = lrelu(a * (a >= 0 ? max_p z_p : min_p z_p) + b),
// and the kernel keeps only max_p z, min_p z per (object, channel) plus the per-channel sum / sum of squares of z (the
// batch statistics of that BatchNorm in train mode) -- 8 KB out per object instead of 2 MB.
//
// Channels on TMEM lanes (the max / min / sums over points are thread-local): D[ch, pts] = W[ch, 512] X^T.  One CTA
// owns a block of 128 output channels for the whole launch: W.hi lives in TENSOR MEMORY (256 columns of packed bf16
// pairs; the A operand of two of the three bf16x3 passes costs no shared-memory bandwidth), W.lo in shared memory
// (128 KiB).  X^T streams through a 3-slot ring of [128 points x 64 channels] hi/lo half-chunks that the 8 compute
// warps produce from the fp32 activations (x4 = x3 + relu(after_norm(t4)) is formed on the fly, pct.py:230), 12
// tcgen05.mma per half-chunk, two accumulators so that the pooling epilogue of tile g-1 runs under the MMAs of tile g.
#include "pct_common.cuh"

namespace sga {
namespace pct {
namespace {

namespace ct {
constexpr uint32_t WLO = 0;                               // 8 blocks (kc * 2 + blk) x 16 KiB
constexpr uint32_t RING = WLO + 8 * kBlk;                 // 3 slots x {hi 16 KiB, lo 16 KiB}
constexpr uint32_t SLOT = 2 * kBlk;
constexpr uint32_t AB4 = RING + 3 * SLOT;                 // a4[128], b4[128]
constexpr uint32_t BARS = AB4 + 1024;
constexpr uint32_t TMEMPTR = BARS + 128;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;
constexpr uint32_t D_COL = 0, WHI_COL = 256;
enum { BAR_FULL = 0 /*..2*/, BAR_FREE = 3 /*..5*/, BAR_D_FULL = 6 /*,7*/, BAR_D_FREE = 8 /*,9*/, kNumBars = 10 };
}  // namespace ct

// kMode 0: the forward above.  kMode 1: the training forward -- the epilogue also tracks WHICH point holds the maximum /
// minimum of every (object, channel) (the backward routes the pooled gradient there).  kMode 2: the dense half of the
// backward through this layer, as the same streaming product with another weight and another epilogue:
//     g_x[a][n, p, :] = -u[a] - (M (xcat[n, p, :] - xbar))[a-th block of 128 channels],   M = W^T diag(f) W  (512 x 512, symmetric)
// (xbar = the batch mean of xcat, subtracted in the loader BEFORE the split: M xcat and M xbar would cancel digits)
// (train-mode BatchNorm sends a gradient -e_c - f_c z[n,p,c] to EVERY point; W^T of that is a 512 -> 512 pointwise map of
// the concatenated activations -- see sgaligner_b200/pct.py).  The fourth source is the plain x4 then (`t4`); M is of
// gradient magnitude, so it is multiplied by the power of two scale[0] before the fp16 split and the product by scale[1]
// (sga_pct_pow2_scale); grid.y = 4, and a thread stores its channel's 32 points per TMEM load: a warp writes 32
// consecutive channels of one point = one 128-byte line per instruction.
struct CatBwdOut { float* g[4]; const float* u; const float* scale; const float* xbar; };

// kImg: the operand tiles come PRE-SPLIT from HBM (pct_cat_pack_kernel writes every [128 points x 64 channels] half-chunk
// once, hi | lo, already in the swizzled image the MMA reads) and a producer warp streams them with cp.async.bulk (TMA,
// 32 KiB per copy, completion on the slot's mbarrier): no registers, no conversion, no shared-memory stores in this
// kernel.  Without it each of the 8 channel-block CTAs re-reads and re-converts the same activations -- 8x the
// conversion work, and the first use of the loaded rows was 31 % of the stall samples (profiles/r2_ncu_pct_cat_lines.txt).
template <int kMode, bool kImg>
__global__ void __launch_bounds__(kImg ? kThreads + 32 : kThreads, 1)
pct_cat_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ x3,
               const float* __restrict__ t4, const float* __restrict__ a4, const float* __restrict__ b4, int64_t N, int P,
               const float* __restrict__ WL, float* __restrict__ zmax, float* __restrict__ zmin, double* __restrict__ stats,
               int32_t* __restrict__ imax, int32_t* __restrict__ imin, const CatBwdOut bo, const unsigned char* __restrict__ img) {
  using namespace ct;
  const int nthr = kImg ? kThreads + 32 : kThreads;
  constexpr int F = 0;
  const float wsc = (kMode == 2) ? bo.scale[0] : 1.f;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  float* ab4 = reinterpret_cast<float*>(sm + AB4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb0 = blockIdx.y * 128;                     // first output channel of this CTA

  // ---- setup: W.lo image, a4/b4, barriers, TMEM, W.hi -> TMEM
  for (int i = tid; i < 128 * 64; i += nthr) {        // [128 rows][64 chunks of 8 k]
    const int r = i >> 6, j = i & 63;
    const float4* src = reinterpret_cast<const float4*>(WL + (int64_t)(cb0 + r) * 512 + j * 8);
    const float4 x = src[0], y = src[1];
    const float f[8] = {x.x * wsc, x.y * wsc, x.z * wsc, x.w * wsc, y.x * wsc, y.y * wsc, y.z * wsc, y.w * wsc};
    uint4 hi, lo;
    split8f<F>(f, hi, lo);
    st_chunk(sm_base + WLO + (uint32_t)(j >> 3) * kBlk + ptx::sw128_offset(r, j & 7), lo);
  }
  if (kMode != 2) {
    for (int i = tid; i < 128; i += nthr) {
      ab4[i] = a4[i];
      ab4[128 + i] = b4[i];
    }
  }
  if (tid == 0) {
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(&bars[BAR_FULL + s], kImg ? 1 : kComputeThreads);
      ptx::mbar_init(&bars[BAR_FREE + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars[BAR_D_FULL + b], 1);
      ptx::mbar_init(&bars[BAR_D_FREE + b], kComputeThreads);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 4) {       // W.hi: output channel on the TMEM lane, 512 k packed two per 32-bit column (lower k in the low half)
    const int r = 32 * warp + lane;
    const float4* src = reinterpret_cast<const float4*>(WL + (int64_t)(cb0 + r) * 512);
#pragma unroll 1
    for (int grp = 0; grp < 16; ++grp) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = src[grp * 8 + j];
        w[2 * j] = pack2f<F>(a.x * wsc, a.y * wsc);
        w[2 * j + 1] = pack2f<F>(a.z * wsc, a.w * wsc);
      }
      ptx::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + WHI_COL + grp * 16, w);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  const int T = (P + kTile - 1) / kTile;
  const int64_t nobj = (N > (int64_t)blockIdx.x) ? (N - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t G = nobj * T;                             // point tiles this CTA walks, as one stream

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = ptx::make_idesc(F, 128, 128);
    const uint64_t dWlo = ptx::smem_desc_sw128(sm_base + WLO);
    const uint64_t dRing = ptx::smem_desc_sw128(sm_base + RING);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t hg = 0;
    for (int64_t g = 0; g < G; ++g) {
      const uint32_t b = (uint32_t)(g & 1);
      if (g >= 2) ptx::mbar_wait(&bars[BAR_D_FREE + b], (uint32_t)(((g - 2) >> 1) & 1));
      for (int h = 0; h < 8; ++h, ++hg) {
        const uint32_t slot = hg % 3;
        ptx::mbar_wait(&bars[BAR_FULL + slot], (hg / 3) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d = tmem_u + D_COL + b * 128;
          const uint64_t xh = dRing + (uint64_t)slot * (SLOT >> 4), xl = xh + (uint64_t)(kBlk >> 4);
          const uint32_t wh = tmem_u + WHI_COL + (uint32_t)h * 32;
          const uint64_t wl = dWlo + (uint64_t)h * (kBlk >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16_ts(d, wh + ks * 8, xh + (uint64_t)(ks * 2), idesc, (h | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16_ts(d, wh + ks * 8, xl + (uint64_t)(ks * 2), idesc, 1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16(d, wl + (uint64_t)(ks * 2), xh + (uint64_t)(ks * 2), idesc, 1);
          ptx::umma_commit(&bars[BAR_FREE + slot]);
          if (h == 7) ptx::umma_commit(&bars[BAR_D_FULL + b]);
        }
        __syncwarp();
      }
    }
  } else if (kImg && warp == 9) {
    // =============================== producer: one lane streams the packed half-chunks ===============================
    if (lane == 0) {
      uint32_t hg = 0;
      int64_t n = blockIdx.x;
      int t = 0;
      for (int64_t g = 0; g < G; ++g) {
        const unsigned char* tile = img + ((n * T + t) * 8) * (int64_t)SLOT;
        for (int h = 0; h < 8; ++h, ++hg) {
          const uint32_t slot = hg % 3;
          if (hg >= 3) ptx::mbar_wait(&bars[BAR_FREE + slot], (hg / 3 - 1) & 1);
          ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + slot], SLOT);
          ptx::bulk_g2s(sm + RING + slot * SLOT, tile + (int64_t)h * SLOT, SLOT, &bars[BAR_FULL + slot]);
        }
        if (++t == T) {
          t = 0;
          n += gridDim.x;
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== compute warps ===============================
    const int cc = tid & 7, r0 = tid >> 3;                 // loader: 8-channel chunk cc of rows r0 + 32 q
    const int q = warp & 3, hc = warp >> 2;                // epilogue: lane quarter (channel), column half (points)
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int ch = cb0 + 32 * q + lane;
    float vmax = -INFINITY, vmin = INFINITY;
    int pmax = 0, pmin = 0;
    const float u_ch = (kMode == 2) ? bo.u[ch] : 0.f;
    const float winv = (kMode == 2) ? bo.scale[1] : 1.f;
    float* const gout = (kMode == 2) ? bo.g[blockIdx.y] : nullptr;
    double dsum = 0, dsq = 0;
    uint32_t hg = 0;
    int e_t = 0;                                           // epilogue cursor: tile within object
    int64_t e_n = blockIdx.x;

    auto epilogue = [&](int64_t gp) {
      const uint32_t b = (uint32_t)(gp & 1);
      ptx::mbar_wait(&bars[BAR_D_FULL + b], (uint32_t)((gp >> 1) & 1));
      ptx::tc_fence_after();
      const int valid = min(kTile, P - e_t * kTile) - hc * 64;     // valid columns of this half
      float s = 0.f, sq = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + lane_addr + D_COL + b * 128 + (uint32_t)(hc * 64 + h * 32), v);
        ptx::tmem_ld_wait();
        if (kMode == 2) {
          float* dst = gout + (e_n * (int64_t)P + (int64_t)e_t * kTile + hc * 64 + h * 32) * 128 + (32 * q + lane);
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (h * 32 + e < valid) dst[(int64_t)e * 128] = -fmaf(__uint_as_float(v[e]), winv, u_ch);
        } else {
          const int pt0 = e_t * kTile + hc * 64 + h * 32;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float f = __uint_as_float(v[e]);
            const bool ok = h * 32 + e < valid;
            if (kMode == 1) {
              if (ok && f > vmax) { vmax = f; pmax = pt0 + e; }
              if (ok && f < vmin) { vmin = f; pmin = pt0 + e; }
            } else {
              vmax = fmaxf(vmax, ok ? f : -INFINITY);
              vmin = fminf(vmin, ok ? f : INFINITY);
            }
            const float g0 = ok ? f : 0.f;
            s += g0;
            sq = fmaf(g0, g0, sq);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_D_FREE + b]);
      dsum += (double)s;
      dsq += (double)sq;
      if (++e_t == T) {                                    // object complete: this half's max / min
        if (kMode != 2) {
          zmax[(e_n * 2 + hc) * 1024 + ch] = vmax;
          zmin[(e_n * 2 + hc) * 1024 + ch] = vmin;
          if (kMode == 1) {
            imax[(e_n * 2 + hc) * 1024 + ch] = pmax;
            imin[(e_n * 2 + hc) * 1024 + ch] = pmin;
          }
        }
        vmax = -INFINITY;
        vmin = INFINITY;
        e_t = 0;
        e_n += gridDim.x;
      }
    };

    int64_t n = blockIdx.x;
    int t = 0;
    for (int64_t g = 0; g < G; ++g) {
      const int64_t rowbase = n * (int64_t)P + (int64_t)t * kTile;
      const int valid = min(kTile, P - t * kTile);
#pragma unroll 1
      for (int h = 0; h < (kImg ? 0 : 8); ++h, ++hg) {
        const int kc = h >> 1;
        const uint32_t slot = hg % 3;
        const float* src = (kc == 0) ? x1 : (kc == 1) ? x2 : (kMode == 2 && kc == 3) ? t4 : x3;
        const int chb = (h & 1) * 64 + cc * 8;             // channel of this thread's chunk inside the 128-channel source
        float4 u[4][2], w[4][2];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int row = r0 + 32 * qq;
          const bool ok = row < valid;
          const float4* s1 = reinterpret_cast<const float4*>(src + (rowbase + row) * 128 + chb);
          u[qq][0] = ok ? __ldg(s1) : make_float4(0.f, 0.f, 0.f, 0.f);
          u[qq][1] = ok ? __ldg(s1 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (kMode != 2 && kc == 3) {
            const float4* s2 = reinterpret_cast<const float4*>(t4 + (rowbase + row) * 128 + chb);
            w[qq][0] = ok ? __ldg(s2) : make_float4(0.f, 0.f, 0.f, 0.f);
            w[qq][1] = ok ? __ldg(s2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (hg >= 3) ptx::mbar_wait(&bars[BAR_FREE + slot], (hg / 3 - 1) & 1);   // the slot's previous half-chunk is consumed
        // per-channel constants of this thread's 8 channels, fetched ONCE per half-chunk as vectors: 64 scalar shared-memory
        // reads per thread (2-way bank conflicts: lanes 32 bytes apart) made the LSU data pipe the bound of this kernel
        // (ncu: lsu wavefronts 75 % of peak, 41 % of them bank conflicts)
        float av[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, mb[8];
        if (kMode != 2 && kc == 3) {
          const float4 a0 = *reinterpret_cast<const float4*>(ab4 + chb), a1 = *reinterpret_cast<const float4*>(ab4 + chb + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(ab4 + 128 + chb), b1 = *reinterpret_cast<const float4*>(ab4 + 128 + chb + 4);
          av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
          bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        }
        if (kMode == 2) {           // centre: xbar (2 KB, L1 resident; shared memory is full)
          const float4 m0 = __ldg(reinterpret_cast<const float4*>(bo.xbar + kc * 128 + chb));
          const float4 m1 = __ldg(reinterpret_cast<const float4*>(bo.xbar + kc * 128 + chb) + 1);
          mb[0] = m0.x; mb[1] = m0.y; mb[2] = m0.z; mb[3] = m0.w; mb[4] = m1.x; mb[5] = m1.y; mb[6] = m1.z; mb[7] = m1.w;
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int row = r0 + 32 * qq;
          float f[8] = {u[qq][0].x, u[qq][0].y, u[qq][0].z, u[qq][0].w, u[qq][1].x, u[qq][1].y, u[qq][1].z, u[qq][1].w};
          if (kMode == 2) {
            const bool okr = row < valid;
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = okr ? f[e] - mb[e] : 0.f;
          }
          if (kMode != 2 && kc == 3) {
            const float tv[8] = {w[qq][0].x, w[qq][0].y, w[qq][0].z, w[qq][0].w, w[qq][1].x, w[qq][1].y, w[qq][1].z, w[qq][1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float r = fmaf(av[e], tv[e], bv[e]);
              f[e] += r > 0.f ? r : 0.f;
            }
            if (row >= valid) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = 0.f;
            }
          }
          uint4 hi, lo;
          split8f<F>(f, hi, lo);
          const uint32_t off = RING + slot * SLOT + ptx::sw128_offset(row, cc);
          st_chunk(sm_base + off, hi);
          st_chunk(sm_base + off + kBlk, lo);
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[BAR_FULL + slot]);
      }
      if (kImg) {
        epilogue(g);
        continue;
      }
      if (g > 0) epilogue(g - 1);
      if (++t == T) {
        t = 0;
        n += gridDim.x;
      }
    }
    if (!kImg && G > 0) epilogue(G - 1);
    if (kMode != 2) {
      atomicAdd(&stats[ch], dsum);
      atomicAdd(&stats[1024 + ch], dsq);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<512>(tmem);
}

// The operand image of pct_cat_kernel<., true>: per point tile and half-chunk h (source kc = h / 2, channels (h & 1) * 64 ..)
// one 32 KiB block [hi 16 KiB | lo 16 KiB] of 128-byte rows with the 128B swizzle -- exactly the bytes of a ring slot.
// x4 = x3 + relu(a4 t4 + b4) is formed here; rows past the object's last point are zero.  One CTA per tile, HBM bound
// (4 B in, 4 B out per element).
__global__ void __launch_bounds__(256) pct_cat_pack_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                           const float* __restrict__ x3, const float* __restrict__ t4,
                                                           const float* __restrict__ a4, const float* __restrict__ b4, int P, int T,
                                                           unsigned char* __restrict__ img) {
  const int64_t tile = blockIdx.x;
  const int64_t n = tile / T;
  const int t = (int)(tile - n * T);
  const int64_t rowbase = n * (int64_t)P + (int64_t)t * kTile;
  const int valid = min(kTile, P - t * kTile);
  const int tid = threadIdx.x, cc = tid & 7, r0 = tid >> 3;
  unsigned char* out = img + tile * 8 * (int64_t)ct::SLOT;
#pragma unroll 1
  for (int h = 0; h < 8; ++h) {
    const int kc = h >> 1;
    const float* src = (kc == 0) ? x1 : (kc == 1) ? x2 : x3;
    const int chb = (h & 1) * 64 + cc * 8;
    float4 u[4][2], w[4][2];
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      const int row = r0 + 32 * qq;
      const bool ok = row < valid;
      const float4* s1 = reinterpret_cast<const float4*>(src + (rowbase + row) * 128 + chb);
      u[qq][0] = ok ? __ldg(s1) : make_float4(0.f, 0.f, 0.f, 0.f);
      u[qq][1] = ok ? __ldg(s1 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (kc == 3) {
        const float4* s2 = reinterpret_cast<const float4*>(t4 + (rowbase + row) * 128 + chb);
        w[qq][0] = ok ? __ldg(s2) : make_float4(0.f, 0.f, 0.f, 0.f);
        w[qq][1] = ok ? __ldg(s2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float av[8], bv[8];
    if (kc == 3) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(a4 + chb)), a1 = __ldg(reinterpret_cast<const float4*>(a4 + chb) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b4 + chb)), b1 = __ldg(reinterpret_cast<const float4*>(b4 + chb) + 1);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    }
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      const int row = r0 + 32 * qq;
      float f[8] = {u[qq][0].x, u[qq][0].y, u[qq][0].z, u[qq][0].w, u[qq][1].x, u[qq][1].y, u[qq][1].z, u[qq][1].w};
      if (kc == 3) {
        const float tv[8] = {w[qq][0].x, w[qq][0].y, w[qq][0].z, w[qq][0].w, w[qq][1].x, w[qq][1].y, w[qq][1].z, w[qq][1].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float r = fmaf(av[e], tv[e], bv[e]);
          f[e] += r > 0.f ? r : 0.f;
        }
        if (row >= valid) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
      }
      uint4 hi, lo;
      split8(f, hi, lo);
      unsigned char* dst = out + (int64_t)h * ct::SLOT + ptx::sw128_offset(row, cc);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + kBlk) = lo;
    }
  }
}

// pooled[n, c] = lrelu_0.2(a_c * (a_c >= 0 ? max : min) + b_c), the two column halves combined (pct.py:288-289, :310);
// with the tracked indices also the point that holds it (pstar) and the selected pre-BatchNorm value (zsel)
__global__ void pct_pool_act_kernel(const float* __restrict__ zmax, const float* __restrict__ zmin, const int32_t* __restrict__ imax,
                                    const int32_t* __restrict__ imin, const float* __restrict__ a, const float* __restrict__ b,
                                    int64_t N, int P, float* __restrict__ out, int32_t* __restrict__ pstar, float* __restrict__ zsel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 1024) return;
  const int64_t n = i >> 10;
  const int c = (int)(i & 1023);
  const int64_t i0 = (n * 2) * 1024 + c, i1 = i0 + 1024;
  float mx = zmax[i0], mn = zmin[i0];
  int px = imax ? imax[i0] : 0, pn = imin ? imin[i0] : 0;
  if (P > 64) {                                           // the second column half holds points only then
    const float mx1 = zmax[i1], mn1 = zmin[i1];
    if (mx1 > mx || (imax && mx1 == mx && imax[i1] < px)) { mx = mx1; if (imax) px = imax[i1]; }
    if (mn1 < mn || (imin && mn1 == mn && imin[i1] < pn)) { mn = mn1; if (imin) pn = imin[i1]; }
  }
  const float ac = a[c];
  const float z = ac >= 0.f ? mx : mn;
  const float y = fmaf(ac, z, b[c]);
  out[i] = y > 0.f ? y : 0.2f * y;
  if (pstar) pstar[i] = ac >= 0.f ? px : pn;
  if (zsel) zsel[i] = z;
}

}  // namespace
}  // namespace pct
}  // namespace sga

static int cat_launch(int mode, const float* x1, const float* x2, const float* x3, const float* t4, const float* a4, const float* b4,
                      int64_t N, int P, const float* WL, float* zmax, float* zmin, double* stats, int32_t* imax, int32_t* imin,
                      const sga::pct::CatBwdOut& bo, const unsigned char* img, void* stream) {
  using namespace sga::pct;
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_cat_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_cat_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_cat_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_cat_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_cat_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct::SMEM_BYTES));
    attr_done = true;
  }
  const int gy = mode == 2 ? 4 : 8;
  int gx = sga::sm_count() / gy;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > N) gx = (int)N;
  const dim3 grid(gx, gy);
  cudaStream_t st = (cudaStream_t)stream;
  if (img && mode == 0) pct_cat_kernel<0, true><<<grid, kThreads + 32, ct::SMEM_BYTES, st>>>(x1, x2, x3, t4, a4, b4, N, P, WL, zmax, zmin, stats, imax, imin, bo, img);
  else if (img && mode == 1) pct_cat_kernel<1, true><<<grid, kThreads + 32, ct::SMEM_BYTES, st>>>(x1, x2, x3, t4, a4, b4, N, P, WL, zmax, zmin, stats, imax, imin, bo, img);
  else if (mode == 0) pct_cat_kernel<0, false><<<grid, kThreads, ct::SMEM_BYTES, st>>>(x1, x2, x3, t4, a4, b4, N, P, WL, zmax, zmin, stats, imax, imin, bo, nullptr);
  else if (mode == 1) pct_cat_kernel<1, false><<<grid, kThreads, ct::SMEM_BYTES, st>>>(x1, x2, x3, t4, a4, b4, N, P, WL, zmax, zmin, stats, imax, imin, bo, nullptr);
  else pct_cat_kernel<2, false><<<grid, kThreads, ct::SMEM_BYTES, st>>>(x1, x2, x3, t4, a4, b4, N, P, WL, zmax, zmin, stats, imax, imin, bo, nullptr);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_cat_linear(const float* x1, const float* x2, const float* x3, const float* t4, const float* a4,
                                  const float* b4, int64_t N, int P, const float* WL, float* zmax, float* zmin,
                                  double* stats, int32_t* imax, int32_t* imin, const void* img, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x1 && x2 && x3 && t4 && a4 && b4 && WL && zmax && zmin && stats && P >= 1, "sga_pct_cat_linear: bad arguments");
  SGA_REQUIRE((imax == nullptr) == (imin == nullptr), "sga_pct_cat_linear: imax / imin come together");
  SGA_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)x3 | (uintptr_t)t4 | (uintptr_t)WL) & 15) == 0,
              "sga_pct_cat_linear: activations / weights must be 16-byte aligned");
  sga::pct::CatBwdOut bo{};
  SGA_REQUIRE(((uintptr_t)img & 127) == 0, "sga_pct_cat_linear: img must be 128-byte aligned");
  return cat_launch(imax ? 1 : 0, x1, x2, x3, t4, a4, b4, N, P, WL, zmax, zmin, stats, imax, imin, bo, (const unsigned char*)img, stream);
}

extern "C" size_t sga_pct_cat_image_bytes(int64_t N, int P) {
  return (size_t)N * (size_t)((P + sga::pct::kTile - 1) / sga::pct::kTile) * 8 * sga::pct::ct::SLOT;
}

extern "C" int sga_pct_cat_pack(const float* x1, const float* x2, const float* x3, const float* t4, const float* a4, const float* b4,
                                int64_t N, int P, void* img, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x1 && x2 && x3 && t4 && a4 && b4 && img && P >= 1, "sga_pct_cat_pack: bad arguments");
  SGA_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)x3 | (uintptr_t)t4 | (uintptr_t)a4 | (uintptr_t)b4 | (uintptr_t)img) & 15) == 0,
              "sga_pct_cat_pack: 16-byte alignment");
  const int T = (P + sga::pct::kTile - 1) / sga::pct::kTile;
  SGA_REQUIRE(N * T < ((int64_t)1 << 31), "sga_pct_cat_pack: too many tiles");
  sga::pct::pct_cat_pack_kernel<<<(unsigned)(N * T), 256, 0, (cudaStream_t)stream>>>(x1, x2, x3, t4, a4, b4, P, T, (unsigned char*)img);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_cat_dense_bwd(const float* x1, const float* x2, const float* x3, const float* x4, int64_t N, int P,
                                     const float* M, const float* scale, const float* u, const float* xbar, float* g1, float* g2,
                                     float* g3, float* g4, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x1 && x2 && x3 && x4 && M && scale && u && xbar && g1 && g2 && g3 && g4 && P >= 1, "sga_pct_cat_dense_bwd: bad arguments");
  SGA_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)x3 | (uintptr_t)x4 | (uintptr_t)M) & 15) == 0,
              "sga_pct_cat_dense_bwd: activations / weights must be 16-byte aligned");
  sga::pct::CatBwdOut bo{};
  bo.g[0] = g1; bo.g[1] = g2; bo.g[2] = g3; bo.g[3] = g4; bo.u = u; bo.scale = scale; bo.xbar = xbar;
  return cat_launch(2, x1, x2, x3, x4, nullptr, nullptr, N, P, M, nullptr, nullptr, nullptr, nullptr, nullptr, bo, nullptr, stream);
}

extern "C" int sga_pct_pool_act(const float* zmax, const float* zmin, const int32_t* imax, const int32_t* imin, const float* a,
                                const float* b, int64_t N, int P, float* out, int32_t* pstar, float* zsel, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(zmax && zmin && a && b && out, "sga_pct_pool_act: null pointer");
  const int64_t total = N * 1024;
  sga::pct::pct_pool_act_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(zmax, zmin, imax, imin, a, b, N, P, out,
                                                                                                  pstar, zsel);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
