// NaivePCT self-attention on the tensor cores (reference: SA.forward, src/aligner/networks/pct.py:211-232).
//   x_q = x_k^T (ONE shared weight, :199)   energy = x_k^T x_k / sqrt(32)   attention = softmax(energy, dim=-1)
//   x_s[:, j] = sum_i x_v[:, i] attention[i, j]            (torch.bmm(x_v, attention), :224)
// The softmax normalises over j while the product contracts i, so the normaliser of row i must be known before any
// output column can be finished: two kernels, and the [P x P] map never leaves the SM.
//
//   pct_attn_stats_kernel  per object: S = K K^T (tcgen05, K = 32 channels as fp16 hi|lo side by side in one 128-byte
//       swizzled row, so the hi/lo partial products are descriptor offsets into the same image), a whole 128-row
//       block of S (<= 512 columns) in tensor memory, thread-local row max and sum of exponentials ->
//       (max_i * a, log2(sum_i))   (a = log2(e) / sqrt(32); max = +inf for padding rows)
//   pct_attn_kernel        per (object, 128-column block j): for every 128-row block i:
//       S^T[j, i] (tensor memory) -> E = exp2((S a - max_i a) - log2 sum_i) = attention[i, j] -> fp16 hi / lo BACK INTO TENSOR MEMORY as the
//       A operand of  O[j, :] += E V[i, :]   (tcgen05.mma A-from-TMEM; V tile read MN-major from shared memory);
//       S double-buffered, the V tile of block i+1 is loaded while block i is in the tensor pipe.
// Split fp16 operands (pct_common.cuh), fp32 accumulate: three passes for x_v * attention, all four partial products for
// the scores.  P <= 512 (the [128 x P] score block must fit the 512 columns of tensor memory).
#include "pct_common.cuh"
#include <stdlib.h>

namespace sga {
namespace pct {
namespace {

constexpr int kMaxT = 4;                               // P <= 512
constexpr float kAlpha = 1.4426950408889634f * 0.17677669529663687f;   // log2(e) / sqrt(32)

// ---- K image: per tile of 128 points one 16 KiB block of 128-byte rows [k.hi (32 fp16) | k.lo (32 fp16)]
__device__ __forceinline__ void load_k_image(const float* __restrict__ k, int64_t rowbase0, int P, int T, uint32_t kimg_addr, int tid) {
  // loads of all tiles first (the asm stores below are ordering barriers for the compiler: interleaved, every tile
  // would pay its own DRAM latency)
  float4 x[kMaxT][2][2];
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + 256 * u;
      const int row = idx >> 2, j = idx & 3;
      const bool ok = t < T && t * kTile + row < P;
      const float4* src = reinterpret_cast<const float4*>(k + (rowbase0 + (int64_t)t * kTile + row) * 32 + j * 8);
      x[t][u][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[t][u][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
    if (t >= T) break;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + 256 * u;
      const int row = idx >> 2, j = idx & 3;
      const float f[8] = {x[t][u][0].x, x[t][u][0].y, x[t][u][0].z, x[t][u][0].w, x[t][u][1].x, x[t][u][1].y, x[t][u][1].z, x[t][u][1].w};
      uint4 hi, lo;
      split8(f, hi, lo);
      const uint32_t base = kimg_addr + (uint32_t)t * kBlk;
      st_chunk(base + ptx::sw128_offset(row, j), hi);
      st_chunk(base + ptx::sw128_offset(row, 4 + j), lo);
    }
  }
}

// D[128 x 128] = K_a K_b^T with split operands: (A offset, B offset) pairs into the hi|lo images, K = 32.  The
// scores sit in an exponent, so all FOUR partial products are taken here (lo*lo included: 2 of 8 cheap K = 16 steps).
__device__ __forceinline__ void issue_s_block(uint32_t d_tmem, uint64_t dKa, uint64_t dKb, uint32_t idesc) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const uint64_t ao = (pass >= 2) ? 4 : 0;        // lo half of the row starts 64 bytes in
    const uint64_t bo = (pass & 1) ? 4 : 0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ptx::umma_bf16(d_tmem, dKa + ao + (uint64_t)(ks * 2), dKb + bo + (uint64_t)(ks * 2), idesc, (pass | ks) != 0);
  }
}

// =====================================================================================================================
namespace st {
constexpr uint32_t KIMG = 0;
constexpr uint32_t XCH = KIMG + kMaxT * kBlk;          // float[2][2][128]
constexpr uint32_t BARS = XCH + 2 * 2 * 128 * 4;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;
enum { BAR_K_FULL = 0, BAR_S_FULL = 1, BAR_S_FREE = 2 };
}  // namespace st

__global__ void __launch_bounds__(kThreads, 1)
pct_attn_stats_kernel(const float* __restrict__ k, int64_t N, int P, float* __restrict__ c2) {
  using namespace st;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  float* xch = reinterpret_cast<float*>(sm + XCH);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = (P + kTile - 1) / kTile;
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_K_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_S_FULL], 1);
    ptx::mbar_init(&bars[BAR_S_FREE], kComputeThreads);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    const uint32_t idesc = ptx::make_idesc(kFmt, 128, 128);
    const uint64_t dK = ptx::smem_desc_sw128(sm_base + KIMG);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t obj_it = 0, blk = 0;
    for (int64_t n = blockIdx.x; n < N; n += gridDim.x, ++obj_it) {
      ptx::mbar_wait(&bars[BAR_K_FULL], obj_it & 1);
      for (int it = 0; it < T; ++it, ++blk) {
        if (blk > 0) ptx::mbar_wait(&bars[BAR_S_FREE], (blk - 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          for (int jt = 0; jt < T; ++jt)
            issue_s_block(tmem_u + (uint32_t)jt * 128, dK + (uint64_t)it * (kBlk >> 4), dK + (uint64_t)jt * (kBlk >> 4), idesc);
          ptx::umma_commit(&bars[BAR_S_FULL]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3, hc = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int Ppad = T * kTile;
    const int ncols = T * 64;                       // this warp's column half
    uint32_t blk = 0;
    for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
      load_k_image(k, n * (int64_t)P, P, T, sm_base + KIMG, tid);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_K_FULL]);
      if (n + gridDim.x < N) {        // the next object's k (P rows of 128 B) -> L2 under this object's score blocks
        for (int r = tid; r < P; r += kComputeThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(k + ((n + gridDim.x) * (int64_t)P + r) * 32));
      }
      for (int it = 0; it < T; ++it, ++blk) {
        ptx::mbar_wait(&bars[BAR_S_FULL], blk & 1);
        ptx::tc_fence_after();
        // ---- pass 1: row maximum over the valid columns
        float m = -INFINITY;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem + lane_addr + (uint32_t)(hc * ncols + c0), v);
          ptx::tmem_ld_wait();
          const int jbase = hc * ncols + c0;
#pragma unroll
          for (int e = 0; e < 32; ++e) m = fmaxf(m, (jbase + e < P) ? __uint_as_float(v[e]) : -INFINITY);
        }
        xch[hc * 128 + row] = m;
        compute_barrier();
        m = fmaxf(xch[row], xch[128 + row]);
        // ---- pass 2: sum of exp2((s - m) a)
        float l = 0.f;
        const float ma = m * kAlpha;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem + lane_addr + (uint32_t)(hc * ncols + c0), v);
          ptx::tmem_ld_wait();
          const int jbase = hc * ncols + c0;
#pragma unroll
          for (int e = 0; e < 32; ++e) l += (jbase + e < P) ? ex2(fmaf(__uint_as_float(v[e]), kAlpha, -ma)) : 0.f;
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[BAR_S_FREE]);
        xch[256 + hc * 128 + row] = l;
        compute_barrier();
        if (hc == 0) {
          l = xch[256 + row] + xch[256 + 128 + row];
          const int i = it * kTile + row;
          // kept apart: their sum would be rounded at the magnitude of ma (energies of 10^4 -> ulp 4e-3 -> every
          // weight of the row off by the same 0.3 %); (s a - ma) is one fused rounding of a small difference
          c2[(n * 2) * Ppad + i] = (i < P) ? ma : INFINITY;
          c2[(n * 2 + 1) * Ppad + i] = (i < P) ? lg2(l) : 0.f;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<512>(tmem);
}

// =====================================================================================================================
namespace at {
constexpr uint32_t KIMG = 0;
constexpr uint32_t VHI = KIMG + kMaxT * kBlk;          // 2 buffers x 2 channel blocks
constexpr uint32_t VLO = VHI + 4 * kBlk;
constexpr uint32_t STAGE = VLO + 4 * kBlk;             // 196608
constexpr uint32_t C2S = STAGE + 8 * kStageFloats * 4;
constexpr uint32_t BARS = C2S + 2 * kMaxT * kTile * 4;
constexpr uint32_t TMEMPTR = BARS + 128;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;
constexpr uint32_t S_COL = 0, O_COL = 256, EHI_COL = 384, ELO_COL = 448;
enum { BAR_K_FULL = 0, BAR_V_FULL = 1 /*,2*/, BAR_S_FULL = 3 /*,4*/, BAR_S_FREE = 5 /*,6*/, BAR_E_FULL = 7, BAR_PV_DONE = 8, BAR_O_FREE = 9, kNumBars = 10 };
}  // namespace at

// kDv = false: the forward above.  kDv = true: the FIRST backward product, same structure with the roles of the two
// indices swapped --  dv[i, :] = sum_j attention[i, j] dxs[j, :]  (x_s[:, j] = sum_i x_v[:, i] attention[i, j], pct.py:224):
// work item (object, ROW block i), loop over the column blocks j, E[i, j] normalised by ITS OWN row (lane) instead of by
// the contracted index, `v` = dxs -- a gradient operand, multiplied by the object's power-of-two scale[n][0] before the
// fp16 split, the result by scale[n][1] (pct_common.cuh).
template <bool kDv>
__global__ void __launch_bounds__(kThreads, 1)
pct_attn_kernel(const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ c2, int64_t N, int P,
                float* __restrict__ xs, const float* __restrict__ scale) {
  using namespace at;
  constexpr int F = 0;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  float* c2s = reinterpret_cast<float*>(sm + C2S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = (P + kTile - 1) / kTile;
  const int Ppad = T * kTile;
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_K_FULL], kComputeThreads);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars[BAR_V_FULL + b], kComputeThreads);
      ptx::mbar_init(&bars[BAR_S_FULL + b], 1);
      ptx::mbar_init(&bars[BAR_S_FREE + b], kComputeThreads);
    }
    ptx::mbar_init(&bars[BAR_E_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_PV_DONE], 1);
    ptx::mbar_init(&bars[BAR_O_FREE], kComputeThreads);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t W = N * T;                           // work items (object, column block), column block fastest

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc_s = ptx::make_idesc(kFmt, 128, 128);
    const uint32_t idesc_pv = ptx::make_idesc(F, 128, 128) | (1u << 16);         // B (= V tile) read MN-major
    const uint64_t dK = ptx::smem_desc_sw128(sm_base + KIMG);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t u = 0, wi = 0;                          // running i-block counter, work-item counter
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x, ++wi) {
      const int jt = (int)(w % T);
      ptx::mbar_wait(&bars[BAR_K_FULL], wi & 1);
      auto issue_s = [&](uint32_t uu, int it) {      // S^T[j block, i block it] into buffer uu & 1
        const uint32_t b = uu & 1;
        if (uu >= 2) ptx::mbar_wait(&bars[BAR_S_FREE + b], ((uu - 2) >> 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_s_block(tmem_u + S_COL + b * 128, dK + (uint64_t)jt * (kBlk >> 4), dK + (uint64_t)it * (kBlk >> 4), idesc_s);
          ptx::umma_commit(&bars[BAR_S_FULL + b]);
        }
        __syncwarp();
      };
      issue_s(u, 0);
      for (int it = 0; it < T; ++it, ++u) {
        if (it + 1 < T) issue_s(u + 1, it + 1);
        const uint32_t b = u & 1;
        ptx::mbar_wait(&bars[BAR_E_FULL], u & 1);
        ptx::mbar_wait(&bars[BAR_V_FULL + b], (u >> 1) & 1);
        if (it == 0 && wi > 0) ptx::mbar_wait(&bars[BAR_O_FREE], (wi - 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t mVhi = desc_mn_sw128(sm_base + VHI + b * 2 * kBlk, kBlk);
          const uint64_t mVlo = desc_mn_sw128(sm_base + VLO + b * 2 * kBlk, kBlk);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_tm = tmem_u + ((pass == 2) ? ELO_COL : EHI_COL);
            const uint64_t bd = (pass == 1) ? mVlo : mVhi;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::umma_bf16_ts(tmem_u + O_COL, a_tm + (uint32_t)(ks * 8), bd + (uint64_t)(ks * 128), idesc_pv, (it | pass | ks) != 0);
          }
          ptx::umma_commit(&bars[BAR_PV_DONE]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, hc = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int cc = tid & 15, r0 = tid >> 4;
    float* stage = reinterpret_cast<float*>(sm + STAGE) + warp * kStageFloats;
    uint32_t u = 0;
    // PV_DONE completions this thread has already observed.  An mbarrier can only be waited on for its LATEST
    // completed phase (the parity test cannot tell phase x from phase x + 2), so every wait goes through here.
    uint32_t pv_seen = 0;
    auto wait_pv = [&](uint32_t x) {
      if (x + 1 > pv_seen) {
        ptx::mbar_wait(&bars[BAR_PV_DONE], x & 1);
        pv_seen = x + 1;
      }
    };
    auto load_v = [&](int64_t n, int it, uint32_t uu) {     // V rows of i block `it` -> buffer uu & 1 (hi / lo, 2 channel blocks)
      const uint32_t b = uu & 1;
      if (uu >= 2) wait_pv(uu - 2);
      const int64_t rowbase = n * (int64_t)P + (int64_t)it * kTile;
      const int valid = min(kTile, P - it * kTile);
      const uint32_t blk_off = (uint32_t)(b * 2 + (cc >> 3)) * kBlk;
      // all 16 loads of the thread in flight together: the tile comes from DRAM / L2 and a thread that waits for it does
      // nothing else (two warps per scheduler), so what the loop pays is ONE latency instead of four
      float4 x[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = r0 + 16 * i;
        const float4* src = reinterpret_cast<const float4*>(v + (rowbase + row) * 128 + cc * 8);
        const bool ok = row < valid;
        x[i][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[i][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float sc = kDv ? __ldg(scale + 2 * n) : 1.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = r0 + 16 * i;
        const float f[8] = {x[i][0].x * sc, x[i][0].y * sc, x[i][0].z * sc, x[i][0].w * sc, x[i][1].x * sc, x[i][1].y * sc, x[i][1].z * sc, x[i][1].w * sc};
        uint4 hi, lo;
        split8f<F>(f, hi, lo);
        const uint32_t off = blk_off + ptx::sw128_offset(row, cc & 7);
        st_chunk(sm_base + VHI + off, hi);
        st_chunk(sm_base + VLO + off, lo);
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_V_FULL + b]);
    };

    for (int64_t w = blockIdx.x; w < W; w += gridDim.x) {
      const int64_t n = w / T;
      const int jt = (int)(w - n * T);
      // the previous item's score MMAs have all completed (every S_FULL was waited for): the K image is free
      load_k_image(k, n * (int64_t)P, P, T, sm_base + KIMG, tid);
      for (int i = tid; i < 2 * Ppad; i += kComputeThreads) c2s[(i < Ppad) ? i : i - Ppad + kMaxT * kTile] = c2[n * 2 * Ppad + i];
      ptx::fence_proxy_async_smem();
      compute_barrier();                              // c2s visible to every compute thread
      ptx::mbar_arrive(&bars[BAR_K_FULL]);
      load_v(n, 0, u);
      for (int it = 0; it < T; ++it, ++u) {
        if (it + 1 < T) load_v(n, it + 1, u + 1);
        const uint32_t b = u & 1;
        // ---- E = exp2(S a - c_i) for this thread's row j and 64 columns i
        ptx::mbar_wait(&bars[BAR_S_FULL + b], (u >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t eh[32], el[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t sv[32];
          ptx::tmem_ld32(tmem + lane_addr + S_COL + b * 128 + (uint32_t)(hc * 64 + h * 32), sv);
          ptx::tmem_ld_wait();
          const float* cp = c2s + it * kTile + hc * 64 + h * 32;
          const float* cl = cp + kMaxT * kTile;
          const float cpr = c2s[jt * kTile + 32 * q + lane], clr = c2s[kMaxT * kTile + jt * kTile + 32 * q + lane];   // own row (kDv)
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float e0 = ex2(fmaf(__uint_as_float(sv[e]), kAlpha, kDv ? -cpr : -cp[e]) - (kDv ? clr : cl[e]));
            const float e1 = ex2(fmaf(__uint_as_float(sv[e + 1]), kAlpha, kDv ? -cpr : -cp[e + 1]) - (kDv ? clr : cl[e + 1]));
            split2f<F>(e0, e1, eh[h * 16 + e / 2], el[h * 16 + e / 2]);
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[BAR_S_FREE + b]);
        if (u >= 1) wait_pv(u - 1);      // the previous block's MMAs have consumed E
        ptx::tc_fence_after();
        ptx::tmem_st32(tmem + lane_addr + EHI_COL + (uint32_t)(hc * 32), eh);
        ptx::tmem_st32(tmem + lane_addr + ELO_COL + (uint32_t)(hc * 32), el);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[BAR_E_FULL]);
      }
      // ---- O[j, :] -> x_s rows
      wait_pv(u - 1);
      ptx::tc_fence_after();
      const int64_t rowbase = n * (int64_t)P + (int64_t)jt * kTile;
      const int nvalid = max(0, min(32, P - jt * kTile - 32 * q));
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t ov[16];
        ptx::tmem_ld16(tmem + lane_addr + O_COL + (uint32_t)(hc * 64 + ch * 16), ov);
        ptx::tmem_ld_wait();
        float f[16];
        const float osc = kDv ? __ldg(scale + 2 * n + 1) : 1.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(ov[e]) * osc;
        float s = 0.f, qv = 0.f;
        stage_store16(stage, f, xs + (rowbase + 32 * q) * 128 + hc * 64 + ch * 16, 128, nvalid, lane, s, qv, false);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_O_FREE]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<512>(tmem);
}


// =====================================================================================================================
// Second backward product of the SA layer: the gradient of the shared q / k projection output.
//   energy = k k^T / sqrt(32) (symmetric),  A[i,j] = softmax_j,  dA[i,j] = v_i . dxs_j,  delta_i = sum_j A[i,j] dA[i,j] = v_i . dv_i
//   dS[i,j] = A[i,j] (dA[i,j] - delta_i)                     dk_i = (1/sqrt(32)) sum_j (dS[i,j] + dS[j,i]) k_j
// Both halves are produced in "row i" orientation, so nothing is transposed or scattered:
//   kCol = false:  T[i,j] = exp2(S a - M_i) (v_i . dxs_j - delta_i)        fixed rows = v,   streamed rows = dxs
//   kCol = true :  T[i,j] = exp2(S a - M_j) (dxs_i . v_j - delta_j)        fixed rows = dxs, streamed rows = v   (= dS[j,i])
// and dk_i += T[i, block b] k_b.  Work item = (object, row block); per column block b the tensor pipe computes S (fp16
// pairs, four partial products: exponent) and the 128-deep product (fp16 pairs, three passes; the dxs rows scaled by the
// object's power of two) into tensor memory, the
// compute warps form T, split it into fp16 pairs BACK INTO TENSOR MEMORY as the A operand of the last product, whose B
// operand is the [hi | lo] image of k_b read MN-major with N = 64: columns 0-31 of the accumulator collect T k.hi,
// columns 32-63 T k.lo (two passes: T.hi, T.lo), added in the epilogue.  Single-buffered: correctness first.
// delta_i.  Mathematically v_i . dv_i, but it must NOT be computed that way: with a peaked softmax (A[i,j*] ~ 1)
// dA[i,j*] - delta_i is a small difference of two numbers that each carry the 1e-5 error of a split-operand product, and
// only if delta is summed from the SAME dA values the kernel subtracts it from do those errors cancel (the error of the
// difference is then eps (1 - A), not eps).  So the kCol = false kernel runs the S / dA products twice per work item:
// a first sweep over the column blocks accumulates delta_i = sum_j E[i,j] dA[i,j] thread-locally (and writes it out
// for the kCol = true launch, whose dA[j,i] = v_j . dxs_i is the same set of partial products), the second forms T.
namespace dk {
constexpr uint32_t KA = 0;                               // own block's k rows, fp16 [hi | lo]
constexpr uint32_t KB16 = KA + kBlk;                     // block b's k rows, fp16 [hi | lo]
constexpr uint32_t FHI = KB16 + kBlk;                    // fixed 128-channel rows (2 channel blocks), fp16 hi / lo
constexpr uint32_t FLO = FHI + 2 * kBlk;
constexpr uint32_t GHI = FLO + 2 * kBlk;                 // streamed rows
constexpr uint32_t GLO = GHI + 2 * kBlk;
constexpr uint32_t STAGE = GLO + 2 * kBlk;               // 163840
constexpr uint32_t NRM = STAGE + 8 * kStageFloats * 4;   // float[3][512]: max * a, log2 sum, delta
constexpr uint32_t XCH = NRM + 3 * kMaxT * kTile * 4;     // float[2][128]
constexpr uint32_t BARS = XCH + 2 * 128 * 4;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;
constexpr uint32_t S_COL = 0, DA_COL = 128, THI_COL = 256, TLO_COL = 320, DK_COL = 384;
enum { BAR_F_FULL = 0, BAR_G_FULL = 1, BAR_SD_FULL = 2, BAR_T_FULL = 3, BAR_DK_DONE = 4, kNumBars = 5 };
}  // namespace dk

// one tile of 128 rows x 32 channels of k -> [hi | lo] image (format kF)
template <int kF>
__device__ __forceinline__ void load_k_tile(const float* __restrict__ k, int64_t rowbase, int valid, uint32_t addr, int tid) {
  float4 x[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 2, j = idx & 3;
    const float4* src = reinterpret_cast<const float4*>(k + (rowbase + row) * 32 + j * 8);
    const bool ok = row < valid;
    x[u][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[u][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 2, j = idx & 3;
    const float f[8] = {x[u][0].x, x[u][0].y, x[u][0].z, x[u][0].w, x[u][1].x, x[u][1].y, x[u][1].z, x[u][1].w};
    uint4 hi, lo;
    split8f<kF>(f, hi, lo);
    st_chunk(addr + ptx::sw128_offset(row, j), hi);
    st_chunk(addr + ptx::sw128_offset(row, 4 + j), lo);
  }
}

// one tile of 128 rows x 128 channels -> K-major hi / lo images (2 channel blocks each), format kF; the 16 loads of a
// thread are all in flight together (the tile usually comes from L2: latency, not bandwidth, is what the loop pays)
template <int kF>
__device__ __forceinline__ void load_tile128(const float* __restrict__ src, int64_t rowbase, int valid, uint32_t hi_addr, uint32_t lo_addr,
                                             int tid, float mul) {
  const int cc = tid & 15, r0 = tid >> 4;
  const uint32_t blk_off = (uint32_t)(cc >> 3) * kBlk;
  float4 x[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + 16 * i;
    const float4* p = reinterpret_cast<const float4*>(src + (rowbase + row) * 128 + cc * 8);
    const bool ok = row < valid;
    x[i][0] = ok ? __ldg(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[i][1] = ok ? __ldg(p + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + 16 * i;
    const float f[8] = {x[i][0].x * mul, x[i][0].y * mul, x[i][0].z * mul, x[i][0].w * mul,
                        x[i][1].x * mul, x[i][1].y * mul, x[i][1].z * mul, x[i][1].w * mul};
    uint4 hi, lo;
    split8f<kF>(f, hi, lo);
    const uint32_t off = blk_off + ptx::sw128_offset(row, cc & 7);
    st_chunk(hi_addr + off, hi);
    st_chunk(lo_addr + off, lo);
  }
}

template <bool kCol>
__global__ void __launch_bounds__(kThreads, 1)
pct_attn_dk_kernel(const float* __restrict__ k, const float* __restrict__ fixed, const float* __restrict__ streamed,
                   const float* __restrict__ c2, const float* __restrict__ delta_in, float* __restrict__ delta_out,
                   const float* __restrict__ scale, int64_t N, int P, float* __restrict__ dk_out, int sweeps) {
  using namespace dk;
  const int kPhases = kCol ? 1 : sweeps;       // kCol = false: 2 = delta from a first sweep (written to delta_out), 1 = delta_in
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  float* nrm = reinterpret_cast<float*>(sm + NRM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = (P + kTile - 1) / kTile;
  const int Ppad = T * kTile;
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_F_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_G_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_SD_FULL], 1);
    ptx::mbar_init(&bars[BAR_T_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_DK_DONE], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t W = N * T;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc_s = ptx::make_idesc(0, 128, 128);
    const uint32_t idesc_da = ptx::make_idesc(0, 128, 128);
    const uint32_t idesc_dk = ptx::make_idesc(0, 128, 64) | (1u << 16);          // B (= k_b image) read MN-major
    const uint64_t dKA = ptx::smem_desc_sw128(sm_base + KA), dKB = ptx::smem_desc_sw128(sm_base + KB16);
    const uint64_t dFhi = ptx::smem_desc_sw128(sm_base + FHI), dFlo = ptx::smem_desc_sw128(sm_base + FLO);
    const uint64_t dGhi = ptx::smem_desc_sw128(sm_base + GHI), dGlo = ptx::smem_desc_sw128(sm_base + GLO);
    const uint64_t mKB = desc_mn_sw128(sm_base + KB16, kBlk);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t u = 0, ud = 0, wi = 0;
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x, ++wi) {
      ptx::mbar_wait(&bars[BAR_F_FULL], wi & 1);
      for (int ph = 0; ph < kPhases; ++ph) {
        const bool dk_phase = ph == kPhases - 1;
        for (int b = 0; b < T; ++b, ++u) {
          ptx::mbar_wait(&bars[BAR_G_FULL], u & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_s_block(tmem_u + S_COL, dKA, dKB, idesc_s);
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint64_t ad = (pass == 1) ? dFlo : dFhi;
              const uint64_t bd = (pass == 2) ? dGlo : dGhi;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t o = (uint64_t)((ks >> 2) * (kBlk >> 4) + (ks & 3) * 2);
                ptx::umma_bf16(tmem_u + DA_COL, ad + o, bd + o, idesc_da, (pass | ks) != 0);
              }
            }
            ptx::umma_commit(&bars[BAR_SD_FULL]);
          }
          __syncwarp();
          if (!dk_phase) continue;
          ptx::mbar_wait(&bars[BAR_T_FULL], ud & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const uint32_t a_tm = tmem_u + (pass ? TLO_COL : THI_COL);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16_ts(tmem_u + DK_COL, a_tm + (uint32_t)(ks * 8), mKB + (uint64_t)(ks * 128), idesc_dk, (b | pass | ks) != 0);
            }
            ptx::umma_commit(&bars[BAR_DK_DONE]);
          }
          __syncwarp();
          ++ud;
        }
      }
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, hc = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    float* stage = reinterpret_cast<float*>(sm + STAGE) + warp * kStageFloats;
    const float* ma_s = nrm;
    const float* lg_s = nrm + kMaxT * kTile;
    const float* de_s = nrm + 2 * kMaxT * kTile;
    float* xch = reinterpret_cast<float*>(sm + XCH);
    uint32_t u = 0, ud = 0;
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x) {
      const int64_t n = w / T;
      const int a = (int)(w - n * T);
      const int64_t obase = n * (int64_t)P;
      // every MMA of the previous item has completed (its last DK_DONE was waited for in the epilogue)
      load_k_tile<0>(k, obase + (int64_t)a * kTile, min(kTile, P - a * kTile), sm_base + KA, tid);
      const float gsc = __ldg(scale + 2 * n), ginv = __ldg(scale + 2 * n + 1);     // the dxs operand carries gsc
      load_tile128<0>(fixed, obase + (int64_t)a * kTile, min(kTile, P - a * kTile), sm_base + FHI, sm_base + FLO, tid, kCol ? gsc : 1.f);
      for (int i = tid; i < Ppad; i += kComputeThreads) {
        nrm[i] = c2[(n * 2) * Ppad + i];
        nrm[kMaxT * kTile + i] = c2[(n * 2 + 1) * Ppad + i];
        nrm[2 * kMaxT * kTile + i] = ((kCol || kPhases == 1) && i < P) ? delta_in[obase + i] : 0.f;
      }
      ptx::fence_proxy_async_smem();
      compute_barrier();
      ptx::mbar_arrive(&bars[BAR_F_FULL]);
      const float mr = ma_s[a * kTile + row], lr = lg_s[a * kTile + row];
      float dr = de_s[a * kTile + row];                 // delta of this thread's row (two sweeps: replaced after the first)
      for (int ph = 0; ph < kPhases; ++ph) {
        const bool dk_phase = ph == kPhases - 1;
        float dacc = 0.f;
        for (int b = 0; b < T; ++b, ++u) {
          if (dk_phase && b >= 1) ptx::mbar_wait(&bars[BAR_DK_DONE], (ud - 1) & 1);      // k_b images, G and T are free again
          const int validb = min(kTile, P - b * kTile);
          load_k_tile<0>(k, obase + (int64_t)b * kTile, validb, sm_base + KB16, tid);
          load_tile128<0>(streamed, obase + (int64_t)b * kTile, validb, sm_base + GHI, sm_base + GLO, tid, kCol ? 1.f : gsc);
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&bars[BAR_G_FULL]);
          ptx::mbar_wait(&bars[BAR_SD_FULL], u & 1);
          ptx::tc_fence_after();
          uint32_t th[32], tl[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t sv[32], dv[32];
            ptx::tmem_ld32(tmem + lane_addr + S_COL + (uint32_t)(hc * 64 + h * 32), sv);
            ptx::tmem_ld32(tmem + lane_addr + DA_COL + (uint32_t)(hc * 64 + h * 32), dv);
            ptx::tmem_ld_wait();
            const int cb = b * kTile + hc * 64 + h * 32;
            if (!dk_phase) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                dacc = fmaf(ex2(fmaf(__uint_as_float(sv[e]), kAlpha, -mr) - lr), __uint_as_float(dv[e]), dacc);
            } else {
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                float t0, t1;
                if (kCol) {
                  t0 = ex2(fmaf(__uint_as_float(sv[e]), kAlpha, -ma_s[cb + e]) - lg_s[cb + e]) * (__uint_as_float(dv[e]) - de_s[cb + e]);
                  t1 = ex2(fmaf(__uint_as_float(sv[e + 1]), kAlpha, -ma_s[cb + e + 1]) - lg_s[cb + e + 1]) * (__uint_as_float(dv[e + 1]) - de_s[cb + e + 1]);
                } else {
                  t0 = ex2(fmaf(__uint_as_float(sv[e]), kAlpha, -mr) - lr) * (__uint_as_float(dv[e]) - dr);
                  t1 = ex2(fmaf(__uint_as_float(sv[e + 1]), kAlpha, -mr) - lr) * (__uint_as_float(dv[e + 1]) - dr);
                }
                split2f<0>(t0, t1, th[h * 16 + e / 2], tl[h * 16 + e / 2]);
              }
            }
          }
          if (!dk_phase) {
            ptx::tc_fence_before();         // the next G_FULL arrival orders the next S / dA products after these reads
            continue;
          }
          ptx::tmem_st32(tmem + lane_addr + THI_COL + (uint32_t)(hc * 32), th);
          ptx::tmem_st32(tmem + lane_addr + TLO_COL + (uint32_t)(hc * 32), tl);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bars[BAR_T_FULL]);
          ++ud;
        }
        if (!dk_phase) {                    // delta_i: the two column halves of the row
          xch[hc * 128 + row] = dacc;
          compute_barrier();
          dr = xch[row] + xch[128 + row];
          if (hc == 0 && a * kTile + row < P) delta_out[obase + (int64_t)a * kTile + row] = dr;
          compute_barrier();                // xch is rewritten by the next item
        }
      }
      // ---- dk rows of this block: (T k.hi) + (T k.lo), scaled by 1/sqrt(32)
      ptx::mbar_wait(&bars[BAR_DK_DONE], (ud - 1) & 1);
      ptx::tc_fence_after();
      uint32_t d0[16], d1[16];
      ptx::tmem_ld16(tmem + lane_addr + DK_COL + (uint32_t)(hc * 16), d0);
      ptx::tmem_ld16(tmem + lane_addr + DK_COL + (uint32_t)(32 + hc * 16), d1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] = (__uint_as_float(d0[e]) + __uint_as_float(d1[e])) * (0.17677669529663687f * ginv);
      const int nvalid = max(0, min(32, P - a * kTile - 32 * q));
      float s_ = 0.f, q_ = 0.f;
      stage_store16(stage, f, dk_out + (obase + (int64_t)a * kTile + 32 * q) * 32 + hc * 16, 32, nvalid, lane, s_, q_, false);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<512>(tmem);
}

}  // namespace
}  // namespace pct
}  // namespace sga

namespace sga {
namespace pct {
// pct_attn2.cu: two CTAs per SM (default); SGA_PCT_ATTN=v1 selects the kernels of this file
int attn2_fwd(const float* k, const float* v, const float* c2, int64_t N, int P, float* xs, cudaStream_t st);
int attn2_dv(const float* k, const float* dxs, const float* c2, const float* scale, int64_t N, int P, float* dv, double* colsum,
             float* absmax, cudaStream_t st);
int attn2_dk(const float* k, const float* fixed, const float* streamed, const float* c2, float* delta, const float* scale, int64_t N,
             int P, int by_col, int delta_sweep, float* dk_out, cudaStream_t st);
static bool attn_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGA_PCT_ATTN");
    v = (e && e[0] == 'v' && e[1] == '1') ? 1 : 0;
  }
  return v == 1;
}
}  // namespace pct
}  // namespace sga

extern "C" int sga_pct_attn_stats(const float* k, int64_t N, int P, float* c2, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(k && c2 && P >= 1 && P <= sga::pct::kMaxT * sga::pct::kTile, "sga_pct_attn_stats: P=%d (1..512)", P);
  SGA_REQUIRE(((uintptr_t)k & 15) == 0, "sga_pct_attn_stats: k must be 16-byte aligned");
  using namespace sga::pct;
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st::SMEM_BYTES));
    attr_done = true;
  }
  int grid = sga::sm_count();
  if ((int64_t)grid > N) grid = (int)N;
  pct_attn_stats_kernel<<<grid, kThreads, st::SMEM_BYTES, (cudaStream_t)stream>>>(k, N, P, c2);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_pct_attn(const float* k, const float* v, const float* c2, int64_t N, int P, float* xs, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(k && v && c2 && xs && P >= 1 && P <= sga::pct::kMaxT * sga::pct::kTile, "sga_pct_attn: P=%d (1..512)", P);
  SGA_REQUIRE((((uintptr_t)k | (uintptr_t)v) & 15) == 0, "sga_pct_attn: k / v must be 16-byte aligned");
  using namespace sga::pct;
  if (!attn_v1()) return attn2_fwd(k, v, c2, N, P, xs, (cudaStream_t)stream);
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at::SMEM_BYTES));
    attr_done = true;
  }
  const int T = (P + kTile - 1) / kTile;
  int64_t W = N * T;
  int grid = sga::sm_count();
  if ((int64_t)grid > W) grid = (int)W;
  pct_attn_kernel<false><<<grid, kThreads, at::SMEM_BYTES, (cudaStream_t)stream>>>(k, v, c2, N, P, xs, nullptr);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

/* dv [N,P,128] = attention dxs per object (first product of the SA backward; see pct_attn_kernel<true>) */
extern "C" int sga_pct_attn_bwd_dv(const float* k, const float* dxs, const float* c2, const float* scale, int64_t N, int P, float* dv,
                                   double* dv_colsum, float* dv_absmax, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(k && dxs && c2 && dv && scale && P >= 1 && P <= sga::pct::kMaxT * sga::pct::kTile, "sga_pct_attn_bwd_dv: P=%d (1..512)", P);
  SGA_REQUIRE((((uintptr_t)k | (uintptr_t)dxs) & 15) == 0, "sga_pct_attn_bwd_dv: k / dxs must be 16-byte aligned");
  using namespace sga::pct;
  if (!attn_v1()) return attn2_dv(k, dxs, c2, scale, N, P, dv, dv_colsum, dv_absmax, (cudaStream_t)stream);
  SGA_REQUIRE(!dv_colsum && !dv_absmax, "sga_pct_attn_bwd_dv: the fused column sums / maxima need the two-CTA kernel (unset SGA_PCT_ATTN)");
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at::SMEM_BYTES));
    attr_done = true;
  }
  const int T = (P + kTile - 1) / kTile;
  int64_t W = N * T;
  int grid = sga::sm_count();
  if ((int64_t)grid > W) grid = (int)W;
  pct_attn_kernel<true><<<grid, kThreads, at::SMEM_BYTES, (cudaStream_t)stream>>>(k, dxs, c2, N, P, dv, scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

/* dk halves of the SA backward (see pct_attn_dk_kernel): by_col = 0: fixed = v, streamed = dxs (row half; also WRITES
 * delta [N,P], in units of the object's scale, when delta_sweep != 0 -- otherwise it READS it like by_col = 1 does);
 * by_col = 1: fixed = dxs, streamed = v (column half; READS delta).
 * scale [N,2] = sga_pct_pow2_scale(dxs, v).  dk_out [N,P,32] is overwritten. */
extern "C" int sga_pct_attn_bwd_dk(const float* k, const float* fixed, const float* streamed, const float* c2, float* delta,
                                   const float* scale, int64_t N, int P, int by_col, int delta_sweep, float* dk_out, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(k && fixed && streamed && c2 && delta && scale && dk_out && P >= 1 && P <= sga::pct::kMaxT * sga::pct::kTile,
              "sga_pct_attn_bwd_dk: P=%d (1..512)", P);
  SGA_REQUIRE((((uintptr_t)k | (uintptr_t)fixed | (uintptr_t)streamed) & 15) == 0, "sga_pct_attn_bwd_dk: operands must be 16-byte aligned");
  using namespace sga::pct;
  if (!attn_v1()) return attn2_dk(k, fixed, streamed, c2, delta, scale, N, P, by_col, delta_sweep, dk_out, (cudaStream_t)stream);
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn_dk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dk::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_attn_dk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dk::SMEM_BYTES));
    attr_done = true;
  }
  const int T = (P + kTile - 1) / kTile;
  int64_t W = N * T;
  int grid = sga::sm_count();
  if ((int64_t)grid > W) grid = (int)W;
  if (by_col) pct_attn_dk_kernel<true><<<grid, kThreads, dk::SMEM_BYTES, (cudaStream_t)stream>>>(k, fixed, streamed, c2, delta, nullptr, scale, N, P, dk_out, 1);
  else pct_attn_dk_kernel<false><<<grid, kThreads, dk::SMEM_BYTES, (cudaStream_t)stream>>>(k, fixed, streamed, c2, delta, delta, scale, N, P, dk_out, delta_sweep ? 2 : 1);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
