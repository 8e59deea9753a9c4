// NaivePCT self-attention, second generation of the x_v * attention product (forward, and the dv product of the
// backward): TWO co-resident CTAs per SM.
//
// ncu of the first kernel (pct_attn.cu: one CTA per SM, S double-buffered, whole-object K image) showed issue slots 20 %
// busy and the long scoreboard as the top stall: per 128 x 128 block the chain  tile load -> S product -> exp / split
// epilogue -> E V product  is serial in every thread, and with two warps per scheduler nothing covers a thread that waits
// (profiles/r2_ncu_pct_attention_stalls.txt).  Loader warps of their own did not help (single-buffered operands keep the
// chain serial whoever loads).  What does: a second, independent chain on the same SM.  To fit two CTAs
//   * shared memory: only the k tiles of the own and the current block (2 x 16 KiB) and ONE V tile (64 KiB); the output
//     staging patch reuses the V tile -> 101 KiB per CTA;
//   * tensor memory: 256 columns per CTA -- the score block S (128 columns) and the output accumulator (128).  The
//     attention weights E (fp16 hi / lo, the A operand of the second product) are written OVER S: a thread reads the two
//     32-column halves of its 64 score columns before it overwrites them, and each epilogue warp's E lands inside its own
//     64-column range, so no warp waits for another;
//   * registers: 112 per thread (65536 / 576); the epilogue works on 32 columns at a time.
// Same arithmetic as pct_attn.cu (fp16 pairs, four partial products for the scores, three for E V; softmax normaliser kept
// in two parts), bit-identical results are not required but the tests hold both to the same 2e-5.
#include "pct_common.cuh"

namespace sga {
namespace pct {
namespace {

constexpr int kMaxT2 = 4;                                  // P <= 512
constexpr float kAlpha2 = 1.4426950408889634f * 0.17677669529663687f;   // log2(e) / sqrt(32)

namespace a2 {
constexpr uint32_t KA = 0;                                 // own block's k rows, fp16 [hi | lo]
constexpr uint32_t KB = KA + kBlk;                         // current block's k rows
constexpr uint32_t VHI = KB + kBlk;                        // current block's V rows: 2 channel blocks hi, 2 lo
constexpr uint32_t VLO = VHI + 2 * kBlk;
constexpr uint32_t C2S = VLO + 2 * kBlk;                   // 98304: float[2][512]
constexpr uint32_t BARS = C2S + 2 * kMaxT2 * kTile * 4;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;       // 103504
constexpr uint32_t SE_COL = 0, O_COL = 128;
enum { BAR_LD_FULL = 0, BAR_S_FULL = 1, BAR_E_FULL = 2, BAR_PV_DONE = 3, kNumBars = 4 };
static_assert(8 * kStageFloats * 4 <= 4 * kBlk, "the output staging patch must fit the V tile");
}  // namespace a2

__device__ __forceinline__ void issue_s_block2(uint32_t d_tmem, uint64_t dKa, uint64_t dKb, uint32_t idesc) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const uint64_t ao = (pass >= 2) ? 4 : 0;        // lo half of the row starts 64 bytes in
    const uint64_t bo = (pass & 1) ? 4 : 0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ptx::umma_bf16(d_tmem, dKa + ao + (uint64_t)(ks * 2), dKb + bo + (uint64_t)(ks * 2), idesc, (pass | ks) != 0);
  }
}

// k tile (128 x 32) and, optionally, a 128 x 128 tile: every load of the thread in flight before its first store
template <bool kWithTile>
__device__ __forceinline__ void load_tiles2(const float* __restrict__ k, int64_t rowbase, int valid, uint32_t k_addr,
                                            const float* __restrict__ v, uint32_t vhi_addr, uint32_t vlo_addr, float mul, int tid) {
  float4 kx[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 2, j = idx & 3;
    const float4* src = reinterpret_cast<const float4*>(k + (rowbase + row) * 32 + j * 8);
    const bool ok = row < valid;
    kx[u][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
    kx[u][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int cc = tid & 15, r0 = tid >> 4;
  float4 x[kWithTile ? 8 : 1][2];
  if (kWithTile) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + 16 * i;
      const float4* src = reinterpret_cast<const float4*>(v + (rowbase + row) * 128 + cc * 8);
      const bool ok = row < valid;
      x[i][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[i][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 2, j = idx & 3;
    const float f[8] = {kx[u][0].x, kx[u][0].y, kx[u][0].z, kx[u][0].w, kx[u][1].x, kx[u][1].y, kx[u][1].z, kx[u][1].w};
    uint4 hi, lo;
    split8(f, hi, lo);
    st_chunk(k_addr + ptx::sw128_offset(row, j), hi);
    st_chunk(k_addr + ptx::sw128_offset(row, 4 + j), lo);
  }
  if (kWithTile) {
    const uint32_t blk_off = (uint32_t)(cc >> 3) * kBlk;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + 16 * i;
      const float f[8] = {x[i][0].x * mul, x[i][0].y * mul, x[i][0].z * mul, x[i][0].w * mul,
                          x[i][1].x * mul, x[i][1].y * mul, x[i][1].z * mul, x[i][1].w * mul};
      uint4 hi, lo;
      split8(f, hi, lo);
      const uint32_t off = blk_off + ptx::sw128_offset(row, cc & 7);
      st_chunk(vhi_addr + off, hi);
      st_chunk(vlo_addr + off, lo);
    }
  }
}

// kDv = false: xs[a, :] = sum_b softmax(energy)[b, a] v[b, :]     (lanes = own block a, normaliser of the CONTRACTED row b)
// kDv = true : dv[a, :] = sum_b softmax(energy)[a, b] dxs[b, :]   (normaliser of the own row a; `v` = dxs, scaled per object)
template <bool kDv>
__global__ void __launch_bounds__(kThreads, 2)
pct_attn2_kernel(const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ c2, int64_t N, int P,
                 float* __restrict__ out, const float* __restrict__ scale) {
  using namespace a2;
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
    ptx::mbar_init(&bars[BAR_LD_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_S_FULL], 1);
    ptx::mbar_init(&bars[BAR_E_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_PV_DONE], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t W = N * T;                           // work items (object, own block), block fastest

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc_s = ptx::make_idesc(kFmt, 128, 128);
    const uint32_t idesc_pv = ptx::make_idesc(kFmt, 128, 128) | (1u << 16);      // B (= V tile) read MN-major
    const uint64_t dKA = ptx::smem_desc_sw128(sm_base + KA), dKB = ptx::smem_desc_sw128(sm_base + KB);
    const uint64_t mVhi = desc_mn_sw128(sm_base + VHI, kBlk), mVlo = desc_mn_sw128(sm_base + VLO, kBlk);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t u = 0;
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x) {
      for (int b = 0; b < T; ++b, ++u) {
        ptx::mbar_wait(&bars[BAR_LD_FULL], u & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_s_block2(tmem_u + SE_COL, dKA, dKB, idesc_s);
          ptx::umma_commit(&bars[BAR_S_FULL]);
        }
        __syncwarp();
        ptx::mbar_wait(&bars[BAR_E_FULL], u & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t lo_off = (pass == 2) ? 32u : 0u;           // E.lo of a 64-column range sits 32 columns after its E.hi
            const uint64_t bd = (pass == 1) ? mVlo : mVhi;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::umma_bf16_ts(tmem_u + O_COL, tmem_u + SE_COL + (uint32_t)((ks >> 2) * 64 + (ks & 3) * 8) + lo_off,
                                bd + (uint64_t)(ks * 128), idesc_pv, (b | pass | ks) != 0);
          }
          ptx::umma_commit(&bars[BAR_PV_DONE]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, hc = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const uint32_t c0 = (uint32_t)(hc * 64);
    float* stage = reinterpret_cast<float*>(sm + VHI) + warp * kStageFloats;     // reuses the V tile once the item's products are done
    uint32_t u = 0;
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x) {
      const int64_t n = w / T;
      const int a = (int)(w - n * T);
      const int64_t obase = n * (int64_t)P;
      const float gsc = kDv ? __ldg(scale + 2 * n) : 1.f;
      // the previous item's products have all completed (its last PV_DONE was waited for); c2s / k_a are free.  The V
      // tile -- the previous item's output staging patch -- is rewritten only after the barrier below.
      for (int i = tid; i < 2 * Ppad; i += kComputeThreads) c2s[(i < Ppad) ? i : i - Ppad + kMaxT2 * kTile] = c2[n * 2 * Ppad + i];
      load_tiles2<false>(k, obase + (int64_t)a * kTile, min(kTile, P - a * kTile), sm_base + KA, nullptr, 0, 0, 1.f, tid);
      compute_barrier();                               // c2s visible to every compute thread; every warp has stored its rows
      const float cpr = c2s[a * kTile + row], clr = c2s[kMaxT2 * kTile + a * kTile + row];     // own row (kDv)
      for (int b = 0; b < T; ++b, ++u) {
        if (b >= 1) ptx::mbar_wait(&bars[BAR_PV_DONE], (u - 1) & 1);            // the k_b / V tiles and the S / E columns are free
        load_tiles2<true>(k, obase + (int64_t)b * kTile, min(kTile, P - b * kTile), sm_base + KB, v, sm_base + VHI, sm_base + VLO, gsc, tid);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[BAR_LD_FULL]);
        ptx::mbar_wait(&bars[BAR_S_FULL], u & 1);
        ptx::tc_fence_after();
        // ---- E = exp2(S a - normaliser) for this thread's row and its 64 columns, written over S:
        //      [c0, c0+32) = E.hi of the 64 columns (two per 32-bit column), [c0+32, c0+64) = E.lo
        const float* cp = c2s + b * kTile + c0;
        const float* cl = cp + kMaxT2 * kTile;
        uint32_t s1[32], eh[16], el[16];
        ptx::tmem_ld32(tmem + lane_addr + SE_COL + c0 + 32, s1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float e0 = ex2(fmaf(__uint_as_float(s1[e]), kAlpha2, kDv ? -cpr : -cp[32 + e]) - (kDv ? clr : cl[32 + e]));
          const float e1 = ex2(fmaf(__uint_as_float(s1[e + 1]), kAlpha2, kDv ? -cpr : -cp[32 + e + 1]) - (kDv ? clr : cl[32 + e + 1]));
          split2f<0>(e0, e1, eh[e / 2], el[e / 2]);
        }
        ptx::tmem_ld32(tmem + lane_addr + SE_COL + c0, s1);                     // the first half, before anything is overwritten
        ptx::tmem_ld_wait();
        ptx::tmem_st16(tmem + lane_addr + SE_COL + c0 + 16, eh);
        ptx::tmem_st16(tmem + lane_addr + SE_COL + c0 + 48, el);
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float e0 = ex2(fmaf(__uint_as_float(s1[e]), kAlpha2, kDv ? -cpr : -cp[e]) - (kDv ? clr : cl[e]));
          const float e1 = ex2(fmaf(__uint_as_float(s1[e + 1]), kAlpha2, kDv ? -cpr : -cp[e + 1]) - (kDv ? clr : cl[e + 1]));
          split2f<0>(e0, e1, eh[e / 2], el[e / 2]);
        }
        ptx::tmem_st16(tmem + lane_addr + SE_COL + c0, eh);
        ptx::tmem_st16(tmem + lane_addr + SE_COL + c0 + 32, el);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[BAR_E_FULL]);
      }
      // ---- O[a, :] -> output rows (through the staging patch in the V tile: every product has completed)
      ptx::mbar_wait(&bars[BAR_PV_DONE], (u - 1) & 1);
      ptx::tc_fence_after();
      const int64_t rowbase = obase + (int64_t)a * kTile;
      const int nvalid = max(0, min(32, P - a * kTile - 32 * q));
      const float osc = kDv ? __ldg(scale + 2 * n + 1) : 1.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t ov[16];
        ptx::tmem_ld16(tmem + lane_addr + O_COL + (uint32_t)(hc * 64 + ch * 16), ov);
        ptx::tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(ov[e]) * osc;
        float s = 0.f, qv = 0.f;
        stage_store16(stage, f, out + (rowbase + 32 * q) * 128 + hc * 64 + ch * 16, 128, nvalid, lane, s, qv, false);
      }
      ptx::tc_fence_before();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<256>(tmem);
}

template <bool kDv>
int attn2_launch(const float* k, const float* v, const float* c2, int64_t N, int P, float* out, const float* scale, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn2_kernel<kDv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a2::SMEM_BYTES));
    attr_done = true;
  }
  const int T = (P + kTile - 1) / kTile;
  const int64_t W = N * T;
  int64_t grid = 2 * (int64_t)sm_count();
  if (grid > W) grid = W;
  pct_attn2_kernel<kDv><<<(unsigned)grid, kThreads, a2::SMEM_BYTES, st>>>(k, v, c2, N, P, out, scale);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace

int attn2_fwd(const float* k, const float* v, const float* c2, int64_t N, int P, float* xs, cudaStream_t st) {
  return attn2_launch<false>(k, v, c2, N, P, xs, nullptr, st);
}
int attn2_dv(const float* k, const float* dxs, const float* c2, const float* scale, int64_t N, int P, float* dv, cudaStream_t st) {
  return attn2_launch<true>(k, dxs, c2, N, P, dv, scale, st);
}

}  // namespace pct
}  // namespace sga
