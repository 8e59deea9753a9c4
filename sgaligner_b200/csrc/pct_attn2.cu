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

__device__ int g_attn2_pf = 1;

// kDv = false: xs[a, :] = sum_b softmax(energy)[b, a] v[b, :]     (lanes = own block a, normaliser of the CONTRACTED row b)
// kDv = true : dv[a, :] = sum_b softmax(energy)[a, b] dxs[b, :]   (normaliser of the own row a; `v` = dxs, scaled per object)
template <bool kDv>
__global__ void __launch_bounds__(kThreads, 2)
pct_attn2_kernel(const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ c2, int64_t N, int P,
                 float* __restrict__ out, const float* __restrict__ scale, double* __restrict__ colsum, float* __restrict__ absmax) {
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
    double dcs[4] = {0, 0, 0, 0};                      // kDv: column sums of dv (= d v_conv.bias), this lane's column of each chunk
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
        if (g_attn2_pf) {      // the next step's k / V tile -> L2 while this step's products and exponentials run
          int64_t nb = obase + (int64_t)(b + 1) * kTile;
          int nvalid = min(kTile, P - (b + 1) * kTile);
          if (b + 1 == T) {
            const int64_t wn = w + gridDim.x;
            nb = (wn / T) * (int64_t)P;
            nvalid = (wn < W) ? min(kTile, P) : 0;
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int line = tid + 256 * j;           // 512 lines of 128 B: row = line / 4
            if ((line >> 2) < nvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(v + (nb + (line >> 2)) * 128 + (line & 3) * 32));
          }
          if (tid < nvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(k + (nb + tid) * 32));
        }
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
      float amax = 0.f;
      const bool stats = kDv && colsum != nullptr;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t ov[16];
        ptx::tmem_ld16(tmem + lane_addr + O_COL + (uint32_t)(hc * 64 + ch * 16), ov);
        ptx::tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          f[e] = __uint_as_float(ov[e]) * osc;
          if (kDv && lane < nvalid) amax = fmaxf(amax, fabsf(f[e]));
        }
        float s = 0.f, qv = 0.f;
        stage_store16(stage, f, out + (rowbase + 32 * q) * 128 + hc * 64 + ch * 16, 128, nvalid, lane, s, qv, stats);
        if (stats) dcs[ch] += (double)s;
      }
      if (kDv && absmax != nullptr) {                  // the object's largest |dv|: the scale of the next product (dv W_v) comes from it
        amax = warp_max(amax);
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(absmax + n), __float_as_uint(amax));
      }
      ptx::tc_fence_before();
    }
    if (kDv && colsum != nullptr) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const double t = dcs[ch] + __shfl_xor_sync(0xffffffffu, dcs[ch], 16);
        if (lane < 16) atomicAdd(&colsum[hc * 64 + ch * 16 + lane], t);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<256>(tmem);
}

template <bool kDv>
int attn2_launch(const float* k, const float* v, const float* c2, int64_t N, int P, float* out, const float* scale, double* colsum,
                 float* absmax, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn2_kernel<kDv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a2::SMEM_BYTES));
    attr_done = true;
  }
  const int T = (P + kTile - 1) / kTile;
  const int64_t W = N * T;
  int64_t grid = 2 * (int64_t)sm_count();
  if (grid > W) grid = W;
  static const bool pf_init = [] {
    const char* e = getenv("SGA_PCT_ATTN_PF");
    const int on = (e && e[0] == '0') ? 0 : 1;
    cudaMemcpyToSymbol(g_attn2_pf, &on, sizeof(int));
    return true;
  }();
  (void)pf_init;
  pct_attn2_kernel<kDv><<<(unsigned)grid, kThreads, a2::SMEM_BYTES, st>>>(k, v, c2, N, P, out, scale, colsum, absmax);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace

int attn2_fwd(const float* k, const float* v, const float* c2, int64_t N, int P, float* xs, cudaStream_t st) {
  return attn2_launch<false>(k, v, c2, N, P, xs, nullptr, nullptr, nullptr, st);
}
int attn2_dv(const float* k, const float* dxs, const float* c2, const float* scale, int64_t N, int P, float* dv, double* colsum,
             float* absmax, cudaStream_t st) {
  return attn2_launch<true>(k, dxs, c2, N, P, dv, scale, colsum, absmax, st);
}

}  // namespace pct
}  // namespace sga

// =====================================================================================================================
// dk halves of the SA backward (see pct_attn.cu for the algebra), two CTAs per SM.  What had to shrink to fit (the first
// kernel holds 160 KiB and 448 TMEM columns): column blocks of 64 (the streamed tile is 32 KiB, its k rows 8 KiB), the own
// block's k rows live in TENSOR MEMORY as the A operand of the score product (32 columns), T is written over S and the
// per-step product T k_b over dA, and dk is accumulated in REGISTERS (16 per thread) from the per-step product instead of
// in a resident accumulator: 115 KiB, 192 columns, 96 registers.
namespace sga {
namespace pct {
namespace {

namespace d2 {
constexpr uint32_t kHalf = kBlk / 2;                       // [64 rows x 128 B]
constexpr uint32_t KB = 0;                                 // current block's k rows (64), fp16 [hi | lo]
constexpr uint32_t FHI = KB + kHalf;                       // own block's 128-channel rows: 2 channel blocks hi, 2 lo
constexpr uint32_t FLO = FHI + 2 * kBlk;
constexpr uint32_t GHI = FLO + 2 * kBlk;                   // current block's rows (64): 2 channel half-blocks hi, 2 lo
constexpr uint32_t GLO = GHI + 2 * kHalf;
constexpr uint32_t NRM = GLO + 2 * kHalf;                  // float[3][512]: max * a, log2 sum, delta
constexpr uint32_t XCH = NRM + 3 * kMaxT2 * kTile * 4;     // float[2][128]
constexpr uint32_t BARS = XCH + 2 * 128 * 4;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;       // 114768
constexpr uint32_t KAH_COL = 0, KAL_COL = 16, ST_COL = 64, DA_COL = 128;
enum { BAR_LD_FULL = 0, BAR_SD_FULL = 1, BAR_T_FULL = 2, BAR_DK_DONE = 3, kNumBars = 4 };
static_assert(8 * kStageFloats * 4 <= 4 * kHalf, "the output staging patch must fit the streamed tile");
}  // namespace d2

template <bool kCol>
__global__ void __launch_bounds__(kThreads, 2)
pct_attn2_dk_kernel(const float* __restrict__ k, const float* __restrict__ fixed, const float* __restrict__ streamed,
                    const float* __restrict__ c2, const float* __restrict__ delta_in, float* __restrict__ delta_out,
                    const float* __restrict__ scale, int64_t N, int P, float* __restrict__ dk_out, int sweeps) {
  using namespace d2;
  const int kPhases = kCol ? 1 : sweeps;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  float* nrm = reinterpret_cast<float*>(sm + NRM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = (P + kTile - 1) / kTile;                   // own blocks of 128 rows
  const int Tb = (P + 63) / 64;                            // column blocks of 64
  const int Ppad = T * kTile;
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_LD_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_SD_FULL], 1);
    ptx::mbar_init(&bars[BAR_T_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_DK_DONE], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t W = N * T;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc_s = ptx::make_idesc(kFmt, 128, 64);
    const uint32_t idesc_da = ptx::make_idesc(kFmt, 128, 64);
    const uint32_t idesc_dk = ptx::make_idesc(kFmt, 128, 64) | (1u << 16);        // B (= k_b image) read MN-major
    const uint64_t dKB = ptx::smem_desc_sw128(sm_base + KB);
    const uint64_t dFhi = ptx::smem_desc_sw128(sm_base + FHI), dFlo = ptx::smem_desc_sw128(sm_base + FLO);
    const uint64_t dGhi = ptx::smem_desc_sw128(sm_base + GHI), dGlo = ptx::smem_desc_sw128(sm_base + GLO);
    const uint64_t mKB = desc_mn_sw128(sm_base + KB, kHalf);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t u = 0, ud = 0;
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x) {
      for (int ph = 0; ph < kPhases; ++ph) {
        const bool dk_phase = ph == kPhases - 1;
        for (int b = 0; b < Tb; ++b, ++u) {
          ptx::mbar_wait(&bars[BAR_LD_FULL], u & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            // S[128 x 64] = k_a k_b^T: A from tensor memory (hi / lo, K = 32 = 2 steps of 8 columns), four partial products
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const uint32_t a_tm = tmem_u + ((pass >= 2) ? KAL_COL : KAH_COL);
              const uint64_t bo = (pass & 1) ? 4 : 0;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                ptx::umma_bf16_ts(tmem_u + ST_COL, a_tm + (uint32_t)(ks * 8), dKB + bo + (uint64_t)(ks * 2), idesc_s, (pass | ks) != 0);
            }
            // dA[128 x 64] = F_a G_b^T, K = 128 channels
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint64_t ad = (pass == 1) ? dFlo : dFhi;
              const uint64_t bd = (pass == 2) ? dGlo : dGhi;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t ao = (uint64_t)((ks >> 2) * (kBlk >> 4) + (ks & 3) * 2);
                const uint64_t bo = (uint64_t)((ks >> 2) * (kHalf >> 4) + (ks & 3) * 2);
                ptx::umma_bf16(tmem_u + DA_COL, ad + ao, bd + bo, idesc_da, (pass | ks) != 0);
              }
            }
            ptx::umma_commit(&bars[BAR_SD_FULL]);
          }
          __syncwarp();
          ptx::mbar_wait(&bars[BAR_T_FULL], u & 1);          // S / dA have been read (and, in the dk sweep, T is in place)
          if (!dk_phase) continue;
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            // D[128 x 64] = T [k_b.hi | k_b.lo], K = 64 rows of the block: written over dA
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_bf16_ts(tmem_u + DA_COL, tmem_u + ST_COL + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8 + pass * 16),
                                  mKB + (uint64_t)(ks * 128), idesc_dk, (pass | ks) != 0);
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
    const uint32_t c0 = (uint32_t)(hc * 32);
    float* stage = reinterpret_cast<float*>(sm + GHI) + warp * kStageFloats;      // reuses the streamed tile at the end of an item
    float* xch = reinterpret_cast<float*>(sm + XCH);
    const float* ma_s = nrm;
    const float* lg_s = nrm + kMaxT2 * kTile;
    const float* de_s = nrm + 2 * kMaxT2 * kTile;
    uint32_t u = 0, ud = 0;
    for (int64_t w = blockIdx.x; w < W; w += gridDim.x) {
      const int64_t n = w / T;
      const int a = (int)(w - n * T);
      const int64_t obase = n * (int64_t)P;
      const int valida = min(kTile, P - a * kTile);
      const float gsc = __ldg(scale + 2 * n), ginv = __ldg(scale + 2 * n + 1);
      // every product of the previous item has completed and its rows are stored by THIS warp; the barrier below covers the others
      {   // own block: k rows -> tensor memory (hc == 0 warps, one per lane quarter), fixed tile -> shared memory, statistics
        if (hc == 0) {
          float4 kx[8];
          const float4* src = reinterpret_cast<const float4*>(k + (obase + (int64_t)a * kTile + row) * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) kx[j] = (row < valida) ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          uint32_t kh[16], kl[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            split2f<0>(kx[j].x, kx[j].y, kh[2 * j], kl[2 * j]);
            split2f<0>(kx[j].z, kx[j].w, kh[2 * j + 1], kl[2 * j + 1]);
          }
          ptx::tmem_st16(tmem + lane_addr + KAH_COL, kh);
          ptx::tmem_st16(tmem + lane_addr + KAL_COL, kl);
          ptx::tmem_st_wait();
        }
        const int cc = tid & 15, r0 = tid >> 4;
        float4 x[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const float4* src = reinterpret_cast<const float4*>(fixed + (obase + (int64_t)a * kTile + r) * 128 + cc * 8);
          const bool ok = r < valida;
          x[i][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
          x[i][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float mul = kCol ? gsc : 1.f;
        const uint32_t blk_off = (uint32_t)(cc >> 3) * kBlk;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const float f[8] = {x[i][0].x * mul, x[i][0].y * mul, x[i][0].z * mul, x[i][0].w * mul,
                              x[i][1].x * mul, x[i][1].y * mul, x[i][1].z * mul, x[i][1].w * mul};
          uint4 hi, lo;
          split8(f, hi, lo);
          const uint32_t off = blk_off + ptx::sw128_offset(r, cc & 7);
          st_chunk(sm_base + FHI + off, hi);
          st_chunk(sm_base + FLO + off, lo);
        }
        for (int i = tid; i < Ppad; i += kComputeThreads) {
          nrm[i] = c2[(n * 2) * Ppad + i];
          nrm[kMaxT2 * kTile + i] = c2[(n * 2 + 1) * Ppad + i];
          nrm[2 * kMaxT2 * kTile + i] = ((kCol || kPhases == 1) && i < P) ? delta_in[obase + i] : 0.f;
        }
      }
      compute_barrier();                               // statistics visible; every warp is past its output stores (staging patch = G)
      const float mr = ma_s[a * kTile + row], lr = lg_s[a * kTile + row];
      float dr = de_s[a * kTile + row];
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
      // the per-step product T k_b of step x: added to this thread's 16 dk channels (hi and lo halves of the k image)
      auto drain_dk = [&](uint32_t x) {
        ptx::mbar_wait(&bars[BAR_DK_DONE], x & 1);
        ptx::tc_fence_after();
        uint32_t d0[16], d1[16];
        ptx::tmem_ld16(tmem + lane_addr + DA_COL + (uint32_t)(hc * 16), d0);
        ptx::tmem_ld16(tmem + lane_addr + DA_COL + (uint32_t)(32 + hc * 16), d1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(d0[e]) + __uint_as_float(d1[e]);
      };
      for (int ph = 0; ph < kPhases; ++ph) {
        const bool dk_phase = ph == kPhases - 1;
        float dacc = 0.f;
        for (int b = 0; b < Tb; ++b, ++u) {
          if (dk_phase && b >= 1) drain_dk(ud - 1);           // also: k_b / G / the S and dA columns are free again
          const int validb = min(64, P - b * 64);
          {   // k rows (64 x 32) and streamed rows (64 x 128) of block b; all loads in flight before the first store
            const int64_t rb = obase + (int64_t)b * 64;
            const int krow = tid >> 2, kj = tid & 3;
            const float4* ks = reinterpret_cast<const float4*>(k + (rb + krow) * 32 + kj * 8);
            const bool kok = krow < validb;
            const float4 k0 = kok ? __ldg(ks) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 k1 = kok ? __ldg(ks + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int cc = tid & 15, r0 = tid >> 4;
            float4 x[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r0 + 16 * i;
              const float4* src = reinterpret_cast<const float4*>(streamed + (rb + r) * 128 + cc * 8);
              const bool ok = r < validb;
              x[i][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
              x[i][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            {
              const float f[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
              uint4 hi, lo;
              split8(f, hi, lo);
              st_chunk(sm_base + KB + ptx::sw128_offset(krow, kj), hi);
              st_chunk(sm_base + KB + ptx::sw128_offset(krow, 4 + kj), lo);
            }
            const float mul = kCol ? 1.f : gsc;
            const uint32_t blk_off = (uint32_t)(cc >> 3) * kHalf;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r0 + 16 * i;
              const float f[8] = {x[i][0].x * mul, x[i][0].y * mul, x[i][0].z * mul, x[i][0].w * mul,
                                  x[i][1].x * mul, x[i][1].y * mul, x[i][1].z * mul, x[i][1].w * mul};
              uint4 hi, lo;
              split8(f, hi, lo);
              const uint32_t off = blk_off + ptx::sw128_offset(r, cc & 7);
              st_chunk(sm_base + GHI + off, hi);
              st_chunk(sm_base + GLO + off, lo);
            }
          }
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bars[BAR_LD_FULL]);
          ptx::mbar_wait(&bars[BAR_SD_FULL], u & 1);
          ptx::tc_fence_after();
          // this thread: row `row`, the 32 columns [c0, c0 + 32) of the block, in two halves of 16 (second half first, so
          // that T can be written over S: [c0, c0+16) = T.hi of the 32 columns, [c0+16, c0+32) = T.lo)
          uint32_t th1[8], tl1[8];
#pragma unroll
          for (int h = 1; h >= 0; --h) {
            uint32_t sv[16], dv[16];
            ptx::tmem_ld16(tmem + lane_addr + ST_COL + c0 + (uint32_t)(h * 16), sv);
            ptx::tmem_ld16(tmem + lane_addr + DA_COL + c0 + (uint32_t)(h * 16), dv);
            ptx::tmem_ld_wait();
            const int cb = b * 64 + (int)c0 + h * 16;
            if (!dk_phase) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                dacc = fmaf(ex2(fmaf(__uint_as_float(sv[e]), kAlpha2, -mr) - lr), __uint_as_float(dv[e]), dacc);
            } else {
              uint32_t th[8], tl[8];
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                float t0, t1;
                if (kCol) {
                  t0 = ex2(fmaf(__uint_as_float(sv[e]), kAlpha2, -ma_s[cb + e]) - lg_s[cb + e]) * (__uint_as_float(dv[e]) - de_s[cb + e]);
                  t1 = ex2(fmaf(__uint_as_float(sv[e + 1]), kAlpha2, -ma_s[cb + e + 1]) - lg_s[cb + e + 1]) * (__uint_as_float(dv[e + 1]) - de_s[cb + e + 1]);
                } else {
                  t0 = ex2(fmaf(__uint_as_float(sv[e]), kAlpha2, -mr) - lr) * (__uint_as_float(dv[e]) - dr);
                  t1 = ex2(fmaf(__uint_as_float(sv[e + 1]), kAlpha2, -mr) - lr) * (__uint_as_float(dv[e + 1]) - dr);
                }
                split2f<0>(t0, t1, th[e / 2], tl[e / 2]);
              }
              if (h == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { th1[e] = th[e]; tl1[e] = tl[e]; }
              } else {           // both halves of S are in registers now: write T over them
                ptx::tmem_st8(tmem + lane_addr + ST_COL + c0, th);
                ptx::tmem_st8(tmem + lane_addr + ST_COL + c0 + 8, th1);
                ptx::tmem_st8(tmem + lane_addr + ST_COL + c0 + 16, tl);
                ptx::tmem_st8(tmem + lane_addr + ST_COL + c0 + 24, tl1);
                ptx::tmem_st_wait();
              }
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bars[BAR_T_FULL]);
          if (dk_phase) ++ud;
        }
        if (!dk_phase) {                    // delta_i: the two column halves of the row
          xch[hc * 128 + row] = dacc;
          compute_barrier();
          dr = xch[row] + xch[128 + row];
          if (hc == 0 && a * kTile + row < P) delta_out[obase + (int64_t)a * kTile + row] = dr;
          compute_barrier();
        }
      }
      drain_dk(ud - 1);
      // ---- dk rows of this block, scaled by 1/sqrt(32) and back from the object's scale; staging patch = the streamed tile
      compute_barrier();                               // every warp has drained the last product (which read k_b) before G is reused
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] = acc[e] * (0.17677669529663687f * ginv);
      const int nvalid = max(0, min(32, P - a * kTile - 32 * q));
      float s_ = 0.f, q_ = 0.f;
      stage_store16(stage, f, dk_out + (obase + (int64_t)a * kTile + 32 * q) * 32 + hc * 16, 32, nvalid, lane, s_, q_, false);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<256>(tmem);
}

}  // namespace

int attn2_dk(const float* k, const float* fixed, const float* streamed, const float* c2, float* delta, const float* scale, int64_t N,
             int P, int by_col, int delta_sweep, float* dk_out, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_attn2_dk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pct_attn2_dk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2::SMEM_BYTES));
    attr_done = true;
  }
  const int T = (P + kTile - 1) / kTile;
  const int64_t W = N * T;
  int64_t grid = 2 * (int64_t)sm_count();
  if (grid > W) grid = W;
  if (by_col) pct_attn2_dk_kernel<true><<<(unsigned)grid, kThreads, d2::SMEM_BYTES, st>>>(k, fixed, streamed, c2, delta, nullptr, scale, N, P, dk_out, 1);
  else pct_attn2_dk_kernel<false><<<(unsigned)grid, kThreads, d2::SMEM_BYTES, st>>>(k, fixed, streamed, c2, delta, delta, scale, N, P, dk_out, delta_sweep ? 2 : 1);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace pct
}  // namespace sga
