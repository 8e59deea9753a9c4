// NaivePCT backward: weight gradients and Gram matrices as contractions over ALL points of the batch,
//     C_b[m, n] += sum_r A[r, m] B_b[r, n],      A [R, 128],  B_b [R, 128 or 32],  R = N * P  (2 M rows at C2),
// i.e. dW = dY^T X of every pointwise convolution (pct.py:107,197-202) and the 512 x 512 Gram matrix of the concatenated
// activations that the train-mode BatchNorm of the 512 -> 1024 convolution needs (sgaligner_b200/pct.py, concat stage).
// HBM bound by construction (128 KiB of operands per 1536 tensor cycles), so: one persistent CTA per SM walks row tiles
// of 128 points, the 8 compute warps turn the fp32 rows into bf16 hi / lo images (16 loads of 16 bytes in flight per
// thread, the NEXT operand's while this one is converted), both operands are read MN-major (the contraction index is the ROW of the 128B-swizzled image, the same bytes
// a K-major read of the activations uses), the accumulators of up to four B operands that share one A tile stay in
// tensor memory for the whole launch and are added to C atomically at the end (148 x 16 K atomics per operand).
// bf16 pairs, three passes (hi hi + lo hi + hi lo): no scaling needed -- bf16 has the fp32 range -- and the 2^-17
// rounding errors of 2 M independent rows average out.  A 32-channel B (d k, pct.py:197) is stored as [hi | lo] in ONE
// 128-byte row and read with N = 64: two passes (A.hi, A.lo) give all four partial products, summed in the epilogue.
#include "pct_common.cuh"

namespace sga {
namespace pct {
namespace {

namespace wg {
constexpr uint32_t AHI = 0;
constexpr uint32_t ALO = AHI + 2 * kBlk;
constexpr uint32_t BRING = ALO + 2 * kBlk;               // 2 slots x { hi: 2 blocks, lo: 2 blocks }
constexpr uint32_t BSLOT = 4 * kBlk;
constexpr uint32_t BARS = BRING + 2 * BSLOT;             // 196608
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_BYTES = TMEMPTR + 16 + 1024;
enum { BAR_A_FULL = 0, BAR_A_FREE = 1, BAR_B_FULL = 2 /*,3*/, BAR_B_FREE = 4 /*,5*/, BAR_ACC = 6, kNumBars = 7 };
}  // namespace wg

struct WgArgs {
  const float* A;
  const float* B[4];
  int nb[4];
  int nB;
  float* C[4];
  int64_t ldc[4];
  int transpose;
  int alias;          // 1: a B operand that IS the A tensor reuses the A images (diagonal Gram blocks)
  int64_t R;
};

// 128 rows x 128 channels -> bf16 hi / lo images (2 channel blocks each).  Two halves: all 16 loads of a thread go out
// (wg_issue128), the conversion + shared-memory stores happen one operand later (wg_store128), so the next operand's rows are
// in flight while this one is converted -- inline-asm shared stores are ordering barriers for the compiler, the loads
// have to precede them in program order.
__device__ __forceinline__ void wg_issue128(const float* __restrict__ src, int64_t rowbase, int valid, int tid, float4 (&x)[8][2]) {
  const int cc = tid & 15, r0 = tid >> 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + 16 * i;
    const float4* p = reinterpret_cast<const float4*>(src + (rowbase + row) * 128 + cc * 8);
    const bool ok = row < valid;
    x[i][0] = ok ? __ldg(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[i][1] = ok ? __ldg(p + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void wg_store128(const float4 (&x)[8][2], uint32_t hi_addr, uint32_t lo_addr, int tid) {
  const int cc = tid & 15, r0 = tid >> 4;
  const uint32_t blk_off = (uint32_t)(cc >> 3) * kBlk;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + 16 * i;
    const float f[8] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x, x[i][1].y, x[i][1].z, x[i][1].w};
    uint4 hi, lo;
    split8f<1>(f, hi, lo);
    const uint32_t off = blk_off + ptx::sw128_offset(row, cc & 7);
    st_chunk(hi_addr + off, hi);
    st_chunk(lo_addr + off, lo);
  }
}

// 128 rows x 32 channels -> one block of [hi (32) | lo (32)] rows (registers x[0..1] of the operand buffer)
__device__ __forceinline__ void wg_issue32(const float* __restrict__ src, int64_t rowbase, int valid, int tid, float4 (&x)[8][2]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 2, j = idx & 3;
    const float4* p = reinterpret_cast<const float4*>(src + (rowbase + row) * 32 + j * 8);
    const bool ok = row < valid;
    x[u][0] = ok ? __ldg(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[u][1] = ok ? __ldg(p + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void wg_store32(const float4 (&x)[8][2], uint32_t addr, int tid) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 2, j = idx & 3;
    const float f[8] = {x[u][0].x, x[u][0].y, x[u][0].z, x[u][0].w, x[u][1].x, x[u][1].y, x[u][1].z, x[u][1].w};
    uint4 hi, lo;
    split8f<1>(f, hi, lo);
    st_chunk(addr + ptx::sw128_offset(row, j), hi);
    st_chunk(addr + ptx::sw128_offset(row, 4 + j), lo);
  }
}

__global__ void __launch_bounds__(kThreads, 1) pct_wgrad_kernel(const WgArgs W) {
  using namespace wg;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_A_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_A_FREE], 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars[BAR_B_FULL + s], kComputeThreads);
      ptx::mbar_init(&bars[BAR_B_FREE + s], 1);
    }
    ptx::mbar_init(&bars[BAR_ACC], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (W.R + kTile - 1) / kTile;
  const int64_t ntile = (tiles > (int64_t)blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nB = W.nB;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc128 = ptx::make_idesc(1, 128, 128) | (1u << 15) | (1u << 16);     // both operands MN-major
    const uint32_t idesc64 = ptx::make_idesc(1, 128, 64) | (1u << 15) | (1u << 16);
    const uint64_t mAhi = desc_mn_sw128(sm_base + AHI, kBlk), mAlo = desc_mn_sw128(sm_base + ALO, kBlk);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t u = 0;
    for (int64_t t = 0; t < ntile; ++t) {
      ptx::mbar_wait(&bars[BAR_A_FULL], (uint32_t)(t & 1));
      for (int b = 0; b < nB; ++b) {
        const bool self = W.alias && W.B[b] == W.A;        // a diagonal Gram block: the A images are the B operand too, nothing is loaded
        const uint32_t slot = u & 1;
        if (!self) {
          ptx::mbar_wait(&bars[BAR_B_FULL + slot], (u >> 1) & 1);
          ++u;
        }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sb = sm_base + BRING + slot * BSLOT;
          const uint32_t d = tmem_u + (uint32_t)b * 128;
          if (W.nb[b] == 128) {
            const uint64_t mBhi = self ? mAhi : desc_mn_sw128(sb, kBlk), mBlo = self ? mAlo : desc_mn_sw128(sb + 2 * kBlk, kBlk);
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint64_t ad = (pass == 1) ? mAlo : mAhi;
              const uint64_t bd = (pass == 2) ? mBlo : mBhi;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16(d, ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc128, (t | pass | ks) != 0);
            }
          } else {
            const uint64_t mB = desc_mn_sw128(sb, kBlk);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const uint64_t ad = pass ? mAlo : mAhi;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16(d, ad + (uint64_t)(ks * 128), mB + (uint64_t)(ks * 128), idesc64, (t | pass | ks) != 0);
            }
          }
          if (!self) ptx::umma_commit(&bars[BAR_B_FREE + slot]);
          if (b == nB - 1) ptx::umma_commit(&bars[BAR_A_FREE]);
          if (b == nB - 1 && t == ntile - 1) ptx::umma_commit(&bars[BAR_ACC]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== loaders / epilogue ===============================
    // flat sequence of operands: item = t * (nB + 1) + op, op 0 = the A tile, op b + 1 = B_b.  Item i + 1 is loaded into the
    // other register buffer before item i is converted.
    uint32_t lmap = 0;                          // the B operands that are loaded (B_b == A reuses the A images), 2 bits each
    int nL = 0;
    for (int b = 0; b < nB; ++b)
      if (!(W.alias && W.B[b] == W.A)) lmap |= (uint32_t)b << (2 * nL++);
    const int per = nL + 1;
    const int64_t items = ntile * per;
    auto issue = [&](int64_t item, float4 (&x)[8][2]) {
      const int64_t t = item / per;
      const int op = (int)(item - t * per);
      const int64_t rowbase = ((int64_t)blockIdx.x + t * gridDim.x) * kTile;
      const int valid = (int)min((int64_t)kTile, W.R - rowbase);
      if (op == 0) { wg_issue128(W.A, rowbase, valid, tid, x); return; }
      const int b = (int)((lmap >> (2 * (op - 1))) & 3u);
      if (W.nb[b] == 128) wg_issue128(W.B[b], rowbase, valid, tid, x);
      else wg_issue32(W.B[b], rowbase, valid, tid, x);
    };
    auto store = [&](int64_t item, const float4 (&x)[8][2]) {
      const int64_t t = item / per;
      const int op = (int)(item - t * per);
      const int nB = nL;                        // ring position counts loaded operands only
      const int bsel = op == 0 ? 0 : (int)((lmap >> (2 * (op - 1))) & 3u);
      if (op == 0) {
        if (t >= 1) ptx::mbar_wait(&bars[BAR_A_FREE], (uint32_t)((t - 1) & 1));
        wg_store128(x, sm_base + AHI, sm_base + ALO, tid);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[BAR_A_FULL]);
      } else {
        const uint32_t u = (uint32_t)(t * nB + (op - 1));
        const uint32_t slot = u & 1;
        if (u >= 2) ptx::mbar_wait(&bars[BAR_B_FREE + slot], ((u - 2) >> 1) & 1);
        const uint32_t sb = sm_base + BRING + slot * BSLOT;
        if (W.nb[bsel] == 128) wg_store128(x, sb, sb + 2 * kBlk, tid);
        else wg_store32(x, sb, tid);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[BAR_B_FULL + slot]);
      }
    };
    float4 X0[8][2], X1[8][2];
    if (items > 0) issue(0, X0);
#pragma unroll 1
    for (int64_t i = 0; i < items; i += 2) {
      if (i + 1 < items) issue(i + 1, X1);
      store(i, X0);
      if (i + 1 < items) {
        if (i + 2 < items) issue(i + 2, X0);
        store(i + 1, X1);
      }
    }
    if (ntile > 0) {
      ptx::mbar_wait(&bars[BAR_ACC], 0);
      ptx::tc_fence_after();
      const int q = warp & 3, hc = warp >> 2;
      const int m = 32 * q + lane;
      const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
      for (int b = 0; b < nB; ++b) {
        float* C = W.C[b];
        const int64_t ldc = W.ldc[b];
        if (W.nb[b] == 128) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            ptx::tmem_ld32(tmem + lane_addr + (uint32_t)(b * 128 + hc * 64 + h * 32), v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int n = hc * 64 + h * 32 + e;
              atomicAdd(W.transpose ? (C + (int64_t)n * ldc + m) : (C + (int64_t)m * ldc + n), __uint_as_float(v[e]));
            }
          }
        } else {
          uint32_t d0[16], d1[16];
          ptx::tmem_ld16(tmem + lane_addr + (uint32_t)(b * 128 + hc * 16), d0);
          ptx::tmem_ld16(tmem + lane_addr + (uint32_t)(b * 128 + 32 + hc * 16), d1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = hc * 16 + e;
            atomicAdd(W.transpose ? (C + (int64_t)n * ldc + m) : (C + (int64_t)m * ldc + n), __uint_as_float(d0[e]) + __uint_as_float(d1[e]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<512>(tmem);
}

}  // namespace
}  // namespace pct
}  // namespace sga

extern "C" int sga_pct_wgrad(const float* A, int64_t R, const float* const* B, const int* nb, int nB, float* const* C,
                             const int64_t* ldc, int transpose, void* stream) {
  if (R <= 0 || nB <= 0) return SGA_OK;
  SGA_REQUIRE(A && B && nb && C && ldc && nB <= 4, "sga_pct_wgrad: bad arguments (nB=%d)", nB);
  using namespace sga::pct;
  WgArgs W{};
  W.A = A; W.R = R; W.nB = nB; W.transpose = transpose;
  static const bool no_alias = [] { const char* e = getenv("SGA_PCT_WGRAD_ALIAS"); return e && e[0] == '0'; }();
  W.alias = no_alias ? 0 : 1;
  SGA_REQUIRE(((uintptr_t)A & 15) == 0, "sga_pct_wgrad: A must be 16-byte aligned");
  for (int b = 0; b < nB; ++b) {
    SGA_REQUIRE(B[b] && C[b] && (nb[b] == 128 || nb[b] == 32) && ((uintptr_t)B[b] & 15) == 0, "sga_pct_wgrad: operand %d", b);
    SGA_REQUIRE(B[b] != A || nb[b] == 128, "sga_pct_wgrad: operand %d aliases A with another width", b);
    W.B[b] = B[b]; W.nb[b] = nb[b]; W.C[b] = C[b]; W.ldc[b] = ldc[b];
  }
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pct_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::SMEM_BYTES));
    attr_done = true;
  }
  const int64_t tiles = (R + kTile - 1) / kTile;
  int grid = sga::sm_count();
  if ((int64_t)grid > tiles) grid = (int)tiles;
  pct_wgrad_kernel<<<grid, kThreads, wg::SMEM_BYTES, (cudaStream_t)stream>>>(W);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
