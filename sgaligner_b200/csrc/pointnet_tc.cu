// Fused PointNet feature encoder on the 5th-gen tensor cores (SGA_POINTNET_TC).
// Reference: src/aligner/networks/pointnet.py:140-163 -- conv1(3->64) ReLU, conv2(64->128) ReLU,
// conv3(128->C3) ReLU, max over the points of an object.  The reference materialises three
// activation tensors (3.8 GB at B=32) in HBM; here nothing but the points (12 B/point) and the
// pooled feature (4 B/channel) ever touches HBM.
//
// Numerics: fp32 operands are split x = hi + lo into two bf16 values and every product is
// evaluated as hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM ("bf16x3"); the dropped
// terms are <= 2^-16 relative per product, i.e. ~1e-5 on the pooled feature -- inside the 1e-4
// parity gate, where a single bf16 or tf32 pass (4e-3 / 3e-4) is not.
//
// One persistent CTA per SM; 8 compute warps + 1 MMA-issuing warp.  Per 128-point tile g:
//   S1   (CUDA cores) conv1+ReLU -> A1{hi,lo} [128 pts x 64]  bf16, K-major, 128B-swizzled smem
//   MMA2 D2[pts x 128ch] = A1 * W2^T            3 passes x 4 k-steps   (M=128 N=128 K=16)
//   E2   TMEM->regs (all 128 columns at once, D2 is released immediately), +b2, ReLU, hi/lo split;
//        the four 32-channel chunks of H2{hi,lo} are stored as soon as conv3 of the PREVIOUS tile
//        has consumed the chunk they overwrite
//   MMA3 D3[ch x 128 pts] = W3 * H2^T           2 M-tiles x 3 passes x 8 k-steps, chunk by chunk;
//        W3.hi -- the A operand of two of the three passes -- lives in TMEM (tcgen05.mma A-from-TMEM),
//        which frees 64 KiB of shared memory for a private A1 buffer and halves the shared-memory
//        operand traffic of those passes (SS-mode UMMA at this shape is shared-memory-bandwidth bound)
//   E3   TMEM->regs, running max over columns (= points): channels sit on TMEM lanes, so the
//        max-pool is thread-local; bias + ReLU are applied once per object after the max
//        (max_p relu(z_p + b) == relu(max_p z_p + b)).
// Software pipeline (tensor-pipe order):  ... MMA2(g+1) | MMA3(g) c0..c3 | MMA2(g+2) | MMA3(g+1) ...
// conv2 of the NEXT tile is queued ahead of conv3 of the current one, so E2(g+1) and S1(g+2) run
// on the CUDA cores while the tensor pipe works through MMA3(g), and E3(g) runs under MMA2(g+2):
// the tensor pipe never waits for an epilogue in steady state.
// Shared memory: W2{hi,lo} 32 KiB + W3.lo 64 KiB resident, A1{hi,lo} 32 KiB, H2{hi,lo} 64 KiB.
// TMEM (512 columns): D2 128 | D3 256 | W3.hi 128 (2 M-tiles x 64 columns of packed bf16 pairs).
//
// kMoments (training): the batch statistics of the three pre-ReLU conv outputs that the reference's
// discarded BatchNorm1d calls fold into running_mean/var (pointnet.py:141-142,154-155,158-159) are
// accumulated on the fly: conv3 thread-locally in E3 (a thread owns a channel), conv2 with a
// butterfly transpose-reduce over the 32 points of a warp in E2, conv1 analytically from the first
// and second moments of the points (z1 is affine in the point).
#include "common.cuh"
#include "ptx.cuh"

namespace sga {
namespace {

// test/diagnostic hook: when set (sga_debug_set_trace), CTA 0 stamps clock64() at every pipeline
// event of its first 64 tiles: trace[g*16 + k] (compute thread 0), trace[1024 + g*16 + k] (MMA lane)
__device__ long long* g_trace = nullptr;
#define SGA_TRACE(base, g, k)                                                            \
  do {                                                                                   \
    if (trace && (g) < 64) trace[(base) + (g) * 16 + (k)] = clock64();                   \
  } while (0)

constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;
constexpr int kTile = 128;                    // points per tile
constexpr uint32_t kBlk = 16384;              // one [128 rows x 128 B] swizzle-atom column block

// shared-memory map (byte offsets from the 1024-aligned base)
constexpr uint32_t W2HI = 0;
constexpr uint32_t W2LO = W2HI + kBlk;
constexpr uint32_t W3LO = W2LO + kBlk;        // 4 blocks: (mt*2 + ka); W3.hi lives in TMEM
constexpr uint32_t A1HI = W3LO + 4 * kBlk;
constexpr uint32_t A1LO = A1HI + kBlk;
constexpr uint32_t H2HI = A1LO + kBlk;        // 2 blocks (ka)
constexpr uint32_t H2LO = H2HI + 2 * kBlk;    // 2 blocks
constexpr uint32_t SMALL = H2LO + 2 * kBlk;   // 196608
constexpr uint32_t W1B1 = SMALL;              // float4[64] = {w0,w1,w2,b}
constexpr uint32_t B2 = W1B1 + 1024;          // float[128]
constexpr uint32_t BARS = B2 + 512;           // 13 mbarriers
constexpr uint32_t TMEMPTR = BARS + 128;
constexpr uint32_t SMEM_USED = TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;   // + alignment slack

constexpr uint32_t D2_COL = 0;
constexpr uint32_t D3_COL = 128;              // + mt*128
constexpr uint32_t W3HI_COL = 384;            // + mt*64 + k/2
constexpr int kTmemCols = 512;

enum {
  BAR_A1_FULL = 0,   // S1 wrote A1(g)                     (256 arrivals)
  BAR_D2_FULL = 1,   // MMA2(g) complete (also: A1 free)   (commit)
  BAR_D2_FREE = 2,   // E2 has D2(g) in registers          (256 arrivals)
  BAR_H2_FULL = 3,   // ..6  chunk c of H2(g) stored       (256 arrivals)
  BAR_H2_FREE = 7,   // ..10 MMA3(g) consumed chunk c      (commit)
  BAR_D3_FULL = 11,  // MMA3(g) complete                   (commit)
  BAR_D3_FREE = 12,  // E3 has drained D3(g)               (256 arrivals)
  kNumBars = 13
};

// layout of the raw moment accumulators (doubles) the kMoments variant adds into
constexpr int RAW_PTS = 0;      // 9: sum x,y,z, xx,xy,xz,yy,yz,zz over all points
constexpr int RAW_S2 = 16;      // 128: sum_p d2[p,c]   (d = conv output WITHOUT bias)
constexpr int RAW_Q2 = 144;     // 128: sum_p d2[p,c]^2
constexpr int RAW_S3 = 272;     // C3, then C3 squares

// max-pool near-tie window (see pointnet_tie_fix_kernel): |max - runner-up| <= kTieRel*|max| + kTieAbs
constexpr float kTieRel = 2e-4f;
constexpr float kTieAbs = 2e-5f;
constexpr int kTieMaxP = 32767;          // the runner-up index shares the 32-bit argmax word

__device__ __forceinline__ uint32_t pack2(float a, float b) {   // a -> low half (lower address / lower k)
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// split 8 fp32 values into bf16 hi / lo parts, packed as two 16-byte chunks
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

// 16-byte shared-memory store on a 32-bit shared address: one STS.128 (the generic-pointer form made ptxas
// emit 32-bit generic stores = 4x the shared-memory wavefronts, stolen from the UMMA operand reads)
__device__ __forceinline__ void st_chunk(uint32_t smem_base, uint32_t off, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_base + off), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Butterfly transpose-reduce: every lane holds x[0..63]; on return x[0], x[1] of lane L are the sums
// over all 32 lanes of elements e0(L), e0(L)+1 with e0(L) = 32*b4 + 16*b3 + 8*b2 + 4*b1 + 2*b0 (b_i = bit
// i of L).  62 shuffles instead of 64 x 5.
__device__ __forceinline__ void transpose_reduce64(float (&x)[64], int lane) {
#pragma unroll
  for (int half = 32, off = 16; half >= 2; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? x[i + half] : x[i];
      const float send = up ? x[i] : x[i + half];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}
__device__ __forceinline__ int transpose_reduce64_elem(int lane) {
  return 32 * ((lane >> 4) & 1) + 16 * ((lane >> 3) & 1) + 8 * ((lane >> 2) & 1) + 4 * ((lane >> 1) & 1) + 2 * (lane & 1);
}

template <bool kArgmax, bool kMoments>
__global__ void __launch_bounds__(kThreads, 1)
pointnet_fwd_tc_kernel(const float* __restrict__ pts, int64_t N, int P,
                       const float* __restrict__ W1, const float* __restrict__ b1,
                       const float* __restrict__ W2, const float* __restrict__ b2,
                       const float* __restrict__ W3, const float* __restrict__ b3, int C3,
                       float* __restrict__ out, int32_t* __restrict__ argmax, double* __restrict__ raw) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  const float4* w1b1 = reinterpret_cast<const float4*>(sm + W1B1);
  const float* b2s = reinterpret_cast<const float*>(sm + B2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb0 = blockIdx.y * 256;
  const int nmt = min(2, (C3 - cb0) / 128);

  // ---------------- one-time setup: weights -> bf16 hi/lo swizzled tiles, barriers, TMEM
  for (int i = tid; i < 128 * 8; i += kThreads) {          // W2 [128][64]: 8 chunks per row
    int r = i >> 3, j = i & 7;
    const float4* src = reinterpret_cast<const float4*>(W2 + r * 64 + j * 8);
    float4 a = src[0], b = src[1];
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    split8(f, hi, lo);
    uint32_t off = ptx::sw128_offset(r, j);
    st_chunk(sm_base, W2HI + off, hi);
    st_chunk(sm_base, W2LO + off, lo);
  }
  for (int i = tid; i < nmt * 128 * 16; i += kThreads) {   // W3.lo block rows [nmt*128][128]: 16 chunks per row
    int r = i >> 4, j = i & 15;
    const float4* src = reinterpret_cast<const float4*>(W3 + (int64_t)(cb0 + r) * 128 + j * 8);
    float4 a = src[0], b = src[1];
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    split8(f, hi, lo);
    uint32_t off = (uint32_t)((r >> 7) * 2 + (j >> 3)) * kBlk + ptx::sw128_offset(r & 127, j & 7);
    st_chunk(sm_base, W3LO + off, lo);
  }
  for (int i = tid; i < 64; i += kThreads)
    reinterpret_cast<float4*>(sm + W1B1)[i] = make_float4(W1[i * 3], W1[i * 3 + 1], W1[i * 3 + 2], b1[i]);
  for (int i = tid; i < 128; i += kThreads) reinterpret_cast<float*>(sm + B2)[i] = b2[i];
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_A1_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_D2_FULL], 1);
    ptx::mbar_init(&bars[BAR_D2_FREE], kComputeThreads);
    for (int c = 0; c < 4; ++c) {
      ptx::mbar_init(&bars[BAR_H2_FULL + c], kComputeThreads);
      ptx::mbar_init(&bars[BAR_H2_FREE + c], 1);
    }
    ptx::mbar_init(&bars[BAR_D3_FULL], 1);
    ptx::mbar_init(&bars[BAR_D3_FREE], kComputeThreads);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // W3.hi -> TMEM as the A operand of the two conv3 passes that use it (hi*hi, hi*lo): channel (row of W3)
  // on the TMEM lane, 128 k-values packed two per 32-bit column (lower k in the low half), 64 columns per
  // M-tile.  An A operand in TMEM costs no shared-memory bandwidth, which is what bounds SS-mode UMMA here
  // (M=128,N=128,K=16 reads 8 KB per 64 cycles = the full 128 B/clk of the SM's shared memory).
  if (warp < 8 && (warp >> 2) < nmt) {
    const int mt = warp >> 2, r = 32 * (warp & 3) + lane;
    const float4* src = reinterpret_cast<const float4*>(W3 + (int64_t)(cb0 + mt * 128 + r) * 128);
#pragma unroll 1
    for (int grp = 0; grp < 4; ++grp) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = src[grp * 8 + j];
        w[2 * j] = pack2(a.x, a.y);
        w[2 * j + 1] = pack2(a.z, a.w);
      }
      ptx::tmem_st16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + W3HI_COL + mt * 64 + grp * 16, w);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  const int ntile = (P + kTile - 1) / kTile;
  // objects owned by this CTA: blockIdx.x, +gridDim.x, ...
  const int64_t nobj = (N > (int64_t)blockIdx.x) ? (N - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t G = nobj * ntile;     // tiles this CTA processes, as one stream
  long long* trace = (blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 256)) ? g_trace : nullptr;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the (warp-uniform) control flow; one lane chosen by elect.sync issues.
    // Written this way ptxas keeps descriptors in uniform registers; a plain `if (lane == 0)` makes
    // it wrap every UTCHMMA in an ELECT/R2UR waterfall loop (~100 cycles per MMA).
    const uint32_t idesc = ptx::make_idesc(1, 128, 128);
    // descriptors differ only in the start-address field: base + (byte offset >> 4)
    const uint64_t dA1hi = ptx::smem_desc_sw128(sm_base + A1HI), dA1lo = ptx::smem_desc_sw128(sm_base + A1LO);
    const uint64_t dW2hi = ptx::smem_desc_sw128(sm_base + W2HI), dW2lo = ptx::smem_desc_sw128(sm_base + W2LO);
    const uint64_t dW3lo = ptx::smem_desc_sw128(sm_base + W3LO);
    const uint64_t dH2hi = ptx::smem_desc_sw128(sm_base + H2HI), dH2lo = ptx::smem_desc_sw128(sm_base + H2LO);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    constexpr uint32_t kBlk16 = kBlk >> 4;

    auto conv2 = [&](int64_t g) {      // D2[pts x 128] = A1 * W2^T, tile g
      if (ptx::elect_one()) {
        SGA_TRACE(1024, g, 0);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t ab = (pass == 1) ? dA1lo : dA1hi;
          const uint64_t bb = (pass == 2) ? dW2lo : dW2hi;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16(tmem_u + D2_COL, ab + (uint64_t)(ks * 2), bb + (uint64_t)(ks * 2), idesc, (pass | ks) != 0);
        }
        ptx::umma_commit(&bars[BAR_D2_FULL]);
        SGA_TRACE(1024, g, 1);
      }
      __syncwarp();
    };

    if (G > 0) {
      ptx::mbar_wait(&bars[BAR_A1_FULL], 0);
      ptx::tc_fence_after();
      conv2(0);
    }
    for (int64_t g = 0; g < G; ++g) {
      const uint32_t ph = (uint32_t)(g & 1);
      // ---- conv2 of the NEXT tile goes ahead of conv3 of this one
      if (g + 1 < G) {
        ptx::mbar_wait(&bars[BAR_A1_FULL], ph ^ 1);
        ptx::mbar_wait(&bars[BAR_D2_FREE], ph);
        ptx::tc_fence_after();
        conv2(g + 1);
      }
      // ---- conv3: D3[mt][ch x pts] = W3[mt] * H2^T, chunk by chunk (32 channels of K)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::mbar_wait(&bars[BAR_H2_FULL + c], ph);
        if (c == 0 && g > 0) ptx::mbar_wait(&bars[BAR_D3_FREE], ph ^ 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          SGA_TRACE(1024, g, 2 + c);
          const uint64_t koff = (uint64_t)((c >> 1) * kBlk16 + (c & 1) * 4);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt < nmt) {
              const uint32_t d = tmem_u + D3_COL + mt * 128;
              const uint32_t ah = tmem_u + W3HI_COL + mt * 64 + c * 16;
              const uint64_t al = dW3lo + (uint64_t)(mt * 2 * kBlk16) + koff;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)    // W3.hi (TMEM) * H2.hi
                ptx::umma_bf16_ts(d, ah + ks * 8, dH2hi + koff + (uint64_t)(ks * 2), idesc, (c | ks) != 0);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)    // W3.hi (TMEM) * H2.lo
                ptx::umma_bf16_ts(d, ah + ks * 8, dH2lo + koff + (uint64_t)(ks * 2), idesc, 1);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)    // W3.lo (smem) * H2.hi
                ptx::umma_bf16(d, al + (uint64_t)(ks * 2), dH2hi + koff + (uint64_t)(ks * 2), idesc, 1);
            }
          }
          ptx::umma_commit(&bars[BAR_H2_FREE + c]);
          if (c == 3) {
            ptx::umma_commit(&bars[BAR_D3_FULL]);
            SGA_TRACE(1024, g, 6);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3;            // TMEM lane quarter
    const int wh = warp >> 2;          // 0/1: column half (E2) / M-tile (E3)
    const int row = 32 * q + lane;     // TMEM lane = point (E2) or channel-in-tile (E3)
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const bool stats12 = kMoments && blockIdx.y == 0;   // conv1 / conv2 statistics: once per point

    // conv1 mapping: thread -> 8 channels (warp-uniform group cg) x 4 points (pg + 32 i).  The 8
    // channels' weights live in registers for the whole kernel, so conv1 issues no shared-memory
    // loads (the tensor pipe saturates shared-memory bandwidth while conv3 runs).
    const int pg = tid & 31, cg = tid >> 5;
    float4 wreg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) wreg[e] = w1b1[8 * cg + e];
    float px[4], py[4], pz[4];               // this thread's points of the NEXT tile to encode

    // statistics accumulators (kMoments only; dead code otherwise)
    double pm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};      // warp 0: point moments
    double ds2[2] = {0, 0}, dq2[2] = {0, 0};         // 2 conv2 channels per lane (transpose_reduce64_elem)
    double ds3 = 0, dq3 = 0;                         // this thread's conv3 channel

    // running (object, tile) cursors of the three streams a compute thread walks: point prefetch,
    // conv1 (S1) and the max-pool epilogue (E3); no 64-bit divisions in the tile loop
    const int64_t obj_stride = (int64_t)gridDim.x * P * 3;
    const float* pf_ptr = pts + (int64_t)blockIdx.x * P * 3;
    int pf_t = 0, s1_t = 0, e2_t = 0, e3_t = 0;
    int64_t e3_n = blockIdx.x;
    auto prefetch = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pi = min(pf_t * kTile + pg + 32 * i, P - 1);
        const float* pp = pf_ptr + pi * 3;
        px[i] = __ldg(pp); py[i] = __ldg(pp + 1); pz[i] = __ldg(pp + 2);
      }
      if (++pf_t == ntile) {
        pf_t = 0;
        pf_ptr += obj_stride;
      }
    };
    auto stage1 = [&]() {
      if (stats12 && cg == 0) {
        const int t = s1_t;
        float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (t * kTile + pg + 32 * i < P) {
            const float x = px[i], y = py[i], z = pz[i];
            s[0] += x; s[1] += y; s[2] += z;
            s[3] = fmaf(x, x, s[3]); s[4] = fmaf(x, y, s[4]); s[5] = fmaf(x, z, s[5]);
            s[6] = fmaf(y, y, s[6]); s[7] = fmaf(y, z, s[7]); s[8] = fmaf(z, z, s[8]);
          }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) pm[k] += (double)s[k];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = fmaf(wreg[e].x, px[i], fmaf(wreg[e].y, py[i], fmaf(wreg[e].z, pz[i], wreg[e].w)));
          f[e] = v > 0.f ? v : 0.f;
        }
        uint4 hi, lo;
        split8(f, hi, lo);
        uint32_t off = ptx::sw128_offset(pg + 32 * i, cg);
        st_chunk(sm_base, A1HI + off, hi);
        st_chunk(sm_base, A1LO + off, lo);
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_A1_FULL]);
      if (++s1_t == ntile) s1_t = 0;
    };

    float rmax = -INFINITY;
    // argmax variant: (max, first argmax, runner-up = the largest value STRICTLY below the max, its index).  A branch-free
    // tracker is a chain of ~10 operations per element (four interleaved chains, 4.5 k cycles per tile: longer than the MMAs
    // it should hide under -- 0.67 ms against 0.37 ms for the plain maximum).  The runner-up only matters inside the near-tie
    // window, so an element can only change the state if it reaches `thr` = max - window; after the first columns of an
    // object that is rare (a new record at column n has probability ~1/n), so four elements are rejected together with two
    // FMNMX and one compare, and only the survivors walk the update.
    float tmx = -INFINITY, tr2 = -INFINITY, thr = -INFINITY;
    int tix = 0, tr2i = 0;
    // ---- E3: running max over the 128 points (columns) of tile gp; output at the end of an object
    auto stage_e3 = [&](int64_t gp) {
      const int t = e3_t;
      ptx::mbar_wait(&bars[BAR_D3_FULL], (uint32_t)(gp & 1));
      ptx::tc_fence_after();
      SGA_TRACE(0, gp, 7);
      if (wh < nmt) {
        const uint32_t base = tmem + lane_addr + D3_COL + (uint32_t)wh * 128;
        const int valid = min(kTile, P - t * kTile);
        float ts0 = 0.f, ts1 = 0.f, tq0 = 0.f, tq1 = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          ptx::tmem_ld32(base + cc * 32, v);
          ptx::tmem_ld_wait();
          if (kArgmax) {
            // exact duplicates of the max -- resampled points -- are not rivals: they resolve to the lowest index as in the
            // reference (columns are visited in increasing order, an equal value never replaces the argmax)
            const int col0 = t * kTile + cc * 32;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float f0 = __uint_as_float(v[e]), f1 = __uint_as_float(v[e + 1]);
              const float f2 = __uint_as_float(v[e + 2]), f3 = __uint_as_float(v[e + 3]);
              if (fmaxf(fmaxf(f0, f1), fmaxf(f2, f3)) >= thr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float f = j == 0 ? f0 : (j == 1 ? f1 : (j == 2 ? f2 : f3));
                  if (f >= thr) {
                    if (f > tmx) {             // the old max is the largest value strictly below the new one
                      tr2 = tmx; tr2i = tix;
                      tmx = f; tix = col0 + e + j;
                      thr = tmx - (kTieRel * fabsf(tmx) + kTieAbs);
                    } else if (f < tmx && f > tr2) {
                      tr2 = f; tr2i = col0 + e + j;
                    }
                  }
                }
              }
            }
          } else {
            float m0 = rmax, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[e]));
              m1 = fmaxf(m1, __uint_as_float(v[e + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[e + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[e + 3]));
            }
            rmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          if (kMoments) {
            if (valid == kTile) {
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                const float f0 = __uint_as_float(v[e]), f1 = __uint_as_float(v[e + 1]);
                ts0 += f0; ts1 += f1;
                tq0 = fmaf(f0, f0, tq0); tq1 = fmaf(f1, f1, tq1);
              }
            } else {   // padded columns repeat the last point: they must not be counted
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const float f = (cc * 32 + e < valid) ? __uint_as_float(v[e]) : 0.f;
                ts0 += f;
                tq0 = fmaf(f, f, tq0);
              }
            }
          }
        }
        if (kMoments) {
          ds3 += (double)(ts0 + ts1);
          dq3 += (double)(tq0 + tq1);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_D3_FREE]);
      SGA_TRACE(0, gp, 8);
      if (++e3_t == ntile) {
        e3_t = 0;
        if (wh < nmt) {
          const int64_t n = e3_n;
          const int ch = cb0 + wh * 128 + row;
          int ridx = 0, r2idx = 0;
          float r2 = -INFINITY;
          if (kArgmax) {
            rmax = tmx; ridx = tix; r2 = tr2; r2idx = tr2i;
          }
          const float o = rmax + b3[ch];
          out[n * C3 + ch] = o > 0.f ? o : 0.f;
          if (kArgmax) {
            // Near-tie of the max-pool: the bf16x3 error (~1e-5) could have ordered the two candidates differently
            // from fp32.  The runner-up rides in the upper half of the argmax word; pointnet_tie_fix_kernel
            // re-evaluates both candidates in fp32 (reference summation order) and rewrites the entry.
            // The same re-evaluation (with the max as its own rival) pins the sign of a pooled pre-activation that
            // sits on the ReLU kink, which is the mask the backward applies.
            int code = min(ridx, P - 1);
            const float tol = kTieRel * fabsf(rmax) + kTieAbs;
            if (P <= kTieMaxP && o > -tol) {
              if (rmax - r2 <= tol) code |= (min(r2idx, P - 1) + 1) << 16;
              else if (o <= tol) code |= (code + 1) << 16;
            }
            argmax[n * C3 + ch] = code;
          }
        }
        rmax = -INFINITY;
        tmx = -INFINITY; tr2 = -INFINITY; thr = -INFINITY; tix = 0; tr2i = 0;
        e3_n += gridDim.x;
      }
    };

    if (G > 0) {
      prefetch();
      stage1();
      if (G > 1) prefetch();
    }
    for (int64_t g = 0; g < G; ++g) {
      const uint32_t ph = (uint32_t)(g & 1);
      // ---- E2: the whole conv2 accumulator of this thread's point (64 of the 128 channels: 16 per
      //      32-channel chunk) -> registers; D2 is handed back to the tensor pipe at once
      ptx::mbar_wait(&bars[BAR_D2_FULL], ph);
      ptx::tc_fence_after();
      SGA_TRACE(0, g, 0);
      uint32_t v0[16], v1[16], v2[16], v3[16];
      ptx::tmem_ld16(tmem + lane_addr + D2_COL + 0 + 16 * wh, v0);
      ptx::tmem_ld16(tmem + lane_addr + D2_COL + 32 + 16 * wh, v1);
      ptx::tmem_ld16(tmem + lane_addr + D2_COL + 64 + 16 * wh, v2);
      ptx::tmem_ld16(tmem + lane_addr + D2_COL + 96 + 16 * wh, v3);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_D2_FREE]);
      SGA_TRACE(0, g, 1);

      if (stats12) {
        const bool pvalid = (e2_t * kTile + row) < P;
        if (++e2_t == ntile) e2_t = 0;
        float x[64];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          x[e] = pvalid ? __uint_as_float(v0[e]) : 0.f;
          x[16 + e] = pvalid ? __uint_as_float(v1[e]) : 0.f;
          x[32 + e] = pvalid ? __uint_as_float(v2[e]) : 0.f;
          x[48 + e] = pvalid ? __uint_as_float(v3[e]) : 0.f;
        }
        float xq[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) xq[e] = x[e] * x[e];
        transpose_reduce64(x, lane);
        transpose_reduce64(xq, lane);
        ds2[0] += (double)x[0]; ds2[1] += (double)x[1];
        dq2[0] += (double)xq[0]; dq2[1] += (double)xq[1];
      }

      uint4 Hh[4][2], Hl[4][2];
      auto convert = [&](const uint32_t (&v)[16], int c) {
        const int k0 = 32 * c + 16 * wh;
        const float4* bq = reinterpret_cast<const float4*>(b2s + k0);
        const float4 q0 = bq[0], q1 = bq[1], q2 = bq[2], q3 = bq[3];
        const float bb[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
        float f0[8], f1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a = __uint_as_float(v[e]) + bb[e];
          const float b = __uint_as_float(v[8 + e]) + bb[8 + e];
          f0[e] = a > 0.f ? a : 0.f;
          f1[e] = b > 0.f ? b : 0.f;
        }
        split8(f0, Hh[c][0], Hl[c][0]);
        split8(f1, Hh[c][1], Hl[c][1]);
      };
      convert(v0, 0);
      convert(v1, 1);
      convert(v2, 2);
      convert(v3, 3);
      SGA_TRACE(0, g, 2);

      // ---- conv1 of the next tile (A1 is free: conv2 of this tile has completed)
      if (g + 1 < G) {
        stage1();
        if (g + 2 < G) prefetch();
      }
      SGA_TRACE(0, g, 3);

      // ---- H2 chunks, each as soon as conv3 of the previous tile has consumed what it overwrites
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (g > 0) ptx::mbar_wait(&bars[BAR_H2_FREE + c], ph ^ 1);
        const int k0 = 32 * c + 16 * wh;
        const uint32_t blk = (uint32_t)(k0 >> 6) * kBlk;
        const int j0 = (k0 & 63) >> 3;
        const uint32_t o0 = blk + ptx::sw128_offset(row, j0), o1 = blk + ptx::sw128_offset(row, j0 + 1);
        st_chunk(sm_base, H2HI + o0, Hh[c][0]);
        st_chunk(sm_base, H2HI + o1, Hh[c][1]);
        st_chunk(sm_base, H2LO + o0, Hl[c][0]);
        st_chunk(sm_base, H2LO + o1, Hl[c][1]);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[BAR_H2_FULL + c]);
      }
      SGA_TRACE(0, g, 4);
      // ---- E3 of the previous tile last: its conv3 completes just as the last H2 chunk above is released,
      //      while E2 / S1 above ran under that conv3
      if (g > 0) stage_e3(g - 1);
    }
    if (G > 0) stage_e3(G - 1);

    // ---- statistics: one atomic per (CTA, warp, channel)
    if (kMoments) {
      if (stats12) {
        if (cg == 0) {
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const double s = warp_sum_d(pm[k]);
            if (lane == 0) atomicAdd(&raw[RAW_PTS + k], s);
          }
        }
        // element e of this thread's 64 <-> channel 32*(e/16) + 16*wh + e%16
        const int e0 = transpose_reduce64_elem(lane);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int e = e0 + j, ch = 32 * (e >> 4) + 16 * wh + (e & 15);
          atomicAdd(&raw[RAW_S2 + ch], ds2[j]);
          atomicAdd(&raw[RAW_Q2 + ch], dq2[j]);
        }
      }
      if (wh < nmt) {
        const int ch = cb0 + wh * 128 + row;
        atomicAdd(&raw[RAW_S3 + ch], ds3);
        atomicAdd(&raw[RAW_S3 + C3 + ch], dq3);
      }
    }
  }
  // ---------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<kTmemCols>(tmem);
}


// ---------------------------------------------------------------------------------------------------------------
// fp32 re-check of max-pool near-ties (tensor-core forward, argmax-tracking variant).  An entry of `argmax` whose
// upper half is non-zero carries two candidate points; both pre-activations are recomputed on the FMA pipe in the
// summation order of the fp32 kernel (pointnet_simt.cu: bias first, k ascending) and the entry becomes the fp32
// argmax (lowest index on an exact tie, as torch.max does); the pooled feature is only touched when the fp32 value
// falls on the other side of the ReLU kink (the mask of the backward must be the fp32 mask).  One warp per 32
// entries; W2 is staged in shared memory only by CTAs that found a flagged entry.
__device__ unsigned long long g_tie_stats[2];     // {flagged, reordered} since the last reset (diagnostics)

constexpr int kFixThreads = 256;
constexpr int kW2Ld = 129;     // padded row of the transposed conv2 weights: conflict-free both ways
__global__ void __launch_bounds__(kFixThreads)
pointnet_tie_fix_kernel(const float* __restrict__ pts, int64_t total, int P, const float* __restrict__ W1,
                        const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                        const float* __restrict__ W3, const float* __restrict__ b3, int C3, float* __restrict__ out,
                        int32_t* __restrict__ argmax) {
  __shared__ float W2t[64 * kW2Ld];          // [k][c]
  __shared__ float h2s[kFixThreads / 32][2][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // persistent grid: every CTA stages W2^T once (coalesced reads of the row-major weights), then its warps walk the
  // argmax array in groups of 32 entries and stop at the flagged ones
  for (int j = tid; j < 64 * 128; j += kFixThreads) W2t[(j & 63) * kW2Ld + (j >> 6)] = W2[j];
  __syncthreads();
  unsigned long long nflag = 0, nswap = 0;
  const int64_t ngroup = (total + 31) >> 5;
  for (int64_t grp = (int64_t)blockIdx.x * (kFixThreads / 32) + warp; grp < ngroup; grp += (int64_t)gridDim.x * (kFixThreads / 32)) {
    const int64_t i = grp * 32 + lane;
    const int code = i < total ? argmax[i] : 0;
    unsigned m = __ballot_sync(0xffffffffu, (code >> 16) != 0);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int cd = __shfl_sync(0xffffffffu, code, src);
      const int64_t idx = grp * 32 + src;
      const int64_t n = idx / C3;
      const int ch = (int)(idx - n * C3);
      const int pa = cd & 0xFFFF, pb = (cd >> 16) - 1;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const float* pp = pts + (n * P + (w ? pb : pa)) * 3;
        const float x = pp[0], y = pp[1], z = pp[2];
        float h1[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = lane + 32 * j;
          float v = b1[c];
          v = fmaf(W1[c * 3 + 0], x, v);
          v = fmaf(W1[c * 3 + 1], y, v);
          v = fmaf(W1[c * 3 + 2], z, v);
          h1[j] = v > 0.f ? v : 0.f;
        }
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = b2[lane + 32 * j];
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
          const float a = __shfl_sync(0xffffffffu, k < 32 ? h1[0] : h1[1], k & 31);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(a, W2t[k * kW2Ld + lane + 32 * j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) h2s[warp][w][lane + 32 * j] = acc[j] > 0.f ? acc[j] : 0.f;
      }
      __syncwarp();
      float zv = 0.f;
      if (lane < 2) {
        zv = b3[ch];
        const float* wr = W3 + (int64_t)ch * 128;
        for (int k = 0; k < 128; ++k) zv = fmaf(h2s[warp][lane][k], wr[k], zv);
      }
      const float za = __shfl_sync(0xffffffffu, zv, 0), zb = __shfl_sync(0xffffffffu, zv, 1);
      if (lane == 0) {
        const bool take_b = zb > za || (zb == za && pb < pa);
        const float zm = take_b ? zb : za;
        argmax[idx] = take_b ? pb : pa;
        // the pooled feature keeps its tensor-core value (identical to what the serving variant of the kernel
        // writes) unless the fp32 value lies on the other side of the ReLU kink: then the fp32 mask wins
        const float relu = zm > 0.f ? zm : 0.f;
        if ((relu > 0.f) != (out[idx] > 0.f)) out[idx] = relu;
        ++nflag;
        if (take_b) ++nswap;
      }
      __syncwarp();
    }
  }
  if (lane == 0 && nflag) {
    atomicAdd(&g_tie_stats[0], nflag);
    if (nswap) atomicAdd(&g_tie_stats[1], nswap);
  }
}

// raw accumulators -> {sum1[64], sq1[64], sum2[128], sq2[128], sum3[C3], sq3[C3]} of the pre-ReLU conv
// outputs z = d + b:  sum z = S + n b,  sum z^2 = Q + 2 b S + n b^2;  conv1 is affine in the point.
__global__ void pointnet_moments_finalize_kernel(const double* __restrict__ raw, const float* __restrict__ W1,
                                                 const float* __restrict__ b1, const float* __restrict__ b2,
                                                 const float* __restrict__ b3, int C3, double n,
                                                 double* __restrict__ moments) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64) {
    const double wx = W1[i * 3], wy = W1[i * 3 + 1], wz = W1[i * 3 + 2], b = b1[i];
    const double* m = raw + RAW_PTS;
    const double S = wx * m[0] + wy * m[1] + wz * m[2];
    const double Q = wx * wx * m[3] + wy * wy * m[6] + wz * wz * m[8] + 2.0 * (wx * wy * m[4] + wx * wz * m[5] + wy * wz * m[7]);
    moments[i] += S + n * b;
    moments[64 + i] += Q + 2.0 * b * S + n * b * b;
  } else if (i < 64 + 128) {
    const int c = i - 64;
    const double b = b2[c], S = raw[RAW_S2 + c], Q = raw[RAW_Q2 + c];
    moments[128 + c] += S + n * b;
    moments[256 + c] += Q + 2.0 * b * S + n * b * b;
  } else if (i < 64 + 128 + C3) {
    const int c = i - 192;
    const double b = b3[c], S = raw[RAW_S3 + c], Q = raw[RAW_S3 + C3 + c];
    moments[384 + c] += S + n * b;
    moments[384 + C3 + c] += Q + 2.0 * b * S + n * b * b;
  }
}

}  // namespace

int debug_set_trace(long long* ptr) {
  SGA_CUDA(cudaMemcpyToSymbol(g_trace, &ptr, sizeof(ptr)));
  return SGA_OK;
}

// diagnostics: {near-ties re-evaluated, of those reordered} since the last reset
int debug_tie_stats(unsigned long long* host_out2, int reset) {
  if (host_out2) SGA_CUDA(cudaMemcpyFromSymbol(host_out2, g_tie_stats, 2 * sizeof(unsigned long long)));
  if (reset) {
    const unsigned long long z[2] = {0, 0};
    SGA_CUDA(cudaMemcpyToSymbol(g_tie_stats, z, sizeof(z)));
  }
  return SGA_OK;
}

size_t pointnet_tc_raw_doubles(int C3) { return (size_t)RAW_S3 + 2 * (size_t)C3; }

void pointnet_tc_set_max_ctas(int n) { set_persistent_cta_cap(n); }

// raw != nullptr: also accumulate the BatchNorm batch statistics (raw: zeroed scratch of
// pointnet_tc_raw_doubles(C3) doubles) and add them into moments[2*(64+128+C3)].
int pointnet_fwd_tc(const float* pts, int64_t N, int P, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* W3, const float* b3, int C3, float* out, int32_t* argmax,
                    double* raw, double* moments, cudaStream_t st) {
  SGA_REQUIRE(C3 >= 128 && C3 % 128 == 0, "sga_pointnet_fwd(TC): C3=%d must be a multiple of 128", C3);
  SGA_REQUIRE(P >= 1, "sga_pointnet_fwd(TC): P=%d", P);
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_done = true;
  }
  const int nby = (C3 + 255) / 256;
  int gx = persistent_ctas() / nby;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > N) gx = (int)N;
  dim3 grid(gx, nby);
#define SGA_PN_LAUNCH(A, M) \
  pointnet_fwd_tc_kernel<A, M><<<grid, kThreads, SMEM_BYTES, st>>>(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax, raw)
  if (raw) {
    if (argmax) SGA_PN_LAUNCH(true, true); else SGA_PN_LAUNCH(false, true);
  } else {
    if (argmax) SGA_PN_LAUNCH(true, false); else SGA_PN_LAUNCH(false, false);
  }
#undef SGA_PN_LAUNCH
  SGA_LAUNCH_CHECK();
  if (argmax && P <= kTieMaxP) {
    const int64_t total = N * (int64_t)C3;
    int64_t nb = (total + kFixThreads - 1) / kFixThreads;
    if (nb > 4 * (int64_t)sm_count()) nb = 4 * (int64_t)sm_count();
    pointnet_tie_fix_kernel<<<(unsigned)nb, kFixThreads, 0, st>>>(
        pts, total, P, W1, b1, W2, b2, W3, b3, C3, out, argmax);
    SGA_LAUNCH_CHECK();
  }
  if (raw) {
    const int n = 64 + 128 + C3;
    pointnet_moments_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(raw, W1, b1, b2, b3, C3, (double)N * (double)P, moments);
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}

}  // namespace sga
