// Fused PointNet feature encoder on the 5th-gen tensor cores (SGA_POINTNET_TC).
// Reference: src/aligner/networks/pointnet.py:140-163 -- conv1(3->64) ReLU, conv2(64->128) ReLU,
// conv3(128->C3) ReLU, max over the points of an object.  The reference materialises three
// activation tensors (3.8 GB at B=32) in HBM; here nothing but the points (12 B/point) and the
// pooled feature (4 B/channel) ever touches HBM.
//
// Numerics: fp32 operands are split x = hi + lo into two bf16 values and every product is
// evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM ("bf16x3"); the dropped
// terms are <= 2^-16 relative per product, i.e. ~1e-5 on the pooled feature -- inside the 1e-4
// parity gate, where a single bf16 or tf32 pass (4e-3 / 3e-4) is not.
//
// One persistent CTA per SM; 8 compute warps + 1 MMA-issuing warp.  Per 128-point tile:
//   S1  (CUDA cores) conv1+ReLU -> A1{hi,lo}  [128 pts x 64]   bf16, K-major, 128B-swizzled smem
//   MMA2 D2[pts x 128ch]   = A1 * W2^T         3 passes x 4 k-steps   (M=128 N=128 K=16)
//   E2  TMEM->regs, +b2, ReLU, split -> H2{hi,lo} [128 pts x 128], released in four 32-channel
//       chunks so that conv3 starts after the first quarter of the epilogue
//   MMA3 D3[ch x 128 pts] = W3 * H2^T          2 M-tiles x 3 passes x 8 k-steps
//   E3  TMEM->regs, running max over columns (= points): channels sit on TMEM lanes, so the
//       max-pool is thread-local; bias + ReLU are applied once per object after the max
//       (max_p relu(z_p + b) == relu(max_p z_p + b)).
// Overlap: the H2 half that doubles as the A1 buffer is released by a tcgen05.commit right after
// conv3 has consumed it (half-way through MMA3), so conv1 of tile g+1 runs under MMA3(g), MMA2(g+1)
// is queued directly behind MMA3(g) and runs under E3(g); the tensor pipe only idles for the first
// quarter of E2.  Points of the next tile are prefetched into registers one stage ahead.
// W2 / W3 (hi and lo) stay resident in shared memory for the life of the CTA (160 KiB).
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
constexpr uint32_t W3HI = W2LO + kBlk;        // 4 blocks: (mt*2 + ka)
constexpr uint32_t W3LO = W3HI + 4 * kBlk;
constexpr uint32_t H2HI = W3LO + 4 * kBlk;    // 2 blocks (ka); block 0 doubles as A1HI
constexpr uint32_t H2LO = H2HI + 2 * kBlk;    // 2 blocks;      block 0 doubles as A1LO
constexpr uint32_t SMALL = H2LO + 2 * kBlk;   // 229376
constexpr uint32_t W1B1 = SMALL;              // float4[64] = {w0,w1,w2,b}
constexpr uint32_t B2 = W1B1 + 1024;          // float[128]
constexpr uint32_t BARS = B2 + 512;           // 9 mbarriers
constexpr uint32_t TMEMPTR = BARS + 80;
constexpr uint32_t SMEM_USED = TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;   // + alignment slack

constexpr uint32_t D2_COL = 0;
constexpr uint32_t D3_COL = 128;              // + mt*128
constexpr int kTmemCols = 512;

enum { BAR_A1_FULL = 0, BAR_D2_FULL = 1, BAR_H2_FULL = 2 /*..5*/, BAR_D3_FULL = 6, BAR_D3_FREE = 7, BAR_BLK0_FREE = 8 };

__device__ __forceinline__ uint32_t pack2(float a, float b) {   // a -> low half (lower address)
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

__device__ __forceinline__ void st_chunk(unsigned char* base, uint32_t off, const uint4& v) {
  *reinterpret_cast<uint4*>(base + off) = v;
}

template <bool kArgmax>
__global__ void __launch_bounds__(kThreads, 1)
pointnet_fwd_tc_kernel(const float* __restrict__ pts, int64_t N, int P,
                       const float* __restrict__ W1, const float* __restrict__ b1,
                       const float* __restrict__ W2, const float* __restrict__ b2,
                       const float* __restrict__ W3, const float* __restrict__ b3, int C3,
                       float* __restrict__ out, int32_t* __restrict__ argmax) {
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
    st_chunk(sm, W2HI + off, hi);
    st_chunk(sm, W2LO + off, lo);
  }
  for (int i = tid; i < nmt * 128 * 16; i += kThreads) {   // W3 block rows [nmt*128][128]: 16 chunks per row
    int r = i >> 4, j = i & 15;
    const float4* src = reinterpret_cast<const float4*>(W3 + (int64_t)(cb0 + r) * 128 + j * 8);
    float4 a = src[0], b = src[1];
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    split8(f, hi, lo);
    uint32_t off = (uint32_t)((r >> 7) * 2 + (j >> 3)) * kBlk + ptx::sw128_offset(r & 127, j & 7);
    st_chunk(sm, W3HI + off, hi);
    st_chunk(sm, W3LO + off, lo);
  }
  for (int i = tid; i < 64; i += kThreads)
    reinterpret_cast<float4*>(sm + W1B1)[i] = make_float4(W1[i * 3], W1[i * 3 + 1], W1[i * 3 + 2], b1[i]);
  for (int i = tid; i < 128; i += kThreads) reinterpret_cast<float*>(sm + B2)[i] = b2[i];
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_A1_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_D2_FULL], 1);
    for (int c = 0; c < 4; ++c) ptx::mbar_init(&bars[BAR_H2_FULL + c], kComputeThreads);
    ptx::mbar_init(&bars[BAR_D3_FULL], 1);
    ptx::mbar_init(&bars[BAR_D3_FREE], kComputeThreads);
    ptx::mbar_init(&bars[BAR_BLK0_FREE], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

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
    const uint64_t dA1hi = ptx::smem_desc_sw128(sm_base + H2HI), dA1lo = ptx::smem_desc_sw128(sm_base + H2LO);
    const uint64_t dW2hi = ptx::smem_desc_sw128(sm_base + W2HI), dW2lo = ptx::smem_desc_sw128(sm_base + W2LO);
    const uint64_t dW3hi = ptx::smem_desc_sw128(sm_base + W3HI), dW3lo = ptx::smem_desc_sw128(sm_base + W3LO);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    for (int64_t g = 0; g < G; ++g) {
      const uint32_t ph = (uint32_t)(g & 1);
      // ---- conv2: D2[pts x 128] = A1 * W2^T
      ptx::mbar_wait(&bars[BAR_A1_FULL], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        SGA_TRACE(1024, g, 0);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t ab = (pass == 1) ? dA1lo : dA1hi;   // A1 aliases block 0 of H2
          const uint64_t bb = (pass == 2) ? dW2lo : dW2hi;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16(tmem_u + D2_COL, ab + (uint64_t)(ks * 2), bb + (uint64_t)(ks * 2), idesc, (pass | ks) != 0);
        }
        ptx::umma_commit(&bars[BAR_D2_FULL]);
        SGA_TRACE(1024, g, 1);
      }
      __syncwarp();
      // ---- conv3: D3[mt][ch x pts] = W3[mt] * H2^T, released chunk by chunk (32 channels of K)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::mbar_wait(&bars[BAR_H2_FULL + c], ph);
        if (c == 0 && g > 0) ptx::mbar_wait(&bars[BAR_D3_FREE], (uint32_t)((g - 1) & 1));
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          SGA_TRACE(1024, g, 2 + c);
          constexpr uint32_t kBlk16 = kBlk >> 4;
          const uint64_t koff = (uint64_t)((c >> 1) * kBlk16 + (c & 1) * 4);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt < nmt) {
#pragma unroll
              for (int pass = 0; pass < 3; ++pass) {
                const uint64_t ab = ((pass == 2) ? dW3lo : dW3hi) + (uint64_t)(mt * 2 * kBlk16) + koff;
                const uint64_t bb = ((pass == 1) ? dA1lo : dA1hi) + koff;   // H2{hi,lo} base == A1 base
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  ptx::umma_bf16(tmem_u + D3_COL + mt * 128, ab + (uint64_t)(ks * 2), bb + (uint64_t)(ks * 2), idesc, (c | pass | ks) != 0);
              }
            }
          }
          // chunks 0,1 live in H2 block 0, which doubles as the next tile's A1 buffer
          if (c == 1) ptx::umma_commit(&bars[BAR_BLK0_FREE]);
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

    // conv1 mapping: thread -> 8 channels (warp-uniform group cg) x 4 points (pg + 32 i).  The 8
    // channels' weights live in registers for the whole kernel, so conv1 issues no shared-memory
    // loads (the tensor pipe saturates shared-memory bandwidth while conv3 runs).
    const int pg = tid & 31, cg = tid >> 5;
    float4 wreg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) wreg[e] = w1b1[8 * cg + e];
    float px[4], py[4], pz[4];               // this thread's points of the NEXT tile to encode
    auto prefetch = [&](int64_t g) {
      const int64_t n = blockIdx.x + (g / ntile) * (int64_t)gridDim.x;
      const int t = (int)(g % ntile);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pi = min(t * kTile + pg + 32 * i, P - 1);
        const float* pp = pts + (n * P + pi) * 3;
        px[i] = __ldg(pp); py[i] = __ldg(pp + 1); pz[i] = __ldg(pp + 2);
      }
    };
    auto stage1 = [&]() {
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
        st_chunk(sm, H2HI + off, hi);
        st_chunk(sm, H2LO + off, lo);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_A1_FULL]);
    };

    float rmax = -INFINITY;
    int ridx = 0;
    if (G > 0) {
      prefetch(0);
      stage1();
    }
    for (int64_t g = 0; g < G; ++g) {
      const uint32_t ph = (uint32_t)(g & 1);
      const int t = (int)(g % ntile);
      if (g + 1 < G) prefetch(g + 1);
      // ---- E2: conv2 epilogue, 4 chunks of 32 channels; this warp converts 16 of each
      ptx::mbar_wait(&bars[BAR_D2_FULL], ph);
      ptx::tc_fence_after();
      SGA_TRACE(0, g, 0);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int k0 = 32 * c + 16 * wh;
        uint32_t v[16];
        ptx::tmem_ld16(tmem + lane_addr + D2_COL + k0, v);
        ptx::tmem_ld_wait();
        float f0[8], f1[8];
        {
          const float4* bq = reinterpret_cast<const float4*>(b2s + k0);
          const float4 q0 = bq[0], q1 = bq[1], q2 = bq[2], q3 = bq[3];
          const float bb[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = __uint_as_float(v[e]) + bb[e];
            float b = __uint_as_float(v[8 + e]) + bb[8 + e];
            f0[e] = a > 0.f ? a : 0.f;
            f1[e] = b > 0.f ? b : 0.f;
          }
        }
        uint4 h0, l0, h1, l1;
        split8(f0, h0, l0);
        split8(f1, h1, l1);
        const uint32_t blk = (uint32_t)(k0 >> 6) * kBlk;
        const int j0 = (k0 & 63) >> 3;
        const uint32_t o0 = blk + ptx::sw128_offset(row, j0), o1 = blk + ptx::sw128_offset(row, j0 + 1);
        st_chunk(sm, H2HI + o0, h0);
        st_chunk(sm, H2HI + o1, h1);
        st_chunk(sm, H2LO + o0, l0);
        st_chunk(sm, H2LO + o1, l1);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[BAR_H2_FULL + c]);
        SGA_TRACE(0, g, 1 + c);
      }
      // ---- conv1 of the next tile as soon as conv3 has consumed H2 block 0 (= the A1 buffer);
      //      its conv2 is then queued on the tensor pipe directly behind this tile's conv3
      if (g + 1 < G) {
        ptx::mbar_wait(&bars[BAR_BLK0_FREE], ph);
        SGA_TRACE(0, g, 5);
        stage1();
        SGA_TRACE(0, g, 6);
      }
      // ---- conv3 done: D3 is ours
      ptx::mbar_wait(&bars[BAR_D3_FULL], ph);
      ptx::tc_fence_after();
      SGA_TRACE(0, g, 7);
      // ---- E3: running max over the 128 points (columns) of this tile
      if (wh < nmt) {
        const uint32_t base = tmem + lane_addr + D3_COL + (uint32_t)wh * 128;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          ptx::tmem_ld32(base + cc * 32, v);
          ptx::tmem_ld_wait();
          if (kArgmax) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              float f = __uint_as_float(v[e]);
              if (f > rmax) {
                rmax = f;
                ridx = t * kTile + cc * 32 + e;
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
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_D3_FREE]);
      SGA_TRACE(0, g, 8);
      if (t == ntile - 1) {
        if (wh < nmt) {
          const int64_t n = blockIdx.x + (g / ntile) * (int64_t)gridDim.x;
          const int ch = cb0 + wh * 128 + row;
          const float o = rmax + b3[ch];
          out[n * C3 + ch] = o > 0.f ? o : 0.f;
          if (kArgmax) argmax[n * C3 + ch] = min(ridx, P - 1);
        }
        rmax = -INFINITY;
        ridx = 0;
      }
    }
  }
  // ---------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace

int debug_set_trace(long long* ptr) {
  SGA_CUDA(cudaMemcpyToSymbol(g_trace, &ptr, sizeof(ptr)));
  return SGA_OK;
}

int pointnet_fwd_tc(const float* pts, int64_t N, int P, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* W3, const float* b3, int C3, float* out, int32_t* argmax,
                    cudaStream_t st) {
  SGA_REQUIRE(C3 >= 128 && C3 % 128 == 0, "sga_pointnet_fwd(TC): C3=%d must be a multiple of 128", C3);
  SGA_REQUIRE(P >= 1, "sga_pointnet_fwd(TC): P=%d", P);
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_done = true;
  }
  const int nby = (C3 + 255) / 256;
  int gx = sm_count() / nby;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > N) gx = (int)N;
  dim3 grid(gx, nby);
  if (argmax)
    pointnet_fwd_tc_kernel<true><<<grid, kThreads, SMEM_BYTES, st>>>(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax);
  else
    pointnet_fwd_tc_kernel<false><<<grid, kThreads, SMEM_BYTES, st>>>(pts, N, P, W1, b1, W2, b2, W3, b3, C3, out, argmax);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
