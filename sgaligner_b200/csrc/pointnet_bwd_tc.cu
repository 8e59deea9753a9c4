// Backward of the PointNet feature encoder on the tensor cores (autograd of
// src/aligner/networks/pointnet.py:140-163 through the max-pool; BatchNorm layers get no gradient
// because their outputs are discarded).
//
// Only the point that attains the max of channel c of object n receives gradient for that channel,
// so the backward works on "instances" (n, c, p = argmax[n][c]): one tile = one object x one block of
// 128 channels = 128 instances, instance r <-> channel c_r = cb + r.  Per tile (bf16x3 split operands,
// fp32 accumulation in TMEM, exactly like the forward):
//   S    gather x_r = pts[n, p_r], g_r = grad_out if out > 0;  h1 = relu(W1 x + b1) -> A1{hi,lo} [128 x 64]
//   MMA2 D2[r x k2]   = A1 W2^T                           (conv2 recomputed for the 128 argmax points)
//   E2   z2 = D2 + b2; h2 = relu(z2);  dW3[c_r,:] += g_r h2[r,:]   (thread-private: a thread owns a row)
//        dz2[r,:] = (z2 > 0) g_r W3[c_r,:]  -> DZ2{hi,lo} [128 x 128]
//   MMA5 D5[k2 x k1] += DZ2^T A1     (dW2: both operands read MN-major from the tiles above; the accumulator
//                                      lives in TMEM for the whole kernel)
//   MMA8 D8[k2 x 16]  += DZ2^T 1     (db2 = column sums of dz2 as a product with a tile of ones: the same MN-major
//                                      DZ2 operand, N = 16, accumulator resident in TMEM -- replaces four 16-value
//                                      warp transpose-reduces per thread and tile, 19 % of the samples before)
//   MMA6 D6[r x k1]   = DZ2 W2       (dh1)
//   E1   dz1 = (h1 > 0) D6;  dW1 += dz1^T x, db1 += sum_r dz1         (warp transpose-reduce)
// Tile pipeline: A1 and the instance table are double-buffered, and a compute thread does S(t+1) (gather, conv1,
// A1 store) between handing DZ2(t) to the tensor pipe and reading D6(t) back, so MMA5/8/6(t) and MMA2(t+1) run
// under CUDA-core work instead of in front of a waiting CTA.
// The reference-equivalent SIMT kernel (pointnet_bwd.cu) stays as the path for C3 % 128 != 0 and as the
// on-GPU cross-check.
#include "common.cuh"
#include "ptx.cuh"

namespace sga {
namespace {

constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;
constexpr uint32_t kBlk = 16384;

// shared-memory map
constexpr uint32_t W2HI = 0;                   // [128 k2 rows][64 k1]  K-major: B of MMA2
constexpr uint32_t W2LO = W2HI + kBlk;
constexpr uint32_t W2THI = W2LO + kBlk;        // 2 blocks (k2 halves) of [64 k1 rows][64 k2]  K-major: B of MMA6
constexpr uint32_t W2TLO = W2THI + kBlk;       //   (a block is 64 rows x 128 B = 8 KiB)
constexpr uint32_t A1HI = W2TLO + kBlk;        // 2 buffers (tile parity) of [128 r][64 k1]
constexpr uint32_t A1LO = A1HI + 2 * kBlk;
constexpr uint32_t DZHI = A1LO + 2 * kBlk;     // 2 blocks (k2 halves) of [128 r][64 k2]
constexpr uint32_t DZLO = DZHI + 2 * kBlk;
constexpr uint32_t ONES = DZLO + 2 * kBlk;     // 196608: [16 rows][64 x bf16 1.0]: B operand of the db2 column-sum MMA
constexpr uint32_t SMALL = ONES + 2048;
constexpr uint32_t W1B1 = SMALL;               // float4[64] = {w0,w1,w2,b}
constexpr uint32_t B2 = W1B1 + 1024;           // float[128]
constexpr uint32_t XS = B2 + 512;              // 2 buffers of float4[128] = {x0,x1,x2,g} of a tile
constexpr uint32_t BARS = XS + 2 * 2048;
constexpr uint32_t TMEMPTR = BARS + 64;
constexpr uint32_t SMEM_USED = TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;

constexpr uint32_t D2_COL = 0, D5_COL = 128, D6_COL = 192, D8_COL = 256, W3_COL = 288;   // W3_COL: 128 columns, fp32 W3[cb0 + lane][:]
constexpr int kTmemCols = 512;

enum { BAR_A1_FULL = 0, BAR_D2_FULL = 1, BAR_DZ_FULL = 2, BAR_D6_FULL = 3 };

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
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
__device__ __forceinline__ void st_chunk(uint32_t smem_base, uint32_t off, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_base + off), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16-byte shared-memory load on a 32-bit shared address (pointers derived from the re-aligned dynamic shared memory
// base are generic to the compiler: LD.E instead of LDS, and a float4 read came out as 8 + 4 bytes with 4x the wavefronts)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// MN-major view of a [rows = K index][64 x bf16 = 128 B] SWIZZLE_128B tile (the same bytes MMA2 reads
// K-major): 64-element MN atoms `lbo` bytes apart, groups of 8 K rows 1024 B apart.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Butterfly transpose-reduce over the 32 lanes (see pointnet_tc.cu): 16 values per lane -> the total of
// element e16(L) on lanes 2j, 2j+1.
__device__ __forceinline__ float transpose_reduce16(float (&x)[16], int lane) {
#pragma unroll
  for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? x[i + half] : x[i];
      const float send = up ? x[i] : x[i + half];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return x[0] + __shfl_xor_sync(0xffffffffu, x[0], 1);
}
__device__ __forceinline__ int e16(int lane) { return 8 * ((lane >> 4) & 1) + 4 * ((lane >> 3) & 1) + 2 * ((lane >> 2) & 1) + ((lane >> 1) & 1); }

__global__ void __launch_bounds__(kThreads, 1)
pointnet_bwd_tc_kernel(const float* __restrict__ pts, int64_t N, int P,
                       const float* __restrict__ W1, const float* __restrict__ b1,
                       const float* __restrict__ W2, const float* __restrict__ b2,
                       const float* __restrict__ W3, int C3,
                       const float* __restrict__ out, const int32_t* __restrict__ argmax,
                       const float* __restrict__ gout,
                       float* __restrict__ gW1, float* __restrict__ gb1, float* __restrict__ gW2,
                       float* __restrict__ gb2, float* __restrict__ gW3, float* __restrict__ gb3) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TMEMPTR);
  const float4* w1b1 = reinterpret_cast<const float4*>(sm + W1B1);
  const float* b2s = reinterpret_cast<const float*>(sm + B2);
  float4* xs = reinterpret_cast<float4*>(sm + XS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb0 = blockIdx.y * 128;

  // ---------------- one-time setup
  for (int i = tid; i < 128 * 8; i += kThreads) {          // W2 [128 k2][64 k1]: 8 chunks per row (B of MMA2)
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
  for (int i = tid; i < 64 * 16; i += kThreads) {          // W2^T [64 k1][128 k2]: 16 chunks per row (B of MMA6)
    int k1 = i >> 4, j = i & 15;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = W2[(j * 8 + e) * 64 + k1];
    uint4 hi, lo;
    split8(f, hi, lo);
    uint32_t off = (uint32_t)(j >> 3) * (kBlk / 2) + ptx::sw128_offset(k1, j & 7);
    st_chunk(sm_base, W2THI + off, hi);
    st_chunk(sm_base, W2TLO + off, lo);
  }
  for (int i = tid; i < 64; i += kThreads)
    reinterpret_cast<float4*>(sm + W1B1)[i] = make_float4(W1[i * 3], W1[i * 3 + 1], W1[i * 3 + 2], b1[i]);
  for (int i = tid; i < 128; i += kThreads) reinterpret_cast<float*>(sm + B2)[i] = b2[i];
  for (int i = tid; i < 2048 / 4; i += kThreads) reinterpret_cast<uint32_t*>(sm + ONES)[i] = 0x3F803F80u;   // bf16 1.0 pairs
  if (tid == 0) {
    ptx::mbar_init(&bars[BAR_A1_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_D2_FULL], 1);
    ptx::mbar_init(&bars[BAR_DZ_FULL], kComputeThreads);
    ptx::mbar_init(&bars[BAR_D6_FULL], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // W3[cb0 + r][0..127] (fp32) -> tensor memory, lane r: dz2 = g W3 reads 64 of them per thread and tile; from global
  // memory those loads missed the (shared-memory-squeezed) L1 every tile and the epilogue waited on L2
  if (warp < 4) {
    const int r = 32 * warp + lane;
    const float4* src = reinterpret_cast<const float4*>(W3 + (int64_t)(cb0 + r) * 128);
#pragma unroll 1
    for (int grp = 0; grp < 8; ++grp) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = src[grp * 4 + j];
        w[4 * j] = __float_as_uint(a.x); w[4 * j + 1] = __float_as_uint(a.y); w[4 * j + 2] = __float_as_uint(a.z); w[4 * j + 3] = __float_as_uint(a.w);
      }
      ptx::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + W3_COL + grp * 16, w);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  const int64_t ntile = (N > (int64_t)blockIdx.x) ? (N - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc2 = ptx::make_idesc(1, 128, 128);
    const uint32_t idesc5 = ptx::make_idesc(1, 128, 64) | (1u << 15) | (1u << 16);   // both operands MN-major
    const uint32_t idesc6 = ptx::make_idesc(1, 128, 64);
    const uint64_t dA1hi = ptx::smem_desc_sw128(sm_base + A1HI), dA1lo = ptx::smem_desc_sw128(sm_base + A1LO);
    const uint64_t dW2hi = ptx::smem_desc_sw128(sm_base + W2HI), dW2lo = ptx::smem_desc_sw128(sm_base + W2LO);
    const uint64_t dDZhi = ptx::smem_desc_sw128(sm_base + DZHI), dDZlo = ptx::smem_desc_sw128(sm_base + DZLO);
    const uint64_t dWThi = ptx::smem_desc_sw128(sm_base + W2THI), dWTlo = ptx::smem_desc_sw128(sm_base + W2TLO);
    const uint64_t mDZhi = desc_mn_sw128(sm_base + DZHI, kBlk), mDZlo = desc_mn_sw128(sm_base + DZLO, kBlk);
    const uint64_t mA1hi = desc_mn_sw128(sm_base + A1HI, kBlk), mA1lo = desc_mn_sw128(sm_base + A1LO, kBlk);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc8 = ptx::make_idesc(1, 128, 16) | (1u << 15);               // A (DZ2) MN-major, B (ones) K-major
    const uint64_t dOnes = ptx::smem_desc_sw128(sm_base + ONES);
    for (int64_t t = 0; t < ntile; ++t) {
      const uint32_t ph = (uint32_t)(t & 1);
      const uint64_t abuf = (uint64_t)(ph * (kBlk >> 4));                            // A1 buffer of this tile
      ptx::mbar_wait(&bars[BAR_A1_FULL], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t ab = ((pass == 1) ? dA1lo : dA1hi) + abuf;
          const uint64_t bb = (pass == 2) ? dW2lo : dW2hi;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_bf16(tmem_u + D2_COL, ab + (uint64_t)(ks * 2), bb + (uint64_t)(ks * 2), idesc2, (pass | ks) != 0);
        }
        ptx::umma_commit(&bars[BAR_D2_FULL]);
      }
      __syncwarp();
      ptx::mbar_wait(&bars[BAR_DZ_FULL], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        // MMA5: D5[k2 x k1] += sum_r dz2[r,k2] h1[r,k1]   (K = 128 instances: 8 steps of 16 rows = 2048 B)
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t ab = (pass == 1) ? mDZlo : mDZhi;
          const uint64_t bb = ((pass == 2) ? mA1lo : mA1hi) + abuf;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            ptx::umma_bf16(tmem_u + D5_COL, ab + (uint64_t)(ks * 128), bb + (uint64_t)(ks * 128), idesc5, (t | pass | ks) != 0);
        }
        // MMA8: D8[k2 x 16] += sum_r dz2[r,k2] * 1   (db2; every B element is 1.0, so one 32-byte slice serves all steps)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const uint64_t ab = (pass == 1) ? mDZlo : mDZhi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            ptx::umma_bf16(tmem_u + D8_COL, ab + (uint64_t)(ks * 128), dOnes, idesc8, (t | pass | ks) != 0);
        }
        // MMA6: D6[r x k1] = sum_k2 dz2[r,k2] W2[k2,k1]   (K = 128 channels: 2 blocks x 4 steps of 32 B)
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t ab = (pass == 1) ? dDZlo : dDZhi;
          const uint64_t bb = (pass == 2) ? dWTlo : dWThi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ao = (uint64_t)((ks >> 2) * (kBlk >> 4) + (ks & 3) * 2);
            const uint64_t bo = (uint64_t)((ks >> 2) * (kBlk >> 5) + (ks & 3) * 2);
            ptx::umma_bf16(tmem_u + D6_COL, ab + ao, bb + bo, idesc6, (pass | ks) != 0);
          }
        }
        ptx::umma_commit(&bars[BAR_D6_FULL]);
      }
      __syncwarp();
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, wh = warp >> 2;
    const int row = 32 * q + lane;                       // TMEM lane = instance r
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int pg = tid & 31, cg = tid >> 5;              // conv1 mapping: 8 channels (cg) x 4 rows (pg + 32 i)

    // persistent accumulators
    float acc3[64];                                      // dW3[cb0 + row][64*wh + j]
#pragma unroll
    for (int j = 0; j < 64; ++j) acc3[j] = 0.f;
    float accb3 = 0.f;                                   // db3[cb0 + row]  (wh == 0)
    float accw1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // [16-chunk cc][d = 0,1,2 | bias] of k1 = 32*wh + 16*cc + e16(lane)

    // instance prefetch (threads < 128: one instance each): raw (argmax, out, grad_out) two tiles ahead of
    // their use, the gathered point one tile ahead
    float nx0 = 0.f, nx1 = 0.f, nx2 = 0.f, ng = 0.f;     // tile t   (gathered)
    int pp = 0;                                          // tile t+1 (raw)
    float po = 0.f, pgo = 0.f;
    auto load_raw = [&](int64_t t) {
      if (tid < 128 && t < ntile) {
        const int64_t n = blockIdx.x + t * (int64_t)gridDim.x;
        const int64_t o = n * C3 + cb0 + tid;
        pp = argmax[o]; po = out[o]; pgo = gout[o];
      }
    };
    auto gather = [&](int64_t t) {      // consumes the raw values of tile t
      if (tid < 128 && t < ntile) {
        const int64_t n = blockIdx.x + t * (int64_t)gridDim.x;
        const int p = min(max(pp, 0), P - 1);
        const float* src = pts + (n * P + p) * 3;
        nx0 = __ldg(src); nx1 = __ldg(src + 1); nx2 = __ldg(src + 2);
        ng = po > 0.f ? pgo : 0.f;
      }
    };
    // ---- S(k): instance table + conv1 -> A1, both in the buffers of tile parity k & 1; then the prefetch chain moves on
    auto stage_s = [&](int64_t k) {
      const uint32_t buf = (uint32_t)(k & 1);
      float4* xb = xs + 128 * buf;
      if (tid < 128) xb[tid] = make_float4(nx0, nx1, nx2, ng);
      compute_barrier();
      gather(k + 1);
      load_raw(k + 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 xr = lds128(sm_base + XS + (128 * buf + pg + 32 * i) * 16);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 w = lds128(sm_base + W1B1 + (8 * cg + e) * 16);     // broadcast LDS: registers are the scarce resource here (acc3)
          float v = fmaf(w.x, xr.x, fmaf(w.y, xr.y, fmaf(w.z, xr.z, w.w)));
          f[e] = v > 0.f ? v : 0.f;
        }
        uint4 hi, lo;
        split8(f, hi, lo);
        uint32_t off = buf * kBlk + ptx::sw128_offset(pg + 32 * i, cg);
        st_chunk(sm_base, A1HI + off, hi);
        st_chunk(sm_base, A1LO + off, lo);
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars[BAR_A1_FULL]);
    };

    load_raw(0);
    gather(0);
    load_raw(1);
    if (ntile > 0) stage_s(0);

    for (int64_t t = 0; t < ntile; ++t) {
      const uint32_t ph = (uint32_t)(t & 1);
      // ---- E2
      const float4 me = lds128(sm_base + XS + (128 * ph + row) * 16);
      const float g = me.w;
      if (wh == 0) accb3 += g;
      const float tau = 3e-5f * (1.f + fabsf(me.x) + fabsf(me.y) + fabsf(me.z));
      ptx::mbar_wait(&bars[BAR_D2_FULL], ph);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        ptx::tmem_ld16(tmem + lane_addr + D2_COL + 64 * wh + 16 * c, v);
        uint32_t wwu[16];                      // W3[c_r][64 wh + 16 c ..]: resident in tensor memory (lane = row), see setup
        ptx::tmem_ld16(tmem + lane_addr + W3_COL + 64 * wh + 16 * c, wwu);
        const uint32_t bq = sm_base + B2 + (64 * wh + 16 * c) * 4;
        const float4 q0 = lds128(bq), q1 = lds128(bq + 16), q2 = lds128(bq + 32), q3 = lds128(bq + 48);
        const float bb[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
        ptx::tmem_ld_wait();
        float z[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = __uint_as_float(v[e]) + bb[e];
        // ReLU kink: a pre-activation within the bf16x3 error of zero could land on the other side of the
        // kink than in fp32, which would add / drop a whole gradient term.  Those (rare) elements are
        // recomputed on the FMA pipe in the reference summation order, so the mask is the fp32 mask.
        uint32_t near = 0;
#pragma unroll
        for (int e = 0; e < 16; ++e) near |= (fabsf(z[e]) < tau) ? (1u << e) : 0u;
        // ~10 % of the chunks have one: the WARP evaluates it together (lane k takes conv1 channels k and k + 32 of the
        // flagged instance, one shuffle reduction per element) instead of every lane running the 64-term sum
        uint32_t flagged = __ballot_sync(0xffffffffu, near != 0);
        while (flagged) {                                // warp-uniform
          const int L = __ffs(flagged) - 1;
          flagged &= flagged - 1;
          uint32_t nb = __shfl_sync(0xffffffffu, near, L);
          const float mx = __shfl_sync(0xffffffffu, me.x, L), my = __shfl_sync(0xffffffffu, me.y, L), mz = __shfl_sync(0xffffffffu, me.z, L);
          const float4 wa = lds128(sm_base + W1B1 + lane * 16), wb = lds128(sm_base + W1B1 + (lane + 32) * 16);
          float ha = fmaf(wa.x, mx, fmaf(wa.y, my, fmaf(wa.z, mz, wa.w)));
          float hb = fmaf(wb.x, mx, fmaf(wb.y, my, fmaf(wb.z, mz, wb.w)));
          ha = ha > 0.f ? ha : 0.f;
          hb = hb > 0.f ? hb : 0.f;
          while (nb) {
            const int e = __ffs(nb) - 1;
            nb &= nb - 1;
            const int k2 = 64 * wh + 16 * c + e;
            float part = fmaf(ha, __ldg(W2 + k2 * 64 + lane), hb * __ldg(W2 + k2 * 64 + lane + 32));
            part = warp_sum(part);
            const float zz = part + b2s[k2];
            if (lane == L) {
#pragma unroll
              for (int ee = 0; ee < 16; ++ee)
                if (ee == e) z[ee] = zz;
            }
          }
        }
        float dz[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const bool on = z[e] > 0.f;
          acc3[16 * c + e] = fmaf(g, on ? z[e] : 0.f, acc3[16 * c + e]);
          dz[e] = on ? g * __uint_as_float(wwu[e]) : 0.f;
        }
        const float f0[8] = {dz[0], dz[1], dz[2], dz[3], dz[4], dz[5], dz[6], dz[7]};
        const float f1[8] = {dz[8], dz[9], dz[10], dz[11], dz[12], dz[13], dz[14], dz[15]};
        uint4 h0, l0, h1, l1;
        split8(f0, h0, l0);
        split8(f1, h1, l1);
        const uint32_t blk = (uint32_t)wh * kBlk;
        const uint32_t o0 = blk + ptx::sw128_offset(row, 2 * c), o1 = blk + ptx::sw128_offset(row, 2 * c + 1);
        st_chunk(sm_base, DZHI + o0, h0);
        st_chunk(sm_base, DZHI + o1, h1);
        st_chunk(sm_base, DZLO + o0, l0);
        st_chunk(sm_base, DZLO + o1, l1);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[BAR_DZ_FULL]);

      // ---- S(t+1) while the tensor pipe works through MMA5/8/6(t): the other A1 / instance-table buffers were last
      //      read by the MMAs and threads of tile t-1, all complete once D6_FULL(t-1) was observed
      if (t + 1 < ntile) stage_s(t + 1);

      // ---- E1: dz1 = (h1 > 0) dh1;  dW1 / db1 partial sums over the 32 instances of this warp
      ptx::mbar_wait(&bars[BAR_D6_FULL], ph);
      ptx::tc_fence_after();
      {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + lane_addr + D6_COL + 32 * wh, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float dz1[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float4 w = lds128(sm_base + W1B1 + (32 * wh + 16 * cc + e) * 16);
            const float z1 = fmaf(w.x, me.x, fmaf(w.y, me.y, fmaf(w.z, me.z, w.w)));
            dz1[e] = z1 > 0.f ? __uint_as_float(v[16 * cc + e]) : 0.f;
          }
          float x[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) x[e] = dz1[e] * me.x;
          accw1[4 * cc + 0] += transpose_reduce16(x, lane);
#pragma unroll
          for (int e = 0; e < 16; ++e) x[e] = dz1[e] * me.y;
          accw1[4 * cc + 1] += transpose_reduce16(x, lane);
#pragma unroll
          for (int e = 0; e < 16; ++e) x[e] = dz1[e] * me.z;
          accw1[4 * cc + 2] += transpose_reduce16(x, lane);
          accw1[4 * cc + 3] += transpose_reduce16(dz1, lane);
        }
      }
      // tcgen05.ld above is ordered before the next MMA6 by the tcgen05.fence in the next E2.
    }

    // ---------------- flush
    if (ntile > 0) {
#pragma unroll
      for (int j = 0; j < 64; ++j) atomicAdd(&gW3[(int64_t)(cb0 + row) * 128 + 64 * wh + j], acc3[j]);
      if (wh == 0) atomicAdd(&gb3[cb0 + row], accb3);
      if ((lane & 1) == 0) {      // transpose_reduce16: lanes 2j, 2j+1 hold the same total
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int k1 = 32 * wh + 16 * cc + e16(lane);
          atomicAdd(&gW1[k1 * 3 + 0], accw1[4 * cc + 0]);
          atomicAdd(&gW1[k1 * 3 + 1], accw1[4 * cc + 1]);
          atomicAdd(&gW1[k1 * 3 + 2], accw1[4 * cc + 2]);
          atomicAdd(&gb1[k1], accw1[4 * cc + 3]);
        }
      }
      // dW2 and db2 from the TMEM accumulators: lane = k2, columns = k1 (D5) / 16 identical column sums (D8)
      ptx::tc_fence_after();
      if (wh == 0) {
        uint32_t b[16];
        ptx::tmem_ld16(tmem + lane_addr + D8_COL, b);
        ptx::tmem_ld_wait();
        atomicAdd(&gb2[row], __uint_as_float(b[0]));
      }
      uint32_t v[32];
      ptx::tmem_ld32(tmem + lane_addr + D5_COL + 32 * wh, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) atomicAdd(&gW2[row * 64 + 32 * wh + e], __uint_as_float(v[e]));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace

int pointnet_bwd_tc(const float* pts, int64_t N, int P, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* W3, int C3, const float* out, const int32_t* argmax, const float* gout, float* gW1, float* gb1,
                    float* gW2, float* gb2, float* gW3, float* gb3, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(pointnet_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_done = true;
  }
  const int nby = C3 / 128;
  int gx = persistent_ctas() / nby;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > N) gx = (int)N;
  dim3 grid(gx, nby);
  pointnet_bwd_tc_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(pts, N, P, W1, b1, W2, b2, W3, C3, out, argmax, gout, gW1, gb1, gW2, gb2, gW3, gb3);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
