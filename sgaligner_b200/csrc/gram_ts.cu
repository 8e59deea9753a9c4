// Loss forward Grams, second generation: operands never pass through registers on their way to the tensor cores.
//
//   F[dir] = P_dir [seg0; seg1; seg2]^T        (src/aligner/losses.py:5-15: the matmuls of calculate_prob_dist)
//
// gemm_tc.cu splits fp32 rows into tf32 hi/lo INSIDE the GEMM (global -> registers -> shared memory).  What bounds that
// design: an SS-mode tcgen05.mma at M = N = 128 already uses the SM's whole shared-memory bandwidth, so
// the loaders' hi/lo stores add directly to the MMA time, and every generic-proxy hand-off costs a MEMBAR that drains
// the loader's own prefetch.  Here instead:
//   * `pack_rows_kernel` L2-normalises every row ONCE per step and writes it, already split (round-to-nearest tf32
//     hi / lo) and already in the 128B-swizzled K-major tile image the MMA reads, into one image per row set
//     (e1i, e2i, e1j, e2j; tiles of 128 rows x chunks of 32 k; 16 KiB hi + 16 KiB lo per chunk);
//   * the B operand of a tile is streamed by `cp.async.bulk` (TMA, one 32 KiB copy per K chunk, 6-deep ring, one
//     producer lane, completion on an mbarrier: no loader warps, no proxy fences, nothing in registers);
//   * the A operand (128 anchor rows x all of K, hi and lo) is written ONCE per work item into TENSOR MEMORY
//     (tcgen05.st, 256 columns) and stays there while the CTA sweeps the column tiles: A costs no shared-memory
//     bandwidth at all, so per K chunk only B moves (32 KiB in + 3 x 16 KiB read = 80 KiB < the 98 KiB the three
//     MMA passes' 768 cycles allow).
// Three passes per k-step as everywhere (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM): fp32-faithful results.
// K <= 128 (the 100-d modality embeddings); wider embeddings (the joint) stay on gemm_tc.cu.
// Epilogue as in gemm_tc.cu: exp-sums of the non-anchor segments from registers, output rows through a padded
// shared-memory transpose so that every store is one contiguous 128-byte request.
#include "gram_ts.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"
#include "umma_tf32.cuh"

namespace sga {
namespace {

constexpr uint32_t kChunkBytes = 32768;            // one K chunk of one 128-row tile: hi 16 KiB | lo 16 KiB
// Two variants.  Narrow (K <= 128): A lives in tensor memory, a ring stage is one B chunk.  Wide (K <= 512, the joint
// embedding): A does not fit next to two accumulators, so a ring stage carries an A chunk AND a B chunk and the MMA
// reads both from shared memory (SS mode: shared-memory bandwidth caps it near 60 % tensor-active).
template <bool kWide>
struct Cfg {
  static constexpr int kStages = kWide ? 3 : 5;
  static constexpr uint32_t kStageBytes = kWide ? 2 * kChunkBytes : kChunkBytes;
  static constexpr int kXPitch = kWide ? 32 : 36;    // floats per row of an epilogue warp's [32 x 32] transpose tile
  static constexpr uint32_t BAR_OFF = kStages * kStageBytes;
  static constexpr uint32_t XPOSE_OFF = BAR_OFF + 256;
  static constexpr uint32_t SMEM_BYTES = XPOSE_OFF + 8 * 32 * kXPitch * 4 + 1024;
};
constexpr int kEpiWarps = 8;                       // two per TMEM lane quarter: each takes two of the four 32-column slabs
constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp 0 producer, warp 1 MMA issuer, warps 2..9 A staging + epilogue
constexpr uint32_t ACC_COL = 0, AHI_COL = 256, ALO_COL = 384;

// ------------------------------------------------------------------------------------------ operand images
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ X, int64_t N, int D, int nkc, const int32_t* __restrict__ slot,
                 unsigned char* img0, unsigned char* img1, unsigned char* img2, unsigned char* img3,
                 float* __restrict__ norms, float* __restrict__ Xh) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const float* x = X + row * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s = fmaf(x[k], x[k], s);
  s = warp_sum(s);
  const float nrm = sqrtf(s);
  const float den = fmaxf(nrm, 1e-12f);        // F.normalize: x / max(||x||, eps)
  if (lane == 0) norms[row] = nrm;
  for (int k = lane; k < D; k += 32) Xh[row * D + k] = x[k] / den;
  const int32_t sl = slot[row];
  if (sl < 0) return;
  const int set = sl >> 28, pos = sl & 0x0FFFFFFF;
  unsigned char* img = set == 0 ? img0 : (set == 1 ? img1 : (set == 2 ? img2 : img3));
  const int tile = pos >> 7, r = pos & 127;
  // 16-byte piece p of the row (4 k-values): chunk p/8, piece p%8
  for (int p = lane; p < 8 * nkc; p += 32) {
    const int k0 = 4 * p;
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = (k0 + e < D) ? x[k0 + e] / den : 0.f;      // K padding must be exact zeros
      h[e] = tf32x3::rn_tf32(v);
      l[e] = tf32x3::rn_tf32(v - __uint_as_float(h[e]));
    }
    unsigned char* dst = img + ((size_t)tile * nkc + (p >> 3)) * kChunkBytes + ptx::sw128_offset(r, p & 7);
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// slot[node] = (set << 28) | position inside the set; ridx / r2idx as build_ridx_kernel (the backward still uses them)
__global__ void __launch_bounds__(256)
build_slots_kernel(const int32_t* __restrict__ e1i, const int32_t* __restrict__ e2i, const int32_t* __restrict__ e1j,
                   const int32_t* __restrict__ e2j, int A, int J1, int J2, int32_t* __restrict__ slot) {
  const int total = 2 * A + J1 + J2;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < total; t += gridDim.x * 256) {
    if (t < A) slot[e1i[t]] = (0 << 28) | t;
    else if (t < 2 * A) slot[e2i[t - A]] = (1 << 28) | (t - A);
    else if (t < 2 * A + J1) slot[e1j[t - 2 * A]] = (2 << 28) | (t - 2 * A);
    else slot[e2j[t - 2 * A - J1]] = (3 << 28) | (t - 2 * A - J1);
  }
}

// ------------------------------------------------------------------------------------------ work decomposition
struct Item {
  int g;            // problem
  int m0;           // first anchor row
  int nt_beg, nt_end;
};
__device__ __forceinline__ Item find_item(const GramGroup& G, int item) {
  int g = 0, beg = 0;
  while (g + 1 < G.n && item >= G.item_end[g]) {
    beg = G.item_end[g];
    ++g;
  }
  const GramProblem& P = G.p[g];
  const int local = item - beg;
  const int S = G.nsplit[g];
  const int NT = (P.seg_rows[0] + 127) / 128 + (P.seg_rows[1] + 127) / 128 + (P.seg_rows[2] + 127) / 128;
  const int mt = local / S, part = local % S;
  Item it;
  it.g = g;
  it.m0 = mt * 128;
  it.nt_beg = (int)((int64_t)part * NT / S);
  it.nt_end = (int)((int64_t)(part + 1) * NT / S);
  return it;
}
struct NTile {
  int seg, lt, col0, valid;
};
__device__ __forceinline__ NTile find_ntile(const GramProblem& P, int j) {
  const int t0 = (P.seg_rows[0] + 127) / 128, t1 = (P.seg_rows[1] + 127) / 128;
  NTile t;
  if (j < t0) { t.seg = 0; t.lt = j; t.col0 = 0; }
  else if (j < t0 + t1) { t.seg = 1; t.lt = j - t0; t.col0 = P.seg_rows[0]; }
  else { t.seg = 2; t.lt = j - t0 - t1; t.col0 = P.seg_rows[0] + P.seg_rows[1]; }
  t.col0 += t.lt * 128;
  t.valid = min(128, P.seg_rows[t.seg] - t.lt * 128);
  return t;
}

template <bool kWide>
__global__ void __launch_bounds__(kThreads, 1)
gram_ts_kernel(const __grid_constant__ GramGroup G) {
  using C_ = Cfg<kWide>;
  constexpr int kStages = C_::kStages;
  constexpr uint32_t kStageBytes = C_::kStageBytes;
  constexpr int kXPitch = C_::kXPitch;
  constexpr uint32_t BAR_OFF = C_::BAR_OFF, XPOSE_OFF = C_::XPOSE_OFF;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + BAR_OFF);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;    // [2]
  uint64_t* acc_free = acc_full + 2;       // [2]
  uint64_t* a_ready = acc_free + 2;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
  float* xpose = reinterpret_cast<float*>(sm + XPOSE_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_free[i], 32 * kEpiWarps);
    }
    ptx::mbar_init(a_ready, 32 * kEpiWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nitems = G.item_end[G.n - 1];

  if (warp == 0) {
    // ------------------------------- producer: B chunks by bulk async copy
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const Item I = find_item(G, item);
        const GramProblem& P = G.p[I.g];
        for (int j = I.nt_beg; j < I.nt_end; ++j) {
          const NTile T = find_ntile(P, j);
          const unsigned char* src = P.b_img[T.seg] + (size_t)T.lt * P.nkc * kChunkBytes;
          const unsigned char* asrc = P.a_img + (size_t)(I.m0 >> 7) * P.nkc * kChunkBytes;
          for (int c = 0; c < P.nkc; ++c, ++it) {
            const int s = it % kStages;
            if (it >= kStages) ptx::mbar_wait(&empty[s], (uint32_t)(((it / kStages) - 1) & 1));
            ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
            unsigned char* dst = sm + (size_t)s * kStageBytes;
            if (kWide) {
              ptx::bulk_g2s(dst, asrc + (size_t)c * kChunkBytes, kChunkBytes, &full[s]);
              dst += kChunkBytes;
            }
            ptx::bulk_g2s(dst, src + (size_t)c * kChunkBytes, kChunkBytes, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer: A from tensor memory, B from the ring
    const uint32_t idesc = ptx::make_idesc(2, 128, 128);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    int it = 0, ti = 0, ii = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ii) {
      const Item I = find_item(G, item);
      const GramProblem& P = G.p[I.g];
      if (!kWide) {
        ptx::mbar_wait(a_ready, (uint32_t)(ii & 1));
        ptx::tc_fence_after();
      }
      for (int j = I.nt_beg; j < I.nt_end; ++j, ++ti) {
        const int ab = ti & 1;
        if (ti >= 2) {
          ptx::mbar_wait(&acc_free[ab], (uint32_t)(((ti >> 1) - 1) & 1));
          ptx::tc_fence_after();
        }
        for (int c = 0; c < P.nkc; ++c, ++it) {
          const int s = it % kStages;
          ptx::mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t st = sm_base + (uint32_t)s * kStageBytes;
            const uint32_t d = tmem_u + ACC_COL + ab * 128;
            if (kWide) {
              const uint64_t dAhi = ptx::smem_desc_sw128(st), dAlo = ptx::smem_desc_sw128(st + 16384);
              const uint64_t dBhi = ptx::smem_desc_sw128(st + kChunkBytes), dBlo = ptx::smem_desc_sw128(st + kChunkBytes + 16384);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) ptx::umma_tf32(d, dAhi + (uint64_t)(ks * 2), dBhi + (uint64_t)(ks * 2), idesc, (c | ks) != 0);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) ptx::umma_tf32(d, dAlo + (uint64_t)(ks * 2), dBhi + (uint64_t)(ks * 2), idesc, 1);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) ptx::umma_tf32(d, dAhi + (uint64_t)(ks * 2), dBlo + (uint64_t)(ks * 2), idesc, 1);
            } else {
              const uint64_t dBhi = ptx::smem_desc_sw128(st), dBlo = ptx::smem_desc_sw128(st + 16384);
              const uint32_t ahi = tmem_u + AHI_COL + c * 32, alo = tmem_u + ALO_COL + c * 32;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) ptx::umma_tf32_ts(d, ahi + ks * 8, dBhi + (uint64_t)(ks * 2), idesc, (c | ks) != 0);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) ptx::umma_tf32_ts(d, alo + ks * 8, dBhi + (uint64_t)(ks * 2), idesc, 1);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) ptx::umma_tf32_ts(d, ahi + ks * 8, dBlo + (uint64_t)(ks * 2), idesc, 1);
            }
            ptx::umma_commit(&empty[s]);
            if (c == P.nkc - 1) ptx::umma_commit(&acc_full[ab]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------- A staging (tensor memory) + epilogue
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;           // which half of the work of that quarter (slabs 2h, 2h+1; A chunks c = h mod 2)
    const int r = 32 * q + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    float s_acc[4] = {0.f, 0.f, 0.f, 0.f};   // {s01 seg1, s01 seg2, s1 seg1, s1 seg2} of problem cur_g
    int cur_g = -1;
    auto flush_sums = [&]() {
      if (cur_g < 0) return;
      const GramProblem& Q = G.p[cur_g];
      double* dst[4] = {Q.s01[0], Q.s01[1], Q.s1[0], Q.s1[1]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = warp_sum(s_acc[i]);
        if (lane == 0 && dst[i] && s != 0.f) atomicAdd(dst[i], (double)s);
        s_acc[i] = 0.f;
      }
    };
    int ti = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const Item I = find_item(G, item);
      if (I.g != cur_g) {
        flush_sums();
        cur_g = I.g;
      }
      const GramProblem& P = G.p[I.g];
      const int row = I.m0 + r;
      const bool row_ok = row < P.M;
      // ---- A rows -> tensor memory.  Every MMA of the previous item has completed (this warp waited for the
      //      accumulator of its last tile), so the columns are free.
      if (!kWide) {
        const unsigned char* src = P.a_img + (size_t)(I.m0 >> 7) * P.nkc * kChunkBytes;
        for (int c = h; c < P.nkc; c += 2) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 w = make_uint4(0u, 0u, 0u, 0u);
              if (row_ok) w = *reinterpret_cast<const uint4*>(src + (size_t)c * kChunkBytes + half * 16384 + ptx::sw128_offset(r, j));
              v[4 * j] = w.x; v[4 * j + 1] = w.y; v[4 * j + 2] = w.z; v[4 * j + 3] = w.w;
            }
            ptx::tmem_st32(lane_base + (half ? ALO_COL : AHI_COL) + c * 32, v);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(a_ready);
      }
      float* crow = P.C + (int64_t)row * P.ldc;
      for (int j = I.nt_beg; j < I.nt_end; ++j, ++ti) {
        const NTile T = find_ntile(P, j);
        const int ab = ti & 1;
        ptx::mbar_wait(&acc_full[ab], (uint32_t)((ti >> 1) & 1));
        ptx::tc_fence_after();
        const uint32_t base = lane_base + ACC_COL + ab * 128;
        // both slabs of this warp are fetched before either is processed
        uint32_t v0[32], v1[32];
        ptx::tmem_ld32(base + (2 * h) * 32, v0);
        ptx::tmem_ld32(base + (2 * h + 1) * 32, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&acc_free[ab]);           // the accumulator is in registers: hand it back before the stores
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const uint32_t (&v)[32] = sl == 0 ? v0 : v1;
          const int cc = 2 * h + sl;
          const int vc = T.valid - cc * 32;          // valid columns of this 32-wide slab (warp-uniform)
          if (vc <= 0 || !row_ok) continue;
          if (T.seg > 0) {
            // cosines: |x| <= 1.  exp(x) by ex2.approx; exp(x / 0.1) = exp(x)^10 by four multiplications
            float a01 = 0.f, a1 = 0.f;
            if (vc >= 32) {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const float e1 = __expf(__uint_as_float(v[e]));
                const float e2 = e1 * e1, e4 = e2 * e2;
                a01 += e4 * e4 * e2;
                a1 += e1;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                if (e < vc) {
                  const float e1 = __expf(__uint_as_float(v[e]));
                  const float e2 = e1 * e1, e4 = e2 * e2;
                  a01 += e4 * e4 * e2;
                  a1 += e1;
                }
              }
            }
            if (T.seg == 1) { s_acc[0] += a01; s_acc[2] += a1; }
            else { s_acc[1] += a01; s_acc[3] += a1; }
          }
          (void)0;
        }
        // ---- stores.  Straight from the registers a warp-level 128-bit store touches 32 different rows with 16
        // bytes each, and the LSU / L2 request rate (4096 half-sector requests per tile) becomes the bound of the
        // whole kernel.  A full slab whose first column is 16-byte aligned goes through a [32 x 32] shared-memory
        // transpose at float4 granularity instead: every warp-level store then writes 4 rows x 128 contiguous
        // bytes (8x fewer requests, each a full line).
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const uint32_t (&v)[32] = sl == 0 ? v0 : v1;
          const int cc = 2 * h + sl;
          const int vc = T.valid - cc * 32;
          if (vc <= 0) continue;                      // warp-uniform
          float* slab = P.C + (int64_t)(I.m0 + 32 * q) * P.ldc + T.col0 + cc * 32;     // row 0 of this warp's band
          const bool fast = vc >= 32 && ((reinterpret_cast<uintptr_t>(slab) & 15) == 0) && ((P.ldc & 3) == 0);
          if (fast) {
            float* xt = xpose + (warp - 2) * (32 * kXPitch);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(xt + lane * kXPitch + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int rows_here = min(32, P.M - (I.m0 + 32 * q));
            const int rl = lane >> 3, c4 = 4 * (lane & 7);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + rl;
              const uint4 w = *reinterpret_cast<const uint4*>(xt + rr * kXPitch + c4);
              if (rr < rows_here) *reinterpret_cast<uint4*>(slab + (int64_t)rr * P.ldc + c4) = w;
            }
            __syncwarp();
          } else if (row_ok) {
            store_row32(crow + T.col0 + cc * 32, v, vc);
          }
        }
      }
    }
    flush_sums();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem);
}

}  // namespace

size_t gram_image_bytes(int rows, int D) {
  const size_t tiles = (size_t)(rows + 127) / 128, nkc = (size_t)(D + 31) / 32;
  return tiles * nkc * kChunkBytes;
}

int launch_pack_rows(const float* X, int64_t N, int D, const int32_t* slot, unsigned char* const* img4, float* norms, float* Xh,
                     cudaStream_t st) {
  const int nkc = (D + 31) / 32;
  if (nkc > 16) {
    set_error("gram_ts: D=%d > 512", D);
    return SGA_EINVAL;
  }
  pack_rows_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(X, N, D, nkc, slot, img4[0], img4[1], img4[2], img4[3], norms, Xh);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int launch_build_slots(const int32_t* e1i, const int32_t* e2i, const int32_t* e1j, const int32_t* e2j, int A, int J1, int J2,
                       int32_t* slot, int64_t N, cudaStream_t st) {
  SGA_CUDA(cudaMemsetAsync(slot, 0xFF, 4 * (size_t)N, st));
  const int total = 2 * A + J1 + J2;
  if (total > 0) {
    build_slots_kernel<<<(total + 255) / 256, 256, 0, st>>>(e1i, e2i, e1j, e2j, A, J1, J2, slot);
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}

namespace {
template <bool kWide>
int launch_gram_variant(const GramProblem* problems, int n, cudaStream_t st) {
  // problems of the other variant are skipped: narrow = K <= 128 (nkc <= 4), wide = the rest
  auto mine = [](const GramProblem& P) { return (P.nkc > 4) == kWide && P.M > 0; };
  int i = 0;
  while (i < n) {
    GramGroup G;
    memset(&G, 0, sizeof(G));
    int mt_total = 0, cnt = 0;
    for (int k = i; k < n && cnt < kGramMaxGroup; ++k)
      if (mine(problems[k])) {
        mt_total += (problems[k].M + 127) / 128;
        ++cnt;
      }
    int total = 0;
    for (; i < n && G.n < kGramMaxGroup; ++i) {
      const GramProblem& P = problems[i];
      if (!mine(P)) continue;
      const int NT = (P.seg_rows[0] + 127) / 128 + (P.seg_rows[1] + 127) / 128 + (P.seg_rows[2] + 127) / 128;
      if (NT <= 0) continue;
      // split the column tiles of a row block so that the group has ~3 items per SM, but keep >= 4 tiles per
      // item (narrow: the A operand is re-staged into tensor memory per item)
      int S = (3 * sm_count() + mt_total - 1) / (mt_total > 0 ? mt_total : 1);
      if (S > NT / 4) S = NT / 4;
      if (S < 1) S = 1;
      {
        // the items of a group are equal-sized, so the launch runs in ceil(items / SMs) rounds: among a few larger
        // splits take the one that wastes the least of its last round
        const int sms = sm_count();
        const int smax = NT / 4 < S + 8 ? NT / 4 : S + 8;
        double best = 0.0;
        int bestS = S;
        for (int c = S; c <= smax; ++c) {
          const int items = mt_total * c;
          const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
          if (eff > best + 0.02) { best = eff; bestS = c; }
        }
        S = bestS;
      }
      total += ((P.M + 127) / 128) * S;
      G.p[G.n] = P;
      G.item_end[G.n] = total;
      G.nsplit[G.n] = S;
      ++G.n;
    }
    if (G.n == 0) continue;
    const int grid = total < sm_count() ? total : sm_count();
    gram_ts_kernel<kWide><<<grid, kThreads, Cfg<kWide>::SMEM_BYTES, st>>>(G);
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}
}  // namespace

int launch_gram_ts(const GramProblem* problems, int n, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(gram_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<false>::SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(gram_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<true>::SMEM_BYTES));
    attr_done = true;
  }
  int rc = launch_gram_variant<false>(problems, n, st);
  if (rc != SGA_OK) return rc;
  return launch_gram_variant<true>(problems, n, st);
}

}  // namespace sga
