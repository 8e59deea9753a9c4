// Persistent tcgen05 GEMM with fp32-faithful tf32x3 split operands (umma_tf32.cuh) for the loss
// Grams and their backward (src/aligner/losses.py:5-15 -- the matmuls inside calculate_prob_dist --
// and the autograd of them).
//
//   C[M,N] (op)= sum_k A(m,k) * B(n,k)
// Each operand is read straight from a row-major fp32 matrix, either
//   K-major : X(r,k) = P[row(r)*ld + k]          (row gather through idx, division by div[row])
//   MN-major: X(r,k) = P[row(k)*ld + r]          (the CONTRACTION index is the gathered row)
// so that  F = Zn[e1i] Zn[R]^T  (forward),  dZn[e1i] += dF Zn[R]  and  dZn[R] += dF^T Zn[e1i]
// (backward) are all one launch each without gather / transpose passes: L2-normalisation and the
// e1i/e2i/e1j/e2j gathers happen in the operand loader, the scatter-add in the epilogue.
//
// 128x128 output tiles, K chunks of 32, 3-stage shared-memory ring (4 x 16 KiB per stage), two TMEM
// accumulators so the epilogue of tile i overlaps the main loop of tile i+1.  13 warps: 8 operand
// loaders (half an operand row / half an MN atom each, the global loads of chunk c+1 in flight while
// chunk c is normalised, split and stored), 1 MMA issuer (elect.sync), 4 epilogue warps (one TMEM lane
// quarter each).
#include "gemm_tc.cuh"
#include "umma_tf32.cuh"
#include "epilogue.cuh"
#include <stdlib.h>

namespace sga {

namespace {

constexpr int kStages = 3;
constexpr int kLoaders = 256;
constexpr int kThreads = kLoaders + 32 + 128;
constexpr uint32_t BAR_OFF = kStages * tf32x3::kStageBytes;
constexpr uint32_t XPOSE_OFF = BAR_OFF + 256;                 // 4 epilogue warps x [32][33] fp32 transpose tiles
constexpr uint32_t SMEM_BYTES = XPOSE_OFF + 4 * 32 * 33 * 4 + 1024;

// One loader thread's share of a [128 x 32] operand tile for one K chunk: 16 fp32 values held in registers
// between the global load and the conversion (so that the next chunk's loads overlap this chunk's work).
struct Piece {
  float4 v[4];
  float d;        // divisor of these values (1 when the operand is not normalised)
};

// Where one loader thread's values come from: the (gathered) source row, resolved once per tile for a K-major
// operand (the row does not change along K) and one chunk AHEAD for an MN-major operand (the gathered row is the
// contraction index), so that the index -> divisor -> data dependency chain is never on the critical path.
struct RowSrc {
  const float* p;
  float d;
  bool valid;
};

// K-major operand: thread t -> tile row t/2, features [16*(t&1), +16) of the chunk.
__device__ __forceinline__ RowSrc src_k(const GemmOperand& X, int row0, int nrows, int t) {
  const int r = t >> 1;
  RowSrc s;
  s.valid = r < nrows;
  s.d = 1.f;
  int64_t sr = 0;
  if (s.valid) {
    sr = X.idx ? (int64_t)X.idx[row0 + r] : (int64_t)(row0 + r);
    if (X.div) s.d = X.div[sr];
  }
  s.p = X.p + sr * X.ld + 16 * (t & 1);
  return s;
}

__device__ __forceinline__ void fetch_k(Piece& f, const RowSrc& s, int k0, int K, int t, bool vec_ok) {
  const int kb = k0 + 16 * (t & 1);
  f.d = s.d;
  const float* p = s.p + k0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = kb + 4 * j;
    if (s.valid) {
      if (vec_ok && k + 3 < K) {
        q = *reinterpret_cast<const float4*>(p + 4 * j);
      } else {
        if (k < K) q.x = p[4 * j];
        if (k + 1 < K) q.y = p[4 * j + 1];
        if (k + 2 < K) q.z = p[4 * j + 2];
        if (k + 3 < K) q.w = p[4 * j + 3];
      }
    }
    f.v[j] = q;
  }
}

__device__ __forceinline__ void split4(const float4& q, float d, bool use_div, uint32_t (&h)[4], uint32_t (&l)[4]) {
  const float x[4] = {use_div ? q.x / d : q.x, use_div ? q.y / d : q.y, use_div ? q.z / d : q.z, use_div ? q.w / d : q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = tf32x3::rn_tf32(x[e]);
    l[e] = tf32x3::rn_tf32(x[e] - __uint_as_float(h[e]));
  }
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&w)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}

__device__ __forceinline__ void store_k(uint32_t hi, uint32_t lo, const Piece& f, bool use_div, int t) {
  const int r = t >> 1, half = t & 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t h[4], l[4];
    split4(f.v[j], f.d, use_div, h, l);
    const uint32_t off = ptx::sw128_offset(r, 4 * half + j);
    sts128(hi + off, h);
    sts128(lo + off, l);
  }
}

// MN-major [128 mn x 32 k] tile.  For 32-bit operands the only MN-major shared-memory layout tcgen05
// accepts is SWIZZLE_128B_BASE32B (descriptor layout type 1): atoms of [4 k-rows x 128 B], the
// 32-byte chunk index XORed with (k & 3).  Here: the 32 k-rows of one 32-wide MN atom are contiguous
// (128 B apart; 4-row atoms 512 B apart = SBO), MN atoms 4096 B apart (= LBO).
// Thread t -> k row (t & 31), MN atom (t >> 5) & 3, half (16 mn values) t >> 7.
__device__ __forceinline__ RowSrc src_mn(const GemmOperand& X, int k0, int K, int t) {
  const int krow = k0 + (t & 31);
  RowSrc s;
  s.valid = krow < K;
  s.d = 1.f;
  int64_t sr = 0;
  if (s.valid) {
    sr = X.idx ? (int64_t)X.idx[krow] : (int64_t)krow;
    if (X.div) s.d = X.div[sr];
  }
  s.p = X.p + sr * X.ld;
  return s;
}

__device__ __forceinline__ void fetch_mn(Piece& f, const RowSrc& s, int mn0, int MN, int t, bool vec_ok) {
  const int qd = (t >> 5) & 3, half = t >> 7;
  f.d = s.d;
  const int c0 = mn0 + 32 * qd + 16 * half;
  const float* p = s.p + c0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = c0 + 4 * j;
    if (s.valid) {
      if (vec_ok && c + 3 < MN) {
        q = *reinterpret_cast<const float4*>(p + 4 * j);
      } else {
        if (c < MN) q.x = p[4 * j];
        if (c + 1 < MN) q.y = p[4 * j + 1];
        if (c + 2 < MN) q.z = p[4 * j + 2];
        if (c + 3 < MN) q.w = p[4 * j + 3];
      }
    }
    f.v[j] = q;
  }
}

__device__ __forceinline__ void store_mn(uint32_t hi, uint32_t lo, const Piece& f, bool use_div, int t) {
  const int kk = t & 31, qd = (t >> 5) & 3, half = t >> 7;
  const uint32_t base = (uint32_t)qd * 4096u + (uint32_t)kk * 128u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t h[4], l[4];
    split4(f.v[j], f.d, use_div, h, l);
    const int jj = 4 * half + j;      // 16-byte chunk of the 128-byte k row
    const uint32_t off = base + (uint32_t)((((jj >> 1) ^ (kk & 3)) << 5) + ((jj & 1) << 4));
    sts128(hi + off, h);
    sts128(lo + off, l);
  }
}

__device__ __forceinline__ uint64_t desc_mn_major(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(4096u >> 4) << 16;      // LBO: next 32-element atom along MN
  d |= (uint64_t)(512u >> 4) << 32;       // SBO: next group of 4 k rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
  return d;
}

template <bool A_MN, bool B_MN>
__device__ __forceinline__ void issue_stage_any(uint32_t d_tmem, uint32_t stage_addr, uint32_t idesc, bool first) {
  using namespace tf32x3;
  const uint64_t dAhi = A_MN ? desc_mn_major(stage_addr) : ptx::smem_desc_sw128(stage_addr);
  const uint64_t dAlo = A_MN ? desc_mn_major(stage_addr + kTileBytes) : ptx::smem_desc_sw128(stage_addr + kTileBytes);
  const uint64_t dBhi = B_MN ? desc_mn_major(stage_addr + 2 * kTileBytes) : ptx::smem_desc_sw128(stage_addr + 2 * kTileBytes);
  const uint64_t dBlo = B_MN ? desc_mn_major(stage_addr + 3 * kTileBytes) : ptx::smem_desc_sw128(stage_addr + 3 * kTileBytes);
  constexpr uint64_t aStep = A_MN ? (1024 >> 4) : 2;   // one k-step = 8 k: 8 rows of 128 B / +32 bytes
  constexpr uint64_t bStep = B_MN ? (1024 >> 4) : 2;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t a = (pass == 1) ? dAlo : dAhi;
    const uint64_t b = (pass == 2) ? dBlo : dBhi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ptx::umma_tf32(d_tmem, a + aStep * ks, b + bStep * ks, idesc, (first && pass == 0 && ks == 0) ? 0u : 1u);
  }
}

// Work item `work` of the group -> (problem, item inside the problem)
struct WorkItem {
  int g, local;
};
__device__ __forceinline__ WorkItem find_work(const GemmGroup& G, int work) {
  int g = 0, beg = 0;
  while (g + 1 < G.n && work >= G.work_end[g]) {
    beg = G.work_end[g];
    ++g;
  }
  return {g, work - beg};
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ GemmGroup G) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + BAR_OFF);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;    // [2]
  uint64_t* acc_free = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
  float* xpose = reinterpret_cast<float*>(sm + XPOSE_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], kLoaders / 32);      // one elected arrive per loader warp
      ptx::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_free[i], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nwork = G.work_end[G.n - 1];

  if (warp < 8) {
    // ------------------------------- operand loaders
    int it = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
      const WorkItem wi = find_work(G, work);
      const GemmParams& P = G.p[wi.g];
      const bool a_vec = (P.A.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.A.p) & 15) == 0);
      const bool b_vec = (P.B.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.B.p) & 15) == 0);
      const bool a_div = P.A.div != nullptr, b_div = P.B.div != nullptr;
      const int ntn = (P.N + 127) / 128;
      const int ksplit = P.ksplit > 1 ? P.ksplit : 1;
      const int nkc_all = (P.K + 31) / 32;
      const int kc_per = (nkc_all + ksplit - 1) / ksplit;
      const int tile = wi.local / ksplit, ksl = wi.local % ksplit;
      const int m0 = (tile / ntn) * 128, n0 = (tile % ntn) * 128;
      const int kc_beg = ksl * kc_per, kc_end = min(nkc_all, kc_beg + kc_per);
      if (kc_beg >= kc_end) continue;
      RowSrc sa = A_MN ? src_mn(P.A, kc_beg * 32, P.K, tid) : src_k(P.A, m0, P.M - m0, tid);
      RowSrc sb = B_MN ? src_mn(P.B, kc_beg * 32, P.K, tid) : src_k(P.B, n0, P.N - n0, tid);
      auto fetch = [&](Piece& fa, Piece& fb, int kc) {
        if (A_MN) fetch_mn(fa, sa, m0, P.M, tid, a_vec);
        else fetch_k(fa, sa, kc * 32, P.K, tid, a_vec);
        if (B_MN) fetch_mn(fb, sb, n0, P.N, tid, b_vec);
        else fetch_k(fb, sb, kc * 32, P.K, tid, b_vec);
      };
      auto advance_src = [&](int kc) {       // MN-major: resolve the gathered rows of chunk kc (one chunk ahead of its data)
        if (kc < kc_end) {
          if (A_MN) sa = src_mn(P.A, kc * 32, P.K, tid);
          if (B_MN) sb = src_mn(P.B, kc * 32, P.K, tid);
        }
      };
      Piece fa, fb;
      const int dbg = P.dbg;
      if (dbg & 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) fa.v[j] = fb.v[j] = make_float4(0.1f * tid, 0.2f, 0.3f, 0.4f);
        fa.d = fb.d = 1.f;
      }
      if (!(dbg & 1)) fetch(fa, fb, kc_beg);
      advance_src(kc_beg + 1);
      for (int kc = kc_beg; kc < kc_end; ++kc, ++it) {
        const int s = it % kStages;
        const Piece ca = fa, cb = fb;
        if (kc + 1 < kc_end) {                // in flight during the conversion below
          if (!(dbg & 1)) fetch(fa, fb, kc + 1);
          advance_src(kc + 2);
        }
        if (it >= kStages) ptx::mbar_wait(&empty[s], (uint32_t)(((it / kStages) - 1) & 1));
        const uint32_t st = sm_base + s * tf32x3::kStageBytes;
        if (!(dbg & 2)) {
          if (A_MN) store_mn(st, st + tf32x3::kTileBytes, ca, a_div, tid);
          else store_k(st, st + tf32x3::kTileBytes, ca, a_div, tid);
          if (B_MN) store_mn(st + 2 * tf32x3::kTileBytes, st + 3 * tf32x3::kTileBytes, cb, b_div, tid);
          else store_k(st + 2 * tf32x3::kTileBytes, st + 3 * tf32x3::kTileBytes, cb, b_div, tid);
        } else if (ca.v[0].x == 1234.5f && cb.v[3].w == 5432.1f) {
          sts128(st, reinterpret_cast<const uint32_t(&)[4]>(ca.v[1]));      // keep the loads alive
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_relaxed(&full[s]);
      }
    }
  } else if (warp == 8) {
    // ------------------------------- MMA issuer
    const uint32_t idesc = ptx::make_idesc(2, 128, 128) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    int it = 0, ti = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++ti) {
      const WorkItem wi = find_work(G, work);
      const GemmParams& P = G.p[wi.g];
      const int ksplit = P.ksplit > 1 ? P.ksplit : 1;
      const int nkc_all = (P.K + 31) / 32;
      const int kc_per = (nkc_all + ksplit - 1) / ksplit;
      const int ksl = wi.local % ksplit;
      const int kc_beg = ksl * kc_per, kc_end = min(nkc_all, kc_beg + kc_per);
      const int ab = ti & 1;
      if (ti >= 2) {
        ptx::mbar_wait(&acc_free[ab], (uint32_t)(((ti >> 1) - 1) & 1));
        ptx::tc_fence_after();
      }
      for (int kc = kc_beg; kc < kc_end; ++kc, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (!(P.dbg & 8)) issue_stage_any<A_MN, B_MN>(tmem_u + ab * 128, sm_base + s * tf32x3::kStageBytes, idesc, kc == kc_beg);
          ptx::umma_commit(&empty[s]);
          if (kc == kc_end - 1) ptx::umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
      if (kc_end <= kc_beg && ptx::elect_one()) ptx::umma_commit(&acc_full[ab]);   // empty slice: nothing to add
      __syncwarp();
    }
  } else {
    // ------------------------------- epilogue
    const int q = warp & 3;
    const int r = 32 * q + lane;
    float s_acc[4] = {0.f, 0.f, 0.f, 0.f};   // {s01_lo, s01_hi, s1_lo, s1_hi} of problem `cur_g`
    int cur_g = -1;
    auto flush_sums = [&]() {
      if (cur_g < 0) return;
      const GemmParams& Q = G.p[cur_g];
      if (Q.mode != 1) return;
      double* dst[4] = {Q.s01_lo, Q.s01_hi, Q.s1_lo, Q.s1_hi};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = warp_sum(s_acc[i]);
        if (lane == 0 && dst[i] && s != 0.f) atomicAdd(dst[i], (double)s);
        s_acc[i] = 0.f;
      }
    };
    int ti = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++ti) {
      const WorkItem wi = find_work(G, work);
      if (wi.g != cur_g) {
        flush_sums();
        cur_g = wi.g;
      }
      const GemmParams& P = G.p[wi.g];
      const int ntn = (P.N + 127) / 128;
      const int ksplit = P.ksplit > 1 ? P.ksplit : 1;
      const int nkc_all = (P.K + 31) / 32;
      const int kc_per = (nkc_all + ksplit - 1) / ksplit;
      const int tile = wi.local / ksplit, ksl = wi.local % ksplit;
      const int m0 = (tile / ntn) * 128, n0 = (tile % ntn) * 128;
      const int ab = ti & 1;
      ptx::mbar_wait(&acc_full[ab], (uint32_t)((ti >> 1) & 1));
      ptx::tc_fence_after();
      const int row = m0 + r;
      const bool row_ok = row < P.M && (ksl * kc_per < nkc_all);
      float* crow = nullptr;
      if (row_ok) crow = P.C + ((P.mode == 2 && P.c_idx) ? (int64_t)P.c_idx[row] : (int64_t)row) * P.ldc;
      const uint32_t base = tmem + ((uint32_t)(32 * q) << 16) + ab * 128;
      const bool slice_ok = ksl * kc_per < nkc_all;
      const int rows_here = slice_ok ? min(32, P.M - (m0 + 32 * q)) : 0;      // valid rows of this warp's 32-row band
      float* xt = xpose + q * (32 * 33);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld32(base + cc * 32, v);
        ptx::tmem_ld_wait();
        const int c0 = n0 + cc * 32;
        if (c0 >= P.N || rows_here <= 0 || (P.dbg & 4)) continue;       // warp-uniform
        if (P.mode == 1 && row_ok && c0 + 31 >= P.es_c0) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int c = c0 + e;
            if (c >= P.es_c0 && c < P.N) {
              // cosines: |x| <= 1.  exp(x) by ex2.approx (~2 ulp); exp(x / 0.1) = exp(x)^10 by four multiplications
              // (~20 ulp), half the SFU work of a second exponential: exact enough for sums of ~1e6 positive terms
              const float e1 = __expf(__uint_as_float(v[e]));
              const float e2 = e1 * e1, e4 = e2 * e2;
              const float e01 = e4 * e4 * e2;
              if (c < P.es_split) { s_acc[0] += e01; s_acc[2] += e1; }
              else { s_acc[1] += e01; s_acc[3] += e1; }
            }
          }
        }
        if (P.mode == 2) {
          // scatter-add through a shared-memory transpose: lanes own consecutive COLUMNS of one output row, so each
          // warp-level reduction is one contiguous 128-byte request instead of 32 scattered ones
#pragma unroll
          for (int e = 0; e < 32; ++e) xt[lane * 33 + e] = __uint_as_float(v[e]);
          __syncwarp();
          const bool col_ok = c0 + lane < P.N;
          const unsigned long long cptr = reinterpret_cast<unsigned long long>(crow);   // 0: this lane's row is out of range
#pragma unroll 4
          for (int rr = 0; rr < rows_here; ++rr) {
            float* dst = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, cptr, rr));
            if (col_ok) atomicAdd(dst + c0 + lane, xt[rr * 33 + lane]);
          }
          __syncwarp();
        } else if (row_ok) {
          store_row32(crow + c0, v, P.N - c0);        // 128-bit stores straight from the registers (epilogue.cuh)
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_free[ab]);
    }
    flush_sums();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<256>(tmem);
}

}  // namespace

namespace {
// profiling switches (tools/gemm_probe.py): read once per process
inline int gemm_dbg_flags() {
  static int flags = -1;
  if (flags < 0) {
    const char* e = getenv("SGA_GEMM_DBG");
    flags = e ? atoi(e) : 0;
  }
  return flags;
}
inline int work_items(const GemmParams& P) {
  return ((P.M + 127) / 128) * ((P.N + 127) / 128) * (P.ksplit > 1 ? P.ksplit : 1);
}
}  // namespace

int launch_gemm_tc_group(const GemmParams* problems, int n, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SGA_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_done = true;
  }
  int i = 0;
  while (i < n) {
    GemmGroup G;
    memset(&G, 0, sizeof(G));
    int total = 0;
    const int a_mn = problems[i].A.mn_major, b_mn = problems[i].B.mn_major;
    for (; i < n && G.n < kGemmMaxGroup; ++i) {
      const GemmParams& P = problems[i];
      if (P.M <= 0 || P.N <= 0 || P.K <= 0) continue;
      if (P.A.mn_major != a_mn || P.B.mn_major != b_mn) {
        set_error("gemm_tc: the problems of one group must share the operand layouts");
        return SGA_EINVAL;
      }
      if (P.ksplit > 1 && P.mode != 2) {
        set_error("gemm_tc: split-K needs the scatter-add epilogue");
        return SGA_EINVAL;
      }
      total += work_items(P);
      G.p[G.n] = P;
      G.p[G.n].dbg = gemm_dbg_flags();
      G.work_end[G.n] = total;
      ++G.n;
    }
    if (G.n == 0) continue;
    const int grid = total < sm_count() ? total : sm_count();
    if (!a_mn && !b_mn) gemm_tf32x3_kernel<false, false><<<grid, kThreads, SMEM_BYTES, st>>>(G);
    else if (!a_mn && b_mn) gemm_tf32x3_kernel<false, true><<<grid, kThreads, SMEM_BYTES, st>>>(G);
    else if (a_mn && b_mn) gemm_tf32x3_kernel<true, true><<<grid, kThreads, SMEM_BYTES, st>>>(G);
    else {
      set_error("gemm_tc: (A MN-major, B K-major) is not instantiated");
      return SGA_EINVAL;
    }
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}

int launch_gemm_tc(const GemmParams& P, cudaStream_t st) { return launch_gemm_tc_group(&P, 1, st); }

// Dense product without gathers: C[M,N] = (or +=) sum_k A(m,k) B(n,k).  accumulate: C already holds the value to
// add to and the K range is cut into slices that add their partial products atomically (weight gradients: small
// output, contraction over all nodes of the batch).
int launch_gemm_tc_dense(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, int M, int N, int K,
                         float* C, int64_t ldc, int accumulate, cudaStream_t st) {
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.A = {A, lda, nullptr, nullptr, a_mn};
  P.B = {B, ldb, nullptr, nullptr, b_mn};
  P.M = M; P.N = N; P.K = K;
  P.C = C; P.ldc = ldc;
  P.mode = accumulate ? 2 : 0;
  P.ksplit = 1;
  if (accumulate) {
    const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
    const int chunks = (K + 31) / 32;
    int want = (2 * sm_count() + tiles - 1) / tiles;
    const int cap = chunks / 4;
    if (want > cap) want = cap;
    P.ksplit = want < 1 ? 1 : want;
  }
  return launch_gemm_tc_group(&P, 1, st);
}

}  // namespace sga

// Test / generic entry: C (mode 0: =, mode 2: scatter-add through c_idx) A B^T with optional gathers.
extern "C" int sga_gemm_tf32x3(const float* A, int64_t lda, int a_mn_major, const int32_t* a_idx, const float* a_div,
                               const float* B, int64_t ldb, int b_mn_major, const int32_t* b_idx, const float* b_div,
                               int M, int N, int K, float* C, int64_t ldc, const int32_t* c_idx, int ksplit, void* stream) {
  sga::GemmParams P;
  memset(&P, 0, sizeof(P));
  P.A = {A, lda, a_idx, a_div, a_mn_major};
  P.B = {B, ldb, b_idx, b_div, b_mn_major};
  P.M = M; P.N = N; P.K = K;
  P.C = C; P.ldc = ldc;
  P.mode = c_idx ? 2 : 0;
  P.c_idx = c_idx;
  P.ksplit = c_idx ? ksplit : 1;
  return sga::launch_gemm_tc(P, (cudaStream_t)stream);
}

// Grouped weight-gradient products (NaivePCT backward): C_i [M_i,N_i] += A_i^T B_i, A_i [K,M_i], B_i [K,N_i] row-major --
// both operands MN-major, the contraction runs over all K rows (points of the batch), cut into slices that add their
// partial products atomically.  All problems of the call share one persistent launch per 16.
extern "C" int sga_wgrad_group(const float* const* A, const int64_t* lda, const int* M, const float* const* B, const int64_t* ldb,
                               const int* Nn, float* const* C, const int64_t* ldc, int n, int64_t K, void* stream) {
  if (n <= 0 || K <= 0) return SGA_OK;
  SGA_REQUIRE(A && lda && M && B && ldb && Nn && C && ldc && K < (int64_t)1 << 31, "sga_wgrad_group: bad arguments");
  SGA_REQUIRE(n <= 64, "sga_wgrad_group: at most 64 problems per call (%d)", n);
  sga::GemmParams ps[64];
  int tiles = 0;
  for (int i = 0; i < n; ++i) tiles += ((M[i] + 127) / 128) * ((Nn[i] + 127) / 128);
  const int chunks = (int)((K + 31) / 32);
  int want = (3 * sga::sm_count() + tiles - 1) / tiles;
  const int cap = chunks / 4;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  for (int i = 0; i < n; ++i) {
    SGA_REQUIRE(A[i] && B[i] && C[i], "sga_wgrad_group: null operand %d", i);
    sga::GemmParams& P = ps[i];
    memset(&P, 0, sizeof(P));
    P.A = {A[i], lda[i], nullptr, nullptr, 1};
    P.B = {B[i], ldb[i], nullptr, nullptr, 1};
    P.M = M[i]; P.N = Nn[i]; P.K = (int)K;
    P.C = C[i]; P.ldc = ldc[i];
    P.mode = 2;
    P.ksplit = want;
  }
  return sga::launch_gemm_tc_group(ps, n, (cudaStream_t)stream);
}
