// Matching head on the tensor cores: per sub-scan pair, sim = 1 - E E^T (rows L2-normalised) as a
// tcgen05 tf32x3 Gram fused with the per-row top-k of the ranking -- the similarity matrix only
// reaches HBM when the caller asks for it.
// Reference: src/inference/sgaligner/inference_align_reg.py:125-128 (normalise, mm, argsort) and the
// first-k reads of utils/alignment.py.
//
// One CTA per (pair, 128-row block): 8 worker warps + 1 MMA-issuing warp.
//   staging : worker t owns half a row (16 of the 32 features of a K chunk): gather + normalise +
//             hi/lo split into 128B-swizzled operand tiles, 2-stage ring, the global loads of chunk
//             c+1 in flight while chunk c is converted.  When the row block and the column block
//             coincide (every pair with <= 128 nodes) the B operand aliases the A tiles.
//   epilogue: rows sit on TMEM lanes; warp w reads lane quarter w%4 and the column half w/4 of the
//             tile, keeps its k best (sim, column) pairs in registers (lexicographic order = the
//             stable argsort), the two halves of a row are merged through shared memory at the end;
//             the similarity tile goes out through a padded shared-memory stage as coalesced rows.
#include "common.cuh"
#include "umma_tf32.cuh"

namespace sga {
namespace {

constexpr int kStages = 2;
constexpr int kWorkers = 256;
constexpr int kThreads = kWorkers + 32;
constexpr int KT = 8;   // top-k capacity per row
constexpr int SIM_LD = 129;
constexpr uint32_t SIM_OFF = kStages * tf32x3::kStageBytes;               // float[128][SIM_LD]
constexpr uint32_t MRG_OFF = SIM_OFF + 128 * SIM_LD * 4;                   // {float s; int c}[128][KT]
constexpr uint32_t BAR_OFF = ((MRG_OFF + 128 * KT * 8 + 15) / 16) * 16;
constexpr uint32_t SMEM_BYTES = BAR_OFF + 128 + 1024;

// (sim, column) lexicographic "a before b": the deterministic version of torch.argsort's order
__device__ __forceinline__ bool before(float sa, int ca, float sb, int cb) { return sa < sb || (sa == sb && ca < cb); }

__device__ __forceinline__ void topk_insert(float (&best_s)[KT], int (&best_c)[KT], float sv, int c) {
  if (before(sv, c, best_s[KT - 1], best_c[KT - 1])) {
    float cs = sv;
    int ci = c;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const bool sw = before(cs, ci, best_s[j], best_c[j]);
      const float ts = sw ? best_s[j] : cs;
      const int tc = sw ? best_c[j] : ci;
      best_s[j] = sw ? cs : best_s[j];
      best_c[j] = sw ? ci : best_c[j];
      cs = ts;
      ci = tc;
    }
  }
}

// half a row (16 features) of an operand tile, held in registers between the global load and the
// conversion so that the next chunk's loads overlap this chunk's conversion
struct HalfRow {
  float4 v[4];
};

__device__ __forceinline__ void fetch_half(HalfRow& f, const float* __restrict__ emb, int D, int64_t grow, bool valid, int k0,
                                           int half, bool vec_ok) {
  const int kb = k0 + 16 * half;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = kb + 4 * j;
    if (valid) {
      const float* p = emb + grow * D + k;
      if (vec_ok && k + 3 < D) {
        q = *reinterpret_cast<const float4*>(p);
      } else {
        if (k < D) q.x = p[0];
        if (k + 1 < D) q.y = p[1];
        if (k + 2 < D) q.z = p[2];
        if (k + 3 < D) q.w = p[3];
      }
    }
    f.v[j] = q;
  }
}

__device__ __forceinline__ void store_half(uint32_t hi_addr, uint32_t lo_addr, const HalfRow& f, int row, int half, float inv_is_div) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x[4] = {f.v[j].x / inv_is_div, f.v[j].y / inv_is_div, f.v[j].z / inv_is_div, f.v[j].w / inv_is_div};
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = tf32x3::rn_tf32(x[e]);
      l[e] = tf32x3::rn_tf32(x[e] - __uint_as_float(h[e]));
    }
    const uint32_t off = ptx::sw128_offset(row, 4 * half + j);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
  }
}

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
match_topk_tc_kernel(const float* __restrict__ emb, const float* __restrict__ norms, int D,
                     const int32_t* __restrict__ pair_off, const int64_t* __restrict__ sim_off, int K,
                     int32_t* __restrict__ topk_idx, float* __restrict__ topk_dist, float* __restrict__ sim_out) {
  const int b = blockIdx.y;
  const int o0 = pair_off[b], n = pair_off[b + 1] - o0;
  const int m0 = blockIdx.x * 128;
  if (m0 >= n) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  float* sim_st = reinterpret_cast<float*>(sm + SIM_OFF);
  float* mrg_s = reinterpret_cast<float*>(sm + MRG_OFF);
  int* mrg_c = reinterpret_cast<int*>(sm + MRG_OFF + 128 * KT * 4);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + BAR_OFF);   // [kStages]
  uint64_t* empty = full + kStages;                             // [kStages]
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_free = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], kWorkers);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_free, kWorkers);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc<128>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int nkc = (D + tf32x3::kTileK - 1) / tf32x3::kTileK;
  const int ntile = (n + 127) / 128;
  const bool vec_ok = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(emb) & 15) == 0);

  if (warp == 8) {
    // ------------------------------- MMA issuer
    const uint32_t idesc = ptx::make_idesc(2, 128, 128);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    int it = 0;
    for (int nt = 0; nt < ntile; ++nt) {
      const bool alias = (nt * 128 == m0);
      if (nt > 0) {
        ptx::mbar_wait(acc_free, (uint32_t)((nt - 1) & 1));
        ptx::tc_fence_after();
      }
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t st = sm_base + s * tf32x3::kStageBytes;
          tf32x3::issue_stage_ab(tmem_u, st, alias ? st : st + 2 * tf32x3::kTileBytes, idesc, kc == 0);
          ptx::umma_commit(&empty[s]);
          if (kc == nkc - 1) ptx::umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- workers: operand staging, then epilogue
    const int lrow = tid >> 1, half = tid & 1;     // staging: half a row of the tile
    const int q = warp & 3, ch = warp >> 2;        // epilogue: TMEM lane quarter, column half
    const int erow_l = 32 * q + lane;              // epilogue: tile row = TMEM lane
    const int erow = m0 + erow_l;                  // pair-local row
    const bool erow_ok = erow < n;
    float best_s[KT];
    int best_c[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) { best_s[j] = INFINITY; best_c[j] = 0x7fffffff; }

    const bool a_valid = (m0 + lrow) < n;
    const float a_norm = a_valid ? norms[o0 + m0 + lrow] : 1.f;
    int it = 0;
    for (int nt = 0; nt < ntile; ++nt) {
      const int n0 = nt * 128;
      const bool alias = (n0 == m0);
      const bool b_valid = (n0 + lrow) < n;
      const float b_norm = (!alias && b_valid) ? norms[o0 + n0 + lrow] : 1.f;
      HalfRow fa, fb;
      fetch_half(fa, emb, D, o0 + m0 + lrow, a_valid, 0, half, vec_ok);
      if (!alias) fetch_half(fb, emb, D, o0 + n0 + lrow, b_valid, 0, half, vec_ok);
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % kStages;
        HalfRow ca = fa, cb = fb;
        if (kc + 1 < nkc) {          // next chunk's loads in flight during this chunk's conversion
          fetch_half(fa, emb, D, o0 + m0 + lrow, a_valid, (kc + 1) * 32, half, vec_ok);
          if (!alias) fetch_half(fb, emb, D, o0 + n0 + lrow, b_valid, (kc + 1) * 32, half, vec_ok);
        }
        if (it >= kStages) ptx::mbar_wait(&empty[s], (uint32_t)(((it / kStages) - 1) & 1));
        const uint32_t st = sm_base + s * tf32x3::kStageBytes;
        store_half(st, st + tf32x3::kTileBytes, ca, lrow, half, a_norm);
        if (!alias) store_half(st + 2 * tf32x3::kTileBytes, st + 3 * tf32x3::kTileBytes, cb, lrow, half, b_norm);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&full[s]);
      }
      ptx::mbar_wait(acc_full, (uint32_t)(nt & 1));
      ptx::tc_fence_after();
      const uint32_t base = tmem + ((uint32_t)(32 * q) << 16) + 64 * ch;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld32(base + cc * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int cl = 64 * ch + cc * 32 + e;
          const int c = n0 + cl;
          const float sv = 1.f - __uint_as_float(v[e]);
          if (sim_out) sim_st[erow_l * SIM_LD + cl] = sv;
          if (erow_ok && c < n) topk_insert(best_s, best_c, sv, c);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(acc_free);
      if (sim_out) {
        worker_barrier();            // the whole tile is staged
        const int ncol = min(128, n - n0);
        for (int r = warp; r < 128 && m0 + r < n; r += 8) {
          float* dst = sim_out + sim_off[b] + (int64_t)(m0 + r) * n + n0;
          for (int c = lane; c < ncol; c += 32) dst[c] = sim_st[r * SIM_LD + c];
        }
        worker_barrier();            // before the next tile overwrites the stage
      }
    }
    // merge the two column halves of every row
    if (ch == 1) {
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        mrg_s[erow_l * KT + j] = best_s[j];
        mrg_c[erow_l * KT + j] = best_c[j];
      }
    }
    worker_barrier();
    if (ch == 0 && erow_ok && topk_idx) {
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const int c = mrg_c[erow_l * KT + j];
        if (c != 0x7fffffff) topk_insert(best_s, best_c, mrg_s[erow_l * KT + j], c);
      }
      const int64_t o = (int64_t)(o0 + erow) * K;
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (j < K) {
          topk_idx[o + j] = best_c[j] == 0x7fffffff ? -1 : best_c[j];
          if (topk_dist) topk_dist[o + j] = best_s[j];
        }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc<128>(tmem);
}

__global__ void __launch_bounds__(256)
row_norm_kernel(const float* __restrict__ emb, int64_t N, int D, float* __restrict__ norms) {
  int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    float v = emb[row * D + k];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) norms[row] = sqrtf(s);
}

}  // namespace
}  // namespace sga

extern "C" int sga_match_topk_tc(const float* emb, int64_t N, int D, const int32_t* pair_off, const int64_t* sim_off,
                                 int B, int max_pair_nodes, int K, float* norms, int32_t* topk_idx, float* topk_dist,
                                 float* sim_out, void* stream) {
  if (N <= 0 || B <= 0) return SGA_OK;
  SGA_REQUIRE(K >= 0 && K <= sga::KT, "sga_match_topk_tc: K=%d must be <= %d (use sga_match_rank beyond that)", K, sga::KT);
  SGA_REQUIRE(D >= 1 && max_pair_nodes >= 1, "sga_match_topk_tc: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_done = false;
  if (!attr_done) {
    SGA_CUDA(cudaFuncSetAttribute(sga::match_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sga::SMEM_BYTES));
    attr_done = true;
  }
  sga::row_norm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(emb, N, D, norms);
  SGA_LAUNCH_CHECK();
  dim3 grid((max_pair_nodes + 127) / 128, B);
  sga::match_topk_tc_kernel<<<grid, sga::kThreads, sga::SMEM_BYTES, st>>>(emb, norms, D, pair_off, sim_off, K, K > 0 ? topk_idx : nullptr,
                                                                       topk_dist, sim_out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
