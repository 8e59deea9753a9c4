// Matching head on the tensor cores: per sub-scan pair, sim = 1 - E E^T (rows L2-normalised) as a
// tcgen05 tf32x3 Gram fused with the per-row top-k of the ranking -- the similarity matrix only
// reaches HBM when the caller asks for it.
// Reference: src/inference/sgaligner/inference_align_reg.py:125-128 (normalise, mm, argsort) and the
// first-k reads of utils/alignment.py.
//
// One CTA per (pair, 128-row block).  4 worker warps gather + normalise + hi/lo-split the embedding
// rows into 128B-swizzled operand tiles (3-stage ring) and later run the epilogue; one warp issues
// the MMAs.  Rows sit on TMEM lanes, so every thread owns one row of the similarity matrix and keeps
// its k best (sim, column) pairs in registers while the column tiles stream by.
#include "common.cuh"
#include "umma_tf32.cuh"

namespace sga {
namespace {

constexpr int kStages = 3;
constexpr int kWorkers = 128;
constexpr int kThreads = kWorkers + 32;
constexpr int KT = 8;   // top-k capacity per row
constexpr uint32_t BAR_OFF = kStages * tf32x3::kStageBytes;
constexpr uint32_t SMEM_BYTES = BAR_OFF + 128 + 1024;

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
  if (warp == 4) ptx::tmem_alloc<128>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int nkc = (D + tf32x3::kTileK - 1) / tf32x3::kTileK;
  const int ntile = (n + 127) / 128;
  const bool vec_ok = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(emb) & 15) == 0);

  if (warp == 4) {
    // ------------------------------- MMA issuer
    const uint32_t idesc = ptx::make_idesc(2, 128, 128);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    int it = 0;
    for (int nt = 0; nt < ntile; ++nt) {
      if (nt > 0) {
        ptx::mbar_wait(acc_free, (uint32_t)((nt - 1) & 1));
        ptx::tc_fence_after();
      }
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          tf32x3::issue_stage(tmem_u, sm_base + s * tf32x3::kStageBytes, idesc, kc == 0);
          ptx::umma_commit(&empty[s]);
          if (kc == nkc - 1) ptx::umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- workers: operand staging, then epilogue
    const int t = tid;                 // row of the tile
    const int row = m0 + t;            // pair-local row
    const bool row_ok = row < n;
    float best_s[KT];
    int best_c[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) { best_s[j] = INFINITY; best_c[j] = -1; }
    float* srow = (sim_out && row_ok) ? sim_out + sim_off[b] + (int64_t)row * n : nullptr;
    int it = 0;
    for (int nt = 0; nt < ntile; ++nt) {
      const int n0 = nt * 128;
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % kStages;
        if (it >= kStages) ptx::mbar_wait(&empty[s], (uint32_t)(((it / kStages) - 1) & 1));
        unsigned char* st = sm + s * tf32x3::kStageBytes;
        tf32x3::load_rows(st, st + tf32x3::kTileBytes, emb, D, nullptr, norms, o0 + m0, n - m0, kc * 32, D, t, vec_ok);
        tf32x3::load_rows(st + 2 * tf32x3::kTileBytes, st + 3 * tf32x3::kTileBytes, emb, D, nullptr, norms, o0 + n0, n - n0, kc * 32, D, t,
                          vec_ok);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&full[s]);
      }
      ptx::mbar_wait(acc_full, (uint32_t)(nt & 1));
      ptx::tc_fence_after();
      const uint32_t base = tmem + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld32(base + cc * 32, v);
        ptx::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int c = n0 + cc * 32 + e;
            if (c < n) {
              const float sv = 1.f - __uint_as_float(v[e]);
              if (srow) srow[c] = sv;
              if (sv < best_s[KT - 1]) {
                float cs = sv;
                int ci = c;
#pragma unroll
                for (int j = 0; j < KT; ++j) {
                  const bool sw = cs < best_s[j];
                  const float ts = sw ? best_s[j] : cs;
                  const int tc = sw ? best_c[j] : ci;
                  best_s[j] = sw ? cs : best_s[j];
                  best_c[j] = sw ? ci : best_c[j];
                  cs = ts;
                  ci = tc;
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(acc_free);
    }
    if (row_ok && topk_idx) {
      const int64_t o = (int64_t)(o0 + row) * K;
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (j < K) {
          topk_idx[o + j] = best_c[j];
          if (topk_dist) topk_dist[o + j] = best_s[j];
        }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<128>(tmem);
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
