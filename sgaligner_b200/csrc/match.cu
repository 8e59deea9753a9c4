// Matching head for all sub-scan pairs of a batch at once (fp32 FMA path).
// Reference: src/inference/sgaligner/inference_align_reg.py:125-128
//   emb = emb / emb.norm(dim=1)[:, None]; sim = 1 - emb @ emb.T; rank_list = argsort(sim, dim=1)
// and utils/alignment.py:3-25 (position of the ground-truth match in a row with the node itself
// removed).  The per-pair Python loop and the 7 host round-trips per pair of the reference become
// three launches over the whole batch.
#include "common.cuh"

namespace sga {
namespace {

constexpr int NT = 256;
constexpr int TS = 64;   // sim tile
constexpr int TK = 16;

__global__ void __launch_bounds__(NT)
row_norm_kernel(const float* __restrict__ emb, int64_t N, int D, float* __restrict__ norms) {
  int64_t row = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
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

// sim[b][i][j] = 1 - <e_i, e_j>, e = emb / ||emb||  (rows of pair b only)
__global__ void __launch_bounds__(NT)
match_sim_kernel(const float* __restrict__ emb, const float* __restrict__ norms, int D,
                 const int32_t* __restrict__ pair_off, const int64_t* __restrict__ sim_off,
                 float* __restrict__ sim) {
  __shared__ float As[TK][TS + 1];
  __shared__ float Bs[TK][TS + 1];
  const int b = blockIdx.y;
  const int o0 = pair_off[b], n = pair_off[b + 1] - o0;
  const int tiles = (n + TS - 1) / TS;
  if ((int)blockIdx.x >= tiles * tiles) return;
  const int ti = blockIdx.x / tiles, tj = blockIdx.x % tiles;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += TK) {
    __syncthreads();
    for (int i = tid; i < TS * TK; i += NT) {
      int r = i / TK, k = i % TK;
      int ra = ti * TS + r, rb = tj * TS + r;
      As[k][r] = (ra < n && k0 + k < D) ? emb[(int64_t)(o0 + ra) * D + k0 + k] / norms[o0 + ra] : 0.f;
      Bs[k][r] = (rb < n && k0 + k < D) ? emb[(int64_t)(o0 + rb) * D + k0 + k] / norms[o0 + rb] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
  float* out = sim + sim_off[b];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = ti * TS + ty + 16 * i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = tj * TS + tx + 16 * j;
      if (c < n) out[(int64_t)r * n + c] = 1.f - acc[i][j];
    }
  }
}

__device__ __forceinline__ unsigned long long sort_key(float v, int idx) {
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned)idx;
}
__device__ __forceinline__ float key_value(unsigned long long key) {
  unsigned u = (unsigned)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(u);
}

// One warp ranks one row: bitonic sort of (sim, column) keys in shared memory.  Ascending in sim,
// ties to the lowest column (= a stable argsort).
__global__ void __launch_bounds__(128)
match_rank_kernel(const float* __restrict__ sim, const int32_t* __restrict__ pair_off,
                  const int64_t* __restrict__ sim_off, const int32_t* __restrict__ node_pair,
                  int64_t N, int npad, int K, int32_t* __restrict__ topk_idx,
                  float* __restrict__ topk_dist, int32_t* __restrict__ rank_full) {
  extern __shared__ unsigned long long keys_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_g = (int64_t)blockIdx.x * 4 + warp;
  if (row_g >= N) return;
  unsigned long long* keys = keys_all + (size_t)warp * npad;
  const int b = node_pair[row_g];
  const int o0 = pair_off[b], n = pair_off[b + 1] - o0;
  const int r = (int)(row_g - o0);
  const float* srow = sim + sim_off[b] + (int64_t)r * n;
  int np2 = 32;
  while (np2 < n) np2 <<= 1;
  for (int j = lane; j < np2; j += 32) keys[j] = (j < n) ? sort_key(srow[j], j) : ~0ull;
  __syncwarp();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int s = k >> 1; s > 0; s >>= 1) {
      for (int t = lane; t < np2 / 2; t += 32) {
        int lo = ((t & ~(s - 1)) << 1) | (t & (s - 1));   // index with bit s cleared
        int hi = lo | s;
        bool up = ((lo & k) == 0);
        unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) {
          keys[lo] = c;
          keys[hi] = a;
        }
      }
      __syncwarp();
    }
  }
  if (rank_full) {
    int32_t* dst = rank_full + sim_off[b] + (int64_t)r * n;
    for (int j = lane; j < n; j += 32) dst[j] = (int32_t)(keys[j] & 0xFFFFFFFFull);
  }
  if (topk_idx) {
    for (int j = lane; j < K; j += 32) {
      bool ok = j < n;
      topk_idx[row_g * K + j] = ok ? (int32_t)(keys[j] & 0xFFFFFFFFull) : -1;
      if (topk_dist) topk_dist[row_g * K + j] = ok ? key_value(keys[j]) : INFINITY;
    }
  }
}

// position of column e2i[t] in row e1i[t] once the node itself is removed (alignment.py:3-25)
__global__ void __launch_bounds__(NT)
anchor_pos_kernel(const float* __restrict__ sim, const int32_t* __restrict__ pair_off,
                  const int64_t* __restrict__ sim_off, const int32_t* __restrict__ node_pair,
                  const int32_t* __restrict__ e1i, const int32_t* __restrict__ e2i, int A,
                  int32_t* __restrict__ pos) {
  int t = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (t >= A) return;
  int lane = threadIdx.x & 31;
  const int g1 = e1i[t], g2 = e2i[t];
  const int b = node_pair[g1];
  const int o0 = pair_off[b], n = pair_off[b + 1] - o0;
  const int r = g1 - o0, tgt = g2 - o0;
  const float* srow = sim + sim_off[b] + (int64_t)r * n;
  const unsigned long long kt = sort_key(srow[tgt], tgt);
  int cnt = 0;
  for (int j = lane; j < n; j += 32)
    if (j != r && sort_key(srow[j], j) < kt) ++cnt;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) pos[t] = cnt;
}


// SGAR, alignment score and the top-1 node correspondences of every pair (utils/alignment.py:27-89) in one launch,
// one CTA per pair.
//   phase 1  top1[i] = first entry of row i's ranking once i itself is removed (ties to the lowest column)
//   phase 2  alignment score = #{src node i : top1[i] is a reference node} / n_ref           (alignment.py:80-89)
//   phase 3  SGAR: anchors sorted by the distance of their top-1 prediction; all-correct over the first 2 /
//            the first half / all of them (alignment.py:27-58).  Done without a sort: a wrong anchor's rank
//            is the number of anchors that sort before it, and only the smallest such rank matters.
// pair_out[b] = {sgar '2', sgar '50', sgar '100', alignment score}; the three SGAR values are -1 for a pair
// without anchors (the reference skips such pairs, inference_align_reg.py:121).
__global__ void __launch_bounds__(NT)
pair_metrics_kernel(const float* __restrict__ sim, const int32_t* __restrict__ pair_off,
                    const int64_t* __restrict__ sim_off, const int32_t* __restrict__ n_src,
                    const int32_t* __restrict__ e1i, const int32_t* __restrict__ e2i,
                    const int32_t* __restrict__ anchor_off, int32_t* __restrict__ top1_idx,
                    float* __restrict__ top1_dist, float* __restrict__ pair_out) {
  const int b = blockIdx.x;
  const int o0 = pair_off[b], n = pair_off[b + 1] - o0;
  const int ns = n_src[b], nr = n - ns;
  const float* S = sim + sim_off[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int s_cnt;
  __shared__ int s_minwrong;
  if (threadIdx.x == 0) {
    s_cnt = 0;
    s_minwrong = 0x7fffffff;
  }
  for (int r = warp; r < n; r += NT / 32) {
    const float* srow = S + (int64_t)r * n;
    unsigned long long best = ~0ull;
    for (int j = lane; j < n; j += 32)
      if (j != r) {
        const unsigned long long k = sort_key(srow[j], j);
        best = k < best ? k : best;
      }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if (lane == 0) {
      top1_idx[o0 + r] = best == ~0ull ? -1 : (int32_t)(best & 0xFFFFFFFFull);
      top1_dist[o0 + r] = best == ~0ull ? INFINITY : key_value(best);
    }
  }
  __syncthreads();
  int c = 0;
  for (int r = threadIdx.x; r < ns; r += NT) c += top1_idx[o0 + r] >= ns ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0 && c) atomicAdd(&s_cnt, c);
  const int a0 = anchor_off[b], na = anchor_off[b + 1] - a0;
  for (int t = threadIdx.x; t < na; t += NT) {
    const int g1 = e1i[a0 + t];
    if (top1_idx[g1] == e2i[a0 + t] - o0) continue;
    const unsigned long long kt = sort_key(top1_dist[g1], t);
    int rank = 0;
    for (int u = 0; u < na; ++u) rank += sort_key(top1_dist[e1i[a0 + u]], u) < kt ? 1 : 0;
    atomicMin(&s_minwrong, rank);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* out = pair_out + 4 * (int64_t)b;
    const int mw = s_minwrong;
    out[0] = na ? (mw >= (na < 2 ? na : 2) ? 1.f : 0.f) : -1.f;
    out[1] = na ? (mw >= na / 2 ? 1.f : 0.f) : -1.f;
    out[2] = na ? (mw >= na ? 1.f : 0.f) : -1.f;
    out[3] = nr > 0 ? (float)s_cnt / (float)nr : 0.f;
  }
}

}  // namespace
}  // namespace sga

extern "C" int sga_match_sim(const float* emb, int64_t N, int D, const int32_t* pair_off,
                             const int64_t* sim_off, int B, int max_pair_nodes, float* norms,
                             float* sim, void* stream) {
  if (N <= 0 || B <= 0) return SGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  sga::row_norm_kernel<<<(unsigned)((N + 7) / 8), sga::NT, 0, st>>>(emb, N, D, norms);
  SGA_LAUNCH_CHECK();
  int tiles = (max_pair_nodes + sga::TS - 1) / sga::TS;
  dim3 grid(tiles * tiles, B);
  sga::match_sim_kernel<<<grid, sga::NT, 0, st>>>(emb, norms, D, pair_off, sim_off, sim);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_match_rank(const float* sim, int64_t N, const int32_t* pair_off, const int64_t* sim_off,
                              const int32_t* node_pair, int max_pair_nodes, int K, int32_t* topk_idx,
                              float* topk_dist, int32_t* rank_full, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(max_pair_nodes > 0 && max_pair_nodes <= 4096, "sga_match_rank: max_pair_nodes=%d out of range (1..4096)", max_pair_nodes);
  int npad = 32;
  while (npad < max_pair_nodes) npad <<= 1;
  size_t smem = (size_t)4 * npad * sizeof(unsigned long long);
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SGA_CUDA(cudaFuncSetAttribute(sga::match_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  sga::match_rank_kernel<<<(unsigned)((N + 3) / 4), 128, smem, (cudaStream_t)stream>>>(sim, pair_off, sim_off, node_pair, N, npad, K, topk_idx,
                                                                                        topk_dist, rank_full);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_match_anchor_pos(const float* sim, const int32_t* pair_off, const int64_t* sim_off,
                                    const int32_t* node_pair, const int32_t* e1i, const int32_t* e2i, int A,
                                    int32_t* anchor_pos, void* stream) {
  if (A <= 0) return SGA_OK;
  sga::anchor_pos_kernel<<<(A + 7) / 8, sga::NT, 0, (cudaStream_t)stream>>>(sim, pair_off, sim_off, node_pair, e1i, e2i, A, anchor_pos);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_match_pair_metrics(const float* sim, const int32_t* pair_off, const int64_t* sim_off,
                                      const int32_t* n_src, const int32_t* e1i, const int32_t* e2i,
                                      const int32_t* anchor_off, int B, int32_t* top1_idx, float* top1_dist,
                                      float* pair_out, void* stream) {
  if (B <= 0) return SGA_OK;
  sga::pair_metrics_kernel<<<B, sga::NT, 0, (cudaStream_t)stream>>>(sim, pair_off, sim_off, n_src, e1i, e2i, anchor_off, top1_idx,
                                                                     top1_dist, pair_out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
