// Backward of the per-modality projection + fusion slice (project_fuse.cu):
//   joint[:, col:col+d] = w_m * emb / max(||emb||, 1e-12),  w = softmax(fusion.weight)
// (src/aligner/sg_aligner.py:30-35 and :112-122).
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace sga {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ float softmax_w(const float* __restrict__ fw, int M, int m) {
  float mx = -INFINITY;
  for (int i = 0; i < M; ++i) mx = fmaxf(mx, fw[i]);
  float s = 0.f;
  for (int i = 0; i < M; ++i) s += expf(fw[i] - mx);
  return expf(fw[m] - mx) / s;
}

// g_total[n,:] = g_emb[n,:] + w_m * (gj - eh <eh, gj>) / ||emb||;   t_acc += sum_n <gj, eh>
__global__ void __launch_bounds__(NT)
fuse_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ g_emb, const float* __restrict__ g_joint,
                int joint_ld, int joint_col, const float* __restrict__ fusion_w, int M, int m, int64_t N, int D,
                float* __restrict__ g_total, float* __restrict__ t_acc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float wm = g_joint ? softmax_w(fusion_w, M, m) : 0.f;
  float tsum = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * (NT / 32) + warp; row < N; row += (int64_t)gridDim.x * (NT / 32)) {
    const float* e = emb + row * D;
    float ss = 0.f, dot = 0.f;
    if (g_joint) {
      const float* gj = g_joint + row * joint_ld + joint_col;
      for (int k = lane; k < D; k += 32) {
        float v = e[k];
        ss = fmaf(v, v, ss);
        dot = fmaf(v, gj[k], dot);
      }
      ss = warp_sum(ss);
      dot = warp_sum(dot);
    }
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, 1e-12f);
    const bool clamped = nrm < 1e-12f;
    const float dh = dot / den;   // <eh, gj>
    tsum += dh;
    for (int k = lane; k < D; k += 32) {
      float g = g_emb ? g_emb[row * D + k] : 0.f;
      if (g_joint) {
        float gj = g_joint[row * joint_ld + joint_col + k];
        float eh = e[k] / den;
        g += wm * (gj - (clamped ? 0.f : eh * dh)) / den;
      }
      g_total[row * D + k] = g;
    }
  }
  if (g_joint && lane == 0) atomicAdd(t_acc, tsum);
}

// softmax backward for this modality's contribution: g_fw[i] += t * w_m * (delta_mi - w_i)
__global__ void fusion_w_bwd_kernel(const float* __restrict__ fusion_w, int M, int m, const float* __restrict__ t_acc,
                                    float* __restrict__ g_fw) {
  int i = threadIdx.x;
  if (i >= M) return;
  float wm = softmax_w(fusion_w, M, m), wi = softmax_w(fusion_w, M, i);
  g_fw[i] += t_acc[0] * wm * ((i == m ? 1.f : 0.f) - wi);
}

// gb[c] += sum_n g[n,c]: 64 rows per CTA (one wave of CTAs at N = 4096..9472), 4 row groups of 16 rows
// combined through shared memory, one atomic per (CTA, column)
constexpr int CS_ROWS = 64;
__global__ void __launch_bounds__(NT)
colsum_kernel(const float* __restrict__ g, int64_t N, int D, float* __restrict__ gb) {
  __shared__ float part[4][64];
  const int64_t r0 = (int64_t)blockIdx.x * CS_ROWS;
  const int cl = threadIdx.x & 63, rg = threadIdx.x >> 6;     // column within a 64-column slab, row group
  for (int c0 = 0; c0 < D; c0 += 64) {
    const int c = c0 + cl;
    float s = 0.f;
    if (c < D) {
#pragma unroll 4
      for (int r = rg * 16; r < rg * 16 + 16; ++r) {
        const int64_t n = r0 + r;
        if (n < N) s += g[n * D + c];
      }
    }
    part[rg][cl] = s;
    __syncthreads();
    if (rg == 0 && c < D) atomicAdd(&gb[c], part[0][cl] + part[1][cl] + part[2][cl] + part[3][cl]);
    __syncthreads();
  }
}

}  // namespace
}  // namespace sga

extern "C" int sga_project_fuse_bwd(const float* x, int64_t N, int in_dim, const float* W, int out_dim,
                                    const float* emb, const float* g_emb, const float* g_joint, int joint_ld,
                                    int joint_col, const float* fusion_w, int M, int m, float* gW, float* gb,
                                    float* g_fusion_w, float* gx, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(g_emb || g_joint, "sga_project_fuse_bwd: no upstream gradient");
  const size_t need = ((size_t)N * out_dim + 64) * sizeof(float);
  if (workspace_bytes < need) {
    sga::set_error("sga_project_fuse_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return SGA_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* g_total = (float*)workspace;
  float* t_acc = g_total + (size_t)N * out_dim;
  SGA_CUDA(cudaMemsetAsync(t_acc, 0, sizeof(float), st));
  int64_t blocks = (N + 7) / 8;
  if (blocks > (int64_t)sga::sm_count() * 8) blocks = (int64_t)sga::sm_count() * 8;
  sga::fuse_bwd_kernel<<<(unsigned)blocks, sga::NT, 0, st>>>(emb, g_emb, g_joint, joint_ld, joint_col, fusion_w, M, m, N, out_dim, g_total, t_acc);
  SGA_LAUNCH_CHECK();
  if (g_joint && g_fusion_w) {
    sga::fusion_w_bwd_kernel<<<1, 32, 0, st>>>(fusion_w, M, m, t_acc, g_fusion_w);
    SGA_LAUNCH_CHECK();
  }
  // gW[c][k] += sum_n g[n][c] x[n][k]: contraction over the nodes, both operands read MN-major by the tcgen05 GEMM
  const bool tc = sga::gemm_tc_worth(out_dim, in_dim);
  if (tc) {
    int rc = sga::launch_gemm_tc_dense(g_total, out_dim, 1, x, in_dim, 1, out_dim, in_dim, (int)N, gW, in_dim, 1, st);
    if (rc != SGA_OK) return rc;
  } else {
    SGA_CUDA(sga::launch_gemm(g_total, 1, out_dim, x, in_dim, 1, gW, in_dim, out_dim, in_dim, (int)N, 1, st,
                              sga::splitk_for(out_dim, in_dim, (int)N)));
  }
  sga::colsum_kernel<<<(unsigned)((N + sga::CS_ROWS - 1) / sga::CS_ROWS), sga::NT, 0, st>>>(g_total, N, out_dim, gb);
  SGA_LAUNCH_CHECK();
  // gx[n][k] = sum_c g[n][c] W[c][k]
  if (gx && tc) {
    int rc = sga::launch_gemm_tc_dense(g_total, out_dim, 0, W, in_dim, 1, (int)N, in_dim, out_dim, gx, in_dim, 0, st);
    if (rc != SGA_OK) return rc;
  } else if (gx) {
    SGA_CUDA(sga::launch_gemm(g_total, out_dim, 1, W, in_dim, 1, gx, in_dim, (int)N, in_dim, out_dim, 0, st));
  }
  return SGA_OK;
}
