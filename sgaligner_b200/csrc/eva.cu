// EVA baseline path (reference: src/aligner/eva.py:9-96, MultiGCN of
// src/aligner/networks/gat.py:6-25 over torch_geometric 2.2.0 GCNConv, NCALoss / OverallNCALoss of
// src/aligner/losses.py:154-205).  The dense contractions of this path (GCN layer 2, the NCA score matrix and its two
// gradient products) run on the tf32x3 tcgen05 GEMM of gemm_tc.cu; this file holds the graph / row-wise kernels around
// them.  All of them are HBM / latency bound fp32 kernels: one warp per node row, float4 where the width allows.
//
//   gcn_aggregate      out_i = sum_{j -> i} xw_j / sqrt(deg_j deg_i) (+ bias) (ReLU)     over the block-diagonal CSR of
//                      csr.cu (self loops removed, one added per node = PyG's add_remaining_self_loops; deg = in-degree
//                      incl. the self loop = the CSR row length of the FORWARD graph).  The backward is the same kernel
//                      over the CSR of the reversed edges (the weights are symmetric in (deg_i, deg_j)).
//   linear_smallk / wgrad_smallk   the 3 -> 200 first-layer linear map and its weight gradient (K <= 8: no GEMM)
//   fuse_rows fwd/bwd  MultiModalFusion (sg_aligner.py:23-35) for modalities of DIFFERENT widths (EVA fuses the raw
//                      400-d GCN output and the 200-d PointNet feature with the two 100-d meta embeddings, eva.py:72-76)
//   nca_*              NCALoss (losses.py:161-176): row / column sums of exp(alpha (S - ep)) off the diagonal, the loss
//                      value, and dS in place
//   row_l2norm / normalize_bwd     F.normalize (eps 1e-12) and its backward
#include "common.cuh"

namespace sga {
namespace {

constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(256)
gcn_aggregate_kernel(const float* __restrict__ h, int64_t N, int C, const int32_t* __restrict__ row_beg,
                     const int32_t* __restrict__ row_cnt, const int32_t* __restrict__ col, const int32_t* __restrict__ deg,
                     const float* __restrict__ bias, int relu, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= N) return;
  const int lane = threadIdx.x & 31;
  const int beg = row_beg[i], cnt = row_cnt[i];
  const float di = 1.0f / sqrtf((float)deg[i]);
  if ((C & 3) == 0) {
    for (int c = lane * 4; c < C; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < cnt; ++k) {
        const int j = col[beg + k];
        const float w = di * (1.0f / sqrtf((float)deg[j]));
        const float4 v = *reinterpret_cast<const float4*>(h + (int64_t)j * C + c);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      if (bias) { acc.x += bias[c]; acc.y += bias[c + 1]; acc.z += bias[c + 2]; acc.w += bias[c + 3]; }
      if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      *reinterpret_cast<float4*>(out + i * C + c) = acc;
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int k = 0; k < cnt; ++k) {
        const int j = col[beg + k];
        acc = fmaf(di * (1.0f / sqrtf((float)deg[j])), h[(int64_t)j * C + c], acc);
      }
      if (bias) acc += bias[c];
      if (relu) acc = fmaxf(acc, 0.f);
      out[i * C + c] = acc;
    }
  }
}

// Y[i, c] = sum_k X[i, k] W[c, k]   (K <= 8), thread per output element
__global__ void __launch_bounds__(256)
linear_smallk_kernel(const float* __restrict__ X, int64_t N, int K, const float* __restrict__ W, int C, float* __restrict__ Y) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= N * C) return;
  const int64_t i = t / C;
  const int c = (int)(t - i * C);
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(X[i * K + k], W[c * K + k], acc);
  Y[t] = acc;
}

// gW[c, k] += sum_i G[i, c] X[i, k]   (K <= 8): a block owns 256 rows, a thread a channel (strided)
__global__ void __launch_bounds__(256)
wgrad_smallk_kernel(const float* __restrict__ G, const float* __restrict__ X, int64_t N, int K, int C, float* __restrict__ gW) {
  __shared__ float xs[256 * 8];
  const int64_t i0 = (int64_t)blockIdx.x * 256;
  const int rows = (int)min((int64_t)256, N - i0);
  for (int t = threadIdx.x; t < rows * K; t += 256) xs[t] = X[i0 * K + t];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < rows; ++r) {
      const float g = G[(i0 + r) * C + c];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < K) acc[k] = fmaf(g, xs[r * K + k], acc[k]);
    }
    for (int k = 0; k < K; ++k) atomicAdd(&gW[c * K + k], acc[k]);
  }
}

// out = g * (y > 0)
__global__ void __launch_bounds__(256)
relu_mask_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t n, float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t < n) out[t] = y[t] > 0.f ? g[t] : 0.f;
}

// s[c] += sum_i x[i, c]: a block owns 64 rows, threads own channels
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const float* __restrict__ x, int64_t N, int C, float* __restrict__ s) {
  const int64_t i0 = (int64_t)blockIdx.x * 64;
  const int rows = (int)min((int64_t)64, N - i0);
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += x[(i0 + r) * C + c];
    atomicAdd(&s[c], acc);
  }
}

// norms[i] = max(||x_i||_2, eps)      (F.normalize's denominator)
__global__ void __launch_bounds__(256)
row_l2norm_kernel(const float* __restrict__ x, int64_t N, int D, float eps, float* __restrict__ norms) {
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= N) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = x[i * D + k];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) norms[i] = fmaxf(sqrtf(s), eps);
}

// g_x_i = (g_i - xh_i (xh_i . g_i)) / norm_i with xh_i = x_i / norm_i; in place on g.  (Exact for ||x|| > eps; at the
// clamp F.normalize is x / eps, whose gradient is g / eps: the projection term is dropped there.)
__global__ void __launch_bounds__(256)
normalize_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ norms, int64_t N, int D, float eps, float* __restrict__ g) {
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= N) return;
  const int lane = threadIdx.x & 31;
  const float nr = norms[i], inv = 1.0f / nr;
  float dot = 0.f;
  for (int k = lane; k < D; k += 32) dot = fmaf(x[i * D + k] * inv, g[i * D + k], dot);
  dot = warp_sum(dot);
  if (nr <= eps) dot = 0.f;
  for (int k = lane; k < D; k += 32) g[i * D + k] = (g[i * D + k] - x[i * D + k] * inv * dot) * inv;
}

// ---------------------------------------------------------------------------------------------------------------------
// MultiModalFusion for modalities of different widths.  w = softmax(fusion_w); joint[:, off_m : off_m + d_m] =
// w_m * x_m / max(||x_m||, 1e-12).
constexpr int kMaxFuse = 8;
struct FuseArgs {
  const float* x[kMaxFuse];
  float* gx[kMaxFuse];
  int d[kMaxFuse];
  int off[kMaxFuse];
  int M;
};

__device__ __forceinline__ float softmax_w(const float* __restrict__ fw, int M, int m) {
  float mx = -INFINITY;
  for (int k = 0; k < M; ++k) mx = fmaxf(mx, fw[k]);
  float s = 0.f;
  for (int k = 0; k < M; ++k) s += expf(fw[k] - mx);
  return expf(fw[m] - mx) / s;
}

__global__ void __launch_bounds__(256)
fuse_rows_fwd_kernel(FuseArgs A, const float* __restrict__ fw, int64_t N, int ld, float* __restrict__ joint) {
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= N) return;
  const int m = blockIdx.y, lane = threadIdx.x & 31;
  const int d = A.d[m];
  const float* x = A.x[m] + i * d;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) s = fmaf(x[k], x[k], s);
  s = warp_sum(s);
  const float sc = softmax_w(fw, A.M, m) / fmaxf(sqrtf(s), 1e-12f);
  float* o = joint + i * ld + A.off[m];
  for (int k = lane; k < d; k += 32) o[k] = sc * x[k];
}

// gx_m = w_m / norm (g - xh (xh . g));  sdot[m] += sum_i xh_i . g_i   (for the fusion-weight gradient)
__global__ void __launch_bounds__(256)
fuse_rows_bwd_kernel(FuseArgs A, const float* __restrict__ fw, int64_t N, int ld, const float* __restrict__ g_joint,
                     float* __restrict__ sdot) {
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int m = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float part[kWarpsPerBlock];
  float dot = 0.f;
  if (i < N) {
    const int d = A.d[m];
    const float* x = A.x[m] + i * d;
    const float* g = g_joint + i * ld + A.off[m];
    float s = 0.f;
    for (int k = lane; k < d; k += 32) {
      s = fmaf(x[k], x[k], s);
      dot = fmaf(x[k], g[k], dot);
    }
    s = warp_sum(s);
    dot = warp_sum(dot);
    const float nr = fmaxf(sqrtf(s), 1e-12f), inv = 1.0f / nr;
    dot *= inv;                                       // xh . g
    const float w = softmax_w(fw, A.M, m);
    float* gx = A.gx[m] + i * d;
    const float proj = (sqrtf(s) > 1e-12f) ? dot : 0.f;
    for (int k = lane; k < d; k += 32) gx[k] = w * inv * (g[k] - x[k] * inv * proj);
  }
  if (lane == 0) part[warp] = (i < N) ? dot : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kWarpsPerBlock; ++k) t += part[k];
    atomicAdd(&sdot[m], t);
  }
}

// g_fw[m] += w_m (s_m - sum_k w_k s_k)       (softmax backward; one thread)
__global__ void fuse_w_grad_kernel(const float* __restrict__ fw, const float* __restrict__ sdot, int M, float* __restrict__ g_fw) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mean = 0.f;
  for (int k = 0; k < M; ++k) mean += softmax_w(fw, M, k) * sdot[k];
  for (int m = 0; m < M; ++m) atomicAdd(&g_fw[m], softmax_w(fw, M, m) * (sdot[m] - mean));
}

// ---------------------------------------------------------------------------------------------------------------------
// NCALoss over the score matrix S [A, A] (row-major).
__global__ void __launch_bounds__(256)
nca_rows_kernel(const float* __restrict__ S, int A, float alpha, float ep, float* __restrict__ rs, float* __restrict__ diag) {
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= A) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int j = lane; j < A; j += 32)
    if (j != i) s += expf(alpha * (S[(int64_t)i * A + j] - ep));
  s = warp_sum(s);
  if (lane == 0) {
    rs[i] = s;
    diag[i] = S[(int64_t)i * A + i];
  }
}

__global__ void __launch_bounds__(128)
nca_cols_kernel(const float* __restrict__ S, int A, float alpha, float ep, float* __restrict__ cs) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int i0 = blockIdx.y * 128, i1 = min(A, i0 + 128);
  if (j >= A) return;
  float s = 0.f;
  for (int i = i0; i < i1; ++i)
    if (i != j) s += expf(alpha * (S[(int64_t)i * A + j] - ep));
  atomicAdd(&cs[j], s);
}

// loss = mean_j log(1 + cs_j) / alpha + mean_i log(1 + rs_i) / alpha - beta mean_j log(1 + relu(diag_j))
__global__ void __launch_bounds__(256)
nca_loss_kernel(const float* __restrict__ rs, const float* __restrict__ cs, const float* __restrict__ diag, int A, float alpha,
                float beta, float* __restrict__ loss) {
  __shared__ double part[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < A; i += 256)
    acc += (double)(log1pf(cs[i]) / alpha) + (double)(log1pf(rs[i]) / alpha) - (double)beta * (double)log1pf(fmaxf(diag[i], 0.f));
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(part[0] / (double)A);
}

// S -> dS in place:  i != j: exp(alpha (S - ep)) (1 / (1 + cs_j) + 1 / (1 + rs_i)) / A;   i == j: -beta [S_jj > 0] / (A (1 + S_jj))
__global__ void __launch_bounds__(256)
nca_coef_kernel(float* __restrict__ S, int A, float alpha, float ep, float beta, const float* __restrict__ rs,
                const float* __restrict__ cs) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (int64_t)A * A) return;
  const int i = (int)(t / A), j = (int)(t - (int64_t)i * A);
  const float s = S[t], invA = 1.0f / (float)A;
  float g;
  if (i != j) g = expf(alpha * (s - ep)) * (1.0f / (1.0f + cs[j]) + 1.0f / (1.0f + rs[i])) * invA;
  else g = s > 0.f ? -beta * invA / (1.0f + s) : 0.f;
  S[t] = g;
}

}  // namespace
}  // namespace sga

using namespace sga;

extern "C" int sga_gcn_aggregate(const float* h, int64_t N, int C, const int32_t* row_beg, const int32_t* row_cnt,
                                 const int32_t* col, const int32_t* deg, const float* bias, int relu, float* out, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(h && row_beg && row_cnt && col && deg && out && C >= 1, "sga_gcn_aggregate: null pointer or C=%d", C);
  SGA_REQUIRE((C & 3) != 0 || ((((uintptr_t)h | (uintptr_t)out) & 15) == 0), "sga_gcn_aggregate: h / out must be 16-byte aligned");
  gcn_aggregate_kernel<<<(unsigned)((N + kWarpsPerBlock - 1) / kWarpsPerBlock), 256, 0, (cudaStream_t)stream>>>(h, N, C, row_beg, row_cnt, col,
                                                                                                                  deg, bias, relu, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_linear_smallk(const float* X, int64_t N, int K, const float* W, int C, float* Y, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(X && W && Y && K >= 1 && K <= 8 && C >= 1, "sga_linear_smallk: K=%d (1..8) C=%d", K, C);
  linear_smallk_kernel<<<(unsigned)((N * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, N, K, W, C, Y);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_wgrad_smallk(const float* G, const float* X, int64_t N, int K, int C, float* gW, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(G && X && gW && K >= 1 && K <= 8 && C >= 1, "sga_wgrad_smallk: K=%d (1..8) C=%d", K, C);
  wgrad_smallk_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(G, X, N, K, C, gW);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_relu_mask(const float* g, const float* y, int64_t n, float* out, void* stream) {
  if (n <= 0) return SGA_OK;
  SGA_REQUIRE(g && y && out, "sga_relu_mask: null pointer");
  relu_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, y, n, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_colsum_rows(const float* x, int64_t N, int C, float* s, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && s && C >= 1, "sga_colsum_rows: null pointer");
  colsum_rows_kernel<<<(unsigned)((N + 63) / 64), 256, 0, (cudaStream_t)stream>>>(x, N, C, s);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_row_l2norm(const float* x, int64_t N, int D, float eps, float* norms, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && norms && D >= 1, "sga_row_l2norm: null pointer");
  row_l2norm_kernel<<<(unsigned)((N + kWarpsPerBlock - 1) / kWarpsPerBlock), 256, 0, (cudaStream_t)stream>>>(x, N, D, eps, norms);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_normalize_bwd_rows(const float* x, const float* norms, int64_t N, int D, float eps, float* g, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(x && norms && g && D >= 1, "sga_normalize_bwd_rows: null pointer");
  normalize_bwd_rows_kernel<<<(unsigned)((N + kWarpsPerBlock - 1) / kWarpsPerBlock), 256, 0, (cudaStream_t)stream>>>(x, norms, N, D, eps, g);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

static int fill_fuse(FuseArgs& A, const float* const* x_host, float* const* gx_host, const int* dims_host, int M, int ld) {
  SGA_REQUIRE(M >= 1 && M <= kMaxFuse, "fuse_rows: M=%d (1..%d)", M, kMaxFuse);
  int off = 0;
  for (int m = 0; m < M; ++m) {
    SGA_REQUIRE(x_host[m] && dims_host[m] >= 1, "fuse_rows: bad modality %d", m);
    A.x[m] = x_host[m];
    A.gx[m] = gx_host ? gx_host[m] : nullptr;
    A.d[m] = dims_host[m];
    A.off[m] = off;
    off += dims_host[m];
  }
  A.M = M;
  SGA_REQUIRE(ld >= off, "fuse_rows: joint leading dimension %d < %d", ld, off);
  return SGA_OK;
}

extern "C" int sga_fuse_rows_fwd(const float* const* x_host, const int* dims_host, int M, const float* fusion_w, int64_t N,
                                 float* joint, int ld, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(fusion_w && joint, "sga_fuse_rows_fwd: null pointer");
  FuseArgs A;
  int rc = fill_fuse(A, x_host, nullptr, dims_host, M, ld);
  if (rc) return rc;
  dim3 grid((unsigned)((N + kWarpsPerBlock - 1) / kWarpsPerBlock), M);
  fuse_rows_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, fusion_w, N, ld, joint);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_fuse_rows_bwd(const float* const* x_host, const int* dims_host, int M, const float* fusion_w, int64_t N,
                                 const float* g_joint, int ld, float* const* gx_host, float* g_fusion_w, float* scratch_M,
                                 void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(fusion_w && g_joint && gx_host && g_fusion_w && scratch_M, "sga_fuse_rows_bwd: null pointer");
  FuseArgs A;
  int rc = fill_fuse(A, x_host, gx_host, dims_host, M, ld);
  if (rc) return rc;
  for (int m = 0; m < M; ++m) SGA_REQUIRE(gx_host[m], "sga_fuse_rows_bwd: null gradient buffer %d", m);
  cudaStream_t st = (cudaStream_t)stream;
  SGA_CUDA(cudaMemsetAsync(scratch_M, 0, sizeof(float) * M, st));
  dim3 grid((unsigned)((N + kWarpsPerBlock - 1) / kWarpsPerBlock), M);
  fuse_rows_bwd_kernel<<<grid, 256, 0, st>>>(A, fusion_w, N, ld, g_joint, scratch_M);
  SGA_LAUNCH_CHECK();
  fuse_w_grad_kernel<<<1, 32, 0, st>>>(fusion_w, scratch_M, M, g_fusion_w);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_nca_forward(const float* S, int A, float alpha, float beta, float ep, float* rs, float* cs, float* diag,
                               float* loss, void* stream) {
  SGA_REQUIRE(S && rs && cs && diag && loss && A >= 1 && alpha != 0.f, "sga_nca_forward: bad arguments (A=%d)", A);
  cudaStream_t st = (cudaStream_t)stream;
  SGA_CUDA(cudaMemsetAsync(cs, 0, sizeof(float) * A, st));
  nca_rows_kernel<<<(A + kWarpsPerBlock - 1) / kWarpsPerBlock, 256, 0, st>>>(S, A, alpha, ep, rs, diag);
  SGA_LAUNCH_CHECK();
  dim3 grid((A + 127) / 128, (A + 127) / 128);
  nca_cols_kernel<<<grid, 128, 0, st>>>(S, A, alpha, ep, cs);
  SGA_LAUNCH_CHECK();
  nca_loss_kernel<<<1, 256, 0, st>>>(rs, cs, diag, A, alpha, beta, loss);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_nca_coef(float* S, int A, float alpha, float beta, float ep, const float* rs, const float* cs, void* stream) {
  SGA_REQUIRE(S && rs && cs && A >= 1, "sga_nca_coef: bad arguments");
  const int64_t n = (int64_t)A * A;
  nca_coef_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(S, A, alpha, ep, beta, rs, cs);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
