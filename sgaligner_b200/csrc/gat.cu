// Graph-attention structure encoder (forward): all 2B graphs of the batch in one launch per stage.
// Reference call sites: src/aligner/networks/gat.py:35-47 (two GATConv layers, ELU in between),
// src/aligner/sg_aligner.py:86-110 (per-graph Python loop this batching replaces); arithmetic of
// torch_geometric 2.2.0 nn/conv/gat_conv.py + utils/softmax.py (see oracle/sgaligner_oracle.py).
#include "common.cuh"
#include "linear.cuh"
#include "ptx.cuh"

namespace sga {
namespace {

constexpr int NT = 256;

// ---------------------------------------------------------------------------------------------
// xs[h][n][c] = sum_k x[n][k] W[h*C+c][k];  a_src[n][h] = <xs[h][n][:], att_src[h][:]>, same dst.
// grid (ceil(N/32), H); 256 threads: ty = tid/32 -> nodes ty*4+i, tx -> channels tx+32j (+128q).
// ---------------------------------------------------------------------------------------------
constexpr int LN = 32;   // nodes per CTA
constexpr int LK = 32;   // k chunk

__global__ void __launch_bounds__(NT)
gat_linear_kernel(const void* __restrict__ x, int x_is_f64, int64_t N, int in_dim,
                  const float* __restrict__ W, const float* __restrict__ att_src,
                  const float* __restrict__ att_dst, int H, int C, float* __restrict__ xs,
                  float* __restrict__ a_src, float* __restrict__ a_dst) {
  __shared__ float xt[LN][LK + 1];
  __shared__ float wt[LK][128 + 1];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int h = blockIdx.y;
  const int64_t n0 = (int64_t)blockIdx.x * LN;
  float as_acc[4] = {0, 0, 0, 0}, ad_acc[4] = {0, 0, 0, 0};
  for (int cq = 0; cq < C; cq += 128) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < in_dim; k0 += LK) {
      __syncthreads();
      for (int i = tid; i < LN * LK; i += NT) {
        int r = i / LK, k = i % LK;
        int64_t n = n0 + r;
        xt[r][k] = (n < N && k0 + k < in_dim) ? load_as_float<float>(x, n * in_dim + k0 + k, x_is_f64) : 0.f;
      }
      for (int i = tid; i < 128 * LK; i += NT) {
        int c = i / LK, k = i % LK;
        wt[k][c] = (cq + c < C && k0 + k < in_dim) ? W[(int64_t)(h * C + cq + c) * in_dim + k0 + k] : 0.f;
      }
      __syncthreads();
      const int kmax = min(LK, in_dim - k0);
      for (int k = 0; k < kmax; ++k) {
        float a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = xt[ty * 4 + i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = wt[k][tx + 32 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = cq + tx + 32 * j;
      if (c < C) {
        float ws = att_src[h * C + c], wd = att_dst[h * C + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int64_t n = n0 + ty * 4 + i;
          if (n < N) xs[((int64_t)h * N + n) * C + c] = acc[i][j];
          as_acc[i] = fmaf(acc[i][j], ws, as_acc[i]);
          ad_acc[i] = fmaf(acc[i][j], wd, ad_acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = warp_sum(as_acc[i]), d = warp_sum(ad_acc[i]);
    int64_t n = n0 + ty * 4 + i;
    if (tx == 0 && n < N) {
      a_src[n * H + h] = s;
      a_dst[n * H + h] = d;
    }
  }
}

// Same contract for C <= 128 (every reference configuration) on the shared double-buffered tile
// (linear.cuh): grid (ceil(N/32), H), one CTA = 32 nodes x one head.
__global__ void __launch_bounds__(linear::NT)
gat_linear_tile_kernel(const void* __restrict__ x, int x_is_f64, int64_t N, int in_dim,
                       const float* __restrict__ W, const float* __restrict__ att_src,
                       const float* __restrict__ att_dst, int H, int C, float* __restrict__ xs,
                       float* __restrict__ a_src, float* __restrict__ a_dst) {
  __shared__ linear::Smem sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int h = blockIdx.y;
  const int64_t n0 = (int64_t)blockIdx.x * linear::BM;
  float acc[2][8];
  linear::tile_mma(sm, x, x_is_f64, N, in_dim, W + (int64_t)h * C * in_dim, C, n0, 0, acc);
  float as_acc[2] = {0.f, 0.f}, ad_acc[2] = {0.f, 0.f};
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int c0 = linear::col_of(tx, 4 * g);
    if (c0 < C) {     // C % 4 == 0 (checked by the launcher): the 4-column group is all in or all out
      const float4 ws = *reinterpret_cast<const float4*>(att_src + h * C + c0);
      const float4 wd = *reinterpret_cast<const float4*>(att_dst + h * C + c0);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float* a = &acc[i][4 * g];
        as_acc[i] += a[0] * ws.x + a[1] * ws.y + a[2] * ws.z + a[3] * ws.w;
        ad_acc[i] += a[0] * wd.x + a[1] * wd.y + a[2] * wd.z + a[3] * wd.w;
        const int64_t n = n0 + ty * 2 + i;
        if (n < N) *reinterpret_cast<float4*>(xs + ((int64_t)h * N + n) * C + c0) = make_float4(a[0], a[1], a[2], a[3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float s = linear::row_sum16(as_acc[i]), d = linear::row_sum16(ad_acc[i]);
    const int64_t n = n0 + ty * 2 + i;
    if (tx == 0 && n < N) {
      a_src[n * H + h] = s;
      a_dst[n * H + h] = d;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Aggregation: one CTA per (graph, head).  The graph's [n, C] tile of this head is contiguous in
// xs (head-major layout) and is pulled into shared memory with one bulk async copy (TMA, UBLKCP)
// when it fits; warps then own destination rows: two passes over the row's incoming edges
// (max, then exp / sum / weighted accumulate) exactly as PyG's softmax does, + bias (+ ELU).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
gat_aggregate_kernel(const float* __restrict__ xs, const float* __restrict__ a_src,
                     const float* __restrict__ a_dst, const int32_t* __restrict__ row_beg,
                     const int32_t* __restrict__ row_cnt, const int32_t* __restrict__ col,
                     const int32_t* __restrict__ node_off, int64_t N, int H, int C,
                     const float* __restrict__ bias, int apply_elu, float* __restrict__ out,
                     int smem_nodes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* tile = reinterpret_cast<float*>(smem_raw);
  const int g = blockIdx.x, h = blockIdx.y;
  const int n0 = node_off[g], n = node_off[g + 1] - n0;
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gsrc = xs + ((int64_t)h * N + n0) * C;
  const bool staged = (n <= smem_nodes);
  if (staged) {
    if (tid == 0) {
      ptx::mbar_init(&bar, 1);
      ptx::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      // n*C*4 bytes, issued in <=64 KiB pieces (the mbarrier transaction count is 20 bits)
      const uint32_t total = (uint32_t)n * C * 4u;
      ptx::mbar_arrive_expect_tx(&bar, total);
      for (uint32_t o = 0; o < total; o += 65536u) {
        uint32_t b = min(65536u, total - o);
        ptx::bulk_g2s(smem_raw + o, reinterpret_cast<const unsigned char*>(gsrc) + o, b, &bar);
      }
    }
    ptx::mbar_wait(&bar, 0);
  }
  const float* src = staged ? tile : gsrc;
  const int nq = (C + 127) / 128;   // 128-channel groups per lane (lane owns 4 channels per group)
  for (int i = warp; i < n; i += NT / 32) {
    const int beg = row_beg[n0 + i], cnt = row_cnt[n0 + i];
    const float ad = a_dst[(int64_t)(n0 + i) * H + h];
    float m = -INFINITY;
    for (int b = 0; b < cnt; b += 32) {
      int k = b + lane;
      if (k < cnt) {
        float z = a_src[(int64_t)col[beg + k] * H + h] + ad;
        z = z > 0.f ? z : 0.2f * z;
        m = fmaxf(m, z);
      }
    }
    m = warp_max(m);
    float s = 0.f;
    float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int b = 0; b < cnt; b += 32) {
      int k = b + lane;
      int j = 0;
      float p = 0.f;
      if (k < cnt) {
        j = col[beg + k];
        float z = a_src[(int64_t)j * H + h] + ad;
        z = z > 0.f ? z : 0.2f * z;
        p = expf(z - m);
      }
      s += p;
      const int lim = min(32, cnt - b);
      for (int t = 0; t < lim; ++t) {
        float pt = __shfl_sync(0xffffffffu, p, t);
        int jt = __shfl_sync(0xffffffffu, j, t) - n0;
        for (int q = 0; q < nq && q < 2; ++q) {
          int c = q * 128 + lane * 4;
          if (c < C) {
            float4 v = *reinterpret_cast<const float4*>(src + (int64_t)jt * C + c);
            acc[q][0] = fmaf(pt, v.x, acc[q][0]);
            acc[q][1] = fmaf(pt, v.y, acc[q][1]);
            acc[q][2] = fmaf(pt, v.z, acc[q][2]);
            acc[q][3] = fmaf(pt, v.w, acc[q][3]);
          }
        }
      }
    }
    s = warp_sum(s) + 1e-16f;
    const float inv = 1.f / s;
    for (int q = 0; q < nq && q < 2; ++q) {
      int c = q * 128 + lane * 4;
      if (c < C) {
        float4 bv = *reinterpret_cast<const float4*>(bias + h * C + c);
        float o[4] = {acc[q][0] * inv + bv.x, acc[q][1] * inv + bv.y, acc[q][2] * inv + bv.z, acc[q][3] * inv + bv.w};
        if (apply_elu) {
#pragma unroll
          for (int u = 0; u < 4; ++u) o[u] = o[u] > 0.f ? o[u] : expm1f(o[u]);
        }
        *reinterpret_cast<float4*>(out + (int64_t)(n0 + i) * (H * C) + h * C + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

}  // namespace
}  // namespace sga

extern "C" int sga_gat_linear(const void* x, int x_is_f64, int64_t N, int in_dim, const float* W,
                              const float* att_src, const float* att_dst, int H, int C, float* xs,
                              float* a_src, float* a_dst, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(in_dim > 0 && H > 0 && C > 0, "sga_gat_linear: bad dims in=%d H=%d C=%d", in_dim, H, C);
  if (C <= sga::linear::BN && C % 4 == 0 && ((reinterpret_cast<uintptr_t>(att_src) | reinterpret_cast<uintptr_t>(att_dst) |
                                               reinterpret_cast<uintptr_t>(xs)) & 15) == 0) {
    dim3 grid((unsigned)((N + sga::linear::BM - 1) / sga::linear::BM), H);
    sga::gat_linear_tile_kernel<<<grid, sga::linear::NT, 0, (cudaStream_t)stream>>>(x, x_is_f64, N, in_dim, W, att_src, att_dst, H, C, xs,
                                                                                  a_src, a_dst);
    SGA_LAUNCH_CHECK();
    return SGA_OK;
  }
  dim3 grid((unsigned)((N + sga::LN - 1) / sga::LN), H);
  sga::gat_linear_kernel<<<grid, sga::NT, 0, (cudaStream_t)stream>>>(x, x_is_f64, N, in_dim, W, att_src, att_dst, H, C, xs, a_src, a_dst);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_gat_aggregate(const float* xs, const float* a_src, const float* a_dst,
                                 const int32_t* row_beg, const int32_t* row_cnt, const int32_t* col,
                                 const int32_t* node_off, int G, int max_graph_nodes, int64_t N, int H,
                                 int C, const float* bias, int apply_elu, float* out, void* stream) {
  if (N <= 0 || G <= 0) return SGA_OK;
  SGA_REQUIRE(C % 4 == 0 && C <= 256, "sga_gat_aggregate: C=%d must be a multiple of 4 and <= 256", C);
  // stage the graph tile in shared memory when the largest graph fits (<= 200 KiB)
  const size_t cap = 200 * 1024;
  size_t need = (size_t)max_graph_nodes * C * sizeof(float);
  int smem_nodes = 0;
  size_t smem = 0;
  if (need <= cap && (C * sizeof(float)) % 16 == 0) {
    smem_nodes = max_graph_nodes;
    smem = need;
  }
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SGA_CUDA(cudaFuncSetAttribute(sga::gat_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  dim3 grid(G, H);
  sga::gat_aggregate_kernel<<<grid, sga::NT, smem, (cudaStream_t)stream>>>(xs, a_src, a_dst, row_beg, row_cnt, col, node_off, N, H, C, bias,
                                                                           apply_elu, out, smem_nodes);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
