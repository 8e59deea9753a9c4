// Backward of the batched graph-attention layer (autograd of PyG 2.2.0 GATConv as called from
// src/aligner/networks/gat.py:40-48).  Same decomposition as the forward (gat.cu):
//   aggregate_bwd : one CTA per (graph, head); the head's feature tile and its gradient tile live
//                   in shared memory (bulk-copied in, written back once), so the scatter to source
//                   nodes needs shared-memory atomics only
//   linear_bwd    : folds the attention-logit gradients into g_xs, reduces g_att, then three small
//                   GEMMs (gW, gx) on the fp32 FMA path.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace sga {
namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT)
gat_aggregate_bwd_kernel(const float* __restrict__ xs, const float* __restrict__ a_src,
                         const float* __restrict__ a_dst, const int32_t* __restrict__ row_beg,
                         const int32_t* __restrict__ row_cnt, const int32_t* __restrict__ col,
                         const int32_t* __restrict__ node_off, int64_t N, int H, int C, int apply_elu,
                         const float* __restrict__ out, const float* __restrict__ gout,
                         float* __restrict__ g_xs, float* __restrict__ g_a_src, float* __restrict__ g_a_dst,
                         float* __restrict__ g_bias, int smem_nodes, int dense) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int g = blockIdx.x, h = blockIdx.y;
  const int n0 = node_off[g], n = node_off[g + 1] - n0;
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HC = H * C;
  const float* gsrc = xs + ((int64_t)h * N + n0) * C;
  float* gdst = g_xs + ((int64_t)h * N + n0) * C;
  const bool staged = (n <= smem_nodes);
  // smem: [tile n*C] [gtile n*C] [gas n] [gbias C]
  float* tile = reinterpret_cast<float*>(smem_raw);
  float* gtile = tile + (size_t)smem_nodes * C;
  float* gas = gtile + (size_t)smem_nodes * C;
  float* gbias = gas + smem_nodes;
  // dense mode (graphs small enough for an n x n coefficient matrix in shared memory): instead of scattering
  // alpha_ij * go_i into the source rows with four shared-memory atomics per lane and edge, the attention
  // coefficients go into Aij and the upstream rows into gtile; a second phase gathers
  // g_xs[j] = sum_i Aij[i][j] go_i per source row without atomics.
  float* Aij = gbias + C;
  if (staged) {
    if (tid == 0) {
      ptx::mbar_init(&bar, 1);
      ptx::fence_mbar_init();
    }
    if (dense) {
      for (int i = tid; i < n * n; i += NT) Aij[i] = 0.f;
    } else {
      for (int i = tid; i < n * C; i += NT) gtile[i] = 0.f;
    }
    for (int i = tid; i < n; i += NT) gas[i] = 0.f;
    __syncthreads();
    if (tid == 0) {
      const uint32_t total = (uint32_t)n * C * 4u;
      ptx::mbar_arrive_expect_tx(&bar, total);
      for (uint32_t o = 0; o < total; o += 65536u) {
        uint32_t b = min(65536u, total - o);
        ptx::bulk_g2s(smem_raw + o, reinterpret_cast<const unsigned char*>(gsrc) + o, b, &bar);
      }
    }
    ptx::mbar_wait(&bar, 0);
  }
  for (int i = tid; i < C; i += NT) gbias[i] = 0.f;
  __syncthreads();
  const float* src = staged ? tile : gsrc;
  const int nq = (C + 127) / 128;
  for (int i = warp; i < n; i += NT / 32) {
    const int beg = row_beg[n0 + i], cnt = row_cnt[n0 + i];
    const float ad = a_dst[(int64_t)(n0 + i) * H + h];
    // upstream gradient on the pre-activation output of this (row, head)
    float go[2][4];
    for (int q = 0; q < 2; ++q) {
      int c = q * 128 + lane * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) go[q][u] = 0.f;
      if (q < nq && c < C) {
        const int64_t o = (int64_t)(n0 + i) * HC + h * C + c;
        float4 gv = *reinterpret_cast<const float4*>(gout + o);
        float gg[4] = {gv.x, gv.y, gv.z, gv.w};
        if (apply_elu) {
          float4 ov = *reinterpret_cast<const float4*>(out + o);
          float oo[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) gg[u] *= (oo[u] > 0.f ? 1.f : oo[u] + 1.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          go[q][u] = gg[u];
          atomicAdd(&gbias[c + u], gg[u]);
        }
        if (dense) *reinterpret_cast<float4*>(gtile + (int64_t)i * C + c) = make_float4(gg[0], gg[1], gg[2], gg[3]);
      }
    }
    // softmax statistics (as in the forward)
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
    for (int b = 0; b < cnt; b += 32) {
      int k = b + lane;
      if (k < cnt) {
        float z = a_src[(int64_t)col[beg + k] * H + h] + ad;
        z = z > 0.f ? z : 0.2f * z;
        s += expf(z - m);
      }
    }
    s = warp_sum(s) + 1e-16f;
    const float inv = 1.f / s;
    // pass A: D = sum_k alpha_k <go, xs_jk>  (= <go, aggregated row>)
    float dpart = 0.f;
    for (int b = 0; b < cnt; b += 32) {
      int k = b + lane;
      int j = 0;
      float al = 0.f;
      if (k < cnt) {
        j = col[beg + k];
        float z = a_src[(int64_t)j * H + h] + ad;
        z = z > 0.f ? z : 0.2f * z;
        al = expf(z - m) * inv;
      }
      const int lim = min(32, cnt - b);
      for (int t = 0; t < lim; ++t) {
        float at = __shfl_sync(0xffffffffu, al, t);
        int jt = __shfl_sync(0xffffffffu, j, t) - n0;
        for (int q = 0; q < nq && q < 2; ++q) {
          int c = q * 128 + lane * 4;
          if (c < C) {
            float4 v = *reinterpret_cast<const float4*>(src + (int64_t)jt * C + c);
            dpart += at * (go[q][0] * v.x + go[q][1] * v.y + go[q][2] * v.z + go[q][3] * v.w);
          }
        }
      }
    }
    const float D = warp_sum(dpart);
    // pass B: per edge d_alpha, softmax + leaky-ReLU backward, scatter
    float gad = 0.f;
    for (int b = 0; b < cnt; b += 32) {
      int k = b + lane;
      int j = 0;
      float al = 0.f, zraw = 0.f;
      if (k < cnt) {
        j = col[beg + k];
        zraw = a_src[(int64_t)j * H + h] + ad;
        float z = zraw > 0.f ? zraw : 0.2f * zraw;
        al = expf(z - m) * inv;
      }
      const int lim = min(32, cnt - b);
      float my_dalpha = 0.f;
      for (int t = 0; t < lim; ++t) {
        float at = __shfl_sync(0xffffffffu, al, t);
        int jg = __shfl_sync(0xffffffffu, j, t);
        int jt = jg - n0;
        float dot = 0.f;
        for (int q = 0; q < nq && q < 2; ++q) {
          int c = q * 128 + lane * 4;
          if (c < C) {
            float4 v = *reinterpret_cast<const float4*>(src + (int64_t)jt * C + c);
            dot += go[q][0] * v.x + go[q][1] * v.y + go[q][2] * v.z + go[q][3] * v.w;
            if (!dense) {
              float* gp = staged ? (gtile + (int64_t)jt * C + c) : (gdst + (int64_t)jt * C + c);
              atomicAdd(gp + 0, at * go[q][0]);
              atomicAdd(gp + 1, at * go[q][1]);
              atomicAdd(gp + 2, at * go[q][2]);
              atomicAdd(gp + 3, at * go[q][3]);
            }
          }
        }
        dot = warp_sum(dot);
        if (lane == t) my_dalpha = dot;
      }
      if (k < cnt) {
        if (dense) atomicAdd(&Aij[i * n + (j - n0)], al);      // duplicates of an edge add up
        float de = al * (my_dalpha - D);
        float dz = de * (zraw > 0.f ? 1.f : 0.2f);
        gad += dz;
        if (staged) atomicAdd(&gas[j - n0], dz);
        else atomicAdd(&g_a_src[(int64_t)j * H + h], dz);
      }
    }
    gad = warp_sum(gad);
    if (lane == 0) g_a_dst[(int64_t)(n0 + i) * H + h] = gad;
  }
  __syncthreads();
  if (dense) {
    for (int j = warp; j < n; j += NT / 32) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int i = 0; i < n; ++i) {
        const float a = Aij[i * n + j];
        if (a != 0.f) {                                          // warp-uniform: sparse graphs skip most rows
          for (int q = 0; q < nq && q < 2; ++q) {
            const int c = q * 128 + lane * 4;
            if (c < C) {
              const float4 gv = *reinterpret_cast<const float4*>(gtile + (int64_t)i * C + c);
              acc[q][0] = fmaf(a, gv.x, acc[q][0]);
              acc[q][1] = fmaf(a, gv.y, acc[q][1]);
              acc[q][2] = fmaf(a, gv.z, acc[q][2]);
              acc[q][3] = fmaf(a, gv.w, acc[q][3]);
            }
          }
        }
      }
      for (int q = 0; q < nq && q < 2; ++q) {
        const int c = q * 128 + lane * 4;
        if (c < C) *reinterpret_cast<float4*>(gdst + (int64_t)j * C + c) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
      }
    }
    for (int i = tid; i < n; i += NT) g_a_src[(int64_t)(n0 + i) * H + h] = gas[i];
  } else if (staged) {
    for (int i = tid; i < n * C; i += NT) gdst[i] = gtile[i];
    for (int i = tid; i < n; i += NT) g_a_src[(int64_t)(n0 + i) * H + h] = gas[i];
  }
  for (int i = tid; i < C; i += NT) atomicAdd(&g_bias[h * C + i], gbias[i]);
}

// g_xs += g_as * att_src + g_ad * att_dst (in place);  g_att_* += sum_n g_a* xs
constexpr int kFoldRows = 32;
__global__ void __launch_bounds__(NT)
gat_fold_kernel(const float* __restrict__ xs, float* __restrict__ g_xs, const float* __restrict__ g_as,
                const float* __restrict__ g_ad, const float* __restrict__ att_src,
                const float* __restrict__ att_dst, int64_t N, int H, int C, float* __restrict__ g_att_src,
                float* __restrict__ g_att_dst) {
  // grid (ceil(N/kFoldRows), H); thread -> channel c = tid % C, row group tid / C (all 256 threads busy at C = 128);
  // four independent rows in flight per thread
  const int h = blockIdx.y;
  const int64_t n0 = (int64_t)blockIdx.x * kFoldRows;
  const int groups = NT / C > 0 ? NT / C : 1;
  const int rg = threadIdx.x / C;
  if (rg >= groups) return;
  for (int c = threadIdx.x % C; c < C; c += NT) {
    const float ws = att_src[h * C + c], wd = att_dst[h * C + c];
    float as = 0.f, adv = 0.f;
    for (int r0 = rg; r0 < kFoldRows; r0 += 4 * groups) {
      float gs[4], gd[4], x[4], gx[4];
      int64_t o[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t n = n0 + r0 + u * groups;
        ok[u] = (r0 + u * groups < kFoldRows) && n < N;
        o[u] = ((int64_t)h * N + (ok[u] ? n : 0)) * C + c;
        gs[u] = ok[u] ? g_as[n * H + h] : 0.f;
        gd[u] = ok[u] ? g_ad[n * H + h] : 0.f;
        x[u] = ok[u] ? xs[o[u]] : 0.f;
        gx[u] = ok[u] ? g_xs[o[u]] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ok[u]) g_xs[o[u]] = gx[u] + (gs[u] * ws + gd[u] * wd);
        as = fmaf(gs[u], x[u], as);
        adv = fmaf(gd[u], x[u], adv);
      }
    }
    atomicAdd(&g_att_src[h * C + c], as);
    atomicAdd(&g_att_dst[h * C + c], adv);
  }
}

__global__ void __launch_bounds__(256)
cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = (float)in[i];
}

}  // namespace
}  // namespace sga

extern "C" int sga_cast_f64_f32(const double* in, float* out, int64_t n, void* stream) {
  if (n <= 0) return SGA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  sga::cast_f64_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_gat_aggregate_bwd(const float* xs, const float* a_src, const float* a_dst,
                                     const int32_t* row_beg, const int32_t* row_cnt, const int32_t* col,
                                     const int32_t* node_off, int G, int max_graph_nodes, int64_t N, int H,
                                     int C, int apply_elu, const float* out, const float* grad_out, float* g_xs,
                                     float* g_a_src, float* g_a_dst, float* g_bias, void* stream) {
  if (N <= 0 || G <= 0) return SGA_OK;
  SGA_REQUIRE(C % 4 == 0 && C <= 256, "sga_gat_aggregate_bwd: C=%d must be a multiple of 4 and <= 256", C);
  const size_t cap = 200 * 1024;
  size_t need = ((size_t)2 * max_graph_nodes * C + max_graph_nodes + C) * sizeof(float);
  int smem_nodes = 0, dense = 0;
  size_t smem = (size_t)C * sizeof(float);
  if (need <= cap) {
    smem_nodes = max_graph_nodes;
    smem = need;
    const size_t need_dense = need + (size_t)max_graph_nodes * max_graph_nodes * sizeof(float);
    if (need_dense <= cap) {
      dense = 1;
      smem = need_dense;
    }
  }
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SGA_CUDA(cudaFuncSetAttribute(sga::gat_aggregate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  dim3 grid(G, H);
  sga::gat_aggregate_bwd_kernel<<<grid, sga::NT, smem, (cudaStream_t)stream>>>(xs, a_src, a_dst, row_beg, row_cnt, col, node_off, N, H, C,
                                                                               apply_elu, out, grad_out, g_xs, g_a_src, g_a_dst, g_bias,
                                                                               smem_nodes, dense);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_gat_linear_bwd(const float* x, int64_t N, int in_dim, const float* W, const float* att_src,
                                  const float* att_dst, int H, int C, const float* xs, float* g_xs,
                                  const float* g_a_src, const float* g_a_dst, float* gW, float* g_att_src,
                                  float* g_att_dst, float* gx, void* stream) {
  if (N <= 0) return SGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((N + sga::kFoldRows - 1) / sga::kFoldRows), H);
  sga::gat_fold_kernel<<<grid, sga::NT, 0, st>>>(xs, g_xs, g_a_src, g_a_dst, att_src, att_dst, N, H, C, g_att_src, g_att_dst);
  SGA_LAUNCH_CHECK();
  for (int h = 0; h < H; ++h) {
    const float* gh = g_xs + (int64_t)h * N * C;
    if (sga::gemm_tc_worth(C, in_dim)) {       // layer 1 (256 -> 2 x 128): tcgen05 GEMMs; layer 0 (in_dim = 3) stays on FMA
      // gW[h*C + c][k] += sum_n gh[n][c] x[n][k]
      int rc = sga::launch_gemm_tc_dense(gh, C, 1, x, in_dim, 1, C, in_dim, (int)N, gW + (int64_t)h * C * in_dim, in_dim, 1, st);
      if (rc != SGA_OK) return rc;
      // gx[n][k] (+)= sum_c gh[n][c] W[h*C + c][k]   (head 0 stores, the following launches add)
      if (gx) {
        rc = sga::launch_gemm_tc_dense(gh, C, 0, W + (int64_t)h * C * in_dim, in_dim, 1, (int)N, in_dim, C, gx, in_dim, h > 0, st);
        if (rc != SGA_OK) return rc;
      }
      continue;
    }
    // gW[h*C + c][k] += sum_n gh[n][c] x[n][k]
    SGA_CUDA(sga::launch_gemm(gh, 1, C, x, in_dim, 1, gW + (int64_t)h * C * in_dim, in_dim, C, in_dim, (int)N, 1, st,
                              sga::splitk_for(C, in_dim, (int)N)));
    // gx[n][k] (+)= sum_c gh[n][c] W[h*C + c][k]
    if (gx) SGA_CUDA(sga::launch_gemm(gh, C, 1, W + (int64_t)h * C * in_dim, in_dim, 1, gx, in_dim, (int)N, in_dim, C, h > 0, st));
  }
  return SGA_OK;
}
