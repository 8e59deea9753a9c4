// Per-modality projection (nn.Linear) fused with its slice of the modality fusion.
// Reference: src/aligner/sg_aligner.py:112-122 (object_embedding / structure_embedding /
// meta_embedding_rel / meta_embedding_attr) and :30-35 (MultiModalFusion: softmax over the
// modality weights, F.normalize(eps=1e-12), scale, concatenate).
#include "common.cuh"
#include "linear.cuh"

namespace sga {
namespace {

constexpr int NT = 256;
constexpr int PN = 32;   // nodes per CTA
constexpr int PK = 32;   // k chunk
constexpr int MAXO = 128;

__device__ __forceinline__ float softmax_weight(const float* __restrict__ fw, int M, int m) {
  float mx = -INFINITY;
  for (int i = 0; i < M; ++i) mx = fmaxf(mx, fw[i]);
  float s = 0.f;
  for (int i = 0; i < M; ++i) s += expf(fw[i] - mx);
  return expf(fw[m] - mx) / s;
}

__global__ void __launch_bounds__(NT)
project_fuse_fwd_kernel(const void* __restrict__ x, int x_is_f64, int64_t N, int in_dim,
                        const float* __restrict__ W, const float* __restrict__ b, int out_dim,
                        float* __restrict__ emb, float* __restrict__ joint, int joint_ld, int joint_col,
                        const float* __restrict__ fusion_w, int M, int m) {
  __shared__ float xt[PN][PK + 1];
  __shared__ float wt[PK][MAXO + 1];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * PN;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = tx + 32 * j;
    float bv = (c < out_dim) ? b[c] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][j] = bv;
  }
  for (int k0 = 0; k0 < in_dim; k0 += PK) {
    __syncthreads();
    for (int i = tid; i < PN * PK; i += NT) {
      int r = i / PK, k = i % PK;
      int64_t n = n0 + r;
      xt[r][k] = (n < N && k0 + k < in_dim) ? load_as_float<float>(x, n * in_dim + k0 + k, x_is_f64) : 0.f;
    }
    for (int i = tid; i < MAXO * PK; i += NT) {
      int c = i / PK, k = i % PK;
      wt[k][c] = (c < out_dim && k0 + k < in_dim) ? W[(int64_t)c * in_dim + k0 + k] : 0.f;
    }
    __syncthreads();
    const int kmax = min(PK, in_dim - k0);
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
  const float wm = joint ? softmax_weight(fusion_w, M, m) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t n = n0 + ty * 4 + i;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = tx + 32 * j;
      if (c < out_dim) ss = fmaf(acc[i][j], acc[i][j], ss);
    }
    ss = warp_sum(ss);
    if (n < N) {
      const float scale = wm / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = tx + 32 * j;
        if (c < out_dim) {
          emb[n * out_dim + c] = acc[i][j];
          if (joint) joint[n * joint_ld + joint_col + c] = acc[i][j] * scale;
        }
      }
    }
  }
}

// All M modalities in ONE launch on the shared double-buffered tile (linear.cuh): grid (ceil(N/32), M);
// blockIdx.y = modality, one CTA = 32 nodes x the (<= 128) output columns of that modality.
constexpr int kMaxModal = 8;
struct ProjectMulti {
  const void* x[kMaxModal];
  const float* W[kMaxModal];
  const float* b[kMaxModal];
  float* emb[kMaxModal];
  int x_is_f64[kMaxModal];
  int in_dim[kMaxModal];
};

__global__ void __launch_bounds__(linear::NT)
project_fuse_fwd_multi_kernel(const ProjectMulti P, int M, int64_t N, int out_dim, float* __restrict__ joint, int joint_ld,
                              const float* __restrict__ fusion_w) {
  __shared__ linear::Smem sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m = blockIdx.y;
  const int64_t n0 = (int64_t)blockIdx.x * linear::BM;
  float acc[2][8];
  linear::tile_mma(sm, P.x[m], P.x_is_f64[m], N, P.in_dim[m], P.W[m], out_dim, n0, 0, acc);
  const float* __restrict__ b = P.b[m];
  float* __restrict__ emb = P.emb[m];
  const float wm = joint ? softmax_weight(fusion_w, M, m) : 0.f;
  const bool vec4 = (out_dim % 4 == 0) && (joint == nullptr || joint_ld % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(joint)) & 15) == 0;
  float ss[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = linear::col_of(tx, j);
    const float bv = (c < out_dim) ? b[c] : 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      acc[i][j] += bv;
      if (c < out_dim) ss[i] = fmaf(acc[i][j], acc[i][j], ss[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float tot = linear::row_sum16(ss[i]);
    const int64_t n = n0 + ty * 2 + i;
    if (n < N) {
      const float scale = wm / fmaxf(sqrtf(tot), 1e-12f);
      if (vec4) {       // out_dim % 4 == 0: a thread's 4-column groups are entirely in or out, rows 16-byte aligned
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = linear::col_of(tx, 4 * g);
          if (c < out_dim) {
            const float* a = &acc[i][4 * g];
            *reinterpret_cast<float4*>(emb + n * out_dim + c) = make_float4(a[0], a[1], a[2], a[3]);
            if (joint)
              *reinterpret_cast<float4*>(joint + n * joint_ld + m * out_dim + c) =
                  make_float4(a[0] * scale, a[1] * scale, a[2] * scale, a[3] * scale);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = linear::col_of(tx, j);
          if (c < out_dim) {
            emb[n * out_dim + c] = acc[i][j];
            if (joint) joint[n * joint_ld + m * out_dim + c] = acc[i][j] * scale;
          }
        }
      }
    }
  }
}

}  // namespace
}  // namespace sga

extern "C" int sga_project_fuse_fwd_multi(const void* const* x_host, const int* x_is_f64_host, const int* in_dim_host,
                                          const float* const* W_host, const float* const* b_host, float* const* emb_host, int M,
                                          int64_t N, int out_dim, float* joint, int joint_ld, const float* fusion_w, void* stream) {
  if (N <= 0 || M <= 0) return SGA_OK;
  SGA_REQUIRE(M <= sga::kMaxModal, "sga_project_fuse_fwd_multi: M=%d must be <= %d", M, sga::kMaxModal);
  SGA_REQUIRE(out_dim > 0 && out_dim <= sga::linear::BN, "sga_project_fuse_fwd_multi: out_dim=%d must be in 1..%d", out_dim, sga::linear::BN);
  SGA_REQUIRE(joint == nullptr || (fusion_w != nullptr && joint_ld >= M * out_dim), "sga_project_fuse_fwd_multi: bad joint arguments");
  sga::ProjectMulti P;
  for (int m = 0; m < M; ++m) {
    SGA_REQUIRE(x_host[m] && W_host[m] && b_host[m] && emb_host[m] && in_dim_host[m] > 0, "sga_project_fuse_fwd_multi: bad modality %d", m);
    P.x[m] = x_host[m]; P.W[m] = W_host[m]; P.b[m] = b_host[m]; P.emb[m] = emb_host[m];
    P.x_is_f64[m] = x_is_f64_host[m]; P.in_dim[m] = in_dim_host[m];
  }
  dim3 grid((unsigned)((N + sga::linear::BM - 1) / sga::linear::BM), M);
  sga::project_fuse_fwd_multi_kernel<<<grid, sga::linear::NT, 0, (cudaStream_t)stream>>>(P, M, N, out_dim, joint, joint_ld, fusion_w);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

extern "C" int sga_project_fuse_fwd(const void* x, int x_is_f64, int64_t N, int in_dim, const float* W,
                                    const float* b, int out_dim, float* emb, float* joint, int joint_ld,
                                    int joint_col, const float* fusion_w, int M, int m, void* stream) {
  if (N <= 0) return SGA_OK;
  SGA_REQUIRE(out_dim > 0 && out_dim <= sga::MAXO, "sga_project_fuse_fwd: out_dim=%d must be in 1..%d", out_dim, sga::MAXO);
  SGA_REQUIRE(in_dim > 0, "sga_project_fuse_fwd: in_dim=%d", in_dim);
  SGA_REQUIRE(joint == nullptr || (fusion_w != nullptr && M > 0 && M <= 16 && m >= 0 && m < M), "sga_project_fuse_fwd: bad fusion args M=%d m=%d", M, m);
  unsigned grid = (unsigned)((N + sga::PN - 1) / sga::PN);
  sga::project_fuse_fwd_kernel<<<grid, sga::NT, 0, (cudaStream_t)stream>>>(x, x_is_f64, N, in_dim, W, b, out_dim, emb, joint, joint_ld,
                                                                           joint_col, fusion_w, M, m);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
