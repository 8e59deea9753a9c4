// OverallLoss forward + analytic gradient w.r.t. every embedding (src/aligner/losses.py:5-152).
//
// Non-redundant formulation.  For an embedding X (a modality or the joint), Xh = F.normalize(X),
//   P1 = Xh[e1i], P2 = Xh[e2i], Q1 = Xh[e1j], Q2 = Xh[e2j],  T = A + J1 + J2
//   F1 = P1 [P2;Q1;Q2]^T   (A x T):  F1[:, :A] = G,    F1[:, A:A+J1] = U11, F1[:, A+J1:] = U12
//   F2 = P2 [P1;Q2;Q1]^T   (A x T):  F2[:, :A] = G^T,  F2[:, A:A+J2] = U22, F2[:, A+J2:] = U21
// calculate_prob_dist(e1i,e2i,e1j,e2j,t)[a,b] only needs F1[a,b] and the two SCALARS
// sum(exp(U11/t)), sum(exp(U12/t)) (losses.py:10-11: .sum() over the whole matrix); the swapped
// call (losses.py:52) needs F2[a,b], sum(exp(U22/t)), sum(exp(U21/t)).  So all 78 matmuls of the
// reference collapse to two GEMMs per embedding, shared by both temperatures (ICL 0.1 / IAL 1.0)
// and by all M IAL terms that reuse the joint embedding.  The backward is 4 more GEMMs per
// embedding on the in-place gradient of F1/F2.
//
// The GEMMs run on the tensor cores (gemm_tc.cu, tcgen05 tf32x3): L2-normalisation and the index
// gathers are fused into the operand loader, the normaliser sums into the forward epilogue, and the
// backward GEMMs scatter-add straight into d(normalised embedding); F1/F2 are materialised in HBM
// (A x T fp32 each) because the element-wise loss terms need the global normalisers first.
#include "common.cuh"

namespace sga {

struct GemmOperand {
  const float* p;
  int64_t ld;
  const int32_t* idx;
  const float* div;
  int mn_major;
};
struct GemmParams {
  GemmOperand A, B;
  int M, N, K;
  float* C;
  int64_t ldc;
  int mode;
  const int32_t* c_idx;
  int es_c0, es_split;
  double* s01_lo; double* s01_hi; double* s1_lo; double* s1_hi;
  int ksplit;
};
int launch_gemm_tc(const GemmParams& P, cudaStream_t st);   // gemm_tc.cu

namespace {

constexpr int NT = 256;
constexpr float kEps = 1e-9f;

struct Layout {
  // byte offsets into the workspace
  size_t scal;                // doubles: S[n_emb][2][4], dS[n_emb][2][4], icl_raw[n_emb], ial_raw[n_emb]
  size_t ridx, r2idx;         // int32 [T]: rows [e2i;e1j;e2j] and [e1i;e2j;e1j]
  size_t norms[17], den[17], F1[17], F2[17], dXh[17];
  size_t acc1, acc2;
  size_t total;
  int ldF;
};

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

Layout make_layout(int n_emb, const int* dims, int64_t N, int A, int J1, int J2, int want_grad) {
  Layout L;
  memset(&L, 0, sizeof(L));
  size_t o = 0;
  const size_t T = (size_t)A + J1 + J2;
  L.ldF = (int)((T + 3) & ~(size_t)3);     // 16-byte aligned rows for the vectorised operand loads
  L.scal = o;
  o = al256(o + sizeof(double) * (size_t)n_emb * 18);
  L.ridx = o; o = al256(o + 4 * T);
  L.r2idx = o; o = al256(o + 4 * T);
  for (int x = 0; x < n_emb; ++x) {
    const size_t d = dims[x];
    L.norms[x] = o; o = al256(o + 4 * (size_t)N);
    L.den[x] = o; o = al256(o + 4 * (size_t)N);
    L.F1[x] = o; o = al256(o + 4 * (size_t)A * L.ldF);
    L.F2[x] = o; o = al256(o + 4 * (size_t)A * L.ldF);
    if (want_grad) { L.dXh[x] = o; o = al256(o + 4 * (size_t)N * d); }
  }
  if (want_grad && n_emb > 1) {
    L.acc1 = o; o = al256(o + 4 * (size_t)A * A);
    L.acc2 = o; o = al256(o + 4 * (size_t)A * A);
  }
  L.total = o;
  return L;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
row_norm_kernel(const float* __restrict__ X, int64_t N, int D, float* __restrict__ norms, float* __restrict__ den) {
  int64_t row = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (row >= N) return;
  int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    float v = X[row * D + k];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    norms[row] = sqrtf(s);
    den[row] = fmaxf(sqrtf(s), 1e-12f);    // F.normalize: x / max(||x||, eps)
  }
}

// ridx = [e2i; e1j; e2j], r2idx = [e1i; e2j; e1j]
__global__ void __launch_bounds__(NT)
build_ridx_kernel(const int32_t* __restrict__ e1i, const int32_t* __restrict__ e2i, const int32_t* __restrict__ e1j,
                  const int32_t* __restrict__ e2j, int A, int J1, int J2, int32_t* __restrict__ ridx,
                  int32_t* __restrict__ r2idx) {
  const int T = A + J1 + J2;
  for (int r = blockIdx.x * NT + threadIdx.x; r < T; r += gridDim.x * NT) {
    ridx[r] = r < A ? e2i[r] : (r < A + J1 ? e1j[r - A] : e2j[r - A - J1]);
    r2idx[r] = r < A ? e1i[r] : (r < A + J2 ? e2j[r - A] : e1j[r - A - J2]);
  }
}

template <int NV>
__device__ __forceinline__ void block_reduce_add(float (&v)[NV], double* const* dst) {
  __shared__ float red[NT / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = warp_sum(v[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += (double)red[w][threadIdx.x];
    if (dst[threadIdx.x]) atomicAdd(dst[threadIdx.x], s);
  }
}

struct QD { float q, dg, dsa, dsb; };

// calculate_prob_dist element (losses.py:5-15) and its partial derivatives
__device__ __forceinline__ QD cpd(float g, float tau, float sa, float sb) {
  QD r;
  const float mx = expf(g / tau);
  const float r1 = mx / sa, r2 = mx / sb;
  const float u1 = r1 + kEps, u2 = r2 + kEps;
  const float inv = 1.f + 1.f / u1 + 1.f / u2;
  r.q = 1.f / (inv + kEps);
  const float q2 = r.q * r.q;
  const float t1 = q2 / (u1 * u1) * r1, t2 = q2 / (u2 * u2) * r2;
  r.dg = (t1 + t2) / tau;
  r.dsa = -t1 / sa;
  r.dsb = -t2 / sb;
  return r;
}

struct PairArgs {
  float* F1m; float* F2m;             // this embedding (gradient written in place when want_grad)
  const float* F1j; const float* F2j; // joint (modal mode only)
  float* acc1; float* acc2;           // joint gradient accumulators [A,A] (may be null)
  int A, T;
  const double* Sm; const double* Sj; // [2][4]: tau index (0: 0.1, 1: 1.0) x {S11,S12,S22,S21}
  double* dSm; double* dSj;
  double* icl_raw; double* ial_raw;
  const float* lv_icl; const float* lv_ial;   // this modality's log_vars (null -> coefficient 1)
  float zoom;
  int modal;      // 1: modal embedding with IAL against the joint; 0: joint (or single-modality) ICL only
  int want_grad;
};

__global__ void __launch_bounds__(NT)
pair_kernel(PairArgs p) {
  const int A = p.A, T = p.T;
  const float invA2 = 1.f / ((float)A * (float)A);
  const float c_icl = (p.lv_icl ? expf(-p.lv_icl[0]) : 1.f) * invA2;
  const float c_ial = p.modal ? p.zoom * expf(-p.lv_ial[0]) * 0.1f * 0.5f : 0.f;
  float sm[2][4], sj[4];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int s = 0; s < 4; ++s) sm[t][s] = (float)p.Sm[t * 4 + s] + kEps;
#pragma unroll
  for (int s = 0; s < 4; ++s) sj[s] = p.modal ? (float)p.Sj[4 + s] + kEps : 1.f;
  // accumulators: 0 icl, 1 ial, 2..5 dSm[0.1], 6..9 dSm[1.0], 10..13 dSj[1.0]
  float acc[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) acc[i] = 0.f;
  const int64_t total = (int64_t)A * A;
  for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * NT) {
    const int64_t a = t / A, b = t % A;
    const int64_t off = a * T + b;
    const float g1 = p.F1m[off], g2 = p.F2m[off];
    // ---- ICL, tau = 0.1 (losses.py:43-58)
    QD q12 = cpd(g1, 0.1f, sm[0][0], sm[0][1]);
    QD q21 = cpd(g2, 0.1f, sm[0][2], sm[0][3]);
    const float w = 0.5f * q12.q + 0.5f * q21.q;
    acc[0] += -logf(w);
    float d1 = 0.f, d2 = 0.f;
    if (p.want_grad) {
      const float dq = -0.5f / w * c_icl;
      d1 = dq * q12.dg;
      d2 = dq * q21.dg;
      acc[2] += dq * q12.dsa; acc[3] += dq * q12.dsb;
      acc[4] += dq * q21.dsa; acc[5] += dq * q21.dsb;
    }
    if (p.modal) {
      // ---- IAL, tau = 1.0 (losses.py:68-97): target = modal q, input = log(joint q)
      QD o12 = cpd(g1, 1.f, sm[1][0], sm[1][1]);
      QD o21 = cpd(g2, 1.f, sm[1][2], sm[1][3]);
      const float j1 = p.F1j[off], j2 = p.F2j[off];
      QD m12 = cpd(j1, 1.f, sj[0], sj[1]);
      QD m21 = cpd(j2, 1.f, sj[2], sj[3]);
      const float ea = expf(o12.q), eb = expf(o21.q);
      const float la = o12.q - logf(m12.q), lb = o21.q - logf(m21.q);
      acc[1] += 0.5f * (ea * la + eb * lb);
      if (p.want_grad) {
        const float dqo12 = c_ial * ea * (la + 1.f), dqo21 = c_ial * eb * (lb + 1.f);
        const float dqm12 = -c_ial * ea / m12.q, dqm21 = -c_ial * eb / m21.q;
        d1 += dqo12 * o12.dg;
        d2 += dqo21 * o21.dg;
        acc[6] += dqo12 * o12.dsa; acc[7] += dqo12 * o12.dsb;
        acc[8] += dqo21 * o21.dsa; acc[9] += dqo21 * o21.dsb;
        acc[10] += dqm12 * m12.dsa; acc[11] += dqm12 * m12.dsb;
        acc[12] += dqm21 * m21.dsa; acc[13] += dqm21 * m21.dsb;
        p.acc1[a * A + b] += dqm12 * m12.dg;
        p.acc2[a * A + b] += dqm21 * m21.dg;
      }
    } else if (p.want_grad && p.acc1) {
      d1 += p.acc1[a * A + b];
      d2 += p.acc2[a * A + b];
    }
    if (p.want_grad) {
      p.F1m[off] = d1;
      p.F2m[off] = d2;
    }
  }
  __shared__ double* dst[14];
  if (threadIdx.x == 0) {
    dst[0] = p.icl_raw;
    dst[1] = p.modal ? p.ial_raw : nullptr;
    for (int s = 0; s < 4; ++s) {
      dst[2 + s] = p.want_grad ? p.dSm + s : nullptr;
      dst[6 + s] = (p.want_grad && p.modal) ? p.dSm + 4 + s : nullptr;
      dst[10 + s] = (p.want_grad && p.modal) ? p.dSj + 4 + s : nullptr;
    }
  }
  __syncthreads();
  block_reduce_add<14>(acc, dst);
}

// in place over the U blocks of F1 / F2: u -> sum_tau dS[tau][seg] / tau * exp(u / tau)
__global__ void __launch_bounds__(NT)
coef_kernel(float* __restrict__ F, int A, int T, int ld, int c_split, const double* __restrict__ dS, int seg_lo, int seg_hi) {
  const int64_t w = T - A, total = (int64_t)A * w;
  const float lo01 = (float)dS[seg_lo] * 10.f, lo1 = (float)dS[4 + seg_lo];
  const float hi01 = (float)dS[seg_hi] * 10.f, hi1 = (float)dS[4 + seg_hi];
  for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * NT) {
    int64_t a = t / w, c = A + t % w;
    float u = F[a * ld + c];
    float e01 = expf(u / 0.1f), e1 = expf(u);
    F[a * ld + c] = (c < c_split) ? lo01 * e01 + lo1 * e1 : hi01 * e01 + hi1 * e1;
  }
}

// backward of F.normalize: dX = (dXh - Xh <Xh, dXh>) / max(||X||, eps)
__global__ void __launch_bounds__(NT)
normalize_bwd_kernel(const float* __restrict__ X, const float* __restrict__ norms, const float* __restrict__ dXh,
                     int64_t N, int D, float* __restrict__ dX) {
  int64_t row = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (row >= N) return;
  int lane = threadIdx.x & 31;
  const float nrm = norms[row];
  const float den = fmaxf(nrm, 1e-12f);
  float dot = 0.f;
  for (int k = lane; k < D; k += 32) dot = fmaf(X[row * D + k] / den, dXh[row * D + k], dot);
  dot = warp_sum(dot);
  const bool clamped = nrm < 1e-12f;   // below eps the denominator is a constant
  for (int k = lane; k < D; k += 32) {
    float xh = X[row * D + k] / den;
    dX[row * D + k] = (dXh[row * D + k] - (clamped ? 0.f : xh * dot)) / den;
  }
}

// losses_out = {loss, icl_unimodal, icl_multimodal, ial}; gradients of the log_vars
__global__ void finalize_kernel(const double* __restrict__ icl_raw, const double* __restrict__ ial_raw, int M,
                                int n_emb, int A, const float* __restrict__ lv_ial,
                                const float* __restrict__ lv_icl, float zoom, float* __restrict__ losses_out,
                                float* __restrict__ g_lv_ial, float* __restrict__ g_lv_icl) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double a2 = (double)A * (double)A;
  if (n_emb == 1) {
    float icl = (float)(icl_raw[0] / a2);
    losses_out[0] = icl; losses_out[1] = icl; losses_out[2] = 0.f; losses_out[3] = 0.f;
    return;
  }
  float ial_tot = 0.f, icl_uni = 0.f;
  for (int m = 0; m < M; ++m) {
    float ial = 0.1f * (float)ial_raw[m];
    float icl = (float)(icl_raw[m] / a2);
    float pa = expf(-lv_ial[m]), pc = expf(-lv_icl[m]);
    ial_tot += pa * ial + lv_ial[m];
    icl_uni += pc * icl + lv_icl[m];
    if (g_lv_ial) g_lv_ial[m] = zoom * (1.f - pa * ial);
    if (g_lv_icl) g_lv_icl[m] = 1.f - pc * icl;
  }
  ial_tot *= zoom;
  float icl_multi = (float)(icl_raw[M] / a2);
  losses_out[0] = ial_tot + icl_uni + icl_multi;
  losses_out[1] = icl_uni;
  losses_out[2] = icl_multi;
  losses_out[3] = ial_tot;
}

inline unsigned grid_for(int64_t total) {
  int64_t b = (total + NT - 1) / NT;
  int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace sga

extern "C" size_t sga_loss_workspace_bytes(int n_emb, const int* dims_host, int64_t N, int A, int J1, int J2, int want_grad) {
  if (n_emb < 1 || n_emb > 16) return 0;
  return sga::make_layout(n_emb, dims_host, N, A, J1, J2, want_grad).total;
}

extern "C" int sga_loss_fwd_bwd(const float* const* embs_host, const int* dims_host, int n_emb, int64_t N,
                                const int32_t* e1i, const int32_t* e2i, const int32_t* e1j, const int32_t* e2j,
                                int A, int J1, int J2, const float* log_vars_ial, const float* log_vars_icl,
                                float zoom, float* losses_out, int want_grad, float* const* g_embs_host,
                                float* g_log_vars_ial, float* g_log_vars_icl, void* workspace,
                                size_t workspace_bytes, void* stream) {
  using namespace sga;
  SGA_REQUIRE(n_emb >= 1 && n_emb <= 16, "sga_loss_fwd_bwd: n_emb=%d out of range", n_emb);
  SGA_REQUIRE(A > 0, "sga_loss_fwd_bwd: the batch has no anchors (A=%d); the reference loss is NaN there", A);
  SGA_REQUIRE(J1 >= 0 && J2 >= 0 && N > 0, "sga_loss_fwd_bwd: bad sizes");
  const int M = n_emb == 1 ? 1 : n_emb - 1;
  SGA_REQUIRE(n_emb == 1 || (log_vars_ial && log_vars_icl), "sga_loss_fwd_bwd: log_vars required when M > 1");
  Layout L = make_layout(n_emb, dims_host, N, A, J1, J2, want_grad);
  if (workspace_bytes < L.total) {
    set_error("sga_loss_fwd_bwd: workspace %zu < %zu bytes", workspace_bytes, L.total);
    return SGA_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  const int T = A + J1 + J2;
  const int ldF = L.ldF;
  double* scal = (double*)(ws + L.scal);
  double* S = scal;                       // [n_emb][2][4]
  double* dS = scal + (size_t)n_emb * 8;  // [n_emb][2][4]
  double* icl_raw = scal + (size_t)n_emb * 16;
  double* ial_raw = icl_raw + n_emb;
  SGA_CUDA(cudaMemsetAsync(scal, 0, sizeof(double) * (size_t)n_emb * 18, st));
  auto F = [&](size_t off) { return (float*)(ws + off); };
  int32_t* ridx = (int32_t*)(ws + L.ridx);
  int32_t* r2idx = (int32_t*)(ws + L.r2idx);
  build_ridx_kernel<<<(T + NT - 1) / NT, NT, 0, st>>>(e1i, e2i, e1j, e2j, A, J1, J2, ridx, r2idx);
  SGA_LAUNCH_CHECK();

  // ---- forward Grams on the tensor cores; normalise + gather in the loader, exp-sums in the epilogue
  for (int x = 0; x < n_emb; ++x) {
    const int d = dims_host[x];
    const float* X = embs_host[x];
    row_norm_kernel<<<(unsigned)((N + 7) / 8), NT, 0, st>>>(X, N, d, F(L.norms[x]), F(L.den[x]));
    SGA_LAUNCH_CHECK();
    double* Sx = S + (size_t)x * 8;
    for (int dir = 0; dir < 2; ++dir) {
      GemmParams P;
      memset(&P, 0, sizeof(P));
      P.A = {X, d, dir == 0 ? e1i : e2i, F(L.den[x]), 0};
      P.B = {X, d, dir == 0 ? ridx : r2idx, F(L.den[x]), 0};
      P.M = A; P.N = T; P.K = d;
      P.C = F(dir == 0 ? L.F1[x] : L.F2[x]); P.ldc = ldF;
      P.mode = 1;
      P.es_c0 = A;
      P.es_split = A + (dir == 0 ? J1 : J2);
      const int lo = dir == 0 ? 0 : 2, hi = dir == 0 ? 1 : 3;   // {S11,S12} / {S22,S21}
      P.s01_lo = Sx + lo; P.s01_hi = Sx + hi; P.s1_lo = Sx + 4 + lo; P.s1_hi = Sx + 4 + hi;
      int rc = launch_gemm_tc(P, st);
      if (rc != SGA_OK) return rc;
    }
  }
  // ---- element-wise loss terms (+ in-place gradient of the G blocks)
  const int xj = n_emb - 1;   // joint (or the single modality)
  if (want_grad && n_emb > 1) {
    SGA_CUDA(cudaMemsetAsync(ws + L.acc1, 0, 4 * (size_t)A * A, st));
    SGA_CUDA(cudaMemsetAsync(ws + L.acc2, 0, 4 * (size_t)A * A, st));
  }
  for (int x = 0; x < n_emb; ++x) {
    PairArgs p;
    memset(&p, 0, sizeof(p));
    const bool modal = (n_emb > 1 && x < xj);
    p.F1m = F(L.F1[x]); p.F2m = F(L.F2[x]);
    p.F1j = F(L.F1[xj]); p.F2j = F(L.F2[xj]);
    p.acc1 = (want_grad && n_emb > 1) ? F(L.acc1) : nullptr;
    p.acc2 = (want_grad && n_emb > 1) ? F(L.acc2) : nullptr;
    p.A = A; p.T = ldF;
    p.Sm = S + (size_t)x * 8; p.Sj = S + (size_t)xj * 8;
    p.dSm = dS + (size_t)x * 8; p.dSj = dS + (size_t)xj * 8;
    p.icl_raw = icl_raw + x; p.ial_raw = ial_raw + x;
    p.lv_icl = modal ? log_vars_icl + x : nullptr;
    p.lv_ial = modal ? log_vars_ial + x : nullptr;
    p.zoom = zoom;
    p.modal = modal ? 1 : 0;
    p.want_grad = want_grad;
    pair_kernel<<<grid_for((int64_t)A * A), NT, 0, st>>>(p);
    SGA_LAUNCH_CHECK();
  }
  finalize_kernel<<<1, 32, 0, st>>>(icl_raw, ial_raw, M, n_emb, A, log_vars_ial, log_vars_icl, zoom, losses_out,
                                    want_grad ? g_log_vars_ial : nullptr, want_grad ? g_log_vars_icl : nullptr);
  SGA_LAUNCH_CHECK();
  if (!want_grad) return SGA_OK;

  // ---- backward: coefficient blocks in place, then 4 scatter-add GEMMs per embedding:
  //   dXh[e1i]  += dF1   Xh[R]      dXh[R]  += dF1^T Xh[e1i]
  //   dXh[e2i]  += dF2   Xh[R']     dXh[R'] += dF2^T Xh[e2i]
  auto ksplit_for = [&](int Mr, int Nc, int Kc) {
    int tiles = ((Mr + 127) / 128) * ((Nc + 127) / 128);
    int chunks = (Kc + 31) / 32;
    int want = (2 * sm_count() + tiles - 1) / tiles;
    int cap = chunks / 4;
    if (want > cap) want = cap;
    return want < 1 ? 1 : want;
  };
  for (int x = 0; x < n_emb; ++x) {
    const int d = dims_host[x];
    const float* X = embs_host[x];
    double* dSx = dS + (size_t)x * 8;
    if (T > A) {
      coef_kernel<<<grid_for((int64_t)A * (T - A)), NT, 0, st>>>(F(L.F1[x]), A, T, ldF, A + J1, dSx, 0, 1);
      coef_kernel<<<grid_for((int64_t)A * (T - A)), NT, 0, st>>>(F(L.F2[x]), A, T, ldF, A + J2, dSx, 2, 3);
      SGA_LAUNCH_CHECK();
    }
    SGA_CUDA(cudaMemsetAsync(ws + L.dXh[x], 0, 4 * (size_t)N * d, st));
    for (int dir = 0; dir < 2; ++dir) {
      const float* Fd = F(dir == 0 ? L.F1[x] : L.F2[x]);
      const int32_t* rows_i = dir == 0 ? e1i : e2i;
      const int32_t* rows_r = dir == 0 ? ridx : r2idx;
      GemmParams P;
      memset(&P, 0, sizeof(P));
      // dXh[rows_i[a]] += sum_t Fd[a,t] Xh[rows_r[t]]
      P.A = {Fd, ldF, nullptr, nullptr, 0};
      P.B = {X, d, rows_r, F(L.den[x]), 1};
      P.M = A; P.N = d; P.K = T;
      P.C = F(L.dXh[x]); P.ldc = d; P.mode = 2; P.c_idx = rows_i;
      P.ksplit = ksplit_for(A, d, T);
      int rc = launch_gemm_tc(P, st);
      if (rc != SGA_OK) return rc;
      // dXh[rows_r[t]] += sum_a Fd[a,t] Xh[rows_i[a]]
      P.A = {Fd, ldF, nullptr, nullptr, 1};
      P.B = {X, d, rows_i, F(L.den[x]), 1};
      P.M = T; P.N = d; P.K = A;
      P.c_idx = rows_r;
      P.ksplit = ksplit_for(T, d, A);
      rc = launch_gemm_tc(P, st);
      if (rc != SGA_OK) return rc;
    }
    normalize_bwd_kernel<<<(unsigned)((N + 7) / 8), NT, 0, st>>>(embs_host[x], F(L.norms[x]), F(L.dXh[x]), N, d, g_embs_host[x]);
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}
