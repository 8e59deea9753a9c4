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
#include "gemm_tc.cuh"
#include "gram_ts.cuh"

namespace sga {

namespace {

constexpr int NT = 256;
constexpr float kEps = 1e-9f;

struct Layout {
  // byte offsets into the workspace
  size_t scal;                // doubles: S[n_emb][2][4], dS[n_emb][2][4], icl_raw[n_emb], ial_raw[n_emb]
  size_t ridx, r2idx;         // int32 [T]: rows [e2i;e1j;e2j] and [e1i;e2j;e1j]
  size_t norms[17], Xh[17], F1[17], F2[17], dXh[17];
  size_t slot;                // int32 [N]: (row set << 28) | position, for the packed operand images
  size_t img[17][4];          // operand images of the embeddings that take the gram_ts path (d <= 128), else 0
  size_t acc1, acc2;
  size_t total;
  int ldF;
};

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// Embeddings up to 512 wide take the TMA-fed Gram kernel (gram_ts.cu: A in tensor memory up to 128 wide, both operands
// streamed beyond that); SGA_LOSS_GRAM=legacy forces the in-loader-split GEMM (gemm_tc.cu) for everything (A/B
// comparisons, profiling).
int g_gram_legacy_override = 0;      // sga_loss_set_gram_path(): 1 = the caller's index sets are not a partition
inline bool gram_ts_ok(int d) {
  static int legacy = -1;
  if (legacy < 0) {
    const char* e = getenv("SGA_LOSS_GRAM");
    legacy = (e && strcmp(e, "legacy") == 0) ? 1 : 0;
  }
  return !legacy && !g_gram_legacy_override && d <= 512;
}

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
  L.slot = o; o = al256(o + 4 * (size_t)N);
  for (int x = 0; x < n_emb; ++x) {
    const size_t d = dims[x];
    L.norms[x] = o; o = al256(o + 4 * (size_t)N);
    L.Xh[x] = o; o = al256(o + 4 * (size_t)N * d);
    L.F1[x] = o; o = al256(o + 4 * (size_t)A * L.ldF);
    L.F2[x] = o; o = al256(o + 4 * (size_t)A * L.ldF);
    if (gram_ts_ok((int)d)) {
      const int rows[4] = {A, A, J1, J2};
      for (int s4 = 0; s4 < 4; ++s4) {
        o = (o + 1023) & ~(size_t)1023;      // bulk-copy sources: keep the images 1 KiB aligned
        L.img[x][s4] = o;
        o += gram_image_bytes(rows[s4], (int)d);
      }
      o = al256(o);
    }
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
// norms[row] = ||X[row]||, Xh[row] = F.normalize(X)[row] = X[row] / max(||X[row]||, 1e-12) (losses.py:44,69-70):
// the normalised rows are written ONCE here, so the GEMM operand loaders only gather, split and store.
__global__ void __launch_bounds__(NT)
row_norm_kernel(const float* __restrict__ X, int64_t N, int D, float* __restrict__ norms, float* __restrict__ Xh) {
  int64_t row = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (row >= N) return;
  int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    float v = X[row * D + k];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  const float nrm = sqrtf(s);
  const float den = fmaxf(nrm, 1e-12f);    // F.normalize: x / max(||x||, eps)
  if (lane == 0) norms[row] = nrm;
  for (int k = lane; k < D; k += 32) Xh[row * D + k] = X[row * D + k] / den;
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

// calculate_prob_dist element (losses.py:5-15) and its partial derivatives.  `isa`, `isb` = 1 / (S + 1e-9) (one exact
// division per launch); the per-element reciprocals use the SFU (MUFU.RCP / EX2, ~2 ulp): the element-wise pass is
// ALU bound (6 of these per modal element) and feeds sums checked at 1e-3.
__device__ __forceinline__ QD cpd(float g, float inv_tau, float isa, float isb) {
  QD r;
  const float mx = __expf(g * inv_tau);
  const float r1 = mx * isa, r2 = mx * isb;
  const float iu1 = __fdividef(1.f, r1 + kEps), iu2 = __fdividef(1.f, r2 + kEps);
  const float inv = 1.f + iu1 + iu2;
  r.q = __fdividef(1.f, inv + kEps);
  const float q2 = r.q * r.q;
  const float t1 = q2 * iu1 * iu1 * r1, t2 = q2 * iu2 * iu2 * r2;
  r.dg = (t1 + t2) * inv_tau;
  r.dsa = -t1 * isa;
  r.dsb = -t2 * isb;
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
  float ism[2][4], isj[4];     // reciprocals of the normalisers (+1e-9, losses.py:12-13)
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int s = 0; s < 4; ++s) ism[t][s] = 1.f / ((float)p.Sm[t * 4 + s] + kEps);
#pragma unroll
  for (int s = 0; s < 4; ++s) isj[s] = p.modal ? 1.f / ((float)p.Sj[4 + s] + kEps) : 1.f;
  // accumulators: 0 icl, 1 ial, 2..5 dSm[0.1], 6..9 dSm[1.0], 10..13 dSj[1.0]
  float acc[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) acc[i] = 0.f;
  for (int a = blockIdx.x; a < A; a += gridDim.x) {
   const int64_t rowF = (int64_t)a * T, rowA = (int64_t)a * A;
   for (int b = threadIdx.x; b < A; b += NT) {
    const int64_t off = rowF + b;
    const float g1 = p.F1m[off], g2 = p.F2m[off];
    // ---- ICL, tau = 0.1 (losses.py:43-58)
    QD q12 = cpd(g1, 10.f, ism[0][0], ism[0][1]);
    QD q21 = cpd(g2, 10.f, ism[0][2], ism[0][3]);
    const float w = 0.5f * q12.q + 0.5f * q21.q;
    acc[0] += -__logf(w);
    float d1 = 0.f, d2 = 0.f;
    if (p.want_grad) {
      const float dq = __fdividef(-0.5f * c_icl, w);
      d1 = dq * q12.dg;
      d2 = dq * q21.dg;
      acc[2] += dq * q12.dsa; acc[3] += dq * q12.dsb;
      acc[4] += dq * q21.dsa; acc[5] += dq * q21.dsb;
    }
    if (p.modal) {
      // ---- IAL, tau = 1.0 (losses.py:68-97): target = modal q, input = log(joint q)
      QD o12 = cpd(g1, 1.f, ism[1][0], ism[1][1]);
      QD o21 = cpd(g2, 1.f, ism[1][2], ism[1][3]);
      const float j1 = p.F1j[off], j2 = p.F2j[off];
      QD m12 = cpd(j1, 1.f, isj[0], isj[1]);
      QD m21 = cpd(j2, 1.f, isj[2], isj[3]);
      const float ea = __expf(o12.q), eb = __expf(o21.q);
      const float la = o12.q - __logf(m12.q), lb = o21.q - __logf(m21.q);
      acc[1] += 0.5f * (ea * la + eb * lb);
      if (p.want_grad) {
        const float dqo12 = c_ial * ea * (la + 1.f), dqo21 = c_ial * eb * (lb + 1.f);
        const float dqm12 = __fdividef(-c_ial * ea, m12.q), dqm21 = __fdividef(-c_ial * eb, m21.q);
        d1 += dqo12 * o12.dg;
        d2 += dqo21 * o21.dg;
        acc[6] += dqo12 * o12.dsa; acc[7] += dqo12 * o12.dsb;
        acc[8] += dqo21 * o21.dsa; acc[9] += dqo21 * o21.dsb;
        acc[10] += dqm12 * m12.dsa; acc[11] += dqm12 * m12.dsb;
        acc[12] += dqm21 * m21.dsa; acc[13] += dqm21 * m21.dsb;
        p.acc1[rowA + b] += dqm12 * m12.dg;
        p.acc2[rowA + b] += dqm21 * m21.dg;
      }
    } else if (p.want_grad && p.acc1) {
      d1 += p.acc1[rowA + b];
      d2 += p.acc2[rowA + b];
    }
    if (p.want_grad) {
      p.F1m[off] = d1;
      p.F2m[off] = d2;
    }
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

// All embeddings in ONE pass over the A x A anchor block (M <= kPairMaxModal modalities + the joint): the joint's
// similarities are read once and its tau = 1 probabilities computed once instead of once per modality, and the
// gradient the M IAL terms send into the joint stays in registers (no [A,A] accumulator arrays in HBM).
constexpr int kPairMaxModal = 4;
struct PairAllArgs {
  float* F1[kPairMaxModal + 1]; float* F2[kPairMaxModal + 1];   // modal embeddings, then the joint (or the single embedding)
  int M;                     // modal embeddings (0: a single embedding, ICL only)
  int A, T;
  const double* S; double* dS;            // [M+1][2][4]
  double* icl_raw; double* ial_raw;       // [M+1]
  const float* lv_icl; const float* lv_ial;
  float zoom;
  int want_grad;
};

__global__ void __launch_bounds__(NT)
pair_all_kernel(PairAllArgs p) {
  constexpr int MM = kPairMaxModal;
  constexpr int NV = 10 * MM + 9;         // per modality {icl, ial, dS01[4], dS1[4]}; joint {icl, dS01[4], dSj1[4]}
  const int A = p.A, T = p.T, M = p.M;
  __shared__ float s_is[MM + 1][8];       // reciprocals of the normalisers (+1e-9, losses.py:12-13)
  __shared__ float s_c[MM][2];            // {c_icl, c_ial} per modality
  __shared__ float red[NT / 32][NV];
  const float invA2 = 1.f / ((float)A * (float)A);
  if (threadIdx.x < (M + 1) * 8) s_is[threadIdx.x >> 3][threadIdx.x & 7] = 1.f / ((float)p.S[threadIdx.x] + kEps);
  if (threadIdx.x < M) {
    s_c[threadIdx.x][0] = expf(-p.lv_icl[threadIdx.x]) * invA2;
    s_c[threadIdx.x][1] = p.zoom * expf(-p.lv_ial[threadIdx.x]) * 0.1f * 0.5f;
  }
  __syncthreads();
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  const float* isj = s_is[M];
  for (int a = blockIdx.x; a < A; a += gridDim.x) {
    const int64_t rowF = (int64_t)a * T;
    for (int b = threadIdx.x; b < A; b += NT) {
      const int64_t off = rowF + b;
      const float j1 = p.F1[M][off], j2 = p.F2[M][off];
      QD m12, m21;
      float lm12 = 0.f, lm21 = 0.f, accJ1 = 0.f, accJ2 = 0.f;
      if (M > 0) {
        m12 = cpd(j1, 1.f, isj[4], isj[5]);
        m21 = cpd(j2, 1.f, isj[6], isj[7]);
        lm12 = __logf(m12.q);
        lm21 = __logf(m21.q);
      }
#pragma unroll
      for (int x = 0; x < MM; ++x) {
        if (x < M) {
          const float* is = s_is[x];
          const float c_icl = s_c[x][0], c_ial = s_c[x][1];
          const float g1 = p.F1[x][off], g2 = p.F2[x][off];
          // ---- ICL, tau = 0.1 (losses.py:43-58)
          QD q12 = cpd(g1, 10.f, is[0], is[1]);
          QD q21 = cpd(g2, 10.f, is[2], is[3]);
          const float w = 0.5f * q12.q + 0.5f * q21.q;
          acc[10 * x] += -__logf(w);
          // ---- IAL, tau = 1.0 (losses.py:68-97): target = modal q, input = log(joint q)
          QD o12 = cpd(g1, 1.f, is[4], is[5]);
          QD o21 = cpd(g2, 1.f, is[6], is[7]);
          const float ea = __expf(o12.q), eb = __expf(o21.q);
          const float la = o12.q - lm12, lb = o21.q - lm21;
          acc[10 * x + 1] += 0.5f * (ea * la + eb * lb);
          if (p.want_grad) {
            const float dq = __fdividef(-0.5f * c_icl, w);
            float d1 = dq * q12.dg, d2 = dq * q21.dg;
            acc[10 * x + 2] += dq * q12.dsa; acc[10 * x + 3] += dq * q12.dsb;
            acc[10 * x + 4] += dq * q21.dsa; acc[10 * x + 5] += dq * q21.dsb;
            const float dqo12 = c_ial * ea * (la + 1.f), dqo21 = c_ial * eb * (lb + 1.f);
            const float dqm12 = __fdividef(-c_ial * ea, m12.q), dqm21 = __fdividef(-c_ial * eb, m21.q);
            d1 += dqo12 * o12.dg;
            d2 += dqo21 * o21.dg;
            acc[10 * x + 6] += dqo12 * o12.dsa; acc[10 * x + 7] += dqo12 * o12.dsb;
            acc[10 * x + 8] += dqo21 * o21.dsa; acc[10 * x + 9] += dqo21 * o21.dsb;
            acc[10 * MM + 5] += dqm12 * m12.dsa; acc[10 * MM + 6] += dqm12 * m12.dsb;
            acc[10 * MM + 7] += dqm21 * m21.dsa; acc[10 * MM + 8] += dqm21 * m21.dsb;
            accJ1 += dqm12 * m12.dg;
            accJ2 += dqm21 * m21.dg;
            p.F1[x][off] = d1;
            p.F2[x][off] = d2;
          }
        }
      }
      // ---- the joint (or the single embedding): ICL only; coefficient 1 (losses.py:139-141)
      QD q12 = cpd(j1, 10.f, isj[0], isj[1]);
      QD q21 = cpd(j2, 10.f, isj[2], isj[3]);
      const float w = 0.5f * q12.q + 0.5f * q21.q;
      acc[10 * MM] += -__logf(w);
      if (p.want_grad) {
        const float dq = __fdividef(-0.5f * invA2, w);
        acc[10 * MM + 1] += dq * q12.dsa; acc[10 * MM + 2] += dq * q12.dsb;
        acc[10 * MM + 3] += dq * q21.dsa; acc[10 * MM + 4] += dq * q21.dsb;
        p.F1[M][off] = dq * q12.dg + accJ1;
        p.F2[M][off] = dq * q21.dg + accJ2;
      }
    }
  }
  // block reduction, one double atomic per (block, value)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int i = threadIdx.x;
    double v = 0.0;
    for (int w = 0; w < NT / 32; ++w) v += (double)red[w][i];
    double* dst = nullptr;
    if (i < 10 * MM) {
      const int x = i / 10, k = i % 10;
      if (x < M) dst = k == 0 ? p.icl_raw + x : (k == 1 ? p.ial_raw + x : (p.want_grad ? p.dS + (size_t)x * 8 + (k - 2) : nullptr));
    } else {
      const int k = i - 10 * MM;
      if (k == 0) dst = p.icl_raw + M;
      else if (p.want_grad) dst = p.dS + (size_t)M * 8 + (k - 1);     // k 1..4 -> dS01[0..3], 5..8 -> dS1[0..3]
    }
    if (dst && v != 0.0) atomicAdd(dst, v);
  }
}

// in place over the U blocks of F1 (blockIdx.y = 0) / F2 (1): u -> sum_tau dS[tau][seg] / tau * exp(u / tau)
// (the same ex2.approx exponential as the forward epilogue that produced the sums)
__global__ void __launch_bounds__(NT)
coef_kernel(float* __restrict__ F1, float* __restrict__ F2, int A, int T, int ld, int J1, int J2, const double* __restrict__ dS) {
  const int dir = blockIdx.y;
  float* __restrict__ F = dir == 0 ? F1 : F2;
  const int c_split = A + (dir == 0 ? J1 : J2);
  const int seg_lo = dir == 0 ? 0 : 2, seg_hi = seg_lo + 1;
  const float lo01 = (float)dS[seg_lo] * 10.f, lo1 = (float)dS[4 + seg_lo];
  const float hi01 = (float)dS[seg_hi] * 10.f, hi1 = (float)dS[4 + seg_hi];
  for (int a = blockIdx.x; a < A; a += gridDim.x) {
    float* row = F + (int64_t)a * ld;
    for (int c = A + threadIdx.x; c < T; c += NT) {
      const float u = row[c];
      const float e01 = __expf(u * 10.f), e1 = __expf(u);
      row[c] = (c < c_split) ? lo01 * e01 + lo1 * e1 : hi01 * e01 + hi1 * e1;
    }
  }
}

// backward of F.normalize: dX = (dXh - Xh <Xh, dXh>) / max(||X||, eps)
__global__ void __launch_bounds__(NT)
normalize_bwd_kernel(const float* __restrict__ Xh, const float* __restrict__ norms, const float* __restrict__ dXh,
                     int64_t N, int D, float* __restrict__ dX) {
  int64_t row = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (row >= N) return;
  int lane = threadIdx.x & 31;
  const float nrm = norms[row];
  const float den = fmaxf(nrm, 1e-12f);
  float dot = 0.f;
  for (int k = lane; k < D; k += 32) dot = fmaf(Xh[row * D + k], dXh[row * D + k], dot);
  dot = warp_sum(dot);
  const bool clamped = nrm < 1e-12f;   // below eps the denominator is a constant
  for (int k = lane; k < D; k += 32) {
    float xh = Xh[row * D + k];
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

}  // namespace
}  // namespace sga

extern "C" size_t sga_loss_workspace_bytes(int n_emb, const int* dims_host, int64_t N, int A, int J1, int J2, int want_grad) {
  if (n_emb < 1 || n_emb > 16) return 0;
  return sga::make_layout(n_emb, dims_host, N, A, J1, J2, want_grad).total;
}

// 0 (default): e1i/e2i/e1j/e2j partition the nodes (the dataloader contract), narrow embeddings take the packed-image
// Gram kernel.  1: arbitrary (overlapping / repeated) index sets -- every Gram gathers rows in the GEMM loader instead.
extern "C" void sga_loss_set_gram_path(int legacy) { sga::g_gram_legacy_override = legacy ? 1 : 0; }

// Kernels one sga_loss_fwd_bwd call launches (for the caller's launch accounting; memsets are not kernels).
extern "C" int sga_loss_launch_count(int n_emb, const int* dims_host, int J1, int J2, int want_grad) {
  using namespace sga;
  int n_narrow = 0, n_tswide = 0;
  for (int x = 0; x < n_emb; ++x)
    if (gram_ts_ok(dims_host[x])) (dims_host[x] <= 128 ? n_narrow : n_tswide) += 1;
  const int n_wide = n_emb - n_narrow - n_tswide;          // beyond 512: generic GEMM
  int n = 2 + n_emb + ((n_emb - 1) <= kPairMaxModal ? 1 : n_emb);   // ridx, finalize; per embedding: norm/pack; pair kernel(s)
  if (n_narrow + n_tswide) n += 1;                         // slots
  if (n_narrow) n += (2 * n_narrow + kGramMaxGroup - 1) / kGramMaxGroup;     // grouped gram_ts launches per variant
  if (n_tswide) n += (2 * n_tswide + kGramMaxGroup - 1) / kGramMaxGroup;
  if (n_wide) n += (2 * n_wide + kGemmMaxGroup - 1) / kGemmMaxGroup;
  if (want_grad) n += n_emb * ((J1 + J2 > 0) ? 2 : 1) + 2 * ((2 * n_emb + kGemmMaxGroup - 1) / kGemmMaxGroup);
  return n;
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

  // ---- forward Grams on the tensor cores.  d <= 128: rows normalised, split and laid out as MMA tile images once
  //      (pack_rows_kernel), Grams by the TMA-fed A-in-tensor-memory kernel (gram_ts.cu), all embeddings and both
  //      directions in one grouped launch.  Wider embeddings: rows normalised once, gathers + split in the loader of
  //      the generic GEMM (gemm_tc.cu), again one grouped launch.  Normaliser sums come out of the epilogues.
  GemmParams probs[2 * 16];
  GramProblem gprobs[2 * 16];
  int np = 0, ngp = 0;
  bool slots_built = false;
  for (int x = 0; x < n_emb; ++x) {
    const int d = dims_host[x];
    const float* Xh = F(L.Xh[x]);
    double* Sx = S + (size_t)x * 8;
    if (gram_ts_ok(d)) {
      if (!slots_built) {
        int rc = launch_build_slots(e1i, e2i, e1j, e2j, A, J1, J2, (int32_t*)(ws + L.slot), N, st);
        if (rc != SGA_OK) return rc;
        slots_built = true;
      }
      unsigned char* img[4] = {ws + L.img[x][0], ws + L.img[x][1], ws + L.img[x][2], ws + L.img[x][3]};
      int rc = launch_pack_rows(embs_host[x], N, d, (const int32_t*)(ws + L.slot), img, F(L.norms[x]), F(L.Xh[x]), st);
      if (rc != SGA_OK) return rc;
      for (int dir = 0; dir < 2; ++dir) {
        GramProblem& P = gprobs[ngp++];
        memset(&P, 0, sizeof(P));
        P.a_img = img[dir];                       // P1 = Xh[e1i] / P2 = Xh[e2i]
        P.b_img[0] = img[1 - dir];                // [P2; Q1; Q2] / [P1; Q2; Q1]
        P.b_img[1] = img[dir == 0 ? 2 : 3];
        P.b_img[2] = img[dir == 0 ? 3 : 2];
        P.seg_rows[0] = A; P.seg_rows[1] = dir == 0 ? J1 : J2; P.seg_rows[2] = dir == 0 ? J2 : J1;
        P.M = A; P.nkc = (d + 31) / 32;
        P.C = F(dir == 0 ? L.F1[x] : L.F2[x]); P.ldc = ldF;
        const int lo = dir == 0 ? 0 : 2, hi = dir == 0 ? 1 : 3;   // {S11,S12} / {S22,S21}
        P.s01[0] = Sx + lo; P.s01[1] = Sx + hi; P.s1[0] = Sx + 4 + lo; P.s1[1] = Sx + 4 + hi;
      }
      continue;
    }
    row_norm_kernel<<<(unsigned)((N + 7) / 8), NT, 0, st>>>(embs_host[x], N, d, F(L.norms[x]), F(L.Xh[x]));
    SGA_LAUNCH_CHECK();
    for (int dir = 0; dir < 2; ++dir) {
      GemmParams& P = probs[np++];
      memset(&P, 0, sizeof(P));
      P.A = {Xh, d, dir == 0 ? e1i : e2i, nullptr, 0};
      P.B = {Xh, d, dir == 0 ? ridx : r2idx, nullptr, 0};
      P.M = A; P.N = T; P.K = d;
      P.C = F(dir == 0 ? L.F1[x] : L.F2[x]); P.ldc = ldF;
      P.mode = 1;
      P.es_c0 = A;
      P.es_split = A + (dir == 0 ? J1 : J2);
      const int lo = dir == 0 ? 0 : 2, hi = dir == 0 ? 1 : 3;   // {S11,S12} / {S22,S21}
      P.s01_lo = Sx + lo; P.s01_hi = Sx + hi; P.s1_lo = Sx + 4 + lo; P.s1_hi = Sx + 4 + hi;
    }
  }
  {
    int rc = launch_gram_ts(gprobs, ngp, st);
    if (rc != SGA_OK) return rc;
    rc = launch_gemm_tc_group(probs, np, st);
    if (rc != SGA_OK) return rc;
  }
  // ---- element-wise loss terms (+ in-place gradient of the G blocks)
  const int xj = n_emb - 1;   // joint (or the single modality)
  const bool fused_pairs = (n_emb - 1) <= kPairMaxModal;
  if (fused_pairs) {
    PairAllArgs p;
    memset(&p, 0, sizeof(p));
    for (int x = 0; x < n_emb; ++x) { p.F1[x] = F(L.F1[x]); p.F2[x] = F(L.F2[x]); }
    p.M = n_emb - 1; p.A = A; p.T = ldF;
    p.S = S; p.dS = dS; p.icl_raw = icl_raw; p.ial_raw = ial_raw;
    p.lv_icl = log_vars_icl; p.lv_ial = log_vars_ial;
    p.zoom = zoom; p.want_grad = want_grad;
    pair_all_kernel<<<(unsigned)(A < 8 * sm_count() ? A : 8 * sm_count()), NT, 0, st>>>(p);
    SGA_LAUNCH_CHECK();
  }
  if (!fused_pairs && want_grad && n_emb > 1) {
    SGA_CUDA(cudaMemsetAsync(ws + L.acc1, 0, 4 * (size_t)A * A, st));
    SGA_CUDA(cudaMemsetAsync(ws + L.acc2, 0, 4 * (size_t)A * A, st));
  }
  for (int x = 0; x < n_emb && !fused_pairs; ++x) {
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
    pair_kernel<<<(unsigned)(A < 8 * sm_count() ? A : 8 * sm_count()), NT, 0, st>>>(p);
    SGA_LAUNCH_CHECK();
  }
  finalize_kernel<<<1, 32, 0, st>>>(icl_raw, ial_raw, M, n_emb, A, log_vars_ial, log_vars_icl, zoom, losses_out,
                                    want_grad ? g_log_vars_ial : nullptr, want_grad ? g_log_vars_icl : nullptr);
  SGA_LAUNCH_CHECK();
  if (!want_grad) return SGA_OK;

  // ---- backward: coefficient blocks in place, then 4 scatter-add GEMMs per embedding:
  //   dXh[e1i]  += dF1   Xh[R]      dXh[R]  += dF1^T Xh[e1i]
  //   dXh[e2i]  += dF2   Xh[R']     dXh[R'] += dF2^T Xh[e2i]
  // split K so that the whole group (2 * n_emb problems per launch) fills the SMs for ~4 waves
  auto ksplit_for = [&](int Mr, int Nc, int Kc) {
    int tiles = ((Mr + 127) / 128) * ((Nc + 127) / 128);
    int chunks = (Kc + 31) / 32;
    int per_problem = (4 * sm_count() + 2 * n_emb - 1) / (2 * n_emb);
    int want = (per_problem + tiles - 1) / tiles;
    int cap = chunks / 4;
    if (want > cap) want = cap;
    return want < 1 ? 1 : want;
  };
  // all 4 * n_emb GEMMs in TWO grouped launches (one per operand-layout combination)
  GemmParams rowp[2 * 16], colp[2 * 16];
  int nr = 0, nc = 0;
  for (int x = 0; x < n_emb; ++x) {
    const int d = dims_host[x];
    const float* Xh = F(L.Xh[x]);
    double* dSx = dS + (size_t)x * 8;
    if (T > A) {
      coef_kernel<<<dim3((unsigned)(A < 4 * sm_count() ? A : 4 * sm_count()), 2), NT, 0, st>>>(F(L.F1[x]), F(L.F2[x]), A, T, ldF, J1, J2, dSx);
      SGA_LAUNCH_CHECK();
    }
    SGA_CUDA(cudaMemsetAsync(ws + L.dXh[x], 0, 4 * (size_t)N * d, st));
    for (int dir = 0; dir < 2; ++dir) {
      const float* Fd = F(dir == 0 ? L.F1[x] : L.F2[x]);
      const int32_t* rows_i = dir == 0 ? e1i : e2i;
      const int32_t* rows_r = dir == 0 ? ridx : r2idx;
      // dXh[rows_i[a]] += sum_t Fd[a,t] Xh[rows_r[t]]
      GemmParams& P = rowp[nr++];
      memset(&P, 0, sizeof(P));
      P.A = {Fd, ldF, nullptr, nullptr, 0};
      P.B = {Xh, d, rows_r, nullptr, 1};
      P.M = A; P.N = d; P.K = T;
      P.C = F(L.dXh[x]); P.ldc = d; P.mode = 2; P.c_idx = rows_i;
      P.ksplit = ksplit_for(A, d, T);
      // dXh[rows_r[t]] += sum_a Fd[a,t] Xh[rows_i[a]]
      GemmParams& Q = colp[nc++];
      Q = P;
      Q.A = {Fd, ldF, nullptr, nullptr, 1};
      Q.B = {Xh, d, rows_i, nullptr, 1};
      Q.M = T; Q.N = d; Q.K = A;
      Q.c_idx = rows_r;
      Q.ksplit = ksplit_for(T, d, A);
    }
  }
  {
    int rc = launch_gemm_tc_group(rowp, nr, st);
    if (rc != SGA_OK) return rc;
    rc = launch_gemm_tc_group(colp, nc, st);
    if (rc != SGA_OK) return rc;
  }
  for (int x = 0; x < n_emb; ++x) {
    normalize_bwd_kernel<<<(unsigned)((N + 7) / 8), NT, 0, st>>>(F(L.Xh[x]), F(L.norms[x]), F(L.dXh[x]), N, dims_host[x], g_embs_host[x]);
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}
