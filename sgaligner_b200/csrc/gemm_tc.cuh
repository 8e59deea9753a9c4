// Problem descriptors of the tcgen05 tf32x3 GEMM (gemm_tc.cu), shared with its callers (loss.cu).
#pragma once
#include "common.cuh"

namespace sga {

struct GemmOperand {
  const float* p;
  int64_t ld;
  const int32_t* idx;   // optional row gather
  const float* div;     // optional per-(source)-row divisor (L2 norm, clamped by the caller)
  int mn_major;
};

struct GemmParams {
  GemmOperand A, B;
  int M, N, K;
  float* C;
  int64_t ldc;
  int mode;               // 0 store, 1 store + exp-sums, 2 scatter-add rows (atomicAdd C[c_idx[m]][n])
  const int32_t* c_idx;   // mode 2
  int es_c0, es_split;    // mode 1: columns >= es_c0 feed the sums; < es_split -> S_lo else S_hi
  double* s01_lo; double* s01_hi; double* s1_lo; double* s1_hi;   // sum exp(x/0.1), sum exp(x)
  int ksplit;             // mode 2 only: the K range is cut into ksplit slices, each adds its partial product
  int dbg;                // profiling only (SGA_GEMM_DBG): 1 no global loads, 2 no convert/stores, 4 no C writes, 8 no MMA
};

// Several independent problems with the SAME operand majorness executed by ONE persistent launch: the
// work items (output tile x K slice) of all problems form one queue, so a batch of small Grams fills
// the 148 SMs for several waves instead of paying launch + pipeline ramp per problem.
constexpr int kGemmMaxGroup = 16;
struct GemmGroup {
  GemmParams p[kGemmMaxGroup];
  int work_end[kGemmMaxGroup];   // exclusive prefix sum of the problems' work-item counts
  int n;
};

int launch_gemm_tc(const GemmParams& P, cudaStream_t st);
// `n` problems (any n: launched in chunks of kGemmMaxGroup); all must share (A.mn_major, B.mn_major).
int launch_gemm_tc_group(const GemmParams* problems, int n, cudaStream_t st);
// dense product without gathers; accumulate = split-K with atomic adds into C
int launch_gemm_tc_dense(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, int M, int N, int K,
                         float* C, int64_t ldc, int accumulate, cudaStream_t st);
// shapes worth a tensor-core launch (tiny feature widths stay on the FMA kernel)
inline bool gemm_tc_worth(int feat_a, int feat_b) { return feat_a >= 32 && feat_b >= 32; }

}  // namespace sga
