// Problem descriptors of the TMA-fed, A-in-tensor-memory Gram kernel (gram_ts.cu), shared with loss.cu.
#pragma once
#include "common.cuh"

namespace sga {

struct GramProblem {
  const unsigned char* a_img;      // image of the anchor row set (tiles of 128 rows, see gram_ts.cu)
  const unsigned char* b_img[3];   // images of the three column segments, in output-column order
  int seg_rows[3];                 // rows (= output columns) per segment
  int M;                           // anchor rows
  int nkc;                         // K chunks of 32 (<= 4)
  float* C;                        // [M, ldc], columns: segment 0 | segment 1 | segment 2
  int64_t ldc;
  double* s01[2];                  // sum exp(x / 0.1) over segment 1 / segment 2
  double* s1[2];                   // sum exp(x)
};

constexpr int kGramMaxGroup = 16;
struct GramGroup {
  GramProblem p[kGramMaxGroup];
  int item_end[kGramMaxGroup];     // exclusive prefix of the problems' work items
  int nsplit[kGramMaxGroup];       // column-tile ranges per row block
  int n;
};

size_t gram_image_bytes(int rows, int D);
int launch_build_slots(const int32_t* e1i, const int32_t* e2i, const int32_t* e1j, const int32_t* e2j, int A, int J1, int J2,
                       int32_t* slot, int64_t N, cudaStream_t st);
// norms[N], Xh[N,D] (as row_norm_kernel) + the four row-set images of this embedding
int launch_pack_rows(const float* X, int64_t N, int D, const int32_t* slot, unsigned char* const* img4, float* norms, float* Xh,
                     cudaStream_t st);
int launch_gram_ts(const GramProblem* problems, int n, cudaStream_t st);

}  // namespace sga
