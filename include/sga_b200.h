/*
 * sga_b200.h -- C ABI of libsga_b200.so: the B200 (sm_100a) kernels behind the SGAligner
 * node-embedding / matching / contrastive-loss hot path.
 *
 * The reference (sayands/sgaligner) is pure Python/PyTorch: its "FFI" for this path is the
 * nn.Module call surface (src/aligner/sg_aligner.py:71 MultiModalEncoder.forward,
 * src/aligner/losses.py:114 OverallLoss.forward, src/inference/sgaligner/inference_align_reg.py:125
 * matching head).  Each entry point below replaces the ATen / cuDNN / cuBLAS / PyG call sequence of
 * one reference function (cited per function); the Python mirror of the reference modules
 * (sgaligner_b200/sg_aligner.py, losses.py, matching.py) binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller (PyTorch) allocates every buffer, including workspaces; the library never
 *     allocates or frees device memory and keeps no state besides the last error string and
 *     cached function attributes;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 on success, SGA_E* (< 0) for argument errors, a positive cudaError_t for
 *     launch errors; sga_last_error() gives the text;
 *   - row-major contiguous tensors; "N" is the number of objects (graph nodes) in the batch.
 */
#ifndef SGA_B200_H
#define SGA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGA_OK 0
#define SGA_EINVAL (-1)      /* bad argument / unsupported shape */
#define SGA_EWORKSPACE (-2)  /* workspace too small */
#define SGA_EARCH (-3)       /* not running on an sm_100 device */

/* kernel selection for sga_pointnet_fwd */
#define SGA_POINTNET_SIMT 0  /* fp32 FMA reference path (any shape) */
#define SGA_POINTNET_TC 1    /* tcgen05 bf16x3 split-operand tensor-core path */

const char* sga_last_error(void);
int sga_version(void);
/* device query used by the host side to size persistent grids */
int sga_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- a3: PointNetfeat.forward (src/aligner/networks/pointnet.py:120-175; conv1..3 + ReLU + max)
 * pts [N,P,3] f32 (data_dict['tot_obj_pts'] as collated, NOT permuted); W1 [64,3] b1 [64];
 * W2 [128,64] b2 [128]; W3 [C3,128] b3 [C3]; out [N,C3]; argmax [N,C3] int32 (point index that
 * attains the max, lowest index on ties) or NULL.  mode: SGA_POINTNET_*.  With SGA_POINTNET_TC and argmax != NULL
 * (P <= 32767) a second launch re-evaluates near-ties of the max-pool in fp32, so argmax is the fp32 argmax. */
int sga_pointnet_fwd(const float* pts, int64_t N, int P,
                     const float* W1, const float* b1, const float* W2, const float* b2,
                     const float* W3, const float* b3, int C3,
                     float* out, int32_t* argmax, int mode, void* stream);

/* Cap on the SMs the persistent tensor-core forward occupies (0 = all of them): a serving step that runs the
 * graph branch concurrently on a second stream leaves it a few SMs this way.  Process-wide setting. */
int sga_pointnet_set_max_ctas(int n);

/* backward of the above (autograd of pointnet.py:140-163 through the max-pool): accumulates (+=)
 * into gW1 [64,3] gb1 gW2 [128,64] gb2 gW3 [C3,128] gb3.  `out`/`argmax` are the forward results. */
int sga_pointnet_bwd(const float* pts, int64_t N, int P,
                     const float* W1, const float* b1, const float* W2, const float* b2,
                     const float* W3, const float* b3, int C3,
                     const float* out, const int32_t* argmax, const float* grad_out,
                     float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3,
                     void* stream);

/* the same with kernel selection: SGA_POINTNET_TC runs the backward as tcgen05 GEMMs over tiles of 128
 * (object, channel) instances (conv2 recomputed for the argmax points, dW2 = dz2^T h1 and dh1 = dz2 W2 on the
 * tensor cores, bf16x3 split operands); needs C3 % 128 == 0 and 16-byte aligned W2 / W3. */
int sga_pointnet_bwd_mode(const float* pts, int64_t N, int P,
                          const float* W1, const float* b1, const float* W2, const float* b2,
                          const float* W3, const float* b3, int C3,
                          const float* out, const int32_t* argmax, const float* grad_out,
                          float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3,
                          int mode, void* stream);

/* train-mode side effect of the discarded BatchNorm1d calls (pointnet.py:141-142,154-155,158-159):
 * per-channel sum and sum of squares of the three pre-ReLU conv outputs over all N*P points.
 * moments: f64 [2*(64+128+C3)] = {sum1[64], sq1[64], sum2[128], sq2[128], sum3[C3], sq3[C3]}, zeroed
 * by the caller. */
int sga_pointnet_bn_moments(const float* pts, int64_t N, int P,
                            const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, int C3, double* moments, void* stream);

/* Same statistics as sga_pointnet_bn_moments (added into moments[2*(64+128+C3)], f64), computed from Gram matrices
 * of the ReLU outputs on the tensor cores (csrc/pointnet_gram.cu) instead of summing every conv output: about a
 * third of a forward pass.  scratch >= sga_pointnet_gram_scratch_bytes(), 16-byte aligned; W2 16-byte aligned. */
size_t sga_pointnet_gram_scratch_bytes(void);
int sga_pointnet_bn_moments_gram(const float* pts, int64_t N, int P, const float* W1, const float* b1,
                                 const float* W2, const float* b2, const float* W3, const float* b3, int C3,
                                 double* moments, void* scratch, size_t scratch_bytes, void* stream);

/* The running_mean / running_var / num_batches_tracked update the reference's discarded BatchNorm1d calls perform
 * in train() (pointnet.py:141-142,154-155,158-159; momentum update with the unbiased batch variance) for the three
 * layers in one launch.  moments as produced by sga_pointnet_bn_moments / sga_pointnet_fwd_stats; cnt = N*P. */
int sga_bn_running_update(const double* moments, double cnt, float momentum, int C3, float* rm1, float* rv1,
                          float* rm2, float* rv2, float* rm3, float* rv3, int64_t* nbt1, int64_t* nbt2,
                          int64_t* nbt3, void* stream);

/* Training-mode forward on the tensor cores: sga_pointnet_fwd(mode = SGA_POINTNET_TC) that ALSO accumulates
 * the statistics of sga_pointnet_bn_moments in the same pass (conv3 thread-locally in the max-pool
 * epilogue, conv2 by a warp transpose-reduce, conv1 analytically from the point moments), i.e. the whole
 * train-mode PointNetfeat.forward (pointnet.py:140-163 incl. the BatchNorm side effect) in one launch + a
 * finalize launch.  moments as in sga_pointnet_bn_moments (added into; zeroed by the caller); scratch:
 * zeroed device buffer of >= sga_pointnet_stats_scratch_bytes(C3) bytes.  C3 must be a multiple of 128. */
size_t sga_pointnet_stats_scratch_bytes(int C3);
int sga_pointnet_fwd_stats(const float* pts, int64_t N, int P,
                           const float* W1, const float* b1, const float* W2, const float* b2,
                           const float* W3, const float* b3, int C3,
                           float* out, int32_t* argmax, double* moments, void* scratch, size_t scratch_bytes,
                           void* stream);

/* ---- a2/a7: block-diagonal CSR (by destination) for all 2B graphs of the batch, replacing the
 * per-graph Python slicing of sg_aligner.py:86-104 plus PyG's remove_self_loops/add_self_loops.
 * edges [E,2] int64 graph-local (src,dst) rows as collated (scan3r.py:201); node_off [G+1] int32 and
 * edge_off [G+1] int64 are exclusive prefix sums of graph_per_obj_count / graph_per_edge_count.
 * Outputs: row_beg [N] / row_cnt [N] int32 and col [E+N] int32 (GLOBAL source node ids; the slots of
 * graph g start at edge_off[g]+node_off[g]; within a row: surviving edges in input order, then the
 * added self loop). */
int sga_csr_build(const int64_t* edges, const int32_t* node_off, const int64_t* edge_off, int G,
                  int max_graph_nodes, int32_t* row_beg, int32_t* row_cnt, int32_t* col, void* stream);

/* ---- a7 (first half): GATConv linear + attention logits.  x [N,in_dim] (f32, or f64 when
 * x_is_f64 -- data_dict['tot_rel_pose'] arrives as f64); W [H*C,in_dim]; att_src/att_dst [H*C].
 * xs [H][N][C] (head-major so that one graph's tile of one head is contiguous); a_src/a_dst [N,H]. */
int sga_gat_linear(const void* x, int x_is_f64, int64_t N, int in_dim, const float* W,
                   const float* att_src, const float* att_dst, int H, int C,
                   float* xs, float* a_src, float* a_dst, void* stream);

/* ---- a7 (second half): leaky-ReLU(0.2) edge logits, softmax over incoming edges (+1e-16),
 * weighted aggregation, + bias, optional ELU (gat.py:45-46).  out [N,H*C]. */
int sga_gat_aggregate(const float* xs, const float* a_src, const float* a_dst,
                      const int32_t* row_beg, const int32_t* row_cnt, const int32_t* col,
                      const int32_t* node_off, int G, int max_graph_nodes, int64_t N, int H, int C,
                      const float* bias, int apply_elu, float* out, void* stream);

/* backward of sga_gat_aggregate: given grad_out [N,H*C] (and `out` for the ELU derivative) produces
 * g_xs [H][N][C], g_a_src / g_a_dst [N,H] (all three zeroed by the caller) and accumulates g_bias [H*C]. */
int sga_gat_aggregate_bwd(const float* xs, const float* a_src, const float* a_dst,
                          const int32_t* row_beg, const int32_t* row_cnt, const int32_t* col,
                          const int32_t* node_off, int G, int max_graph_nodes, int64_t N, int H, int C,
                          int apply_elu, const float* out, const float* grad_out, float* g_xs,
                          float* g_a_src, float* g_a_dst, float* g_bias, void* stream);

/* backward of sga_gat_linear (x must be f32 here -- see sga_cast_f64_f32): folds g_a_src/g_a_dst into
 * g_xs (in place), then gW += g_xs^T x, g_att_* +=, and (if gx != NULL) gx [N,in_dim] = g_xs W. */
int sga_gat_linear_bwd(const float* x, int64_t N, int in_dim, const float* W, const float* att_src,
                       const float* att_dst, int H, int C, const float* xs, float* g_xs,
                       const float* g_a_src, const float* g_a_dst, float* gW, float* g_att_src,
                       float* g_att_dst, float* gx, void* stream);

/* the reference's `.float()` on the f64 dataloader tensors (sg_aligner.py:73-75) for the backward
 * kernels, which take f32 inputs only */
int sga_cast_f64_f32(const double* in, float* out, int64_t n, void* stream);

/* ---- a5 + a8: one modality's nn.Linear (sg_aligner.py:112-122) fused with its slice of
 * MultiModalFusion (sg_aligner.py:30-35).  x [N,in_dim] f32/f64; W [out_dim,in_dim]; b [out_dim];
 * emb [N,out_dim].  When joint != NULL: joint[n, joint_col : joint_col+out_dim] =
 * softmax(fusion_w)[m] * emb[n] / max(||emb[n]||, 1e-12), joint row stride joint_ld floats. */
int sga_project_fuse_fwd(const void* x, int x_is_f64, int64_t N, int in_dim, const float* W,
                         const float* b, int out_dim, float* emb, float* joint, int joint_ld,
                         int joint_col, const float* fusion_w, int M, int m, void* stream);

/* The same for ALL M modalities in one launch (MultiModalEncoder.forward, sg_aligner.py:112-135): host arrays
 * of M device pointers / sizes; every modality projects to out_dim (= emb_dim, <= 128) columns and modality m
 * writes joint[:, m*out_dim : (m+1)*out_dim].  joint may be NULL (M == 1: no fusion). */
int sga_project_fuse_fwd_multi(const void* const* x_host, const int* x_is_f64_host, const int* in_dim_host,
                               const float* const* W_host, const float* const* b_host, float* const* emb_host, int M,
                               int64_t N, int out_dim, float* joint, int joint_ld, const float* fusion_w, void* stream);

/* backward: g_emb [N,out_dim] (direct gradient on the modality embedding, may be NULL) and
 * g_joint (gradient on the joint embedding, may be NULL) -> gW +=, gb +=, g_fusion_w [M] +=, and
 * gx [N,in_dim] (may be NULL).  x must be f32.  workspace >= (N*out_dim + 64) floats. */
int sga_project_fuse_bwd(const float* x, int64_t N, int in_dim, const float* W, int out_dim,
                         const float* emb, const float* g_emb, const float* g_joint, int joint_ld,
                         int joint_col, const float* fusion_w, int M, int m, float* gW, float* gb,
                         float* g_fusion_w, float* gx, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- a9: matching head (inference_align_reg.py:125-128) for all pairs of the batch at once.
 * emb [N,D]; pair_off [B+1] int32 node offsets; sim_off [B+1] int64 = prefix sum of n_b^2 (n_b =
 * nodes of pair b, source + reference).  Per pair: rows L2-normalised (division by the norm, no
 * eps), sim = 1 - E E^T over source+reference nodes.  norms [N] is scratch, sim is packed per pair
 * at sim_off[b] as [n_b, n_b]. */
int sga_match_sim(const float* emb, int64_t N, int D, const int32_t* pair_off, const int64_t* sim_off,
                  int B, int max_pair_nodes, float* norms, float* sim, void* stream);

/* Tensor-core matching head: the same sim = 1 - E E^T as a tcgen05 tf32x3 Gram (fp32-faithful,
 * ~1e-6) fused with the per-row top-K (K <= 8) of the ranking, one CTA per (pair, 128-row block).
 * topk_idx/topk_dist [N,K] as in sga_match_rank; sim_out (packed like sga_match_sim) may be NULL, in
 * which case the similarity matrix never reaches HBM.  norms [N] is scratch. */
int sga_match_topk_tc(const float* emb, int64_t N, int D, const int32_t* pair_off, const int64_t* sim_off,
                      int B, int max_pair_nodes, int K, float* norms, int32_t* topk_idx, float* topk_dist,
                      float* sim_out, void* stream);

/* rank_list = argsort(sim, dim=1) made deterministic: ascending by (sim, column).  node_pair [N]
 * int32 maps a node to its pair.  topk_idx/topk_dist [N,K] (pair-local columns, the node itself
 * included exactly as in the reference's rank_list; -1 / +inf padding when n_b < K) may be NULL;
 * rank_full (int32, packed like sim) may be NULL. */
int sga_match_rank(const float* sim, int64_t N, const int32_t* pair_off, const int64_t* sim_off,
                   const int32_t* node_pair, int max_pair_nodes, int K, int32_t* topk_idx,
                   float* topk_dist, int32_t* rank_full, void* stream);

/* ---- a10: utils/alignment.py:3-25 on the device: for anchor t, the 0-based position of e2i[t] in
 * row e1i[t] of the ranking once the node itself is removed (Hits@k <=> pos < k, RR = 1/(pos+1)).
 * e1i/e2i [A] are GLOBAL node ids as collated. */
int sga_match_anchor_pos(const float* sim, const int32_t* pair_off, const int64_t* sim_off,
                         const int32_t* node_pair, const int32_t* e1i, const int32_t* e2i, int A,
                         int32_t* anchor_pos, void* stream);

/* ---- 8(f)2 device-side collation: src/datasets/scan3r.py:99-100 (obj_points - pcl_center) in place on the
 * raw points after the H2D copy.  pts [N,P,3] f32; center [B,3] f32; node_pair [N] int32 maps an object to
 * its pair. */
int sga_center_points(float* pts, int64_t N, int P, const float* center, const int32_t* node_pair, void* stream);

/* ---- a11: utils/alignment.py:27-89 on the device, one launch for all pairs: top1_idx/top1_dist [N] = the
 * best match of every node once the node itself is removed (pair-local column; compute_node_corrs with
 * k = 1 keeps the source nodes whose top1 is a reference node, i.e. top1_idx >= n_src[b]);
 * pair_out [B,4] = {SGAR '2', SGAR '50', SGAR '100', alignment score} (SGAR = -1 for a pair without
 * anchors, which the reference skips).  n_src [B]; e1i/e2i [A] GLOBAL node ids grouped by pair,
 * anchor_off [B+1] their per-pair prefix (cumsum of e1i_count). */
int sga_match_pair_metrics(const float* sim, const int32_t* pair_off, const int64_t* sim_off,
                           const int32_t* n_src, const int32_t* e1i, const int32_t* e2i,
                           const int32_t* anchor_off, int B, int32_t* top1_idx, float* top1_dist,
                           float* pair_out, void* stream);

/* ---- a12-a16: OverallLoss.forward (losses.py:114-152) and its gradient w.r.t. every embedding.
 * embs_host: host array of M+1 device pointers (M modal embeddings in module order, then the joint;
 * for M == 1 pass only the single embedding and n_emb = 1); dims_host [n_emb] feature widths.
 * e1i/e2i [A], e1j [J1], e2j [J2] int32 global row ids.  log_vars_ial/icl [M] (device).
 * losses_out [4] = {loss, icl_unimodal, icl_multimodal, ial}.  When want_grad: g_embs_host gives
 * n_emb device pointers [N,dim] that receive d loss / d emb (overwritten), g_log_vars_ial/icl [M].
 * workspace >= sga_loss_workspace_bytes(). */
size_t sga_loss_workspace_bytes(int n_emb, const int* dims_host, int64_t N, int A, int J1, int J2, int want_grad);
/* number of kernels one sga_loss_fwd_bwd call with these arguments launches (launch accounting of bench.py).
 * e1i/e2i/e1j/e2j must be pairwise disjoint and duplicate-free (they partition the nodes, scan3r.py:101-107). */
int sga_loss_launch_count(int n_emb, const int* dims_host, int J1, int J2, int want_grad);
/* legacy = 1: the index sets of the following sga_loss_fwd_bwd calls may overlap or repeat nodes (not the
 * dataloader's partition): all Grams gather their rows in the GEMM loader.  Process-wide switch; default 0. */
void sga_loss_set_gram_path(int legacy);
int sga_loss_fwd_bwd(const float* const* embs_host, const int* dims_host, int n_emb, int64_t N,
                     const int32_t* e1i, const int32_t* e2i, const int32_t* e1j, const int32_t* e2j,
                     int A, int J1, int J2, const float* log_vars_ial, const float* log_vars_icl,
                     float zoom, float* losses_out, int want_grad, float* const* g_embs_host,
                     float* g_log_vars_ial, float* g_log_vars_icl, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Generic fp32-faithful tensor-core GEMM (tcgen05, tf32x3 split operands) used by the loss and its
 * backward:  C[M,N] = sum_k A(m,k) B(n,k).  An operand is K-major (X(r,k) = P[row(r)*ld + k]) or
 * MN-major (X(r,k) = P[row(k)*ld + r]); row(i) = idx ? idx[i] : i; values are divided by div[row] when
 * div != NULL.  c_idx == NULL: C is stored; otherwise rows are scatter-added (atomicAdd) into
 * C[c_idx[m]*ldc + n] and the K range may be cut into `ksplit` slices.  Supported layout pairs:
 * (K,K), (K,MN), (MN,MN). */
int sga_gemm_tf32x3(const float* A, int64_t lda, int a_mn_major, const int32_t* a_idx, const float* a_div,
                    const float* B, int64_t ldb, int b_mn_major, const int32_t* b_idx, const float* b_div,
                    int M, int N, int K, float* C, int64_t ldc, const int32_t* c_idx, int ksplit, void* stream);

/* ---- a17: torch.optim.Adam step (L2 weight decay folded into the gradient, bias-corrected) over
 * a flat parameter buffer; grad is multiplied by grad_scale first (1/world_size after allreduce). */
int sga_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float grad_scale, void* stream);

/* the same over a flat buffer made of `nseg` parameter segments (seg_off [nseg+1] int64, device): a segment whose
 * gradient is exactly zero -- a parameter that produced no gradient in this step, torch's `.grad is None` -- is
 * skipped entirely, as torch.optim.Adam skips it (no weight-decay drift, moments untouched).  seg_active [nseg]
 * int32 device scratch.  Every seg_off entry is a multiple of 4 floats and grad is 16-byte aligned. */
int sga_adam_step_segments(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           const int64_t* seg_off, int nseg, int32_t* seg_active, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ==== 8(f)1: NaivePCT object encoder (src/aligner/networks/pct.py:275-317; the 'pct' module of
 * MultiModalEncoder, sg_aligner.py:59-60).  Activations are fp32 [N,P,C] (one row of C channels per point); every
 * BatchNorm is applied as a folded per-channel affine pair (a, b) -- BN(z) = a z + b -- by the prologue of the NEXT
 * kernel, so train mode (batch statistics) and eval mode (running statistics) run the same kernels; the kernels that
 * produce a BatchNorm input also accumulate its batch statistics {sum [C], sum of squares [C]} (f64, zeroed by the
 * caller, NULL = not wanted).  All contractions: tcgen05, bf16x3 split operands, fp32 accumulate.  P <= 512. ---- */

/* first / second moments of all N*P points: mom9 = {sum x,y,z, sum xx,xy,xz,yy,yz,zz} (f64, zeroed by the caller) */
int sga_pct_point_moments(const float* pts, int64_t NP, double* mom9, void* stream);
/* batch statistics of z = W p (a bias-free 3 -> C conv, Embedding.conv1 pct.py:106) from the point moments */
int sga_pct_affine_stats(const double* mom9, const float* W, int C, double* stats, void* stream);
/* nn.BatchNorm1d bookkeeping (pct.py:109-110,203,287,293-294): stats {sum, sumsq} [2C] over cnt values per channel of
 * the STORED tensor z; lin_bias (may be NULL): a bias the producing layer has but did not add to z.  training != 0:
 * batch statistics, and running_mean / running_var / num_batches_tracked get torch's momentum update (unbiased
 * variance); else running statistics (stats may be NULL).  a_out/b_out [C]: BN(z + lin_bias) = a z + b. */
int sga_bn_fold(const double* stats, double cnt, const float* lin_bias, const float* gamma, const float* beta,
                float* running_mean, float* running_var, int64_t* num_batches_tracked, int training,
                float momentum, float eps, int C, float* a_out, float* b_out, void* stream);
/* Embedding (pct.py:120-125): z2 [N,P,128] = conv2(relu(bn1(conv1(pts)))); W1 [128,3], (a1,b1) = folded bn1,
 * W2 [128,128]; stats [256] of z2 or NULL. */
int sga_pct_embed(const float* pts, int64_t N, int P, const float* W1, const float* a1, const float* b1,
                  const float* W2, float* z2, double* stats, void* stream);
/* 128-input-channel pointwise convolution with a fused prologue  X = g1(src1) + g2(src2),  g(s) = s (a == NULL) or
 * relu(a s + b); src2 may be NULL.  Y = X W^T + bias, W [Cout,128], Cout = 128, or 160 = k_conv (32 rows) | v_conv (128
 * rows) of one SA layer (pct.py:197-200,213-215): columns [0,c0) go to out0 [N,P,c0], the rest to out1 [N,P,Cout-c0].
 * out_x (may be NULL) receives X [N,P,128] (the SA residual x_l, pct.py:228-230).  stats [2 Cout] of Y or NULL. */
int sga_pct_pointwise(const float* src1, const float* a1, const float* b1, const float* src2, const float* a2,
                      const float* b2, int64_t N, int P, const float* W, const float* bias, int Cout, int c0,
                      float* out_x, float* out0, float* out1, double* stats, void* stream);
/* The fused k | v convolution (Cout = 32 + 128, W rows 0..31 = k_conv = q_conv, pct.py:197-200) that also records
 * v_absmax [N] (zeroed by the caller) = max |v| per object, which the backward's operand scale needs. */
int sga_pct_pointwise_kv(const float* src1, const float* a1, const float* b1, const float* src2, const float* a2,
                         const float* b2, int64_t N, int P, const float* W, const float* bias, float* out_x, float* k,
                         float* v, float* v_absmax, void* stream);
/* SA attention (pct.py:217-224), q and k share one weight: k [N,P,32], v [N,P,128].
 * sga_pct_attn_stats: c2 [N, 2, Ppad] (Ppad = P rounded up to 128) = log2-domain softmax normaliser of every row i of
 * energy = k k^T / sqrt(32), in two parts that are never added in fp32: [n,0,i] = row max * log2e/sqrt(32) (+inf for
 * padding rows), [n,1,i] = log2 of the row's sum of exponentials.
 * sga_pct_attn: xs [N,P,128], xs[j,:] = sum_i softmax(energy)[i,j] v[i,:]  (= torch.bmm(x_v, attention)). */
int sga_pct_attn_stats(const float* k, int64_t N, int P, float* c2, void* stream);
int sga_pct_attn(const float* k, const float* v, const float* c2, int64_t N, int P, float* xs, void* stream);
/* cat(x1..x4) -> Conv1d(512,1024,bias=False) (pct.py:285-289,306-308) with x4 = x3 + relu(a4 t4 + b4) formed on the fly.
 * WL [1024,512].  Per object and channel only max_p z and min_p z are kept (zmax/zmin [N,2,1024]: the two 64-point
 * column halves of the tiles separately) plus the statistics [2048] of z: BN + LeakyReLU + max commute with them.
 * imax/imin [N,2,1024] (both or neither; training): the point index that holds each maximum / minimum.
 * img (may be NULL): the pre-split operand image written by sga_pct_cat_pack (sga_pct_cat_image_bytes bytes, 128-byte
 * aligned); with it the kernel streams its operands with cp.async.bulk instead of converting them 8 times over. */
size_t sga_pct_cat_image_bytes(int64_t N, int P);
int sga_pct_cat_pack(const float* x1, const float* x2, const float* x3, const float* t4, const float* a4, const float* b4,
                     int64_t N, int P, void* img, void* stream);
int sga_pct_cat_linear(const float* x1, const float* x2, const float* x3, const float* t4, const float* a4,
                       const float* b4, int64_t N, int P, const float* WL, float* zmax, float* zmin,
                       double* stats, int32_t* imax, int32_t* imin, const void* img, void* stream);
/* pooled [N,1024] = LeakyReLU_0.2(a * (a >= 0 ? max : min) + b) = max_p LeakyReLU(BN(z)) (pct.py:310); with imax/imin
 * also pstar [N,1024] (the arg-max point, torch.max's index) and zsel [N,1024] (the selected z); NULL to skip. */
int sga_pct_pool_act(const float* zmax, const float* zmin, const int32_t* imax, const int32_t* imin, const float* a,
                     const float* b, int64_t N, int P, float* out, int32_t* pstar, float* zsel, void* stream);
/* head (pct.py:311-316): column statistics of x [N,C] over the objects; out = relu(a x + b) * (mask ? mask*scale : 1) */
int sga_col_stats(const float* x, int64_t N, int C, double* stats, void* stream);
int sga_bn_act_rows(const float* x, const float* a, const float* b, const float* mask, float scale, int64_t N, int C,
                    float* out, void* stream);

/* ---- backward of the NaivePCT encoder (autograd of pct.py:275-317; host orchestration in sgaligner_b200/pct.py).
 * Every BatchNorm is differentiated in closed form: with gy = upstream * activation' (and dropout mask),
 * S1 = sum gy, S2 = sum gy y over the batch,  d y = a gy - e - f y  (train mode: the batch-statistics terms; eval: e = f = 0). */
/* sums [2C] (f64, zeroed by the caller) += {sum gy, sum gy*y} per column of g, y [rows, C]; gy = g * (mask ? mask*scale : 1)
 * * (a y + b > 0 ? 1 : slope).  C in {128, 256, 512, 1024}. */
int sga_bn_bwd_stats(const float* g, const float* y, const float* a, const float* b, const float* mask, float scale,
                     float slope, int64_t rows, int C, double* sums, void* stream);
/* per channel from the sums: d gamma, d beta, and the (e, f, mean) of  dy = a gy - e - f (y - mean).  stats = forward batch statistics
 * {sum y, sum y^2} over cnt elements (training) or NULL (eval: running statistics, lin_bias = the bias folded into BN). */
int sga_bn_bwd_coef(const double* sums, const double* stats, double cnt, const float* lin_bias, const float* gamma,
                    const float* running_mean, const float* running_var, int training, float eps, int C, float* e,
                    float* f, float* mean, float* dgamma, float* dbeta, void* stream);
/* out [rows, C] = a gy - e - f (y - mean)  (e, f, mean may all be NULL = 0) */
int sga_bn_bwd_apply(const float* g, const float* y, const float* a, const float* b, const float* mask, float scale,
                     float slope, const float* e, const float* f, const float* mean, int64_t rows, int C, float* out,
                     void* stream);
/* ... recording absmax [rows / rows_per_object] (zeroed by the caller) = max |out| per object, the operand scale of the
 * tensor-core product that consumes `out` */
int sga_bn_bwd_apply_absmax(const float* g, const float* y, const float* a, const float* b, const float* mask, float scale,
                            float slope, const float* e, const float* f, const float* mean, int64_t rows, int C, float* out,
                            int64_t rows_per_object, float* absmax, void* stream);
/* Per-object power-of-two scale of a gradient operand (the backward's tensor-core products split their operands into
 * fp16 pairs; gradients are far below that range): scale [N,2] = {s, 1/s}, s = 2^floor(log2(target / (max|x_n| max|y_n|)))
 * over the `per` elements of object n (y may be NULL). */
int sga_pct_pow2_scale(const float* x, const float* y, int64_t N, int64_t per, float target, float* scale, void* stream);
/* SA backward, attention part (pct.py:217-224): dv [N,P,128] = attention dxs;  dk halves, see csrc/pct_attn.cu.
 * scale = sga_pct_pow2_scale(dxs, v, ...) */
int sga_pct_attn_bwd_dv(const float* k, const float* dxs, const float* c2, const float* scale, int64_t N, int P, float* dv,
                        double* dv_colsum /* [128], zeroed; or NULL */, float* dv_absmax /* [N], zeroed; or NULL */,
                        void* stream);
/* scale [N,2] = {s, 1/s}, s = 2^floor(log2(target / absmax[n])): the per-object scale from maxima a kernel recorded */
int sga_pct_scale_from_absmax(const float* absmax, int64_t N, float target, float* scale, void* stream);
/* the same from two maxima: s = 2^floor(log2(target / (absmax_x[n] * absmax_y[n]))) (= sga_pct_pow2_scale(x, y, ...)) */
int sga_pct_scale_from_absmax_pair(const float* absmax_x, const float* absmax_y, int64_t N, float target, float* scale,
                                   void* stream);
int sga_pct_attn_bwd_dk(const float* k, const float* fixed, const float* streamed, const float* c2, float* delta,
                        const float* scale, int64_t N, int P, int by_col, int delta_sweep, float* dk_out, void* stream);
/* delta [N,P] = scale[n][0] * sum_c x[n,p,c] y[n,p,c] (C = 128): the softmax-backward row term v_i . dv_i in scaled units */
int sga_pct_rowdot_scaled(const float* x, const float* y, const float* scale, int64_t N, int P, float* out, void* stream);
/* dX = dY Wt^T on the tensor cores: src [N,P,128] (scaled per object by scale [N,2]), Wt [128,128] (pass W^T), out [N,P,128] */
int sga_pct_pointwise_scaled(const float* src, const float* scale, int64_t N, int P, const float* Wt, float* out, void* stream);
/* ... recording out_absmax [N] (zeroed by the caller) = max |out| per object on the way out */
int sga_pct_pointwise_scaled_absmax(const float* src, const float* scale, int64_t N, int P, const float* Wt, float* out,
                                    float* out_absmax, void* stream);
/* gradient of an SA layer's input: out = gx (+ gcat) + dxv + (dk1 + dk2) Wk;  dk1 <- dk1 + dk2.  [rows,128] / [rows,32] */
int sga_pct_sa_input_grad(const float* gx, const float* gcat, const float* dxv, float* dk1, const float* dk2,
                          const float* Wk, int64_t rows, float* out, void* stream);
/* a1 [rows,128] = relu(a (W1 p) + b)  (Embedding.conv1 + bn1 + ReLU, pct.py:122, re-materialised for the weight gradient) */
int sga_pct_embed_a1(const float* pts, const float* W1, const float* a, const float* b, int64_t rows, float* out, void* stream);
/* sums5 [5*128] (f64, zeroed) += per channel {sum gy, sum gy z1, sum gy p_x, sum gy p_y, sum gy p_z}, gy = g * [a z1 + b > 0] */
int sga_pct_embed1_bwd_stats(const float* g, const float* pts, const float* W1, const float* a, const float* b,
                             int64_t rows, double* sums5, void* stream);
/* dW1 [128,3] = a T - e m1 - f (W1 M2 - mean m1) from the sums above and the 9 point moments (cnt = number of points) */
int sga_pct_embed1_wgrad(const double* sums5, const double* mom9, double cnt, const float* W1, const float* a, const float* e,
                         const float* f, float* dW1, void* stream);
/* backward through concat -> conv(512,1024) -> BN -> LeakyReLU -> max (pct.py:306-310).  Dense half: g_x[a] = -u[a] - (M (xcat - xbar))[a]
 * (M [512,512] symmetric, u, xbar [512]; overwrites g1..g4 [N,P,128]); sparse half: coef [N,1024] = a_c gy routed to the arg-max
 * point pstar [N,1024]:  g_x[n,pstar,:] += coef WL[c,:]  and  dWL[c,:] += coef xcat[n,pstar,:]. */
int sga_pct_cat_dense_bwd(const float* x1, const float* x2, const float* x3, const float* x4, int64_t N, int P,
                          const float* M, const float* scale /* [2] = sga_pct_pow2_scale(M) */, const float* u,
                          const float* xbar, float* g1, float* g2, float* g3, float* g4, void* stream);
int sga_pct_cat_sparse_bwd_x(const float* coef, const int32_t* pstar, const float* WL, int64_t N, int P, float* g1,
                             float* g2, float* g3, float* g4, void* stream);
int sga_pct_cat_sparse_bwd_w(const float* coef, const int32_t* pstar, const float* x1, const float* x2, const float* x3,
                             const float* x4, int64_t N, int P, float* dWL, void* stream);
/* out [rows,128] = x + relu(a t + b)  (x4, re-materialised for the backward) */
int sga_pct_residual(const float* x, const float* t, const float* a, const float* b, int64_t rows, float* out, void* stream);
/* dst [R,C] = alpha * dst + beta * rowscale[r] * src[r,c] + gamma * rowscale2[r] * colvec[c]  (small weight-gradient algebra) */
int sga_axpby_rows(float* dst, float alpha, const float* src, float beta, const float* rowscale, float gamma,
                   const float* rowscale2, const double* colvec, int64_t R, int C, void* stream);
/* Weight gradients / Gram blocks as contractions over all R = N*P points (csrc/pct_wgrad.cu; tcgen05, bf16 operand pairs,
 * accumulators resident in tensor memory):  C_b[m, n] += sum_r A[r, m] B_b[r, n]  for up to four B_b [R, nb_b] (nb_b = 128 or
 * 32) that share A [R, 128]; transpose != 0 stores C_b[n * ldc_b + m] instead (dW = dY^T X with A = X).  C must hold the
 * value to add to. */
int sga_pct_wgrad(const float* A, int64_t R, const float* const* B, const int* nb, int nB, float* const* C,
                  const int64_t* ldc, int transpose, void* stream);
/* Grouped weight-gradient products on sga_gemm_tf32x3 (both operands MN-major, split-K, atomic accumulation):
 * C_i [M_i,N_i] += A_i^T B_i with A_i [K,M_i] (row stride lda_i) and B_i [K,N_i]; C must hold the value to add to. */
int sga_wgrad_group(const float* const* A, const int64_t* lda, const int* M, const float* const* B, const int64_t* ldb,
                    const int* Nn, float* const* C, const int64_t* ldc, int n, int64_t K, void* stream);

/* ==== 8(f)4: EVA baseline path (src/aligner/eva.py:9-96; MultiGCN, src/aligner/networks/gat.py:6-25, over
 * torch_geometric 2.2.0 GCNConv; NCALoss / OverallNCALoss, src/aligner/losses.py:154-205).  The dense products of this
 * path (GCN layer 2, the NCA score matrix, its two gradient products) go through sga_gemm_tf32x3. */
/* GCNConv aggregation (gcn_norm + propagate): out_i = sum_{j in row i} h_j / sqrt(deg_i deg_j) (+ bias) (ReLU if
 * relu != 0); row_beg/row_cnt/col: block-diagonal CSR of sga_csr_build (self loops removed, one appended per node =
 * add_remaining_self_loops); deg [N] = in-degree incl. the self loop = row_cnt of the FORWARD (by-destination) CSR.
 * The backward of the aggregation is the same call over the CSR of the reversed edges with the same deg. */
int sga_gcn_aggregate(const float* h, int64_t N, int C, const int32_t* row_beg, const int32_t* row_cnt,
                      const int32_t* col, const int32_t* deg, const float* bias, int relu, float* out, void* stream);
/* Y [N,C] = X [N,K] W^T, W [C,K], K <= 8 (GCN layer 1: 3 -> 200, gat.py:15); gW [C,K] += G^T X. */
int sga_linear_smallk(const float* X, int64_t N, int K, const float* W, int C, float* Y, void* stream);
int sga_wgrad_smallk(const float* G, const float* X, int64_t N, int K, int C, float* gW, void* stream);
/* out = g * (y > 0)  (ReLU backward, gat.py:23);  s [C] += column sums of x [N,C] (bias gradients). */
int sga_relu_mask(const float* g, const float* y, int64_t n, float* out, void* stream);
int sga_colsum_rows(const float* x, int64_t N, int C, float* s, void* stream);
/* F.normalize (losses.py:193): norms [N] = max(||x_i||, eps);  backward in place on g [N,D]:
 * g_i <- (g_i - xh_i (xh_i . g_i)) / norms_i. */
int sga_row_l2norm(const float* x, int64_t N, int D, float eps, float* norms, void* stream);
int sga_normalize_bwd_rows(const float* x, const float* norms, int64_t N, int D, float eps, float* g, void* stream);
/* MultiModalFusion (sg_aligner.py:23-35) for M <= 8 modalities of different widths dims[m] (EVA fuses raw 400-d / 200-d
 * / 100-d embeddings, eva.py:72-76): joint [N, ld] = cat_m softmax(fusion_w)_m * x_m / max(||x_m||, 1e-12).
 * x_host / gx_host: HOST arrays of M device pointers.  Backward: gx_m written, g_fusion_w [M] accumulated (+=);
 * scratch_M: M floats of device scratch. */
int sga_fuse_rows_fwd(const float* const* x_host, const int* dims_host, int M, const float* fusion_w, int64_t N,
                      float* joint, int ld, void* stream);
int sga_fuse_rows_bwd(const float* const* x_host, const int* dims_host, int M, const float* fusion_w, int64_t N,
                      const float* g_joint, int ld, float* const* gx_host, float* g_fusion_w, float* scratch_M, void* stream);
/* NCALoss (losses.py:161-176) over the score matrix S [A,A] = src_emb ref_emb^T: rs/cs [A] = row / column sums of
 * exp(alpha (S - ep)) off the diagonal, diag [A], loss [1] = mean log(1+cs)/alpha + mean log(1+rs)/alpha
 * - beta mean log(1 + relu(diag)).  sga_nca_coef turns S into dLoss/dS in place. */
int sga_nca_forward(const float* S, int A, float alpha, float beta, float ep, float* rs, float* cs, float* diag,
                    float* loss, void* stream);
int sga_nca_coef(float* S, int A, float alpha, float beta, float ep, const float* rs, const float* cs, void* stream);

/* diagnostics only: the tensor-core forward, when it tracks the argmax, re-evaluates near-ties of the max-pool in
 * fp32 (pointnet_tie_fix_kernel, so that the gradient is routed through the point the fp32 reference selects,
 * pointnet.py:162-163).  counts_host[2] (HOST pointer, may be NULL) receives {near-ties re-evaluated, of those
 * reordered} since the last reset; reset != 0 zeroes the counters.  Synchronises the device. */
int sga_debug_tie_stats(unsigned long long* counts_host, int reset);

/* diagnostics only: device buffer of >= 2048 int64 that CTA 0 of the tensor-core PointNet kernel fills
 * with clock64() stamps of its pipeline events (profiles/trace_pointnet.py); NULL switches it off. */
int sga_debug_set_trace(long long* trace);

/* ---- bring-up / self-test kernels (tests only): one 128xN tile D = A B^T through tcgen05 with
 * split operands.  kind 0: bf16x3, kind 1: tf32x3.  A [128,K], B [Ncols,K] f32 row-major,
 * D [128,Ncols] f32. */
int sga_selftest_umma(const float* A, const float* B, float* D, int Ncols, int K, int kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGA_B200_H */
