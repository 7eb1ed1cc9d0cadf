/*
 * odwscl.h -- C ABI of libodwscl_sm100.so: the B200 (sm_100a) kernels behind the OD-WSCL
 * proposal-feature hot path.  This is the drop-in boundary: every entry point replaces a
 * reference interface (cited per function, paths relative to the reference's wetectron/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  All data pointers are DEVICE
 *     pointers to contiguous row-major buffers owned by the caller; the library allocates
 *     nothing and keeps no global mutable state.  Scratch space is passed in (`ws`), sized by
 *     the matching *_ws_bytes() query.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Kernels are only
 *     enqueued; nothing synchronises the host.
 *   - Return value: 0 on success, a cudaError_t value (>0) when a launch fails, or a negative
 *     ODWSCL_E* code for argument errors.  odwscl_strerror() turns either into text.  The Python
 *     binding raises RuntimeError on non-zero (the reference raises through AT_ASSERTM /
 *     THCudaCheck, csrc/cuda/ROIPool_cuda.cu:115-116,151).
 *   - Empty problems (R == 0, n == 0 ...) return 0 without launching
 *     (csrc/cuda/ROIPool_cuda.cu:132-135,180-183).
 */
#ifndef ODWSCL_H_
#define ODWSCL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODWSCL_VERSION 100
#define ODWSCL_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported shape) */
#define ODWSCL_ENOWS  (-2)   /* workspace too small */
#define ODWSCL_SUPCON_SPLITS 8   /* column splits of the SupCon tile kernels (stats scratch size) */
#define ODWSCL_SIM_DIM 128   /* embedding width of Sim_Net (sim_head/sim_net.py:13-16) */

typedef void* odwscl_stream_t;

int odwscl_version(void);
const char* odwscl_strerror(int code);
/* SMs the persistent CTA-pair kernels (conv3x3, fc GEMM) leave unoccupied (rounded down to whole pairs; default 0): with
 * data-parallel training the NCCL gradient all-reduce (tools/train_net.py:50-55) runs beside the backward kernels and
 * otherwise waits for a persistent kernel to retire before it gets an SM.  Process-wide; not a per-call argument. */
int odwscl_set_sm_margin(int sms);

/* ---- A3: ROIPool forward.  Replaces _C.roi_pool_forward (csrc/ROIPool.h:11-24 ->
 * csrc/cuda/ROIPool_cuda.cu:16-77,110-153).  feat [B,C,H,W] fp32 NCHW; rois [R,5] =
 * (batch, x1,y1,x2,y2) image pixels; out [R,C,ph,pw] fp32; argmax [R,C,ph,pw] int32 =
 * h*W+w inside the (b,c) plane or -1.  Bit-exact with the reference rule (Appendix B).
 * The 7x7 / C%4==0 fast path transposes the map to channels-last inside `ws`. */
size_t odwscl_roi_pool_fwd_ws_bytes(int B, int C, int H, int W, int R, int ph, int pw);
int odwscl_roi_pool_fwd_f32(const float* feat, int B, int C, int H, int W, const float* rois, int R,
                            float scale, int ph, int pw, float* out, int32_t* argmax,
                            void* ws, size_t ws_bytes, odwscl_stream_t stream);

/* ---- A4: ROIPool backward.  Replaces _C.roi_pool_backward (csrc/ROIPool.h:26-45 ->
 * csrc/cuda/ROIPool_cuda.cu:79-108,156-202).  grad_in [B,C,H,W] is zeroed inside. */
int odwscl_roi_pool_bwd_f32(const float* grad_out, const int32_t* argmax, const float* rois, int R,
                            int B, int C, int H, int W, int ph, int pw, float* grad_in,
                            odwscl_stream_t stream);

/* Channels-last variants of A3 / A4 for the model path (the conv stack below is NHWC): feat_nhwc and
 * grad_in_nhwc are [B,H,W,C]; out / argmax / grad_out keep [R,C,7,7].  7x7 bins, C % 4 == 0. */
int odwscl_roi_pool_fwd_nhwc_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                 float scale, float* out, int32_t* argmax, odwscl_stream_t stream);
int odwscl_roi_pool_bwd_nhwc_f32(const float* grad_out, const int32_t* argmax, const float* rois, int R, int B,
                                 int C, int H, int W, float* grad_in_nhwc, odwscl_stream_t stream);
/* A4 for a pooled tensor with several consumers (weak_head.py:107-120): up to two dense gradients [R,C,7,7] and
 * one sparse one (srows [S] int64 roi indices, sgrad [S,C,7,7]) are summed while they are scattered.  mask2 (optional,
 * [R,49]) multiplies grad_out2 per (roi, bin): with the block mask x scale written by odwscl_dropblock_mask_f32 the
 * DropBlock backward of the augmented consumer happens inside the scatter.  EINVAL when H*W*4 B exceeds the
 * shared-memory plane (caller sums and uses odwscl_roi_pool_bwd_nhwc_f32). */
int odwscl_roi_pool_bwd_nhwc_multi_f32(const float* grad_out, const float* grad_out2, const float* mask2,
                                       const int64_t* srows, const float* sgrad, int S, const int32_t* argmax, const float* rois, int R,
                                       int B, int C, int H, int W, float* grad_in_nhwc, odwscl_stream_t stream);

/* ---- A5: ROIAlign (legacy, non-aligned).  Replaces _C.roi_align_forward/backward
 * (csrc/ROIAlign.h:11-45 -> csrc/cuda/ROIAlign_cuda.cu:64-122,177-254). */
int odwscl_roi_align_fwd_f32(const float* feat, int B, int C, int H, int W, const float* rois, int R,
                             float scale, int ph, int pw, int sampling_ratio, float* out,
                             odwscl_stream_t stream);
int odwscl_roi_align_bwd_f32(const float* grad_out, const float* rois, int R, float scale, int ph,
                             int pw, int B, int C, int H, int W, int sampling_ratio, float* grad_in,
                             odwscl_stream_t stream);

/* ---- A9: pairwise IoU.  plus_one=1 replaces boxlist_iou (structures/boxlist_ops.py:127-160,
 * legacy +1 pixel convention); plus_one=0 is torchvision's convention.  out [na,nb]. */
int odwscl_box_iou_f32(const float* a, int na, const float* b, int nb, int plus_one, float* out,
                       odwscl_stream_t stream);

/* ---- A10: NMS with torchvision.ops.nms semantics (structures/boxlist_ops.py:57): stable
 * descending sort, IoU without +1, suppress iff IoU > thr, kept ids in descending-score order.
 * keep [n] int64 and n_keep [1] int32 are device buffers; no host sweep, no sync.  n <= 8192. */
int odwscl_nms_f32(const float* boxes, const float* scores, int n, float thr, int64_t* keep,
                   int32_t* n_keep, odwscl_stream_t stream);
/* `_C.nms` (csrc/nms.h:10-28 -> csrc/cuda/nms.cu:23-130): +1 convention, suppress iff IoU > thr,
 * kept ids returned ASCENDING. */
int odwscl_nms_legacy_f32(const float* boxes, const float* scores, int n, float thr, int64_t* keep,
                          int32_t* n_keep, odwscl_stream_t stream);

/* ---- A11: object discovery (roi_heads/weak_head/loss.py:271-345).
 * A "pair" is one (image b, positive class c); pairs are listed image-major, class-ascending.
 *   boxes [R,4]; img_off [B+1] row offsets; scores[3] = the three supervisor score tensors
 *   [R,C] (final_score, softmax(ref1), softmax(ref2)); pair_img / pair_cls [P] (cls 0-based,
 *   background excluded); Ncap = max proposals per image.
 * Phase A (loss.py:281-307): per pair the union over the 3 branches of
 *   { j : IoU+1(P[j], P[argmax_j score]) >= thres }.
 *   out: amax [P,3] int32 (image-local argmax per branch), member [P,Ncap] uint8,
 *        cntA [P] int32, offA [P+1] int32 (exclusive prefix of cntA; offA[P] = K),
 *        rowsA [>= sum cntA] int32 GLOBAL row ids, pair-major ascending,
 *        hardA [same] fp32 = S0[row,c+1] / sum_j S0[j,c+1] (loss.py:294), colsum [P] fp32. */
int odwscl_discover_phase_a_f32(const float* boxes, const int32_t* img_off, int B, int R, int C,
                                const float* s0, const float* s1, const float* s2,
                                const int32_t* pair_img, const int32_t* pair_cls, int P, int Ncap,
                                float thres, int32_t* amax, uint8_t* member, int32_t* cntA,
                                int32_t* offA, int32_t* rowsA, float* hardA, float* colsum,
                                odwscl_stream_t stream);
/* Phase B (loss.py:311-345): per pair and branch: similarity rows of the top proposal(s),
 * tau = mean(F[m] . coll[c]^T), the `>= tau` / bool-vs-float rule, torchvision NMS, fallback,
 * set difference against the running membership.  F [R,128]; E [2K,128] = embeddings of the
 * drop / noise augmented Phase-A positives (first K rows drop, next K rows noise, both in
 * rowsA order).  out: inst [P,3,Ncap] int32 image-local ids in descending-score order +
 * inst_cnt [P,3]; newl [P,3,Ncap] int32 ascending + new_cnt [P,3]; hardB [P,3,Ncap] fp32;
 * tau_out [P,3] fp32 (diagnostic).  sim_rows_in: optional [P,3,Ncap] fp32 override of the
 * m-row similarities (stage-wise parity tests feed the oracle's rows); NULL in production. */
int odwscl_discover_phase_b_f32(const float* boxes, const int32_t* img_off, int B, int R, int C,
                                const float* s0, const float* s1, const float* s2,
                                const int32_t* pair_img, const int32_t* pair_cls, int P, int Ncap,
                                const float* F, const float* E, const int32_t* amax,
                                uint8_t* member, const int32_t* cntA, const int32_t* offA,
                                const int32_t* rowsA, const float* colsum, float nms_thr,
                                int32_t* inst, int32_t* inst_cnt, int32_t* newl, int32_t* new_cnt,
                                float* hardB, float* tau_out, const float* sim_rows_in,
                                odwscl_stream_t stream);
/* Bank assembly for SupConLossV2 (sim_head/sim_loss.py:55-58 + loss.py:290-345 append order):
 * rows class-major, weights in execution order (the reference's misalignment is reproduced).
 * row_src [Mcap] int32: < R -> row of F, else R + row of E; row_lab [Mcap] int32; row_w [Mcap];
 * M_out [2] int32: the rows written (min(M, Mcap)) and the unclamped M. */
int odwscl_bank_assemble(const int32_t* pair_img, const int32_t* pair_cls, int P, int B, int R,
                         int Ncap, int num_fg_classes, const int32_t* img_off, const int32_t* cntA,
                         const int32_t* offA, const int32_t* rowsA, const float* hardA,
                         const int32_t* newl, const int32_t* new_cnt, const float* hardB, int Mcap,
                         int32_t* row_src, int32_t* row_lab, float* row_w, int32_t* M_out,
                         odwscl_stream_t stream);

/* ---- A12: SupConLossV2 forward / backward (sim_head/sim_loss.py:49-80), fused: the M x M
 * similarity tile product, masked exp row sums and the weighted log ratio never leave the SM.
 * V = [F ; E] addressed through row_src; M read from device (M_dev) so no host sync is needed.
 * stats [(1 + ODWSCL_SUPCON_SPLITS) * Mcap, 4] fp32: rows [0,Mcap) = (row max, pos sum, all sum, row loss), the rest is
 * scratch for the per-column-split partials (the column loop is split over ODWSCL_SUPCON_SPLITS CTAs per row tile and
 * merged in a fixed order); loss_out [1] fp32 = mean_r(...). */
int odwscl_supcon_fwd_f32(const float* F, const float* E, int R, const int32_t* row_src,
                          const int32_t* row_lab, const float* row_w, const int32_t* M_dev, int Mcap,
                          float inv_temp, float* stats, float* loss_out, odwscl_stream_t stream);
/* dF [R,128] and dE [nE,128] are ACCUMULATED into (caller zeroes); gscale = upstream grad. */
int odwscl_supcon_bwd_f32(const float* F, const float* E, int R, const int32_t* row_src,
                          const int32_t* row_lab, const float* row_w, const int32_t* M_dev, int Mcap,
                          float inv_temp, const float* stats, const float* gscale_dev, float* dF,
                          float* dE, odwscl_stream_t stream);
/* The same loss and gradient with both M x M contractions on the tensor cores (3xTF32 through odwscl_fc_gemm_tf32: the bank
 * rows are gathered once as [Vh | Vh | Vl] / [Vh | Vl | Vh] so S = V V^T is ONE K = 384 GEMM; dV = H V likewise): for the
 * banks of 8 images per rank (M = 4-6 k), where the tile kernels above grow quadratically on the FFMA pipe.  `ws`
 * (odwscl_supcon_tc_ws_bytes(Mcap) bytes, 16-byte aligned) holds S between the forward and the backward call of a bank;
 * `stats` as above. */
size_t odwscl_supcon_tc_ws_bytes(int Mcap);
int odwscl_supcon_tc_fwd_f32(const float* F, const float* E, int R, const int32_t* row_src, const int32_t* row_lab,
                             const float* row_w, const int32_t* M_dev, int Mcap, float inv_temp, float* ws, size_t ws_bytes,
                             float* stats, float* loss_out, odwscl_stream_t stream);
int odwscl_supcon_tc_bwd_f32(int R, const int32_t* row_src, const int32_t* row_lab, const float* row_w, const int32_t* M_dev,
                             int Mcap, float inv_temp, float* ws, size_t ws_bytes, const float* stats,
                             const float* gscale_dev, float* dF, float* dE, odwscl_stream_t stream);

/* ---- A13: od_layer (weak_head/pseudo_label_generator.py:135-197): per (image, branch) the
 * pseudo-GT set from `inst`, IoU+1 N x G with first-max argmax on device (no numpy round trip),
 * labels (bg iff max <= fg_thr), weights, BoxCoder(10,10,5,5).encode targets
 * (modeling/box_coder.py:22-50).  labels [3,R] int64, weights [3,R], targets [3,R,4]. */
int odwscl_od_layer_f32(const float* boxes, const int32_t* img_off, int B, int R, int C,
                        const float* s0, const float* s1, const float* s2, const int32_t* pair_img,
                        const int32_t* pair_cls, int P, int Ncap, const int32_t* inst,
                        const int32_t* inst_cnt, float fg_thr, int64_t* labels, float* weights,
                        float* targets, odwscl_stream_t stream);

/* ---- A6 (elementwise part): nn.ReLU(True) + nn.Dropout(p) of the fc6 / fc7 outputs (modeling/backbone/vgg16.py:
 * 122-130) as ONE in-place pass: x <- relu(x) * keep / (1-p), keep ~ Bernoulli(1-p) from Philox4x32-10(seed, index/4).
 * n % 4 == 0.  The backward needs no mask: gx = gy / (1-p) where the saved output y > 0, else 0. */
int odwscl_relu_dropout_fwd_f32(float* x, long long n, float p, unsigned long long seed, odwscl_stream_t stream);
int odwscl_relu_dropout_bwd_f32(const float* y, const float* gy, float* gx, long long n, float p,
                                odwscl_stream_t stream);

/* ---- A1 (operand layouts): the reference's [Cout,Cin,3,3] weights -> [Cout,3,3,Cin] (fprop operand, w_krsc) and the
 * tap-flipped [Cin,3,3,Cout] (dgrad operand, w_crsk_flip) in one pass, optionally TF32-rounded (cvt.rna).  Either
 * output may be null. */
int odwscl_conv_weight_xform_f32(const float* w_oihw, int Cout, int Cin, float* w_krsc, float* w_crsk_flip,
                                 int round_tf32, odwscl_stream_t stream);

/* ---- A3 + A15 fused (N1): ROIPool forward that ALSO writes the DropBlock-augmented copy of the pooled features
 * (weak_head.py:107 + :111: out_aug = out * aug_mask[r, bin], aug_mask = block_mask * numel/sum from
 * odwscl_dropblock_prepare_f32) from the same staged values -- the separate DropBlock pass re-read what this kernel wrote. */
int odwscl_roi_pool_fwd_nhwc_aug_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                     float scale, float* out, int32_t* argmax, const float* aug_mask, float* out_aug,
                                     odwscl_stream_t stream);
/* scale_io [2] = (sum(block_mask), numel/sum) and mask_out [R, ph*pw] = block_mask * scale from the centre mask alone. */
int odwscl_dropblock_prepare_f32(const float* centres, int R, int ph, int pw, int block, float* scale_io,
                                 float* mask_out, odwscl_stream_t stream);

/* Measurement aid, not part of the path: loads every cell of every roi's clamped extent ONCE per (roi, 128-channel slab)
 * with the pooling kernel's access pattern and reduces it with a plain max -- the time the L2 -> SM fabric charges for the
 * bytes of a "stage the roi region once" ROIPool design.  out: R * ceil(C/128) * 7 floats. */
int odwscl_probe_roi_stream_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                float scale, float* out, odwscl_stream_t stream);

/* ---- A5, channels-last: legacy ROIAlign (csrc/cuda/ROIAlign_cuda.cu:64-122,177-254) on an NHWC map [B,H,W,C]
 * (C % 4 == 0), 7x7 bins, output / grad_out in the reference's [R,C,7,7] layout.  Forward: 16-byte loads serving 4
 * channels per bilinear corner, sample weights once per (roi, bin, sample).  Backward: shared-memory accumulation of a
 * 4-channel plane per CTA (ODWSCL_ENOWS when 16*H*W bytes exceed shared memory: use the NCHW entry point). */
int odwscl_roi_align_fwd_nhwc_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                  float scale, int sampling_ratio, float* out, odwscl_stream_t stream);
int odwscl_roi_align_bwd_nhwc_f32(const float* grad_out, const float* rois, int R, float scale, int B, int C, int H,
                                  int W, int sampling_ratio, float* grad_in_nhwc, odwscl_stream_t stream);

/* ---- A6 / A7 / N1: the fully-connected block -- fc6 + fc7 (modeling/backbone/vgg16.py:122-130,148-162), Sim_Net
 * (roi_heads/sim_head/sim_net.py:10-26) and the MIST predictor heads (roi_heads/weak_head/roi_weak_predictors.py:
 * 158-165); replaces the cuBLAS GEMMs behind nn.Linear forward / backward plus the separate ReLU, Dropout and
 * gradient-accumulation kernels.  One persistent CTA-pair tcgen05 kernel (TF32 math, fp32 accumulate in TMEM):
 *     C[M,N] (+)= sum_k A(m,k) * B(n,k)
 * A is [M,K] row-major (a_mn_major = 0, pitch lda) or [K,M] row-major (a_mn_major = 1); B is [N,K] (b_mn_major = 0,
 * pitch ldb) or [K,N] (b_mn_major = 1) -- so forward (X W^T), input gradient (dY W) and weight gradient (dY^T X)
 * all read the tensors where they lie.  Pitches in floats, multiples of 4; A, B, C, mask_src 16-byte aligned.
 * flags (fused epilogue, applied in this order):
 *   ODWSCL_FC_BIAS     + bias[n]
 *   ODWSCL_FC_ACCUM    + the previous content of C (3xTF32 passes; folding a second weight gradient, beta = 1)
 *   ODWSCL_FC_RELU     max(., 0)
 *   ODWSCL_FC_DROPOUT  nn.Dropout(dropout_p): Philox4x32-10 keyed by (seed, row * N + col); survivors x 1/(1-p)
 *   ODWSCL_FC_MASK     x mask_scale where mask_src[m, n] > 0, else 0 (the ReLU+Dropout derivative of the layer below:
 *                      mask_src = its saved output, mask_scale = 1/(1-p))
 *   ODWSCL_FC_ROUND    round to TF32 (cvt.rna) -- for outputs another tensor-core GEMM consumes
 * max_pairs > 0 caps the resident CTA pairs (leave SMs to a concurrent NCCL all-reduce); 0 = all 74.
 * (A2, B2, K2 > 0): a second operand pair of the same majors whose contraction is appended to the first's,
 * C = A B^T + A2 B2^T in ONE accumulation -- the weight gradient of a layer applied twice in a step (dW = dY1^T X1 +
 * dY2^T X2: the [2R]-row batch and the augmented positives of loss.py:299-310) without a second gradient tensor. */
#define ODWSCL_FC_BIAS 1
#define ODWSCL_FC_ACCUM 2
#define ODWSCL_FC_RELU 4
#define ODWSCL_FC_DROPOUT 8
#define ODWSCL_FC_MASK 16
#define ODWSCL_FC_ROUND 32
int odwscl_fc_gemm_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, float* C,
                        int ldc, int M, int N, int K, int flags, const float* bias, const float* mask_src,
                        int ld_mask, float mask_scale, float dropout_p, unsigned long long seed, int max_pairs,
                        const float* A2, int lda2, const float* B2, int ldb2, int K2, odwscl_stream_t stream);
/* The same product (no epilogue flags) whose result is ADDED, times out_scale, to every rank's replica of C through the
 * NVSwitch multicast alias C_multicast of a symmetric-memory allocation (multimem.red.add.v4.f32 from the epilogue): the
 * weight gradient and its cross-rank sum in one kernel.  Replaces, for the fc6 / fc7 weights, the bucket all-reduce
 * DistributedDataParallel runs after the cuBLAS weight gradient (tools/train_net.py:50-55).  The caller zeroes C on
 * every rank and synchronises the ranks before the launch and before reading C (od-wscl_b200/sharding.py PeerGradSum). */
int odwscl_fc_gemm_peer_sum_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, float* C,
                                 float* C_multicast, int ldc, int M, int N, int K, float out_scale, int max_pairs,
                                 const float* A2, int lda2, const float* B2, int ldb2, int K2, odwscl_stream_t stream);
/* Reduce-scatter form of the above for larger worlds: rows [r * rows_per_owner, (r + 1) * rows_per_owner) of the product
 * are added (red.add.v4.f32 over NVLink) ONLY to rank r's replica peer_C[r] (host array of n_peers <= 8 device pointers,
 * peer-mapped; rows_per_owner % 32 == 0), so a rank receives (world-1)/world of the gradient instead of (world-1) copies.
 * odwscl_peer_broadcast_f32 is the all-gather half: every rank's dst[i] = src[i] through the multicast alias. */
int odwscl_fc_gemm_peer_scatter_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, float* C,
                                     const void* const* peer_C, int n_peers, int rows_per_owner, int ldc, int M, int N, int K,
                                     float out_scale, int max_pairs, const float* A2, int lda2, const float* B2, int ldb2,
                                     int K2, odwscl_stream_t stream);
int odwscl_peer_broadcast_f32(const float* src, float* dst_multicast, long long n, odwscl_stream_t stream);
/* every rank's dst[i] += scale * src[i] through the multicast alias (n % 4 == 0, 16-byte aligned). */
int odwscl_peer_add_f32(const float* src, float* dst_multicast, long long n, float scale, odwscl_stream_t stream);
/* out[c] (+)= sum_r x[r, c] over a [rows, cols] matrix of pitch ld: the bias gradients of the block above
 * (accumulate != 0 adds to out, which folds a second call's gradient). */
int odwscl_colsum_f32(const float* x, long long rows, int cols, int ld, float* out, int accumulate,
                      odwscl_stream_t stream);

/* ---- A13 / N2: the MIL + refinement side of RoIRegLossComputation.__call__ (roi_heads/weak_head/loss.py:233-259,
 * 349-406) over the ONE logits buffer the eight predictor heads leave: logits [R, ld], column blocks
 * [cls C][det C][ref1 C][bbox1 Q][ref2 C][bbox2 Q][ref3 C][bbox3 Q] (Q = 4*C, or 8 when class-agnostic); C <= 96.
 * img_off [B+1] int32 = first row of every image.  Replaces ~280 eager torch kernels (forward + autograd backward).
 *
 * odwscl_head_scores_f32 (before object discovery): final_score [R,C] = softmax_c(cls) * softmax_over_the_image's_
 * proposals(det) (loss.py:234-246), sm1 / sm2 [R,C] = softmax_c(ref1 / ref2) (the supervisors of branches 1, 2:
 * loss.py:283,313), img_score [B,C] = per-image column sums of final_score (loss.py:352); det_max / det_sum [B,C]
 * (softmax statistics, reused by the loss call) and ref_colsum [3,B,C] (per-image sums of the refinement logits:
 * the accuracy metric of loss.py:25-34) are scratch outputs; ws: odwscl_head_scores_ws_bytes(B, C) bytes (partial
 * column statistics of the two-stage, fixed-order reductions). */
size_t odwscl_head_scores_ws_bytes(int B, int C);
int odwscl_head_scores_f32(const float* logits, int ld, int R, int C, int Q, const int32_t* img_off, int B,
                           float* det_max, float* det_sum, float* ref_colsum, float* final_score, float* sm1,
                           float* sm2, float* img_score, void* ws, size_t ws_bytes, odwscl_stream_t stream);
/* odwscl_head_loss_f32 (after od_layer): pseudo_labels [3,R] int64, label_weights [3,R], reg_targets [3,R,4] as
 * written by odwscl_od_layer_f32; img_labels [B,C] multi-hot.  out11 = (loss_img, loss_ref_cls0, loss_ref_reg0,
 * loss_ref_cls1, loss_ref_reg1, loss_ref_cls2, loss_ref_reg2, acc_img, acc_ref0, acc_ref1, acc_ref2), every entry
 * already divided by B (loss.py:403-406).  grad_logits [R, ld] receives d(sum of the seven losses)/d(logits) in closed
 * form (every column of the 5C+3Q block is written).  partial: scratch of ceil(R/8)*6 floats. */
int odwscl_head_loss_f32(const float* logits, int ld, int R, int C, int Q, int cls_agnostic, const int32_t* img_off,
                         int B, const float* det_max, const float* det_sum, const float* img_score,
                         const float* img_labels, const int64_t* pseudo_labels, const float* label_weights,
                         const float* reg_targets, const float* ref_colsum, float eps, float* grad_logits,
                         float* partial, float* out11, odwscl_stream_t stream);
/* The backward of that node: column block of loss k scaled in place by upstream7[k] (device; order as out11[0..6]). */
int odwscl_head_grad_scale_f32(float* grad_logits, int ld, long long R, int C, int Q, const float* upstream7,
                               odwscl_stream_t stream);

/* ---- A7: F.normalize(x, dim=1) at the end of Sim_Net (roi_heads/sim_head/sim_net.py:26; eps 1e-12), forward and
 * backward as one launch each.  z [R, D] with row pitch ldz; y [R, D] dense; inv_norm [R] = 1 / max(||z||, eps). */
int odwscl_l2norm_fwd_f32(const float* z, int ldz, int R, int D, float eps, float* y, float* inv_norm,
                          odwscl_stream_t stream);
int odwscl_l2norm_bwd_f32(const float* y, const float* g, const float* inv_norm, int R, int D, float eps, float* dz,
                          odwscl_stream_t stream);
/* Index bookkeeping of the sync-free contrastive branch (the augmented-positives batch padded to a bound Kc on the
 * device-resident count *k_dev of Phase-A positives, loss.py:299-310): rows [Kc] = proposal row of every batch slot
 * (0 for padding), sel [sel_n] = for every entry of the [2K] (drop rows, noise rows) layout the discovery kernels
 * address, its row in the padded [2 Kc] embedding matrix, *overflow = (K > Kc). */
int odwscl_spec_index(const int32_t* k_dev, const int32_t* rowsA, int Kc, long long sel_n, int64_t* rows, int64_t* sel,
                      float* overflow, odwscl_stream_t stream);

/* ---- A15: DropBlock2D apply (modeling/dropblock/drop_block.py:29-66) with a device-sampled
 * centre mask [R,ph,pw] (1.0 = drop centre): block mask by block x block dilation, global
 * renormalisation numel/sum, y = x * mask * scale in ONE pass over x [R,C,ph,pw].
 * scale_io [2] fp32: [0] = sum(block_mask), [1] = numel/sum.  reuse_scale != 0 skips the
 * reduction and applies the stored scale (the backward: dx = dy * mask * scale). */
int odwscl_dropblock_f32(const float* x, const float* centres, int R, int C, int ph, int pw,
                         int block, float* y, float* scale_io, int reuse_scale,
                         odwscl_stream_t stream);
/* mask_out [R,ph*pw] = block_mask * scale_io[1] (the per-(roi, bin) factor of the forward AND of the backward). */
int odwscl_dropblock_mask_f32(const float* centres, int R, int ph, int pw, int block, const float* scale_io,
                              float* mask_out, odwscl_stream_t stream);
/* Same over a PADDED batch: only the first *n_valid_dev rows (a device-resident count, <= R) take part in the
 * renormalisation and are written; rows past it are zero-filled.  Lets the caller size the batch from an upper
 * bound without reading the count back (no host synchronisation in the contrastive branch, loss.py:299-310). */
int odwscl_dropblock_rows_f32(const float* x, const float* centres, int R, int C, int ph, int pw,
                              int block, float* y, float* scale_io, int reuse_scale,
                              const int32_t* n_valid_dev, odwscl_stream_t stream);

/* Same over a batch made of P row SEGMENTS (seg_off_dev [P+1] int32, device-resident, ascending; segment p = rows
 * [seg_off[p], seg_off[p+1])): every segment is renormalised by its own numel/sum -- the arithmetic of the reference's
 * per-(image, class) drop_pool calls (roi_heads/weak_head/loss.py:299, modeling/backbone/vgg16.py:173-175) in one
 * launch pair.  Rows >= seg_off[P] are padding and are zero-filled.  scale_seg [P,2] fp32 = (sum, numel/sum) per
 * segment; reuse_scale != 0 applies the stored scales (the backward). */
int odwscl_dropblock_seg_f32(const float* x, const float* centres, int R, int C, int ph, int pw, int block,
                             float* y, const int32_t* seg_off_dev, int P, float* scale_seg, int reuse_scale,
                             odwscl_stream_t stream);

/* ---- A10 for inputs beyond the single-CTA kernels' 8192 boxes (torchvision.ops.nms has no size limit; the UNION
 * merge of test-time views, engine/bbox_aug.py:11-141, concatenates every view's proposals): stable rank by counting,
 * 64 x 64 suppression bit masks, one-CTA sweep -- all on the device.  boxes are read as 4 consecutive floats at
 * boxes + i * box_stride (box_stride % 4 == 0, 16-byte aligned: a class column of a [N, C*4] tensor needs no copy),
 * scores at scores + i * score_stride; only candidates with score > score_thr take part (-INFINITY: all).
 * legacy == 0: torchvision semantics (no +1, IoU > thr), kept indices in DESCENDING score order into keep64 or keep32
 * (exactly one non-null); legacy != 0: `_C.nms` as csrc/cuda/nms.cu (+1, IoU > thr), ascending original indices into keep64.
 * ws: odwscl_nms_large_ws_bytes(n) bytes of scratch (order + n * ceil(n/64) mask words). */
size_t odwscl_nms_large_ws_bytes(int n);
int odwscl_nms_large_f32(const float* boxes, int box_stride, const float* scores, int score_stride, int n,
                         float score_thr, float thr, int legacy, int64_t* keep64, int32_t* keep32, int32_t* n_keep,
                         void* ws, size_t ws_bytes, odwscl_stream_t stream);

/* ---- the augmented positives of the contrastive branch (roi_heads/weak_head/loss.py:296-305) in one pass: for slot k
 * (proposal rows[k] of pooled [R, D], D = C * cells, cells = 7 * pw) out[k] = DropBlock(block x block) view, renormalised
 * per segment (seg_off_dev [P+1]; modeling/backbone/vgg16.py:173-175), out[Kc + k] = eps * x + x, eps ~ N(0,1) (vgg16.py:
 * 177-180; Philox keyed by (seed, k * D + e), or `noise` [Kc, D] when given).  Slots >= seg_off[P] are padding (zeros).
 * compute_scale != 0 fills scale_seg [P,2] from the centres first.  backward != 0: src = gradient of out [2 Kc, D],
 * dst [Kc, D] = gradient w.r.t. the gathered pooled rows (same masks, same noise). */
int odwscl_aug_positives_f32(const float* src, int D, int cells, const int64_t* rows, int Kc, const int32_t* seg_off_dev,
                             int P, const float* centres, int block, float* scale_seg, int compute_scale,
                             const float* noise, unsigned long long seed, int backward, float* dst,
                             odwscl_stream_t stream);

/* ---- N4 (test time): PostProcessor.filter_results (roi_heads/box_head/inference.py:216-258) -- for every foreground
 * class j in [1,C): candidates with scores[i,j] > score_thr, torchvision-semantics NMS at `thr` on boxes[i, 4j..4j+3],
 * all classes in ONE launch (one CTA per class, sort + sweep in shared memory, no host round trip).  boxes [N,C*4],
 * scores [N,C]; keep [C,N] int32 (row j = kept proposal indices, descending score), n_keep [C] (n_keep[0] = 0).
 * N <= 8192. */
int odwscl_nms_per_class_f32(const float* boxes, const float* scores, int n, int C, float score_thr, float thr,
                             int32_t* keep, int32_t* n_keep, odwscl_stream_t stream);

/* ---- A11 (drop-in for loss.py:319): full N x N similarity F F^T on the tensor cores
 * (tcgen05.mma kind::tf32 fed by TMA, accumulators in TMEM; 3xTF32 operand split so the result is
 * fp32-accurate).  out [N,N] fp32; ws sized by odwscl_sim_nxn_ws_bytes(N). */
size_t odwscl_sim_nxn_ws_bytes(int N);
int odwscl_sim_nxn_f32(const float* F, int N, float* out, void* ws, size_t ws_bytes, odwscl_stream_t stream);

/* ---- tensor-core building block: C[M,N] = A[M,K] * B[N,K]^T, single-pass TF32 (what torch's
 * allow_tf32 matmul computes; modeling/backbone/vgg16.py:151,161, sim_net.py:26 run through
 * cuBLAS in the reference).  K % 4 == 0, ldc >= N, 16-byte aligned A and B. */
int odwscl_gemm_nt_tf32(const float* A, const float* B, float* C, int M, int N, int K, int ldc,
                        odwscl_stream_t stream);

/* ---- A1: the VGG16-OICR conv stack (modeling/backbone/vgg16.py:26-36,58-83; cuDNN via nn.Conv2d in
 * the reference), channels-last.  3x3 / stride 1 / padding = dilation ("same") convolution as a tcgen05
 * implicit GEMM fed by TMA (zero-filled out-of-image boxes are the padding):
 *   y[b,h,w,co] = epilogue( sum_{r,s,ci} x[b, h+(r-1)d, w+(s-1)d, ci] * w_krsc[co,r,s,ci] )
 * x [B,H,W,Cin], w_krsc [Cout,3,3,Cin], y [B,H,W,Cout]; Cin % 32 == 0, Cout % 32 == 0.  Single-pass TF32
 * inputs, fp32 accumulation.  flags: 1 = ReLU, 2 = accumulate into y (y += conv; used by the 3-pass
 * strict-fp32 mode), 4 = zero y where mask_src <= 0 (mask_src [B,H,W,Cout]: fused ReLU derivative, which
 * makes this entry point the DGRAD as well: x = dY, w_krsc = flipped/transposed weights).  Order of the
 * epilogue: + bias, + previous y, ReLU, mask, round.  bias may be NULL.  8 = round the result to TF32
 * (round-to-nearest): the tensor core TRUNCATES fp32 operands to 10 mantissa bits, which biases every
 * product low; activations that feed another convolution are therefore stored pre-rounded (cuDNN's TF32
 * kernels round on load -- same arithmetic, no per-layer shrink). */
#define ODWSCL_CONV_RELU  1
#define ODWSCL_CONV_ACCUM 2
#define ODWSCL_CONV_MASK  4
#define ODWSCL_CONV_ROUND 8
int odwscl_conv3x3_nhwc_tf32(const float* x, int B, int H, int W, int Cin, const float* w_krsc,
                             const float* bias, int Cout, int dilation, int flags, const float* mask_src,
                             float* y, odwscl_stream_t stream);
/* WGRAD + bias gradient of the same convolution: dw_krsc [Cout,3,3,Cin] = sum over pixels of
 * dz[b,h,w,co] * x[b, h+(r-1)d, w+(s-1)d, ci] (tcgen05, contraction over pixels with MN-major operands taken
 * straight from the NHWC tensors, split over pixel ranges and combined with vector atomics; dw is zeroed
 * inside); db [Cout] = sum over pixels of dz (may be NULL).  Cout % 32 == 0 (rows beyond Cout of the last
 * 128-row tile are never written), Cin % 32 == 0. */
int odwscl_conv3x3_wgrad_nhwc_tf32(const float* x, const float* dz, int B, int H, int W, int Cin, int Cout,
                                   int dilation, float* dw_krsc, float* db, odwscl_stream_t stream);
/* conv1_1 (Cin = 3, Cout = 64): fp32 FFMA, reads the NCHW image [B,3,H,W] and torch-layout weights
 * [64,3,3,3], writes NHWC [B,H,W,64].  `relu` is a flag mask: ODWSCL_CONV_RELU | ODWSCL_CONV_ROUND. */
int odwscl_conv3x3_c3_f32(const float* x_nchw, int B, int H, int W, const float* w_oihw, const float* bias,
                          int Cout, int relu, float* y_nhwc, odwscl_stream_t stream);
/* 2x2 / stride 2 max-pool, NHWC (C % 4 == 0), and its backward (gradient to the first maximum of each
 * window; relu_mask != 0 additionally zeroes it where the maximum is <= 0). */
int odwscl_maxpool2x2_nhwc_f32(const float* x, int B, int H, int W, int C, float* y, odwscl_stream_t stream);
int odwscl_maxpool2x2_nhwc_bwd_f32(const float* x, const float* gy, int B, int H, int W, int C, int relu_mask,
                                   float* gx, odwscl_stream_t stream);
/* x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi): operands of the 3-pass strict-fp32 mode.
 * lo may be NULL (plain rounding of weights / incoming gradients to TF32; hi may alias x). */
int odwscl_split_tf32(const float* x, long long n, float* hi, float* lo, odwscl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ODWSCL_H_ */
