/*
 * halo_b200.h -- C ABI of libhalo_sm100.so: the B200 (sm_100a) implementation of HALO's per-pixel
 * hyperbolic hot path (Poincare-ball classifier head + active-learning acquisition pass).
 *
 * Plain C types, raw DEVICE pointers and a CUDA stream handle only; no torch / C++ types cross this
 * boundary.  Every entry point enqueues on the caller's stream and returns without synchronising.
 * The library owns no tensor memory: outputs and scratch ("workspace") are caller-provided, sized by
 * the matching *_workspace_bytes() query.  Return value: 0 on success, negative halo_status otherwise;
 * halo_last_error() returns a thread-local human-readable message for the last failure.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * checkout paolomandica/HALO).  The reference has no FFI of its own (it is pure PyTorch); the
 * binding a maintainer would add is the ctypes stub shown in INTEGRATION.md.
 *
 * Layouts: all image-shaped tensors are contiguous NCHW / NHW, channel reduction on dim=1 exactly as
 * the reference (core/models/classifier.py:553, core/utils/hyperbolic.py:136).
 */
#ifndef HALO_B200_H
#define HALO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HALO_ABI_VERSION 2

typedef void* halo_stream_t; /* cudaStream_t / CUstream of the caller; NULL = legacy default stream */

typedef enum {
  HALO_OK = 0,
  HALO_ERR_BAD_ARG = -1,     /* NULL / non-positive dims / inconsistent options */
  HALO_ERR_UNSUPPORTED = -2, /* legal request outside the compiled envelope (e.g. num_classes > 32) */
  HALO_ERR_CUDA = -3,        /* a CUDA runtime call failed; message carries cudaGetErrorString */
  HALO_ERR_WORKSPACE = -4    /* workspace pointer NULL or too small */
} halo_status;

/* what `feat` holds */
typedef enum {
  HALO_FEAT_TANGENT_F32 = 0, /* raw decoder features u (fp32): expmap0+project is fused in  (hyperbolic.py:28-39) */
  HALO_FEAT_BALL_F32 = 1,    /* points already on the Poincare ball, fp32 */
  HALO_FEAT_BALL_F64 = 2     /* points already on the ball, fp64 (what the reference hands HyperMLR) */
} halo_feat_kind;
/* OR-ed into feat_kind: keep the contraction on the fp32 CUDA cores (no tcgen05 path); used by parity tests */
#define HALO_FEAT_FLAG_NO_TENSOR_CORE 0x100

/* per-pixel uncertainty written by the head / logits pass (floating_region.py:70-92,123-127) */
typedef enum {
  HALO_PIXUNC_ENTROPY = 0,   /* -sum p log(p+1e-6) / log(19)   (19 hard-coded, :74-76) */
  HALO_PIXUNC_ONE_MINUS_PGT = 1 /* 1 - p[gt], gt==255 -> argmax   ("oracle_acc", :77-83) */
} halo_pixunc_mode;

typedef enum {
  HALO_LABEL_ARGMAX = 0,     /* argmax_k p_k                    ("ripu", :166) */
  HALO_LABEL_GT_FILLED = 1   /* gt with 255 replaced by argmax  ("oracle_ripu", :172-173) */
} halo_label_mode;

typedef enum {
  HALO_NORM_RADIUS = 0,      /* Poincare distance to origin      (hyperbolic.py:74-83) */
  HALO_NORM_EUCLID = 1       /* ||x||_2 of the ball point        ("euc_norm", floating_region.py:195) */
} halo_norm_mode;

/* region uncertainty (floating_region.py:70-92, 158-163) */
typedef enum {
  HALO_UNC_BOXSUM = 0,       /* k x k zero-padded box SUM of the per-pixel map ("entropy", "oracle_acc") */
  HALO_UNC_PIXEL = 1,        /* the per-pixel map itself        ("pixel_entropy") */
  HALO_UNC_ZERO = 2          /* zeros                            ("none" and every unknown string) */
} halo_unc_mode;

/* region impurity (floating_region.py:165-202) */
typedef enum {
  HALO_PUR_NORM = 0,         /* impurity := the radius / norm plane, count := 1   ("radius", "euc_norm") */
  HALO_PUR_LABEL_HIST = 1,   /* k x k label-histogram entropy, count := window population ("ripu", "oracle_ripu") */
  HALO_PUR_RADIUS_BINS = 2,  /* same over K quantised radius bins   ("hyper", :94-110; window is pk x pk) */
  HALO_PUR_ZERO = 3          /* zeros, count := 1                    ("none") */
} halo_pur_mode;

int halo_abi_version(void);
const char* halo_last_error(void);
const char* halo_source_hash(void); /* "HALO_SRC_SHA256=<hex>": content hash of the sources the library was built from */

/* Which kernel variants the LAST halo_head_fwd / halo_head_bwd call on this thread launched (OR of the bits below).
 * Shapes outside the tensor-core envelope (see each entry point) run fp32 CUDA-core kernels that are several times
 * slower; the library also prints one line on stderr the first time that happens (HALO_QUIET=1 silences it). */
#define HALO_PATH_FWD_TC 0x01       /* head_fwd_tc_kernel: tcgen05 3xTF32, TMA */
#define HALO_PATH_FWD_CUDA_CORE 0x02
#define HALO_PATH_BWD_PIX_TC 0x04   /* head_bwd_tc_kernel */
#define HALO_PATH_BWD_PIX_CUDA_CORE 0x08
#define HALO_PATH_BWD_DW_TC 0x10    /* head_bwd_dw_tc_kernel */
#define HALO_PATH_BWD_DW_CUDA_CORE 0x20
#define HALO_PATH_BWD_STREAM_TC 0x40 /* head_bwd_stream_kernel: du and dW from ONE pass over the features */
#define HALO_PATH_BWD_RECOMPUTE 0x80 /* ... after recomputing the contractions (no saved planes were passed) */
int halo_last_path(void);

/* ---- Poincare-ball classifier head, forward -------------------------------------------------------
 * Replaces HyperMapper.expmap (core/utils/hyperbolic.py:28-39), HyperMLR.forward/_hyper_logits
 * (:120-188), HyperMapper.poincare_distance_origin (:74-83) and the softmax-entropy / argmax prologue of
 * FloatingRegionScore.forward (core/active/floating_region.py:151-166) in ONE pass over the features.
 *   feat   [N,C,H,W]  per `feat_kind`;  P, A [O,C] fp32 (HyperMLR.P_MLR / A_MLR);  c curvature > 0
 * Optional outputs (NULL = skip):
 *   logits [N,O,H,W] f32; radius [N,H,W] f32 (per norm_mode); pixunc [N,H,W] f32 (per pixunc_mode);
 *   label [N,H,W] u8 (per label_mode); stats [N,4] f32 = {min,max of the radius plane, unused, unused}
 *   saved [N, halo_head_saved_rows(C,O,H,W), H*W] f32 | NULL: the per-pixel contractions <u,-p_k>, <u,a_k/|a_k|> and |u|^2 a
 *   TRAINING forward keeps for halo_head_bwd (164 B per pixel at O = 19), so that the backward reads the features once.
 *   Only shapes with halo_head_saved_rows() > 0 can save (HALO_ERR_UNSUPPORTED otherwise: call again with saved = NULL).
 *   gt [N,H,W] u8 is read only by HALO_PIXUNC_ONE_MINUS_PGT / HALO_LABEL_GT_FILLED.
 * ws: halo_head_workspace_bytes(O, C) bytes of device scratch (packed class parameters). */
size_t halo_head_workspace_bytes(int O, int C);
int halo_head_saved_rows(int C, int O, int H, int W); /* 2*round_up(O,4)+1, or 0 when the shape cannot save */
int halo_head_fwd(const void* feat, int feat_kind, const float* P, const float* A, float c,
                  float* logits, float* radius, float* pixunc, uint8_t* label, float* stats, float* saved,
                  const uint8_t* gt, int pixunc_mode, int label_mode, int norm_mode,
                  int N, int C, int O, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream);

/* ---- head backward (core/train_learners.py:238,362,457,557: autograd through expmap + HyperMLR) ----
 *   feat [N,C,H,W] f32 raw features (HALO_FEAT_TANGENT_F32) ; dlogits [N,O,H,W] f32
 *   saved: the planes halo_head_fwd wrote for the SAME feat / P / A / c, or NULL (the contractions are then recomputed
 *          by one more pass over the features)
 *   dfeat [N,C,H,W] f32 ; dP, dA [O,C] f32 (overwritten, not accumulated)
 * Kernel selection (halo_last_path() reports it): C in {64,128,256}, O <= 24, H*W % 4 == 0 -> the streaming tensor-core
 * kernel (du and dW from ONE pass over the features); else C % 32 == 0, 64 <= C <= 256, O <= 24 -> the two-kernel
 * tensor-core path (pixel pass + weight gradient, the latter on tensor cores for C = 128 or 256); else fp32 CUDA-core
 * kernels.  Reductions have a fixed order: results are bitwise reproducible on a given device.  Environment knobs for A/B
 * tests: HALO_BWD_TWO_KERNEL=1, HALO_BWD_CUDA_CORE=1, HALO_BWD_DW_CUDA_CORE=1 pin the older paths. */
size_t halo_head_bwd_workspace_bytes(int N, int C, int O, int H, int W);
int halo_head_bwd(const float* feat, const float* P, const float* A, float c, const float* dlogits,
                  const float* saved, float* dfeat, float* dP, float* dA, int N, int C, int O, int H, int W,
                  void* ws, size_t ws_bytes, halo_stream_t stream);

/* ---- eager pieces of the head (callers that need the materialised tensors) --------------------------
 * halo_expmap0_project: HyperMapper.expmap (hyperbolic.py:28-39) on dim=1; out_f64 selects the output type. */
int halo_expmap0_project(const float* u, void* x_out, int out_f64, float c, int N, int C, int H, int W,
                         halo_stream_t stream);
/* halo_ball_norm: poincare_distance_origin (:74-83) or ||x|| of points on the ball; stats as above (or NULL) */
int halo_ball_norm(const void* x, int x_f64, float c, int norm_mode, float* out, float* stats,
                   int N, int C, int H, int W, halo_stream_t stream);
/* halo_radius_f64: the radius plane in the reference's own precision, for the "hyper" purity
 * (core/active/floating_region.py:94-110 quantises an fp64 poincare_distance_origin, hyperbolic.py:74-83, into K bins with
 * round-half-even; an fp32 radius lands ~1e-4 of the pixels in the neighbouring bin).  feat per halo_feat_kind: raw fp32
 * features (closed form of expmap0 + project + dist0 from an fp64 |u|^2) or ball points fp32/fp64.
 *   radius64 [N,H,W] f64;  stats64 [N,2] f64 = {min,max} of each image's plane. */
int halo_radius_f64(const void* feat, int feat_kind, float c, double* radius64, double* stats64,
                    int N, int C, int H, int W, halo_stream_t stream);
/* halo_logits_stats: softmax -> per-pixel uncertainty + label from explicit logits
 * (floating_region.py:151-152, 70-83, 123-127, 166, 172-173). */
int halo_logits_stats(const float* logits, const uint8_t* gt, int pixunc_mode, int label_mode,
                      float* pixunc, uint8_t* label, int N, int O, int H, int W, halo_stream_t stream);

/* ---- fused bilinear up-sampling in front of the score (core/active/build.py:122-135) ------------------
 * Replaces F.interpolate(logits, size, bilinear, align_corners=True) (:123-125) followed by softmax entropy / argmax
 * (floating_region.py:152,166) and F.interpolate(decoder_out, ...) (:132-135) followed by poincare_distance_origin /
 * norm (floating_region.py:188,195), per OUTPUT pixel, without materialising either up-sampled tensor.
 *   logits_lr [N,O,lh,lw] f32 | NULL;  emb_lr [N,C,eh,ew] per emb_kind (halo_feat_kind: raw features get
 *   expmap0+project at the low-resolution pixels first) | NULL -- the two may come at different resolutions (the
 *   DeepLab v3+ head up-samples only its logits, classifier.py:556-557);  outputs at [N,H,W]: pixunc f32, label u8,
 *   radius f32, stats [N,4]; radius64 [N,H,W] f64 + stats64 [N,2] f64 (both or neither; the fp64 radius of the interpolated
 *   embedding and its per-image extrema, as halo_radius_f64, for the "hyper" purity).  ws: halo_upsample_workspace_bytes(N,eh,ew), required whenever emb_lr is given (it holds the
 *   per-low-resolution-pixel Gram entries and exp-map factors the output pixels evaluate the norm from). */
size_t halo_upsample_workspace_bytes(int N, int h, int w);
int halo_upsample_score_inputs(const float* logits_lr, const void* emb_lr, int emb_kind, float c, const uint8_t* gt,
                               int pixunc_mode, int label_mode, int norm_mode, float* pixunc, uint8_t* label,
                               float* radius, float* stats, double* radius64, double* stats64,
                               int N, int O, int C, int lh, int lw, int eh, int ew,
                               int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream);

/* ---- floating-region score (core/active/floating_region.py:129-217 after the softmax) ---------------
 *   pixunc [N,H,W] f32, radius [N,H,W] f32 (+ its stats [N,4]), label [N,H,W] u8, active [N,H,W] u8|NULL
 *   k = uncertainty window (odd), pk = purity window (odd; 3 when the module was built for "hyper"),
 *   n_bins = class count for LABEL_HIST or K for RADIUS_BINS.
 *   radius64 [N,H,W] f64 + radius_stats64 [N,2] f64 (both or neither, from halo_radius_f64 / halo_upsample_score_inputs):
 *   when given, HALO_PUR_RADIUS_BINS quantises THEM, in fp64 like the reference (floating_region.py:94-110), and the fp32
 *   radius / radius_stats may be NULL in that mode.
 *   normalize: 0 = off; 1 = min-max normalise both maps (floating_region.py:206-208) and return them normalised;
 *              2 = same score, but the impurity/uncertainty planes are left un-normalised (scratch only).
 *   score [N,H,W] f32 (active!=0 -> -inf, build.py:146); impurity / uncertainty [N,H,W] f32 are
 *   REQUIRED scratch+outputs (they receive the maps the reference returns). */
size_t halo_score_workspace_bytes(int N);
int halo_score(const float* pixunc, const float* radius, const float* radius_stats,
               const double* radius64, const double* radius_stats64, const uint8_t* label, const uint8_t* active, int unc_mode, int pur_mode, int normalize, int k, int pk, int n_bins,
               float* score, float* impurity, float* uncertainty, int N, int H, int W,
               void* ws, size_t ws_bytes, halo_stream_t stream);

/* ---- budgeted greedy selection (core/active/build.py:27-64, select_pixels_to_label) -----------------
 * Bit-exact equivalent of the sequential loop: repeat n_regions times { arg-max of score with ties to the
 * smallest w then smallest h; stop at -inf; score/active window of radius mask_radius := -inf/1;
 * selected window of radius active_radius := 1; active_mask window := gt window }.  In place on all four.
 *   score [N,H,W] f32|f64; active, selected, active_mask, gt [N,H,W] u8
 *   n_picked [N] i32 out; picks [N,n_regions] i32 out|NULL (h*W+w in pick order, -1 padded)
 *   flags: HALO_SELECT_KEEP_SCORE leaves `score` untouched (the caller does not need the -inf windows written back;
 *          active / selected / active_mask are updated as usual) */
#define HALO_SELECT_KEEP_SCORE 0x1
size_t halo_select_workspace_bytes(int N, int H, int W, int n_regions);
int halo_select_f32(float* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask, const uint8_t* gt,
                    int n_regions, int active_radius, int mask_radius, int flags, int* n_picked, int* picks,
                    int N, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream);
int halo_select_f64(double* score, uint8_t* active, uint8_t* selected, uint8_t* active_mask, const uint8_t* gt,
                    int n_regions, int active_radius, int mask_radius, int flags, int* n_picked, int* picks,
                    int N, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream);

/* ---- round deltas: compact exchange of a round's mask updates between image shards ------------
 * A round labels the (2*active_radius+1)^2 windows around its picks (build.py:58-62).  pack gathers those labels:
 *   picks [N,cap] i32 (as written by halo_select_*), n_picked [N] i32, gt [N,H,W] u8 ->
 *   lab [N,cap,(2a+1)^2] u8 (gt of the window, row-major; 255 outside the image and for i >= n_picked).
 * apply replays gathered deltas onto replicated masks: row j of the gathered buffers belongs to pool image
 * row_image[j] (< 0: padding row, skipped);  masks [n_images,H,W] u8 receives masks[image][window] = lab. */
int halo_round_delta_pack(const int* picks, const int* n_picked, const uint8_t* gt, uint8_t* lab, int N, int cap,
                          int H, int W, int active_radius, halo_stream_t stream);
int halo_round_delta_apply(uint8_t* masks, const int* row_image, const int* picks, const int* n_picked,
                           const uint8_t* lab, int rows, int cap, int H, int W, int active_radius, halo_stream_t stream);


/* ---- channel reduction + hyperbolic feature re-weighting upstream of the head, evaluation mode --------
 * Replaces core/models/classifier.py:526-550 (and the same block of the v2 head, :187-214) as the acquisition round runs
 * them (classifier.eval(), core/active/build.py:72-73):
 *     y = conv_reduce(f)                                   1x1 conv Cin -> C with bias
 *     wt = clamp(mean_pixels(wn_mlp(y)), 1e-5)             Linear, BatchNorm1d (running statistics), ReLU, Linear
 *     z = F.normalize(y over the pixels of each channel) * wt
 *   feat [N,Cin,H,W] f32; Wr [C,Cin], br [C]|NULL (conv_reduce.weight / .bias);
 *   W1 [C,C], b1 [C], bn_gamma/beta/mean/var [C], bn_eps, W2 [C,C], b2 [C] (wn_mlp[0], [1], [3]); W1 = NULL: no HFR (z = y)
 *   out [N,C,H,W] f32 = z, the features HyperMapper.expmap / halo_head_fwd read next; scale_out [N,C] f32 | NULL = z / y
 * HFR needs C <= 128.  Evaluation-mode forward; the training mode is halo_reduce_hfr_train_fwd / _bwd below. */
size_t halo_reduce_hfr_workspace_bytes(int N, int C, int H, int W);
int halo_reduce_hfr_fwd(const float* feat, const float* Wr, const float* br, const float* W1, const float* b1,
                        const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var,
                        float bn_eps, const float* W2, const float* b2, float* out, float* scale_out,
                        int N, int Cin, int C, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream);

/* Training mode of the same block (the training step's side of classifier.py:526-550): BatchNorm1d normalises with the
 * statistics of the batch (all N*H*W rows), and halo_reduce_hfr_train_bwd is the autograd of the whole block.
 *   fwd: y_out [N,C,H,W] = conv_reduce(f) and a_out [N,C,H,W] = the hidden pre-activations W1 y + b1 (both kept for the
 *        backward); z_out [N,C,H,W] = the re-weighted features (W1 = NULL:
 *        no HFR, z = y, z_out unused); batch_stats [2][C] = batch mean and biased variance of the hidden pre-activations
 *        (the caller updates running_mean / running_var from them, the variance times M/(M-1), M = N*H*W);
 *        small [N][3][C] = per-image {mean hidden activation, re-weighting before the clamp, |y_c|} for the backward.
 *   bwd: dz [N,C,H,W] -> dfeat [N,Cin,H,W] | NULL, dWr [C,Cin], dbr [C] | NULL, and with HFR dW1 [C,C], db1, dgamma, dbeta,
 *        dW2 [C,C], db2 (all overwritten).  Fixed-order reductions: bitwise reproducible.
 *   fixed_stats [2][C] | NULL: a BatchNorm1d left in evaluation mode inside a differentiated step normalises with its
 *        running mean / variance; pass them here (they are copied to batch_stats) and stats_are_batch = 0 to the backward.
 * One workspace size serves both calls. */
size_t halo_reduce_hfr_train_workspace_bytes(int N, int Cin, int C, int H, int W);
int halo_reduce_hfr_train_fwd(const float* feat, const float* Wr, const float* br, const float* W1, const float* b1,
                              const float* bn_gamma, const float* bn_beta, float bn_eps, const float* W2, const float* b2,
                              const float* fixed_stats, float* y_out, float* a_out, float* z_out, float* batch_stats,
                              float* small, int N, int Cin, int C, int H, int W, void* ws, size_t ws_bytes,
                              halo_stream_t stream);
int halo_reduce_hfr_train_bwd(const float* feat, const float* Wr, const float* W1, const float* bn_gamma,
                              const float* bn_beta, float bn_eps, const float* W2, const float* y, const float* a,
                              const float* batch_stats, const float* small, const float* dz, float* dfeat, float* dWr, float* dbr, float* dW1,
                              float* db1, float* dgamma, float* dbeta, float* dW2, float* db2, int stats_are_batch, int N,
                              int Cin, int C, int H, int W, void* ws, size_t ws_bytes, halo_stream_t stream);

/* ---- losses on the head's logits, fused with the up-sampling in front of them and its adjoint -------
 * Replaces, in the training step (core/train_learners.py:343-356 target branch, :232-236 source branch):
 *     out = F.interpolate(logits_lr, (H,W), bilinear, align_corners=True)        core/models/classifier.py:556-557
 *     loss_sup = CrossEntropyLoss(ignore_index=255)(out, labels)                  skipped when no pixel is labelled
 *     negative = NegativeLearningLoss(threshold)(softmax(out)) * neg_weight       core/loss/negative_learning_loss.py:6-16
 *     (loss_sup + negative).backward()  down to logits_lr
 * without materialising `out`, its softmax or their gradients.
 *   logits_lr [N,O,h,w] f32;  labels [N,H,W] u8 (255 = ignore) | NULL (no supervised term);  neg_weight = SOLVER.NEGATIVE_LOSS
 *   (0 disables the negative term);  threshold = 0.05 in the reference.
 *   losses [4] f32 out = {loss_sup + negative, loss_sup, negative (weighted), number of labelled pixels}
 *   dlogits_lr [N,O,h,w] f32 out | NULL = d(loss_sup + negative)/d logits_lr.  Deterministic (gather, fixed-order sums).
 *   ws: halo_seg_loss_workspace_bytes() bytes. */
size_t halo_seg_loss_workspace_bytes(void);
int halo_seg_loss(const float* logits_lr, const uint8_t* labels, float neg_weight, float threshold, float* losses,
                  float* dlogits_lr, int N, int O, int h, int w, int H, int W, void* ws, size_t ws_bytes,
                  halo_stream_t stream);

/* ---- packed round rows: the ONE exchange at the end of a sharded round (the reference runs the round on
 * rank 0 alone, core/train_learners.py:307-326, so this has no reference counterpart beyond build.py:58-62) ------------
 * One row per pool image, row_bytes = halo_round_row_bytes(cap, a) (a multiple of 16):
 *     [ int32 count ][ int32 picks[cap] ][ uint8 lab[cap][(2a+1)^2] ][ pad ]
 * A shard packs its images' rows into one buffer (pack may be called once per batch with `rows` pointing at the batch's
 * first row), all-gathers that buffer ONCE, and every rank replays all rows onto its replica of the pool's masks.
 *   pack : picks [N,pick_stride] i32, n_picked [N] i32 (as written by halo_select_*), gt [N,H,W] u8 -> rows [N,row_bytes]
 *   apply: rows [n_rows,row_bytes]; row j belongs to pool image row_image[j] (< 0: padding row, skipped);
 *          masks [n_images,H,W] u8 receives the labels; n_picked_out [n_images] i32 | NULL receives the counts. */
size_t halo_round_row_bytes(int cap, int active_radius);
int halo_round_rows_pack(const int* picks, const int* n_picked, const uint8_t* gt, uint8_t* rows, int N, int cap,
                         int pick_stride, int H, int W, int active_radius, halo_stream_t stream);
int halo_round_rows_apply(uint8_t* masks, const int* row_image, const uint8_t* rows, int* n_picked_out, int n_rows,
                          int cap, int H, int W, int active_radius, halo_stream_t stream);
/* Position-sensitive 64-bit checksum of a device buffer (sum of 8-byte words times odd position weights, mod 2^64;
 * order-independent, so bitwise reproducible): ranks compare it after the exchange to prove their replicas agree.
 * word_offset chains several buffers as if concatenated; accumulate != 0 adds into *out instead of overwriting it. */
int halo_checksum64(const void* data, size_t nbytes, unsigned long long word_offset, unsigned long long* out,
                    int accumulate, halo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HALO_B200_H */
