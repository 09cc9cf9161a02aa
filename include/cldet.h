/*
 * cldet.h -- C ABI of the B200-native detection-head path (libcldet.so).
 *
 * Every entry point takes plain pointers and sizes; there are no torch types here.  Device
 * pointers are prefixed d_, host pointers h_.  `stream` is a cudaStream_t passed as void*.
 * All functions return a cldet_status (0 = ok); none of them calls exit/abort, none
 * allocates device memory behind the caller's back (workspaces are passed in), none keeps
 * mutable global state, so calls are re-entrant and safe from several host threads
 * (reference: evaluator.py:400-422 runs predict from up to 10 threads).
 *
 * Each function names the reference interface it replaces (paths relative to the
 * reference checkout, EonianCoda/CL_object_detection).
 */
#ifndef CLDET_H_
#define CLDET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cldet_status {
    CLDET_OK = 0,
    CLDET_ERR_INVALID_ARGUMENT = 1,   /* null pointer, negative extent, unsupported size */
    CLDET_ERR_WORKSPACE_TOO_SMALL = 2,
    CLDET_ERR_CUDA = 3,               /* a CUDA runtime call failed; see cldet_last_cuda_error() */
    CLDET_ERR_UNSUPPORTED = 4
} cldet_status;

#define CLDET_ABI_VERSION 1

int cldet_abi_version(void);
const char* cldet_status_string(int status);
/* cudaGetErrorString of the last failing CUDA call made by THIS thread inside libcldet (thread-local). */
const char* cldet_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------
 * Per-anchor assignment word written by cldet_iou_assign and read by the loss kernel.
 *   bits  0..1  state: 0 background (IoU_max < 0.4), 1 positive (IoU_max >= 0.5), 2 ignore,
 *               3 image has no valid GT row (reference branch losses.py:292-307)
 *   bit   2     set by the new_ignore_past_class pre-pass: background anchor whose old-class columns count as
 *               negatives (sum of old-class probabilities < 0.5, losses.py:326-327)
 *   bits  3..15 class label of the assigned GT row (valid when state == 1; 0x1fff = label out of range)
 *   bits 16..31 RAW row index (into annotations[j]) of the assigned GT (first maximal IoU)
 * ------------------------------------------------------------------------------------- */
#define CLDET_STATE_BG 0u
#define CLDET_STATE_POS 1u
#define CLDET_STATE_IGNORE 2u
#define CLDET_STATE_EMPTY 3u
#define CLDET_MAX_GT_ROWS 65536
#define CLDET_MAX_CLASSES 8191
#define CLDET_BAD_LABEL 0x1fffu
#define CLDET_META_OLD_ACTIVE 4u

/* ---- a1: Anchors.forward (retinanet/anchors.py:21-40; generate_anchors :42-73; shift :109-129) ----
 * Levels 3..7, strides 8..128, sizes 32..512, ratios {0.5,1,2}, scales {1,2^(1/3),2^(2/3)}.
 * fp64 arithmetic on the device, ONE final rounding to fp32 (bit-exact with the numpy reference). */
int cldet_num_anchors(int height, int width, int64_t* out_num_anchors);
int cldet_anchors(int height, int width, float* d_anchors /* [A,4] */, void* stream);

/* ---- a2-a4: calc_iou + max/argmax + 0.4/0.5 thresholds (retinanet/losses.py:4-21, 287-288, 309-341) ----
 * annotations: [N,G,5] fp32 rows (x1,y1,x2,y2,label); rows with label == -1 are padding and are
 * skipped in order (the reference compacts them away, losses.py:287-288).  Pseudo-label rows merged by
 * the dataset (dataloader.py:129-136) are ordinary rows here.
 * Outputs (all device):
 *   d_meta    [N,A] uint32   assignment word, see above
 *   d_argmax  [N,A] int32    index into the COMPACTED GT list = reference IoU_argmax (may be NULL); -1 for empty images
 *   d_iou_max [N,A] float    reference IoU_max (may be NULL unless the decrease_positive_by_IOU variant is used)
 *   d_npos    [N]   int32    number of positive anchors per image; MUST be zero on entry (cldet_focal_loss does this)
 *   d_nvalid  [N]   int32    number of valid GT rows per image
 * num_classes is used only to flag out-of-range labels (CLDET_BAD_LABEL). */
int cldet_iou_assign(const float* d_anchors, int64_t num_anchors, const float* d_annotations, int num_images,
                     int gt_rows, int num_classes, uint32_t* d_meta, int32_t* d_argmax, float* d_iou_max,
                     int32_t* d_npos, int32_t* d_nvalid, void* stream);

/* Standalone pairwise IoU matrix [A,G] (calc_iou, losses.py:4-21; also its copies in IL_method/mas.py:15-32,
 * weight_init.py:7-24).  For tests and the "next" callers; the loss path never materialises this matrix. */
int cldet_calc_iou(const float* d_a, int64_t num_a, const float* d_b, int num_b, float* d_iou, void* stream);

/* Row-wise max of the fp64 IoU matrix (calc_iou on float64 inputs followed by .max(dim=1)): the pseudo-label generator's
 * overlap test against the real GT (IL_method/persuado_label.py:59-75).  d_a [num_a,4], d_b [num_b,4] doubles;
 * d_max [num_a]; d_argmax [num_a] int32 (first maximal index) may be NULL. */
int cldet_iou_max_f64(const double* d_a, int64_t num_a, const double* d_b, int num_b, double* d_max, int32_t* d_argmax,
                      void* stream);

/* ---- a5-a8: FocalLoss.forward + its autograd backward (retinanet/losses.py:252-452) ---- */
typedef struct cldet_loss_params {
    float alpha;                       /* params['alpha'], default 0.25 */
    float gamma;                       /* params['gamma'], default 2 (computed as x*x like ATen) */
    int32_t incremental;               /* cur_state > 0 */
    int32_t past_class_num;            /* params.states[cur_state]['num_past_class'] */
    int32_t ignore_past_class;         /* losses.py:319-322 */
    int32_t new_ignore_past_class;     /* :323-328 */
    int32_t decrease_positive_by_iou;  /* :353-362 */
    int32_t enhance_on_new;            /* :380-384 */
    float decrease_positive;           /* :364-366, default 1.0 */
    int32_t image_height, image_width; /* > 0: d_anchors is the standard anchor grid of an image of this size (what
                                          cldet_anchors(height, width) produces): cldet_focal_loss then uses the GT-centric
                                          assignment (a few hundred anchors per GT box instead of every anchor x GT pair).
                                          0: arbitrary anchors, anchor-centric kernel.  Results are identical. */
    int32_t cls_is_logits;             /* 0: d_cls holds probabilities (the reference's FocalLoss input).
                                          1: d_cls holds LOGITS -- the kernel applies ATen's sigmoid 1/(1+exp(-x)) itself and
                                          writes dL/dlogits = dL/dp * (1-p) * p (replaces the separate
                                          Sigmoid at losses.py:566/633 and its backward; d_grad_cls may alias d_cls) */
} cldet_loss_params;

/* Bytes of scratch cldet_focal_loss needs for (N, A).  A workspace must be ZERO before its first use; every call leaves it
 * zero-clean again (counters, npos accumulator, IoU_max scratch), so one memset at allocation time is enough and the steady
 * state has no memset nodes. */
size_t cldet_focal_loss_workspace_bytes(int num_images, int64_t num_anchors);

/* One call = IoU/assign kernel + fused loss/gradient kernel (two launches).
 *   d_cls  [N,A,C] probabilities (post-sigmoid, as the reference passes them)   d_reg [N,A,4]
 *   d_anchors [A,4]      d_annotations [N,G,5]
 *   d_weights [4,N] (row-major, one row per term) upstream gradients per image: row 0 dL/d(bg_j), row 1 dL/d(fg_j),
 *            row 2 dL/d(reg_j), row 3 dL/d(enhance term); reg_j is the per-image regression term, so a caller holding
 *            dL/d(reg_loss[0]) passes that / N.
 *            NULL = no gradients wanted (d_grad_* must then be NULL too).
 *   d_baked_weights [4,N] out, may be NULL: receives a copy of the weights baked into the gradients (the record that
 *            cldet_focal_loss_reweight compares against), so d_weights itself can be a shared read-only tensor.
 *   d_grad_cls [N,A,C], d_grad_reg [N,A,4]: written completely (zeros where the reference's gradient is zero).
 *            d_grad_cls may alias d_cls (in-place variant for a caller that no longer needs the probabilities).
 *   d_losses [4,N]: rows bg_j, fg_j (each already divided by max(npos_j,1)), reg_j, enhance_on_new partial of image j.
 *   d_meta [N,A] uint32 out (kept by the caller for the re-weighting pass); d_iou_max [N,A] out, may be NULL unless
 *            params->decrease_positive_by_iou (on the GT-centric path it is exact wherever IoU_max >= 0.39 -- all the loss
 *            reads -- and 0 elsewhere; background anchors' words carry label/row 0); d_npos / d_nvalid [N] out.
 *   d_bg_mask [N,A] uint8 out, may be NULL: 1 where the anchor is NOT positive (reference 'bg_masks', losses.py:334-335).
 *   d_status: int32[1] out, may be NULL; set non-zero when a positive anchor's label is outside [0,C) (reference raises, Q8). */
int cldet_focal_loss(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                     int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                     const cldet_loss_params* params, const float* d_weights, float* d_baked_weights,
                     float* d_grad_cls, float* d_grad_reg, float* d_losses,
                     uint32_t* d_meta, float* d_iou_max, int32_t* d_npos, int32_t* d_nvalid,
                     uint8_t* d_bg_mask, int32_t* d_status,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- image-sharded runs: fused loss + all-gather over NVLink peer memory ----
 * Every rank owns a gather buffer float[2][world][4][N] and arrival counters uint32[2][world] (both zero-initialised) that
 * all ranks have mapped (CUDA IPC / symmetric memory).  With a peer exchange, the loss kernel's last block per image stores
 * that image's four terms into slot [parity][rank] of EVERY rank's buffer (one lane per destination: peer stores followed by
 * a release-ordered system-scope arrival) -- there is no separate collective launch.  The arrival counters ONLY GROW: the
 * u-th use of a parity (u = 1, 2, ...) is complete when every source rank's counter has reached u * N.  cldet_peer_wait
 * blocks the stream until then and copies the gathered terms out; alternate parity 0/1 between consecutive steps.
 * All ranks must use the same N. */
typedef struct cldet_peer_exchange {
    void* d_peer_terms;   /* device array of `world` float*  : rank p's gather buffer, as mapped in THIS process */
    void* d_peer_flags;   /* device array of `world` uint32* : rank p's arrival counters */
    int32_t rank, world, parity;
    /* Optional FUSED WAIT (d_wait_out != NULL): the loss kernel's block that finishes this rank's last image then does what
     * cldet_peer_wait does -- waits (bounded by timeout_ms) for target_arrivals on this rank's counters, copies this parity's
     * terms into d_wait_out [4][world*N] and writes the global regression mean to the call's d_reg_mean -- so there is no wait
     * launch behind the loss kernel at all.  Arguments as for cldet_peer_wait. */
    int32_t timeout_ms;
    uint32_t target_arrivals;
    int32_t reserved;
    const void* d_flags_local;
    const void* d_terms_local;
    float* d_wait_out;
    int32_t* d_wait_status;
} cldet_peer_exchange;
int cldet_focal_loss_sharded(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                             int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                             const cldet_loss_params* params, const float* d_weights, float* d_baked_weights,
                             float* d_grad_cls, float* d_grad_reg, float* d_losses,
                             uint32_t* d_meta, float* d_iou_max, int32_t* d_npos, int32_t* d_nvalid,
                             uint8_t* d_bg_mask, int32_t* d_status,
                             void* d_workspace, size_t workspace_bytes, const cldet_peer_exchange* peer, float* d_reg_mean,
                             void* stream);
/* d_reg_mean (device float[1], may be NULL; used when peer == NULL or peer->world <= 1): receives reg_loss = mean_j reg_j
 * (losses.py:445), added in image order by the block that finishes the last image, so the caller needs no reduction launch.
 * With a peer exchange the mean over the GLOBAL batch comes from cldet_peer_wait. */
/* Node-shared buffers for the peer exchange.  cldet_peer_alloc: cudaMalloc + zero + export (64-byte cudaIpcMemHandle_t copied
 * to h_handle64); cldet_peer_open: map a peer rank's buffer into this process WITH THE CALLER'S DEVICE CURRENT (lazy peer
 * access), returning a pointer kernels of that device can dereference; cldet_peer_close / cldet_peer_free undo them. */
int cldet_peer_alloc(size_t bytes, void** d_ptr, unsigned char* h_handle64);
int cldet_peer_open(const unsigned char* h_handle64, void** d_ptr);
int cldet_peer_close(void* d_ptr);
int cldet_peer_free(void* d_ptr);
/* Let kernels of the CURRENT device dereference memory of `peer_device` (cudaDeviceEnablePeerAccess; already-enabled is ok). */
int cldet_enable_peer_access(int peer_device);
/* Wait (bounded by timeout_ms) until all `world` source ranks have signalled `target_arrivals` (wrap-safe compare) on
 * this parity's counters of THIS rank (d_flags_local), then copy this parity's terms d_terms_local[parity][world][4][N] into
 * d_out [4][world*N] (row k, image r*N + j: the caller's private copy in global image order; d_out may be NULL = wait only).
 * A source rank that does not arrive in time gets its rows of d_out filled with NaN and *d_status (may be NULL; may point to
 * mapped pinned host memory) is set to 2; the counters are never reset, so later steps stay in sequence regardless. */
int cldet_peer_wait(const void* d_flags_local, const float* d_terms_local, int world, int num_images, int parity,
                    uint32_t target_arrivals, int timeout_ms, float* d_out, float* d_reg_mean, int32_t* d_status, void* stream);
/* d_reg_mean (may be NULL): receives the mean of row 2 (the per-image regression terms) over all world*N images. */

/* Profiling hook (per host thread, one-shot): the next cldet_focal_loss call of THIS thread records the given
 * cudaEvent_t handles before the assign kernel, between the two kernels and after the loss kernel, on its stream.
 * NULL = do not record.  Lets a benchmark time each kernel inside the real fused call. */
int cldet_focal_loss_profile_events(void* ev_begin, void* ev_between, void* ev_end);

/* Loss/gradient stage alone, on an assignment produced earlier by cldet_iou_assign (same argument meaning).
 * d_meta is read; with params->new_ignore_past_class its bit 2 is (re)written by a pre-pass over the old-class columns.
 * Workspace contract as cldet_focal_loss. */
int cldet_focal_loss_from_assignment(const float* d_cls, const float* d_reg, const float* d_anchors,
                                     const float* d_annotations, int num_images, int64_t num_anchors, int num_classes,
                                     int gt_rows, const cldet_loss_params* params, const float* d_weights,
                                     float* d_baked_weights, float* d_grad_cls, float* d_grad_reg, float* d_losses,
                                     uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos, uint8_t* d_bg_mask,
                                     int32_t* d_status, void* d_workspace, size_t workspace_bytes, void* stream);

/* Backward with upstream weights that differ from the ones baked into d_grad_* by the forward call
 * (IL_Loss's clip_loss masking, losses.py:575-581, is only known after the forward).  Per image and per term the
 * kernel compares d_new_weights with d_baked_weights ON THE DEVICE (no host sync): images whose weights are unchanged
 * are skipped; a changed fg/reg weight touches only that image's positive anchors; a changed bg weight recomputes the
 * image.  d_baked_weights is updated to d_new_weights for the images that were patched.  d_workspace: the same (zeroed)
 * workspace contract as cldet_focal_loss. */
int cldet_focal_loss_reweight(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                              int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                              const cldet_loss_params* params, const float* d_new_weights, float* d_baked_weights,
                              float* d_grad_cls, float* d_grad_reg, const uint32_t* d_meta, const float* d_iou_max,
                              const int32_t* d_npos, void* d_workspace, size_t workspace_bytes, void* stream);

/* Same, with the four upstream-weight rows given separately as (pointer, element stride): exactly what autograd hands to
 * backward (dL/dbg[N], dL/dfg[N], dL/dreg_j[N], dL/d enhance) without packing them first.  stride 0 broadcasts one value,
 * a NULL row means zeros. */
int cldet_focal_loss_reweight_rows(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                                   int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                                   const cldet_loss_params* params, const float* d_w_bg, int64_t stride_bg,
                                   const float* d_w_fg, int64_t stride_fg, const float* d_w_reg, int64_t stride_reg,
                                   const float* d_w_enh, int64_t stride_enh, const float* d_w_reg_mean, float reg_mean_scale,
                                   float* d_baked_weights, float* d_grad_cls,
                                   float* d_grad_reg, const uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos,
                                   void* d_workspace, size_t workspace_bytes, void* stream);
/* d_w_reg_mean (device float[1], may be NULL) and reg_mean_scale: a second source of the regression weight for callers that
 * hold dL/d(reg_loss) of reg_loss = mean_j reg_j (the tensor the reference returns): dL/dreg_j = d_w_reg[j*stride] +
 * d_w_reg_mean[0] * reg_mean_scale with reg_mean_scale = 1/N (over the global batch on image-sharded runs). */

/* ---- SURVEY 8(f) row f1, second half: the head's RAW conv outputs in, gradients of the same layout out ----
 * Replaces, for training, the layout work between the output convolutions and FocalLoss: ClassificationModel.forward /
 * RegressionModel.forward (retinanet/model.py:125-130, 170-184: permute(0,2,3,1) + contiguous() + view per level),
 * torch.cat over the five levels in ResNet.forward (model.py:472-474), optionally classifier_act = Sigmoid
 * (params->cls_is_logits), and the mirror image of all of it in autograd's backward.
 *   h_cls_levels / h_reg_levels: HOST arrays of num_levels (= 5, pyramid levels 3..7) DEVICE pointers; level l holds the
 *            output convolution's result as is: classification [N, 9*C, H_l, W_l] (channel = k*C + c for anchor type k and
 *            class c), regression [N, 36, H_l, W_l] (channel = k*4 + i), contiguous NCHW, fp32, with
 *            H_l = ceil(image_height / 2^l), W_l = ceil(image_width / 2^l) (retinanet/anchors.py:25).
 *   h_grad_cls_levels / h_grad_reg_levels: host arrays of device pointers to gradient tensors of the same shapes, written
 *            completely; both NULL (together with d_weights) when no gradients are wanted.
 *   d_anchors [A,4]: the standard grid of (image_height, image_width) (cldet_anchors), A = 9 * sum_l H_l*W_l.
 *   Everything else (annotations, params, weights/baked weights, d_losses [4,N], d_meta / d_iou_max / d_bg_mask [N,A] in the
 *   reference's concatenated anchor order, d_npos, d_nvalid, d_status, workspace of cldet_focal_loss_workspace_bytes(N, A))
 *   means what it means for cldet_focal_loss.  Results: per-image terms equal to cldet_focal_loss on the permuted +
 *   concatenated tensors up to summation order; gradients equal element for element. */
int cldet_focal_loss_head(const float* const* h_cls_levels, const float* const* h_reg_levels, int num_levels, int image_height,
                          int image_width, const float* d_anchors, const float* d_annotations, int num_images, int num_classes,
                          int gt_rows, const cldet_loss_params* params, const float* d_weights, float* d_baked_weights,
                          float* const* h_grad_cls_levels, float* const* h_grad_reg_levels, float* d_losses, uint32_t* d_meta,
                          float* d_iou_max, int32_t* d_npos, int32_t* d_nvalid, uint8_t* d_bg_mask, int32_t* d_status,
                          void* d_workspace, size_t workspace_bytes, void* stream);
/* Backward of cldet_focal_loss_head with upstream weights that differ from the baked ones (rows as in
 * cldet_focal_loss_reweight_rows): images whose weights changed are recomputed, the others are skipped. */
int cldet_focal_loss_head_reweight(const float* const* h_cls_levels, const float* const* h_reg_levels, int num_levels,
                                   int image_height, int image_width, const float* d_anchors, const float* d_annotations,
                                   int num_images, int num_classes, int gt_rows, const cldet_loss_params* params,
                                   const float* d_w_bg, int64_t stride_bg, const float* d_w_fg, int64_t stride_fg,
                                   const float* d_w_reg, int64_t stride_reg, const float* d_w_enh, int64_t stride_enh,
                                   float* d_baked_weights, float* const* h_grad_cls_levels, float* const* h_grad_reg_levels,
                                   const uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos, void* d_workspace,
                                   size_t workspace_bytes, void* stream);

/* ---- head-distillation terms of IL_Loss (retinanet/losses.py:705-737) ----
 * d_cls [N,A,C] current logits, d_prev_cls [N,A,P] previous-model logits (P = past_class_num), d_reg / d_prev_reg [N,A,4],
 * d_bg_mask [N,A] uint8 (FocalLoss 'bg_masks').  prev_fg_mask = sigmoid(prev_cls) > 0.05; reg_mask = bg_mask & any(prev_fg_mask);
 * d_losses[0] = MSE over prev_fg_mask elements (ignore_gd: over reg_mask rows) of logits (distill_logits) or probabilities,
 * d_losses[1] = SmoothL1 (beta 1) over reg_mask rows; d_counts[0..1] = the two element counts (float) for the backward.
 * An empty selection yields NaN like the reference's mean over an empty tensor. */
size_t cldet_distill_workspace_bytes(int num_images, int64_t num_anchors);
int cldet_distill_forward(const float* d_cls, const float* d_prev_cls, const float* d_reg, const float* d_prev_reg,
                          const uint8_t* d_bg_mask, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                          int distill_logits, int ignore_gd, float* d_losses, float* d_counts, void* d_workspace,
                          size_t workspace_bytes, void* stream);
/* d_grad_cls_loss / d_grad_reg_loss: device scalars dL/d(dist_cls_loss), dL/d(dist_reg_loss) (NULL = 0).
 * Writes d_grad_cls [N,A,C] completely (zeros for columns >= P) and d_grad_reg [N,A,4]. */
int cldet_distill_backward(const float* d_cls, const float* d_prev_cls, const float* d_reg, const float* d_prev_reg,
                           const uint8_t* d_bg_mask, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                           int distill_logits, int ignore_gd, const float* d_counts, const float* d_grad_cls_loss,
                           const float* d_grad_reg_loss, float* d_grad_cls, float* d_grad_reg, void* stream);

/* ---- SURVEY 8(f) row f2, second half: enhance_error on replay batches (retinanet/losses.py:590-603) ----
 * d_cls [N,A,C] class PROBABILITIES (the replay branch runs the model with enable_act=True).  Over the new-class columns
 * c >= past_class_num, the elements > 0.05 contribute |p| (method 1 = "L1"), p^2 (2 = "L2") or p^3 (3 = "L3");
 * *d_loss = sum / max(count, 1), *d_count = max(count, 1) (float, kept for the backward pass).
 * Backward writes d_grad_cls [N,A,C] completely (zeros for unselected elements); d_grad_loss: device scalar dL/d(loss), NULL = 0. */
size_t cldet_enhance_error_workspace_bytes(int64_t num_elements);
int cldet_enhance_error_forward(const float* d_cls, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                                int method, float* d_loss, float* d_count, void* d_workspace, size_t workspace_bytes, void* stream);
int cldet_enhance_error_backward(const float* d_cls, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                                 int method, const float* d_count, const float* d_grad_loss, float* d_grad_cls, void* stream);

/* ---- the regression term of MAS Output_norm (IL_method/mas.py:52-55) without per-image gathers ----
 * d_terms[j] = mean(|d_reg[j][positive rows]|) over rows x 4 (0 when image j has no positive anchor), d_counts[j] = number
 * of positive rows (float).  d_positive [N,A] uint8.  Backward: d_grad_reg [N,A,4] written completely, from
 * d_grad_terms[j * grad_stride] (stride 0 broadcasts one value). */
int cldet_masked_abs_mean_forward(const float* d_reg, const uint8_t* d_positive, int num_images, int64_t num_anchors,
                                  float* d_terms, float* d_counts, void* stream);
int cldet_masked_abs_mean_backward(const float* d_reg, const uint8_t* d_positive, int num_images, int64_t num_anchors,
                                   const float* d_counts, const float* d_grad_terms, int64_t grad_stride, float* d_grad_reg,
                                   void* stream);

/* ---- a10-a13: eval-mode detection output (retinanet/utils.py:102-144 BBoxTransform/ClipBoxes;
 *      retinanet/model.py:507-550 ResNet.predict; IL_method/persuado_label.py:99-127 Labeler.predict;
 *      torchvision.ops.batched_nms at model.py:540) ---- */

/* BBoxTransform.forward (utils.py:102-126) for every anchor of every image: [N,A,4]; clip != 0 also applies
 * ClipBoxes.forward (utils.py:134-144: x1,y1 >= 0, x2 <= width, y2 <= height).  For callers that want the dense box
 * tensor; the detection pipeline below decodes only the anchors that pass the score threshold. */
int cldet_decode_boxes(const float* d_anchors, const float* d_reg, int num_images, int64_t num_anchors, int clip, int height,
                       int width, float* d_boxes, void* stream);
/* ClipBoxes.forward alone, in place on [num_boxes,4]. */
int cldet_clip_boxes(float* d_boxes, int64_t num_boxes, int height, int width, void* stream);

typedef struct cldet_candidate {   /* 32 bytes, one per anchor that passes the score threshold */
    float x1, y1, x2, y2;          /* decoded, clipped box */
    float score;                   /* max class probability */
    int32_t label;                 /* first class index attaining it */
    int32_t anchor;                /* anchor index inside the image */
    int32_t pad;
} cldet_candidate;

/* Fused: per-anchor max/argmax over C classes (+sigmoid when is_logits), score > thresh filter, box decode + clip of
 * the survivors only, append to a per-image candidate list.  d_candidates is [N, capacity]; d_keys [N, capacity] uint64
 * receives each candidate's ordering key ((ordered score bits << 32) | ~anchor: larger key = earlier in the reference's
 * stable score-descending order); d_counts [N] int32 must be zero on entry and receives the number of survivors (it can
 * exceed capacity: the excess is dropped and the caller must treat that as an error).  Append order is arbitrary;
 * cldet_sort_candidates orders it. */
int cldet_decode_filter(const float* d_cls, int is_logits, const float* d_reg, const float* d_anchors, int num_images,
                        int64_t num_anchors, int num_classes, int height, int width, float score_thresh,
                        cldet_candidate* d_candidates, uint64_t* d_keys, int64_t capacity, int32_t* d_counts, void* stream);

/* The same filter on the head's RAW conv outputs (eval side): h_cls_levels / h_reg_levels are HOST arrays of
 * num_levels (= 5) device pointers to classification [N, 9*C, H_l, W_l] and regression [N, 36, H_l, W_l] tensors exactly as
 * the output convolutions produce them (layout as in cldet_focal_loss_head), replacing the per-level permute + contiguous +
 * view (retinanet/model.py:125-130, 170-184) and the torch.cat over levels (model.py:472-474) in front of ResNet.predict
 * (model.py:507-550).  d_anchors: the standard grid of (image_height, image_width).  Candidates, keys and counts are
 * written exactly as by cldet_decode_filter (anchor indices in the reference's concatenated order), so
 * cldet_sort_candidates / cldet_nms_sorted / cldet_gather_detections follow unchanged. */
int cldet_decode_filter_head(const float* const* h_cls_levels, const float* const* h_reg_levels, int num_levels,
                             int image_height, int image_width, int is_logits, const float* d_anchors, int num_images,
                             int num_classes, float score_thresh, cldet_candidate* d_candidates, uint64_t* d_keys,
                             int64_t capacity, int32_t* d_counts, void* stream);

/* Order each image's candidates by (score descending, anchor ascending) -- the order torchvision's stable
 * descending sort gives the reference's anchor-ordered candidate list -- and optionally keep only the first
 * `topk` of them (topk <= 0: keep all; the reference has no top-k).  With topk > 0 a 3-pass radix
 * select on the score bits finds the k-th score per image first, so only ~topk candidates are ordered.
 * Writes the ordered list to d_sorted [N, sorted_capacity] and min(count, topk) to d_sorted_counts.
 * max_count is an upper bound on d_counts[j] known to the host (e.g. after reading the counts, or `capacity`). */
size_t cldet_sort_workspace_bytes(int num_images, int64_t max_count, int topk);
int cldet_sort_candidates(const cldet_candidate* d_candidates, const uint64_t* d_keys, const int32_t* d_counts,
                          int num_images, int64_t capacity, int64_t max_count, int topk, cldet_candidate* d_sorted,
                          int64_t sorted_capacity, int32_t* d_sorted_counts, void* d_workspace, size_t workspace_bytes,
                          void* stream);

size_t cldet_nms_workspace_bytes(int num_images, int64_t max_count);

/* Per-class greedy NMS over each image's SORTED candidate list (torchvision.ops.batched_nms semantics):
 *   mode 0 = follow torchvision's rule (coordinate trick unless 4*count > vanilla_numel_limit),
 *   mode 1 = always coordinate trick  (boxes + label*(max_coord+1) in fp32, then plain NMS),
 *   mode 2 = always vanilla           (raw coordinates, only same-label pairs suppress).
 * Suppress iff inter / ((area_i + area_j) - inter) > iou_thresh (strict, fp32, no FMA).
 * Outputs: d_keep [N, capacity] int32 = positions (into the sorted list) of kept boxes in order; d_keep_counts [N]
 * (-1 = error marker: that image's count exceeds max_count, the size the workspace was made for). */
int cldet_nms_sorted(const cldet_candidate* d_sorted, const int32_t* d_sorted_counts, int num_images, int64_t capacity,
                     int64_t max_count, float iou_thresh, int mode, int64_t vanilla_numel_limit, int32_t* d_keep,
                     int32_t* d_keep_counts, void* d_workspace, size_t workspace_bytes, void* stream);

/* cldet_nms_sorted followed by cldet_gather_detections in the SAME launch chain: the block that resolves an image also
 * gathers its kept candidates into d_scores [N,capacity], d_labels [N,capacity] int64, d_boxes [N,capacity,4] (one launch
 * fewer on a latency-bound chain).  A negative d_keep_counts[j] is an error marker: image j held more candidates than
 * max_count, for which the workspace was sized (nothing is truncated silently). */
int cldet_nms_gather_sorted(const cldet_candidate* d_sorted, const int32_t* d_sorted_counts, int num_images, int64_t capacity,
                            int64_t max_count, float iou_thresh, int mode, int64_t vanilla_numel_limit, int32_t* d_keep,
                            int32_t* d_keep_counts, float* d_scores, int64_t* d_labels, float* d_boxes, void* d_workspace,
                            size_t workspace_bytes, void* stream);

/* torchvision.ops.nms / batched_nms on caller-provided boxes (drop-in for model.py:540, persuado_label.py:116):
 * d_boxes [K,4], d_scores [K], d_idxs [K] int64 (NULL = plain nms).  d_keep [K] int64 receives ORIGINAL indices in
 * score-descending order (ties by ascending index); *d_keep_count int32.  Workspace: cldet_batched_nms_workspace_bytes(K). */
size_t cldet_batched_nms_workspace_bytes(int64_t num_boxes);
int cldet_batched_nms(const float* d_boxes, const float* d_scores, const int64_t* d_idxs, int64_t num_boxes,
                      float iou_thresh, int mode, int64_t vanilla_numel_limit, int64_t* d_keep, int32_t* d_keep_count,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* Gather the kept candidates into dense outputs: scores [N,capacity], labels [N,capacity] int64, boxes [N,capacity,4]. */
int cldet_gather_detections(const cldet_candidate* d_sorted, const int32_t* d_keep, const int32_t* d_keep_counts,
                            int num_images, int64_t capacity, int64_t max_keep, float* d_scores, int64_t* d_labels,
                            float* d_boxes, void* stream);

/* ---- evaluator post-processing (evaluator.py:329-361) ----
 * From the padded detections of cldet_gather_detections: divide boxes by the image's resize scale (true fp32 division),
 * convert xyxy -> xywh (COCO), drop score < score_threshold, and write compact records ordered by image then rank.
 * d_offsets [N+1] int32 out: exclusive scan of the per-image record counts (d_offsets[N] = total).
 * d_records must hold num_images * capacity records. */
typedef struct cldet_coco_record {   /* 32 bytes */
    int32_t image;                  /* index into the batch */
    int32_t label;
    float score;
    float x, y, w, h;               /* COCO bbox */
    int32_t pad;
} cldet_coco_record;
int cldet_coco_results(const float* d_scores, const int64_t* d_labels, const float* d_boxes, const int32_t* d_counts,
                       const float* d_scales, int num_images, int64_t capacity, float score_threshold,
                       cldet_coco_record* d_records, int32_t* d_offsets, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLDET_H_ */
