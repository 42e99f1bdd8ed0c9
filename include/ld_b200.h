/* ld_b200 -- C-ABI of the B200-native LD (Localization Distillation) train-step hot path.
 *
 * The reference (HikariTJU/LD, an mmdetection 2.10 fork) has NO C-ABI / FFI boundary: it is 100 %
 * Python and its "plugin API" is mmcv's string->class registry.  This header is
 * the drop-in boundary underneath that registry: every entry point replaces the body of one
 * reference Python function on the path (cited per function as file:line under /root/reference)
 * and is what a binding from the reference side would call (INTEGRATION.md shows the ctypes stub).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call is
 *     asynchronous on that stream and never synchronises the device;
 *   - return value: 0 on success, otherwise a cudaError_t (>0) or an LDB_E* code (<0);
 *     ldb200_last_error() returns a message for the calling thread's last failure;
 *   - "row space": per-anchor arrays are laid out level-major, then image, then (y,x):
 *         row(l, n, a) = lvl_off[l] * N + n * A_l + a,   A_l = H_l*W_l,  lvl_off = prefix sum of A_l
 *     which is exactly the reference's per-level `[N, A_l]` stacking (ld_head.py:433-442) with the
 *     five levels stored back to back, so one kernel launch covers all levels.
 */
#ifndef LD_B200_H_
#define LD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDB_MAX_LEVELS 8
#define LDB_NUM_TERMS 9 /* loss_cls, loss_bbox, loss_dfl, loss_ld, loss_ld_vlr, loss_kd, loss_kd_neg, loss_im, sum_weight_targets */

#define LDB_EINVAL (-1)   /* bad argument */
#define LDB_ENOSUP (-2)   /* unsupported configuration */
#define LDB_EARCH (-3)    /* device is not sm_100 */

enum ldb_dtype { LDB_F32 = 0, LDB_BF16 = 1 };
enum ldb_im_mode { LDB_IM_FINEGRAINED = 0, LDB_IM_CENTER = 1 /* 'fitnet' / 'decouple' region */ };

const char* ldb200_last_error(void);
int ldb200_version(void);
/* 0 if the current device can run the kernels (compute capability 10.x), LDB_EARCH otherwise. */
int ldb200_check_device(void);

/* Pyramid description shared by the target and loss entry points. */
typedef struct {
  int num_levels;                 /* L <= LDB_MAX_LEVELS */
  int num_imgs;                   /* N */
  int h[LDB_MAX_LEVELS];          /* feature map height per level */
  int w[LDB_MAX_LEVELS];          /* feature map width  per level */
  int stride[LDB_MAX_LEVELS];     /* anchor stride per level (8,16,32,64,128) */
} ldb_pyramid;

/* ------------------------------------------------------------------------------------------
 * Target generation.  Replaces LDHead.get_targets/_get_target_single (mmdet/models/dense_heads/
 * ld_head.py:377-577) = AnchorGenerator.grid_anchors/valid_flags (core/anchor/anchor_generator.py:
 * 233-328), ATSSAssigner.assign (core/bbox/assigners/atss_assigner.py:33-181), PseudoSampler
 * (samplers/pseudo_sampler.py:23-41), ATSSAssigner.get_vlr_region (atss_assigner.py:183-298),
 * LDHead.get_im_region (ld_head.py:580-611), unmap + images_to_levels.
 * Integer outputs are bit-exact with the reference CPU path; vlr (an IoU value) is float-equal.
 *
 *   gt_boxes    [sumG,4] f32 xyxy pixels, all images back to back
 *   gt_labels   [sumG]   i32
 *   gt_offsets  [N+1]    i32 prefix offsets into the two arrays above (every image needs G >= 1,
 *                            the reference itself breaks on G == 0: atss_assigner.py:203-216)
 *   pad_hw      [N,2]    i32 img_meta['pad_shape'][:2] (valid-flag extent)
 * outputs (row space, R = N * sum_l A_l rows):
 *   labels      [R] i32 (num_classes = background)     label_weights [R] f32
 *   bbox_targets[R,4] f32 (assigned gt, pixels; 0 for non-positives)
 *   vlr         [R] f32 (IoU with the banded gt, else 0)   im_region [R] f32 {0,1}
 *   gt_inds     [R] i32 (0 = bg, k+1 = k-th gt of the image) -- may be NULL
 *   sparse_rows [R] i32 compacted rows with (positive || vlr > 0), arbitrary order
 *   fg_rows     [R] i32 compacted rows with im_region > 0, arbitrary order
 *   counts      i32[4 + 3*LDB_MAX_LEVELS]: [0]=n_sparse [1]=n_fg [2]=num_total_pos (sum_img max(P,1))
 *               [3]=sum_img P, then per level: pos count, vlr count, fg count
 *   workspace   ldb200_targets_workspace_bytes() bytes
 */
size_t ldb200_targets_workspace_bytes(const ldb_pyramid* pyr, int sum_gt);
int ldb200_ld_targets(const ldb_pyramid* pyr, const float* gt_boxes, const int32_t* gt_labels,
                      const int32_t* gt_offsets, int sum_gt, const int32_t* pad_hw, int topk,
                      int im_mode, int num_classes, int32_t* labels, float* label_weights,
                      float* bbox_targets, float* vlr, float* im_region, int32_t* gt_inds,
                      int32_t* sparse_rows, int32_t* fg_rows, int32_t* counts, void* workspace,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Loss hyper-parameters (config keys of LDHead, ld_head.py:46-71 + gfl_head.py:77-100).
 */
typedef struct {
  int num_classes;      /* C = cls_out_channels (80; 81 for the LDv2 configs) */
  int reg_max;          /* 16 -> 17 bins */
  int feat_channels;    /* 256 (ld_head.py:153-154 hard-codes it) */
  float qfl_weight, qfl_beta;          /* QualityFocalLoss  (gfocal_loss.py:77-135) */
  float giou_weight, giou_eps;         /* GIoULoss          (iou_loss.py:325-360)   */
  float dfl_weight;                    /* DistributionFocalLoss (gfocal_loss.py:138-179) */
  float ld_weight, ld_T;               /* KnowledgeDistillationKLDivLoss (kd_loss.py:39-88) */
  float ld_vlr_weight, ld_vlr_T;
  float kd_weight, kd_T;
  float im_weight;                     /* IMLoss (kd_loss.py:99-120) */
} ldb_loss_cfg;

/* Pre-pass over the positive rows.  Replaces ld_head.py:196-207 (weight_targets, Integral,
 * distance2bbox, IoU score) and the sum feeding ld_head.py:362:
 *   score [R] f32: IoU(decoded student box, target) on positives (must be zero elsewhere: the
 *                  call zero-fills it), partial[0] += sum_i weight_targets_i
 * so that the global normalisers (ld_head.py:338-341, 362-363) can be all-reduced BEFORE the fused
 * kernel emits final gradients. */
int ldb200_ld_loss_prepass(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls,
                           const float* reg, const int32_t* labels, const float* bbox_targets,
                           const int32_t* sparse_rows, const int32_t* counts, float* score,
                           float* wsum, void* stream);

/* Fused LDHead.loss_single x L (ld_head.py:116-282) forward AND gradient in one launch:
 * dense QFL over all R x C logits, and on the compacted rows GIoU + DFL + LD-KL (positives),
 * LD-KL on the Valuable Localization Region, KD-KL on class logits, feature-imitation MSE.
 *
 *   cls, t_cls [R,C] f32 ; reg, t_reg [R,4*(reg_max+1)] f32 (NHWC rows, row space)
 *   x, t_x     [R,F] feat_dtype (student / teacher FPN features)
 *   norm       f32[2] device: [0] num_total_samples (already max(.,1), ld_head.py:341)
 *                             [1] avg_factor (sum weight_targets + 1e-6, reduced: ld_head.py:362-363)
 *   term_grad_host f32[LDB_NUM_TERMS-1][LDB_MAX_LEVELS] upstream d(total)/d(term_l) (all 1 for
 *                  BaseDetector._parse_losses, base.py:198-209); NULL = no gradients
 *   losses     f32[LDB_NUM_TERMS][LDB_MAX_LEVELS] device (zero-filled by the call): the 8 terms per
 *              level with loss_bbox/loss_dfl already divided by avg_factor (ld_head.py:364-365)
 *              and row 8 = sum of weight_targets per level
 *   d_cls [R,C], d_reg [R,4*(reg_max+1)] grad_dtype (d_reg zero-filled by the call)
 *   d_im  [n_fg,F] grad_dtype: gradient w.r.t. x at fg_rows[i] (compact; scatter-add downstream)
 *   d_cls_ld / d_reg_ld: row strides in elements (0 = dense); padding columns of d_cls are left
 *              untouched, those of d_reg are zero (so d_reg can feed a TMA operand with ld % 8 == 0)
 */
int ldb200_ld_loss_fused(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls,
                         const float* reg, const float* t_cls, const float* t_reg, const void* x,
                         const void* t_x, int feat_dtype, const int32_t* labels,
                         const float* label_weights, const float* bbox_targets, const float* vlr,
                         const float* score, const int32_t* sparse_rows, const int32_t* fg_rows,
                         const int32_t* counts, const float* norm, const float* term_grad_host,
                         float* losses, void* d_cls, void* d_reg, void* d_im, int d_cls_ld, int d_reg_ld,
                         int grad_dtype, void* stream);

/* General-Instance imitation region.  Replaces LDHead.get_gi_region (ld_head.py:613-638), the row selection of
 * imitation_method='gibox' (ld_head.py:170-176; the LDHead constructor default): per level, over all N*A_l
 * locations, GI score = max_c |sigmoid(t_cls) - sigmoid(cls)|, GI box = teacher's decoded box where the teacher
 * scores higher at that class else the student's (Integral + distance2bbox, grid units), then the first `topk`
 * keeps of torchvision.ops.nms(boxes, scores, iou_thr) (reference: 0.3, 10).
 *   cls, t_cls [R,C] f32 ; reg, t_reg [R,4*(reg_max+1)] f32 (row space)
 *   counts_in  the counts vector of ldb200_ld_targets
 *   gi_rows    i32[num_levels*topk]: picked rows, levels back to back in pick order (first counts_out[1] valid)
 *   counts_out counts_in with [1] = number of picked rows and the per-level "fg" entries = picks per level: pass
 *              (gi_rows, counts_out) to ldb200_ld_loss_fused / ldb200_scatter_add_rows as (fg_rows, counts)
 *   workspace  ldb200_gi_region_workspace_bytes() bytes, 16-byte aligned */
size_t ldb200_gi_region_workspace_bytes(const ldb_pyramid* pyr, int topk);
int ldb200_gi_region(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls, const float* reg,
                     const float* t_cls, const float* t_reg, float iou_thr, int topk, const int32_t* counts_in,
                     int32_t* gi_rows, int32_t* counts_out, void* workspace, void* stream);

/* GFLv2 Distribution-Guided Quality Predictor on row-space head outputs.  Replaces the tail of
 * GFocalHead.forward_single (gfocal_head.py:198-215): per row, softmax of the four side distributions, their top-
 * `topk` probabilities (+ mean when add_mean), reg_conf = conv1x1(4*(topk+add_mean) -> hid) + ReLU + conv1x1(hid -> 1)
 * + sigmoid = quality, and the joint score sigmoid(cls_feat) * quality.
 *   reg [rows, reg_ld>=4*bins] f32, cls_feat [rows, cls_ld>=C] f32 (ld 0 = dense)
 *   W1 [hid, 4*(topk+add_mean)] = reg_conf.0.weight, b1 [hid], W2 [hid] = reg_conf.2.weight, b2 [1] (device)
 *   quality [rows] f32, score [rows, C] f32
 * Backward (autograd of the same lines): from d_score (d loss / d joint score) and an optional extra gradient on
 * cls_feat (d_kd: the KD term of LDv2Head, ld_gflv2.py:242-246) it writes d_cls_feat, ADDS the predictor's
 * contribution to d_reg and accumulates dW1/db1/dW2/db2 (fp32, atomics).  Gradient tensors are grad_dtype
 * (LDB_F32 / LDB_BF16) with row strides dscore_ld / dreg_ld (0 = dense). */
int ldb200_dgqp_forward(const float* reg, int reg_ld, const float* cls_feat, int cls_ld, int rows, int C, int bins,
                        int topk, int add_mean, int hid, const float* W1, const float* b1, const float* W2,
                        const float* b2, float* quality, float* score, void* stream);
int ldb200_dgqp_backward(const float* reg, int reg_ld, const float* cls_feat, int cls_ld, int rows, int C, int bins,
                         int topk, int add_mean, int hid, const float* W1, const float* b1, const float* W2,
                         const float* b2, const float* quality, const void* d_score, const void* d_kd,
                         int dscore_ld, void* d_cls_feat, void* d_reg, int dreg_ld, int grad_dtype, float* dW1,
                         float* db1, float* dW2, float* db2, void* stream);

/* LDv2Head variants (ld_gflv2.py:116-375, configs/ldv2) of the three loss-side entry points.
 * Differences to the LDHead versions above: `score` [R,C] = joint scores sigmoid(cls_feat) * quality (probabilities,
 * from ldb200_dgqp_forward) feed QFL as F.binary_cross_entropy (gfocal_loss.py:29-49, use_sigmoid=False) and
 * weight_targets = max_c score without sigmoid (ld_gflv2.py:200); the KD term compares the student's cls_feat logits
 * with the teacher's (ld_gflv2.py:242-246, 327); C = cfg->num_classes = cls_out_channels = num_fg + 1 and labels >=
 * num_fg are background; the GI score is |t_cls_feat - score| with no sigmoid on either side.
 * Gradients: d_score (w.r.t. the joint score, dense) and d_kd (KD term w.r.t. cls_feat, zero-filled by the call,
 * non-zero on positive rows) both [R, d_cls_ld]; feed both to ldb200_dgqp_backward. */
int ldb200_ld_loss_prepass_v2(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* score, const float* reg,
                              const int32_t* labels, const float* bbox_targets, const int32_t* sparse_rows,
                              const int32_t* counts, int num_fg, float* iou_score, float* wsum, void* stream);
int ldb200_ld_loss_fused_v2(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* score, const float* cls_feat,
                            const float* reg, const float* t_cls_feat, const float* t_reg, const void* x,
                            const void* t_x, int feat_dtype, const int32_t* labels, const float* label_weights,
                            const float* bbox_targets, const float* vlr, const int32_t* sparse_rows,
                            const int32_t* fg_rows, const int32_t* counts, const float* norm,
                            const float* term_grad_host, int num_fg, float* losses, void* d_score, void* d_kd,
                            void* d_reg, void* d_im, int d_cls_ld, int d_reg_ld, int grad_dtype, void* stream);
int ldb200_gi_region_v2(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* score, const float* reg,
                        const float* t_cls_feat, const float* t_reg, float iou_thr, int topk, const int32_t* counts_in,
                        int32_t* gi_rows, int32_t* counts_out, void* workspace, void* stream);

/* Fused ResNet stem (resnet.py:626-629: conv1 7x7/2 pad 3 + eval-BN + ReLU) for one or two networks that share the
 * input image (kd_one_stage.py:60-66: student and frozen teacher), without the im2col buffer of ldb200_stem_im2col:
 *   img [N,3,H,W] f32 NCHW; weights [64*nets, 192] bf16, row = output channel (network 0 first),
 *   K = (r*7+s)*3+c zero-padded from 147 to 192; scaleX/shiftX [64] f32 = folded BN of network X;
 *   outX [N, HO, WO, 64] bf16 NHWC with HO = (H-1)/2+1, WO = (W-1)/2+1.  nets = 1 ignores the *1 arguments.
 * Opt-in from the host side (LDB200_FUSED_STEM=1) until its first GPU run is on record. */
int ldb200_stem_tc(const float* img, int N, int H, int W, const void* weights, int nets, const float* scale0,
                   const float* shift0, void* out0, const float* scale1, const float* shift1, void* out1, void* stream);

/* dst[rows[i], :] += src[i, :]  (rows unique).  Adds the compact IM gradient into the dense
 * gradient of the student FPN map. */
int ldb200_scatter_add_rows(void* dst, const void* src, const int32_t* rows, const int32_t* n_rows_dev,
                            int max_rows, int channels, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense LD-KL micro-benchmark (BASELINE.json configs[4]); replaces
 * knowledge_distillation_kl_div_loss + weight_reduce_loss forward and autograd backward
 * (kd_loss.py:10-36, losses/utils.py:27-54) on dense [rows, bins] logits:
 *   loss += loss_weight/avg_factor * sum_r w_r * T^2/bins * KL(softmax(t_r/T) || softmax(s_r/T))
 *   d_student[r,k] = grad_out * loss_weight/avg_factor * w_r * T/bins * (p_s - p_t)
 * student, teacher, d_student: [rows, bins] f32; weight [rows] f32; loss f32[1] (accumulated). */
int ldb200_kl_dense_fwd_bwd(const float* student, const float* teacher, const float* weight,
                            int64_t rows, int bins, float T, float loss_weight, float avg_factor,
                            float grad_out, float* loss, float* d_student, void* stream);

/* Input side: Resize(keep_ratio, cv2 INTER_LINEAR) -> RandomFlip -> Normalize(mean, std, to_rgb) ->
 * Pad(size_divisor) -> collate of the reference's train pipeline (mmdet/datasets/pipelines/transforms.py:26-600) for a
 * batch in one launch, bit-identical to the cv2 / numpy arithmetic of mmcv.imrescale / imflip / imnormalize / impad.
 * src_u8: the decoded uint8 HWC (BGR) images packed in one device buffer; imgs: device array of N descriptors
 * (source and resized sizes are computed by the host: mmcv.rescale_size); out [N, 3, Hp, Wp] f32, zero outside the
 * resized image; mean3 / std3: HOST float[3] in the OUTPUT channel order (RGB when to_rgb). */
typedef struct {
  long long src_off;   /* byte offset of the image inside src_u8 */
  int sh, sw;          /* source size */
  int dh, dw;          /* size after Resize */
  int flip;            /* 1: horizontal flip */
} ldb_prep_image;
int ldb200_preprocess_batch(const void* src_u8, const ldb_prep_image* imgs, int N, float* out, int Hp, int Wp,
                            const float* mean3, const float* std3, int to_rgb, void* stream);

/* Inference step after the path: GFLHead._get_bboxes (gfl_head.py:354-451) + multiclass_nms with
 * nms type 'nms' (core/post_processing/bbox_nms.py:70-193 -> mmcv.ops.batched_nms) for a whole batch, on row-space
 * head outputs: per level the nms_pre anchors of highest max_c sigmoid(cls) (exact top-k), Integral * stride,
 * distance2bbox clipped to img_hw (divided by scale_factors when given), score threshold, per-class greedy NMS on
 * mmcv's class-offset boxes, the first max_per_img survivors by score.
 *   cls [R, num_classes], bbox_pred [R, 4*bins] f32; img_hw [N,2] i32 (h, w); scale_factors [N,4] f32 or NULL;
 *   dets [N, max_per_img, 5] (x1,y1,x2,y2,score), labels [N, max_per_img] i32, counts [N] i32 (number of detections;
 *   -1 = more than 16384 (box, class) pairs above score_thr in that image: raise score_thr or lower nms_pre). */
size_t ldb200_get_bboxes_workspace_bytes(const ldb_pyramid* pyr, int num_classes, int nms_pre);
int ldb200_gfl_get_bboxes(const ldb_pyramid* pyr, const float* cls, const float* bbox_pred, int num_classes, int bins,
                          int use_sigmoid /* 1: cls = logits (GFLHead); 0: probabilities (GFocalHead joint scores) */,
                          const int32_t* img_hw, const float* scale_factors, int nms_pre, float score_thr, float iou_thr,
                          int max_per_img, float* dets, int32_t* labels, int32_t* counts, void* workspace, void* stream);

/* Stand-alone loss modules of the registry surface (`build_loss(cfg)(pred, target, weight, avg_factor, ...)` outside
 * LDHead): forward value and d/d pred in one launch, fp32.  row_loss (optional, [rows]) receives scale * w_r * l_r for
 * reduction='none'; loss (optional, f32[1], ACCUMULATED) += scale * sum_r w_r * l_r; d_pred (optional) = scale * w_r *
 * d l_r / d pred.  `scale` folds loss_weight and 1 / avg_factor (losses/utils.py:27-54, 57-100); weight may be NULL.
 *   qfl:  quality_focal_loss (gfocal_loss.py:8-50): pred [rows, C] logits (use_sigmoid = 1) or probabilities
 *         (use_sigmoid = 0, the GFLv2/LDv2 configs), label [rows] i32 (>= C or < 0 = background), score [rows]
 *   dfl:  distribution_focal_loss (gfocal_loss.py:53-74): pred [rows, bins] logits, label [rows] f32 in [0, bins-1)
 *   giou: giou_loss (iou_loss.py:85-102, eps 1e-6): pred, target [rows, 4] xyxy */
int ldb200_qfl_fwd_bwd(const float* pred, const int32_t* label, const float* score, const float* weight, int64_t rows,
                       int C, float beta, int use_sigmoid, float scale, float* row_loss, float* loss, float* d_pred,
                       void* stream);
int ldb200_dfl_fwd_bwd(const float* pred, const float* label, const float* weight, int64_t rows, int bins, float scale,
                       float* row_loss, float* loss, float* d_pred, void* stream);
int ldb200_giou_fwd_bwd(const float* pred, const float* target, const float* weight, int64_t rows, float eps, float scale,
                        float* row_loss, float* loss, float* d_pred, void* stream);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (bf16 x bf16 -> fp32 accumulate), NHWC.
 * Replaces the cuDNN `nn.Conv2d` calls of the reference's ResNet / FPN / GFLHead forward
 * (backbones/resnet.py:65-92,260-299; necks/fpn.py:170-221; dense_heads/gfl_head.py:161-183) and,
 * called with rotated weights, autograd's conv backward-data.
 *
 *   out[n, y*oys+oyo, x*oxs+oxo, co] = epi( sum_t sum_ci in[n, y*sy + tap_dy[t], x*sx + tap_dx[t], ci]
 *                                                        * weight[co, tap_wk[t] + ci] )
 *   epi(v) = relu?( v*scale[co] + bias[co] + add_src[n, y*add_H/HO, x*add_W/WO, co] ) * (mask_src > 0)
 *            (+ previous out when accumulate)
 * Out-of-range input coordinates read as zero (padding).  Cin must be a multiple of 64.
 */
typedef struct {
  const void* in;      /* [N, H, W, in_ld] bf16, channels [0, Cin) used */
  int N, H, W, Cin, in_ld;
  const void* weight;  /* [Cout, w_k] bf16, K-major; tap t / input channel ci at column tap_wk[t] + ci */
  int Cout, w_k;
  int ntaps;           /* <= 9 */
  int tap_dx[9], tap_dy[9], tap_wk[9];
  int sx, sy;          /* stride in input coordinates */
  int HO, WO;          /* logical output extent (tile space) */
  int flat;            /* 1: 1x1 stride-1 conv, pixels flattened to one dimension (no tile padding) */
  void* out;           /* [N, OH, OW, out_ld] bf16 or f32 */
  int OH, OW;          /* physical output extent (0 = HO, WO) */
  int oys, oyo, oxs, oxo; /* output pixel = (y*oys + oyo, x*oxs + oxo)   (0,0,0,0 = identity) */
  int out_ld, out_f32;
  const float* scale;  /* [Cout] or NULL */
  const float* bias;   /* [Cout] or NULL */
  const void* add_src; /* [N, add_H, add_W, add_ld] bf16 or NULL (nearest-resampled to the output grid) */
  int add_H, add_W, add_ld;
  const void* mask_src; /* bf16, same pixel addressing as out with row stride mask_ld, or NULL */
  int mask_ld;
  int relu;
  int accumulate;
  /* fp32-accurate ("parity") mode, 0 = off.  split = P in {2, 3}: `in` holds P stacked bf16 tensors
   * [P][N,H,W,in_ld] whose SUM is the fp32 activation (ldb200_split_bf16), `weight` rows hold P planes of `w_plane`
   * columns each ([Cout][P][w_plane], tap_wk relative to a plane); the kernel accumulates the P*(P+1)/2 plane
   * products of leading magnitude into the same fp32 TMEM accumulator (P = 3: relative error ~2^-23, i.e. the
   * fp32 convolution the reference runs; P = 2: ~2^-16). */
  int split, w_plane;
  int add_f32, mask_f32;  /* add_src / mask_src are fp32 tensors (fp32 activations of the parity mode) */
  /* optional [N, Cout/8, 2] f32, zero-filled by the caller: the epilogue ADDS, per image and group of 8 output
   * channels, the sum and the sum of squares of the values it stores -- the GroupNorm(32) statistics of a 256-channel
   * GFL tower conv (gfl_head.py:102-133) without a separate pass over the output.  Non-flat convs, Cout % 32 == 0. */
  float* gn_stats;
  /* split mode in two launches: split_pass 1 = leading plane product only, 2 = correction products only (0 = all in one
   * chain).  pre_src: optional fp32 tensor addressed like `out`, added to the raw accumulator before scale / bias --
   * pass 2 reads the result of pass 1 through it.  Shortens the accumulation chain of the large terms 6x. */
  int split_pass;
  const float* pre_src;
} ldb_conv_desc;

int ldb200_conv2d_tc(const ldb_conv_desc* desc, void* stream);
/* `count` (<= 5) convolutions sharing weights, taps, stride and epilogue configuration -- the five FPN levels of
 * a level-shared GFLHead conv (gfl_head.py:145-183 `multi_apply(forward_single, feats, scales)`) -- as ONE
 * persistent launch over the concatenated tile list (small levels no longer waste whole waves). */
int ldb200_conv2d_tc_multi(const ldb_conv_desc* descs, int count, void* stream);

/* Weight gradient of the same convolution (autograd's conv backward-weight), split-K over pixel
 * tiles with fp32 atomics:
 *   dw[co, tap_wk[t] + ci] += sum_{n,y,x} dy[n,y,x,co] * x[n, y*sy + tap_dy[t], x*sx + tap_dx[t], ci]
 * dy: [N,HO,WO,dy_ld] bf16, x: [N,H,W,x_ld] bf16, dw: [Cout, w_k] f32 (NOT zeroed by the call: gradients
 * of the level-shared head convs accumulate across the five FPN levels). */
typedef struct {
  const void* dy; int N, HO, WO, Cout, dy_ld;
  const void* x;  int H, W, Cin, x_ld;
  int ntaps; int tap_dx[9], tap_dy[9], tap_wk[9];
  int sx, sy;
  float* dw; int w_k;
  const float* scale;     /* optional [Cout]: dw += scale[co] * G (folded eval-BN / Scale factor) */
  const float* w_master;  /* optional fp32 weights [Cout, w_k], needed for d_scale */
  float* d_scale;         /* optional [Cout]: d_scale[co] += <w_master[co,:], G[co,:]> (gradient of the factor) */
  int split;              /* 0, or P in {2,3}: dy and x are P stacked bf16 planes each (see ldb_conv_desc.split) */
  int split_pass;         /* 0 all plane products, 1 leading product only, 2 corrections only (two accumulating launches) */
} ldb_wgrad_desc;
int ldb200_conv2d_wgrad_tc(const ldb_wgrad_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-bound glue around the convolutions.  Activations are NHWC bf16 (`dtype` = LDB_BF16, the throughput mode) or
 * NHWC fp32 (`dtype` = LDB_F32, the fp32-accurate "parity" mode whose convolutions run on split-bf16 operands).
 */
/* fp32 -> P stacked bf16 planes with x ~= sum of planes (P = 3: exact to 2^-24): x [rows, ld_in] f32, first C columns
 * used; out [P][rows, ld_out] bf16, columns [C, ld_out) zero-filled; ld_out % 8 == 0.  Operand preparation for
 * ldb_conv_desc.split / ldb_wgrad_desc.split. */
int ldb200_split_bf16(const float* x, int64_t rows, int C, int ld_in, int planes, void* out, int ld_out, void* stream);
/* 7x7/2 pad-3 stem conv as a GEMM operand: img [N,3,H,W] f32 (NCHW) -> out [N*HO*WO, 192] bf16 with
 * column (r*7+s)*3+c, zero-padded from 147 (resnet.py:524-570, 626). */
int ldb200_stem_im2col(const float* img, void* out, int N, int H, int W, int out_dtype, void* stream);
/* Deformable convolution v1 forward, sampling half (frozen DCN teachers: resnext.py:64-75 / resnet.py:172-184 ->
 * mmcv.ops.DeformConv2dPack; mmcv is not vendored in the reference, the published DCNv1 algorithm is restated):
 *   col[n,yo,xo, k*C + c] = bilinear(x[n,:,:,c], yo*stride - pad + (k/kw)*dil + dy_k, xo*stride - pad + (k%kw)*dil + dx_k)
 * x [N,H,W,C] (dtype bf16 / f32), offset [N,HO,WO,offset_ld] f32 with channel 2*(g*kh*kw + k) = dy, +1 = dx (the layout
 * of mmcv and torchvision.ops.deform_conv2d), col [N*HO*WO, kh*kw*C] (dtype).  The GEMM half is ldb200_conv2d_tc in
 * flat mode on `col` with the conv's ordinary packed weights [Cout, kh*kw*Cin]. */
int ldb200_deform_im2col(const void* x, const float* offset, void* col, int N, int H, int W, int C, int HO, int WO,
                         int kh, int kw, int stride, int pad, int dilation, int deform_groups, int offset_ld, int dtype,
                         void* stream);
/* 3x3 stride-2 pad-1 max pool (resnet.py:630). */
int ldb200_maxpool3x3s2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream);
/* GroupNorm(G) [+ReLU] of mmcv ConvModule in the GFL towers (gfl_head.py:102-133, 177-180);
 * x,y: [N, HW, C] bf16; stats: [N, G, 2] f32 (sum, sum of squares; zero-filled by gn_stats). */
int ldb200_gn_stats(const void* x, float* stats, int N, int HW, int C, int G, int dtype, void* stream);
/* The same GroupNorm on up to LDB_GN_MAX_SEG tensors sharing gamma / beta -- the pyramid levels of one level-shared
 * tower layer -- in ONE launch each (forward apply; backward = reduction pass + apply pass).  The statistics of the
 * forward pass may come from ldb200_gn_stats or from the producing convolution's epilogue (ldb_conv_desc.gn_stats). */
#define LDB_GN_MAX_SEG 5
typedef struct {
  const void* x;        /* [N, HW, C] conv output (pre-norm) */
  const void* y;        /* backward: relu(GN(x)) of the forward pass, or NULL (mask recomputed from x) */
  const void* dy;       /* backward: gradient w.r.t. y */
  void* out;            /* apply: y; backward: dx */
  const float* stats;   /* [N, G, 2] sum, sum of squares of x */
  float* red_ws;        /* backward: [N, G, 2] f32 scratch */
  int red_zeroed;       /* backward: 1 = the caller zero-filled red_ws already */
  int N, HW;
} ldb_gn_seg;
int ldb200_gn_apply_multi(const ldb_gn_seg* segs, int count, const float* gamma, const float* beta, int C, int G,
                          float eps, int relu, int dtype, void* stream);
/* backward: a segment's `y` may be NULL -- the ReLU mask is then recomputed from x, the statistics and gamma / `beta`
 * (y > 0 <=> x*a + b > 0), which saves reading y in both passes. */
int ldb200_gn_relu_bwd_multi(const ldb_gn_seg* segs, int count, const float* gamma, const float* beta, float* dgamma,
                             float* dbeta, int C, int G, float eps, int dtype, void* stream);
int ldb200_gn_apply(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                    int N, int HW, int C, int G, float eps, int relu, int dtype, void* stream);
/* backward of y = relu(GN(x)): dx [N,HW,C] bf16, dgamma/dbeta [C] f32 ACCUMULATED (shared towers),
 * red_ws: [N, G, 2] f32 scratch. */
int ldb200_gn_relu_bwd(const void* dy, const void* y, const void* x, const float* stats, const float* gamma,
                       float* red_ws, float* dgamma, float* dbeta, void* dx, int N, int HW, int C, int G,
                       float eps, int dtype, void* stream);
/* backward of the FPN top-down `fine += nearest_upsample(coarse)` (fpn.py:182-191): dcoarse += gather-sum(dfine). */
int ldb200_upsample_add_bwd(const void* dfine, void* dcoarse, int N, int Hf, int Wf, int Hc, int Wc, int C,
                            int dtype, void* stream);
/* out = dy * (y > 0). */
int ldb200_relu_bwd(const void* dy, const void* y, void* out, int64_t numel, int dtype, void* stream);
/* out[c] += sum_rows x[row, c]; x [rows, ld] (dtype LDB_BF16 / LDB_F32): conv-bias and BN-beta gradients. */
int ldb200_colsum(const void* x, float* out, int64_t rows, int C, int ld, int dtype, void* stream);
/* eval-mode BatchNorm (norm_eval=True, resnet.py:639-649) folded into the conv epilogue's per-channel affine:
 * scale = gamma*rsqrt(var+eps), shift = beta - mean*scale; and the adjoint for the trainable gamma/beta
 * (bn_fold_bwd ACCUMULATES into dgamma / dbeta, i.e. straight into the parameter gradient buffers). */
int ldb200_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                   float* scale, float* shift, int C, void* stream);
int ldb200_bn_fold_bwd(const float* d_scale, const float* d_shift, const float* mean, const float* var, float eps,
                       float* dgamma, float* dbeta, int C, void* stream);
/* IMLoss stand-alone (kd_loss.py:91-120): loss[0] += loss_weight * mean((x - target)^2); dx = d loss / d x (may be
 * NULL).  Inside LDHead the same term is part of ldb200_ld_loss_fused. */
int ldb200_mse_fwd_bwd(const void* x, const void* target, int64_t numel, float loss_weight, int dtype, float* loss,
                       void* dx, void* stream);
/* torch.optim.SGD (momentum, weight decay, dampening 0) on a flat fp32 parameter buffer;
 * grad_scale folds the data-parallel mean.  configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py:53. */
int ldb200_sgd_flat(float* w, const float* g, float* momentum_buf, int64_t n, float lr, float momentum,
                    float weight_decay, float grad_scale, int first_step, void* stream);
/* same update with the hyper-parameters read from device memory, hyper_dev = f32[5]
 * {lr, momentum, weight_decay, grad_scale, first_step(0/1)}: the launch can sit inside a CUDA graph
 * while the warm-up schedule changes lr every iteration. */
int ldb200_sgd_flat_dev(float* w, const float* g, float* momentum_buf, int64_t n, const float* hyper_dev,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LD_B200_H_ */
