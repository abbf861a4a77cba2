/*
 * dskd_b200 -- C ABI of the B200-native (sm_100a) DSKD distillation hot path.
 *
 * This is the drop-in boundary.  The reference (smilekitty7/DSKD) has no
 * native code: its hot path is inline PyTorch inside
 *   mmdet/models/dense_heads/gfl_deformable_detr_head_il.py  (`loss`, :412-1195)
 * reduced by registry loss modules (mmdet/models/losses/{mse_loss,kd_loss,utils}.py) and fed by
 *   mmdet/core/bbox/assigners/gfl_hungarian_assigner.py.
 * Each entry point below names the reference lines it replaces.  The Python host side
 * (dskd_b200/losses.py, assigner.py) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer adds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - Every pointer named `d_*` is a DEVICE pointer, `h_*` a HOST pointer.  The caller owns and
 *     allocates every input, output and workspace; the library never allocates across the
 *     boundary and keeps no mutable global state except a thread-local error string.
 *   - Device entry points are asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronise, and are re-entrant.  Host entry points (dskd_lsap_*) are pure CPU.
 *   - Return 0 on success, a negative DskdStatus otherwise; dskd_last_error() describes it.
 *     No C++ exception crosses; the library never calls exit().
 *   - All floating point is IEEE fp32 (the reference runs the loss under force_fp32,
 *     head_il.py:411); indices are int32 / int64 as noted; LSAP is float64 like SciPy.
 */
#ifndef DSKD_B200_H_
#define DSKD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSKD_ABI_VERSION 2
#define DSKD_MAX_LEVELS 8

typedef enum {
  DSKD_OK = 0,
  DSKD_EINVAL = -1,            /* bad argument (null pointer, size, alignment, enum) */
  DSKD_ECUDA = -2,             /* a CUDA runtime call / launch failed                */
  DSKD_EUNSUPPORTED_ARCH = -3, /* current device is not compute capability 10.x      */
  DSKD_EINFEASIBLE = -4,       /* LSAP: cost matrix infeasible / has NaN or -inf     */
} DskdStatus;

int dskd_abi_version(void);
const char* dskd_last_error(void);
/* Number of CUDA kernels this library has launched in this process (statistics for bench.py). */
uint64_t dskd_launch_count(void);
/* sizeof() of the argument structs as this library was compiled (0 DskdLevel, 1 DskdDsgfdMseArgs, 2 DskdDsgfdKlArgs,
 * 3 DskdDsgfdStepArgs, 4 DskdQmemArgs; -1 otherwise): lets a foreign-language binding check its struct mirror. */
int64_t dskd_struct_size(int32_t which);
/* DSKD_OK iff the current CUDA device is sm_100 (B200).  There is no fallback arch. */
int dskd_check_device(void);

/* One pyramid level of the 4-level feature stack (head_il.py:680-682 `N,C,H,W = features.shape`). */
typedef struct {
  int32_t H, W;
  int64_t cell_offset; /* first cell of this level inside a per-image cell axis (sum of H*W of lower levels);
                          also the token offset inside encoder memory [S,N,C] (head_il.py:869-880) */
} DskdLevel;

/* Feature storage. */
enum { DSKD_LAYOUT_NCHW = 0,   /* per level [N,C,H,W] contiguous: neck outputs (head_il.py:678-679)      */
       DSKD_LAYOUT_SNC = 1 };  /* one [S,N,C] contiguous tensor: encoder memory (transformer.py:1053)    */

/* ---------------------------------------------------------------------------------------------
 * Mask construction (head_il.py:685-706, :742-756, :883-914, :1107-1122; _fg_bk.py:548-566)
 * ------------------------------------------------------------------------------------------- */
enum { DSKD_MASK_DECODE_V1 = 0,  /* row = softmax_c(|hs_T[id_soft] - hs_S[id_pred]|)  head_il.py:705-706 */
       DSKD_MASK_DECODE_V2 = 1 };/* row = softmax_c(hs_T[id_soft])                     head_il.py:754     */

/* d_rows[p,:] for p < num_pairs.  d_hs_*: [num_rows, C]; d_id_*: int64 [num_pairs] (id_pred unused for V2). */
int dskd_mask_rows(int32_t mode, const float* d_hs_teacher, const float* d_hs_student,
                   const int64_t* d_id_soft, const int64_t* d_id_pred, int32_t num_pairs, int32_t C,
                   float* d_rows, void* stream);

/* Backward of dskd_mask_rows (DECODE_V1 only): given d_grad_rows = dLoss/d rows [num_pairs,C], adds
 * dLoss/d hs_S into d_grad_hs_student[id_pred[p],:] (caller zero-fills it; id_pred entries are distinct). */
int dskd_mask_rows_bwd(const float* d_hs_teacher, const float* d_hs_student, const int64_t* d_id_soft,
                       const int64_t* d_id_pred, const float* d_rows, const float* d_grad_rows,
                       int32_t num_pairs, int32_t C, float* d_grad_hs_student, void* stream);

/* Matched student ids (head_il.py:1453-1455 + :672): ascending indices q < n whose assigned label is
 * flagged in d_prev_mask (uint8[num_classes]); the first max_out of them go to d_ids (int64, the rest
 * of d_ids is set to 0), the total count to d_count[0] (int32).  Replaces `nonzero()` without a host sync:
 * callers that want the reference's IndexError on count < max_out read d_count back. */
int dskd_select_prev_queries(const int64_t* d_labels, int32_t n, const uint8_t* d_prev_mask, int32_t num_classes,
                             int32_t max_out, int64_t* d_ids, int32_t* d_count, void* stream);

/* Cell raster semantics. */
enum { DSKD_RASTER_OWNER_EXCL = 0, /* int32 owner = LAST box (highest pair index) whose half-open rect
                                      [floor(y1/img_h*H), ceil(y2/img_h*H)) x [floor(x1..), ceil(x2..)) holds the
                                      cell, -1 if none: the overwrite order of head_il.py:706               */
       DSKD_RASTER_BINARY_INCL = 1,/* float 1 inside any teacher box (inclusive +1 ends), 0 inside any GT box,
                                      sqrt: sg_out, head_il.py:898-914                                        */
       DSKD_RASTER_AREA_INCL = 2,  /* float sqrt(max_j 1/((dh+1)(dw+1))) inclusive ends: fg_only :1107-1119  */
       DSKD_RASTER_AREA_FGBK = 3 };/* as AREA_INCL but x scaled by H and y by W: _fg_bk.py:550-553 (sic)     */

/* d_boxes: fp32 [num_boxes,4] pixel xyxy, concatenated over images; d_box_start: int32 [N+1] prefix offsets;
 * d_gt_boxes/d_gt_start: same for the GT boxes (BINARY_INCL only, else NULL); d_img_hw: int32 [N,2] (h,w).
 * Output d_out: int32 (OWNER_EXCL) or fp32 (others) laid out [N, cells_per_image]; cells_per_image =
 * sum_l H_l*W_l and cell index = levels[l].cell_offset + h*W_l + w. */
int dskd_raster_cells(int32_t mode, const float* d_boxes, const int32_t* d_box_start,
                      const float* d_gt_boxes, const int32_t* d_gt_start, const int32_t* d_img_hw,
                      int32_t N, int32_t max_boxes_per_image, const DskdLevel* levels, int32_t num_levels,
                      int64_t cells_per_image, void* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DSG-FD, masked MSE, fused forward + backward (head_il.py:707-718 with MSELoss, mse_loss.py:9-57)
 *   loss = sum_{l,i,c,h,w} scale_l * M^2 * (T - S)^2,   dS = -2 * scale_l * M^2 * (T - S)
 * scale_l carries loss_weight, the 1/N of head_il.py:716-717 and the reduction ('sum': 1, 'mean':
 * 1/(C*H_l*W_l)).  Row-mask mode (owner + rows): also accumulates energy[p,c] = sum over the cells
 * box p owns of scale_l*(T-S)^2, from which loss = sum rows^2*energy and dLoss/d rows = 2*rows*energy.
 * Cell-mask mode (cell_weight): M = weight[i,cell]; the loss is accumulated into d_loss[0] (double).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t layout;               /* DSKD_LAYOUT_*                                                      */
  int32_t num_levels, N, C;
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* d_student[DSKD_MAX_LEVELS]; /* NCHW: one pointer per level; SNC: [0] only             */
  const float* d_teacher[DSKD_MAX_LEVELS];
  float* d_grad_student[DSKD_MAX_LEVELS];  /* same shapes; fully overwritten; NULL = forward only      */
  float scale[DSKD_MAX_LEVELS];
  int64_t cells_per_image;
  const int32_t* d_owner;       /* [N,cells_per_image] from DSKD_RASTER_OWNER_EXCL, or NULL           */
  const float* d_rows;          /* [num_pairs,C] from dskd_mask_rows (row-mask mode)                   */
  float* d_energy;              /* [num_pairs,C], caller zero-fills                                     */
  int32_t num_pairs;
  const float* d_cell_weight;   /* [N,cells_per_image] (cell-mask mode), or NULL                       */
  double* d_loss;               /* [1] cell-mask mode: loss accumulator, caller zero-fills             */
} DskdDsgfdMseArgs;

int dskd_dsgfd_mse_fwd_bwd(const DskdDsgfdMseArgs* args, void* stream);

/* Row-mask finish: loss[0] = sum_p,c rows^2*energy; d_grad_rows = 2*rows*energy (may alias energy). */
int dskd_dsgfd_mse_finish(const float* d_rows, const float* d_energy, int32_t num_pairs, int32_t C,
                          float* d_loss, float* d_grad_rows, void* stream);

/* Row-mask finish in one launch (one CTA per pair): criterion 0 (mse): d_loss_acc[0] (double, caller
 * zero-fills) += sum rows^2*energy and g = 2*rows*energy; criterion 1 (kl): g = d_energy_or_grad_rows as
 * accumulated by dskd_dsgfd_kl_fwd_bwd.  If d_grad_hs_student != NULL, adds the softmax/abs backward of g
 * (as dskd_mask_rows_bwd) into it. */
int dskd_dsgfd_rows_finish(int32_t criterion, const float* d_hs_teacher, const float* d_hs_student,
                           const int64_t* d_id_soft, const int64_t* d_id_pred, const float* d_rows,
                           const float* d_energy_or_grad_rows, int32_t num_pairs, int32_t C,
                           double* d_loss_acc, float* d_grad_hs_student, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DSG-FD, KL-over-H (the shipped config: KnowledgeDistillationKLDivLoss(T, 'sum'), kd_loss.py:12-43
 * applied to [C,H,W] so softmax runs over H; target = student*mask, detached; pred = teacher*mask).
 *   loss = sum_{l,i,c,w} scale_l * T^2/H_l * sum_h t_h (log t_h - logp_h)
 * No gradient reaches the student features; row-mask mode accumulates
 * d_grad_rows[p,c] = dLoss/d rows directly.  Both feature layouts.
 * One streaming pass (unshifted exponentials) plus a second tiny launch that redoes, with exact column maxima,
 * the (tile, channel) pairs whose sums left the fp32-safe range; d_workspace holds one channel mask per tile.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t num_levels, N, C;
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* d_student[DSKD_MAX_LEVELS];
  const float* d_teacher[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  float temperature;
  int64_t cells_per_image;
  const int32_t* d_owner;
  const float* d_rows;
  float* d_grad_rows;           /* [num_pairs,C], caller zero-fills; NULL = forward only               */
  int32_t num_pairs;
  const float* d_cell_weight;   /* cell-mask mode (no gradient at all)                                  */
  double* d_loss;               /* [1], caller zero-fills                                               */
  int32_t layout;               /* DSKD_LAYOUT_NCHW (one pointer per level) or DSKD_LAYOUT_SNC ([0] only) */
  void* d_workspace;            /* >= dskd_dsgfd_kl_workspace_bytes(...) bytes, 16-byte aligned            */
  int64_t workspace_bytes;      /* (no initialisation needed)                                             */
} DskdDsgfdKlArgs;

int64_t dskd_dsgfd_kl_workspace_bytes(int32_t N, int32_t num_levels, const DskdLevel* levels, int32_t C);
int dskd_dsgfd_kl_fwd_bwd(const DskdDsgfdKlArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DSG-FD in ONE call: everything gfl_deformable_detr_head_il.py:664-719 (or the sibling mask of
 * mask_mode) does for a batch, forward + backward, as a fixed sequence of launches on `stream`:
 * matched ids -> mask rows -> cell raster -> streaming kernel -> row finish.  This is what the
 * `DSGFeatureDistillLoss` module calls; the fine-grained entry points above remain for tests/tools.
 * ------------------------------------------------------------------------------------------- */
enum { DSKD_CRIT_MSE = 0, DSKD_CRIT_KL = 1 };
enum { DSKD_MODE_DECODE_V1 = 0, DSKD_MODE_DECODE_V2 = 1, DSKD_MODE_SG_OUT = 2, DSKD_MODE_FG_ONLY = 3,
       DSKD_MODE_FG_BK = 4 };
typedef struct {
  int32_t criterion, mask_mode, layout;
  int32_t num_levels, N, C;
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* d_student[DSKD_MAX_LEVELS];
  const float* d_teacher[DSKD_MAX_LEVELS];
  float* d_grad_student[DSKD_MAX_LEVELS]; /* NULL entries: no feature gradient (always for KL)           */
  float scale[DSKD_MAX_LEVELS];
  float temperature;
  int64_t cells_per_image;
  const float* d_hs_student;     /* [num_query_rows, C] (decode_v1)                                      */
  const float* d_hs_teacher;     /* [num_query_rows, C] (decode_v1 / v2)                                 */
  float* d_grad_hs_student;      /* [num_query_rows, C], fully overwritten; NULL = not needed            */
  int32_t num_query_rows;
  const int64_t* d_teacher_keepid;   /* [num_pairs] (decode_*)                                           */
  const int64_t* d_student_labels;   /* [num_query_rows] (decode_v1)                                     */
  const uint8_t* d_prev_mask;        /* [num_classes]   (decode_v1)                                      */
  int32_t num_classes;
  const float* d_boxes;          /* [num_pairs, 4] px xyxy, concatenated over images                     */
  const int32_t* d_box_start;    /* [N+1]                                                                */
  const float* d_gt_boxes;       /* sg_out only                                                          */
  const int32_t* d_gt_start;     /* sg_out only                                                          */
  const int32_t* d_img_hw;       /* [N,2]                                                                */
  int32_t num_pairs, max_boxes_per_image;
  float* d_loss;                 /* [1]                                                                  */
  int32_t* d_matched_count;      /* [1] decode_v1: number of student queries with a previous label       */
  void* d_workspace;             /* >= dskd_dsgfd_step_workspace_bytes(...) bytes, 256-byte aligned       */
  int64_t workspace_bytes;
  void* ev_kernel_begin;         /* optional cudaEvent_t pair recorded on `stream` right before / after   */
  void* ev_kernel_end;           /* the streaming kernel (bench.py's roofline timing); NULL = off         */
} DskdDsgfdStepArgs;

/* Workspace size: an upper bound from the sizes alone, and the exact figure for a filled-in argument struct
 * (criterion, levels, N, C, num_pairs are read). */
int64_t dskd_dsgfd_step_workspace_bytes(int32_t N, int64_t cells_per_image, int32_t num_pairs, int32_t C);
int64_t dskd_dsgfd_step_workspace_bytes_for(const DskdDsgfdStepArgs* args);
int dskd_dsgfd_step(const DskdDsgfdStepArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Query x memory contraction (north_star's "query x memory mask").  NOT in the reference
 * (no matmul exists in the IL head): an extension whose oracle is oracle/qmem.py.
 * Replaces the hard box rectangles of head_il.py:688-706 by a soft ownership computed on tcgen05 tensor
 * cores (tf32, fp32 accumulation in tensor memory, operands staged by TMA):
 *   z[s,j] = <memory[s,i,:], hs_T[keepid[j],:]> / (sqrt(C) * temperature)       j over the K_i detections of image i
 *   w[i,s] = sqrt( sum_j c_j e^z[s,j] / (1 + sum_j e^z[s,j]) )                   c_j = d_scores[j] (NULL: 1)
 * The [S, K_i] score matrix never reaches HBM.  d_cell_weight [N,S] is the `cell-mask` input of
 * dskd_dsgfd_mse_fwd_bwd (token index == cell index of the [S,N,C] layout).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t N, C;
  int64_t S;                      /* tokens per image (sum of H*W over levels)                              */
  const float* d_memory;          /* [S,N,C] teacher encoder memory (transformer.py:1053), 16-byte aligned  */
  const float* d_hs_teacher;      /* [num_query_rows, C] teacher decoder embeddings, last layer             */
  int32_t num_query_rows;
  const int64_t* d_keepid;        /* [num_pairs] rows of d_hs_teacher (deformable_detr_il.py:151)           */
  const float* d_scores;          /* [num_pairs] teacher confidences (`pred_scores`), or NULL               */
  const int32_t* d_box_start;     /* [N+1] prefix offsets of the per-image detections                       */
  int32_t num_pairs, max_per_image;
  float temperature;              /* the head's unused `temp` ctor argument (head_il.py:90,124), default 0.5 */
  float* d_cell_weight;           /* [N,S] out                                                              */
  void* d_workspace;              /* >= dskd_qmem_workspace_bytes(...) bytes, 256-byte aligned               */
  int64_t workspace_bytes;
} DskdQmemArgs;

int64_t dskd_qmem_workspace_bytes(int32_t N, int64_t S, int32_t C, int32_t max_per_image);
int dskd_qmem_cell_weights(const DskdQmemArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * BCDD (head_il.py:525-555, :1197-1222)
 * ------------------------------------------------------------------------------------------- */
/* Prototype sums + counts.  d_proto: [2, num_classes, C+1] (0 = teacher, 1 = student; last column =
 * count), fully overwritten.  Student rows: every q < num_student_rows whose label is flagged in
 * d_prev_mask (uint8[num_classes]) (:530-539); teacher rows: hs_T[keepid[i]] into class tlabel[i] (:548-551).
 * Sums run in ascending index order per class, like the reference's Python loops. */
int dskd_bcdd_prototypes(const float* d_hs_student, const int64_t* d_student_labels, int32_t num_student_rows,
                         const float* d_hs_teacher, const int64_t* d_teacher_keepid,
                         const int64_t* d_teacher_labels, int32_t num_teacher,
                         const uint8_t* d_prev_mask, int32_t num_classes, int32_t C,
                         float* d_proto, void* stream);

/* correlation_mat (:1197-1222) + MSELoss, fused with its backward.  L = number of previous classes
 * (rows 0..L-1).  reduction: 0 none (loss not reduced; d_dist only), 1 mean, 2 sum.
 *   d_dist:  [2,L,L] (teacher, student) distance matrices (direct difference, fp32)
 *   d_loss:  [1]  = loss_weight * reduce((D_T - D_S)^2) / L
 *   d_grad_proto_student: [num_classes, C+1] dLoss/d(student sums) (count column 0); NULL = forward only.
 * grad_scale multiplies the gradient: pass world_size when d_proto was all-reduced (sum) over ranks
 * (every rank holds the same loss, so the all-reduce's adjoint is a local x world_size which DDP's
 * gradient mean then cancels -- SURVEY.md section 8e); 1.0f otherwise. */
int dskd_bcdd_distance_loss(const float* d_proto, int32_t num_classes, int32_t C, int32_t L,
                            int32_t reduction, float loss_weight, float grad_scale, float* d_dist,
                            float* d_loss, float* d_grad_proto_student, void* stream);

/* dskd_bcdd_distance_loss followed by dskd_bcdd_scatter_grad in one call (what the module uses after the
 * optional all-reduce of d_proto).  d_grad_proto_student is a [num_classes, C+1] workspace. */
int dskd_bcdd_loss_and_grad(const float* d_proto, int32_t num_classes, int32_t C, int32_t L, int32_t reduction,
                            float loss_weight, float grad_scale, const int64_t* d_student_labels,
                            int32_t num_student_rows, const uint8_t* d_prev_mask, float* d_dist, float* d_loss,
                            float* d_grad_proto_student, float* d_grad_hs_student, void* stream);

/* d_grad_hs_student[q,:] = d_grad_proto_student[label[q], :C] for flagged q, else 0 (fully overwritten). */
int dskd_bcdd_scatter_grad(const float* d_grad_proto_student, const int64_t* d_student_labels,
                           int32_t num_student_rows, const uint8_t* d_prev_mask, int32_t num_classes,
                           int32_t C, float* d_grad_hs_student, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Assignment (gfl_hungarian_assigner.py:102-160; match_cost.py:34-51,215-230,460-476;
 * iou2d_calculator.py:213-260; head_il.py:54-59,1427-1432)
 * ------------------------------------------------------------------------------------------- */
/* All (layer,image) problems in one launch.  d_cls: [P,Q,num_classes] logits, d_box: [P,Q,2+4*(reg_max+1)]
 * sigmoid outputs (reg_max == 0: d_box is [P,Q,4] already-decoded normalised cxcywh, the form
 * GFLHungarianAssigner.assign receives), P = num_problems; problem p uses the GT set of image (p % N).  d_gt_boxes fp32 [G_total,4]
 * px xyxy, d_gt_labels int64 [G_total], d_gt_start int32 [N+1], d_img_hw int32 [N,2].
 * Output d_cost: fp32, problem p at d_cost + p*Q*max_gt, row-major [Q, G_img] with row stride max_gt. */
int dskd_cost_matrix(const float* d_cls, const float* d_box, int32_t num_problems, int32_t N, int32_t Q,
                     int32_t num_classes, int32_t reg_max, const float* d_gt_boxes,
                     const int64_t* d_gt_labels, const int32_t* d_gt_start, const int32_t* d_img_hw,
                     int32_t max_gt, float w_cls, float w_reg, float w_iou, float* d_cost, void* stream);

/* Targets from the matching (head_il.py:1765-1797, pseudo_sampler.py:35-41, sampling_result.py:35):
 * d_assigned_gt int64 [P,Q] (1-based GT index within the image's GT set, 0 = background) ->
 * d_labels int64 [P,Q] (bg = num_classes), d_bbox_targets fp32 [P,Q,4] (normalised cxcywh of the matched
 * GT, 0 elsewhere), d_bbox_weights fp32 [P,Q,4] (1 on positives), d_teacher_only fp32 [P,Q] (1 where the
 * label is flagged in d_prev_mask, :1453-1455).  Any output may be NULL. */
int dskd_assign_targets(const int64_t* d_assigned_gt, int32_t num_problems, int32_t N, int32_t Q,
                        int32_t num_classes, const float* d_gt_boxes, const int64_t* d_gt_labels,
                        const int32_t* d_gt_start, const int32_t* d_img_hw, const uint8_t* d_prev_mask,
                        int64_t* d_labels, float* d_bbox_targets, float* d_bbox_weights, float* d_teacher_only,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Teacher keep-ids: gfl_deformable_detr_head_il.py:1622-1668 (`_get_bboxes_single`,
 * sigmoid branch, need_logits=True) + core/utils/misc.py:143-152 (`filter_scores_and_topk`) +
 * detectors/deformable_detr_il.py:138-151 (keep-id flattening), for all images in one launch.
 * ------------------------------------------------------------------------------------------- */
/* d_cls: [N,Q,num_classes] logits and d_box: [N,Q,2+4*(reg_max+1)] sigmoid outputs of the teacher's LAST
 * decoder layer (reg_max == 0: d_box is [N,Q,4] normalised cxcywh); d_img_hw int32 [N,2].  Per image the
 * (query, class) pairs with sigmoid(logit) > score_thr are ordered by descending score (ties: lower
 * q*num_classes + class first) and the first max_per_img (<= 1024) kept:
 *   d_count  int32 [N]                  K_i
 *   d_bboxes fp32  [N,max_per_img,4]    px xyxy, clamped to the image (:1658-1663); 16-byte aligned
 *   d_scores fp32  [N,max_per_img]      sigmoid score
 *   d_labels int64 [N,max_per_img]      class
 *   d_keepid int64 [N,max_per_img]      q + Q*i  (deformable_detr_il.py:151)
 *   d_logits fp32  [N,max_per_img,num_classes]  sigmoid(cls)[q,:] (`det_logits`, :1636); rows >= K_i untouched
 * Slots r >= K_i hold 0 / -1.  Every output except d_count may be NULL. */
int dskd_teacher_decode(const float* d_cls, const float* d_box, int32_t N, int32_t Q, int32_t num_classes,
                        int32_t reg_max, const int32_t* d_img_hw, float score_thr, int32_t max_per_img,
                        int32_t* d_count, float* d_bboxes, float* d_scores, int64_t* d_labels, int64_t* d_keepid,
                        float* d_logits, void* stream);

/* Ragged -> concatenated form the losses take: d_start int32 [N+1] prefix sums of d_count, and the first K_i
 * rows of every image copied behind each other into d_cat_* (capacity N*max_per_img rows; NULL = skip). */
int dskd_teacher_compact(const int32_t* d_count, int32_t N, int32_t max_per_img, const float* d_bboxes,
                         const float* d_scores, const int64_t* d_labels, const int64_t* d_keepid, int32_t* d_start,
                         float* d_cat_bboxes, float* d_cat_scores, int64_t* d_cat_labels, int64_t* d_cat_keepid,
                         void* stream);

/* scipy.optimize.linear_sum_assignment (rectangular_lsap, modified Jonker-Volgenant), float64, minimise.
 * h_cost row-major [rows, cols]; writes k = min(rows, cols) pairs sorted by row.  Bit-exact with SciPy. */
int dskd_lsap_f64(const double* h_cost, int32_t rows, int32_t cols, int64_t* h_row_ind, int64_t* h_col_ind);

/* Batch over fp32 cost matrices as produced by dskd_cost_matrix (host copy): problem p is
 * [rows, h_cols[p]] with row stride `ld` at h_cost + p*rows*ld; writes h_assigned_gt[p*rows + r] =
 * 1-based matched column or 0 (gfl_hungarian_assigner.py:153-158).  num_threads <= 0: hardware conc. */
int dskd_lsap_batch_f32(const float* h_cost, int32_t num_problems, int32_t rows, int32_t ld,
                        const int32_t* h_cols, int64_t* h_assigned_gt, int32_t num_threads);

/* The same solver on the DEVICE for a batch laid out like dskd_cost_matrix's output (problem p: [rows, cols_p]
 * with row stride ld, cols_p = d_gt_start[p % N + 1] - d_gt_start[p % N] <= max_cols): one CTA per problem, float64
 * arithmetic in SciPy's operation order and tie-breaking, so the indices are identical to dskd_lsap_batch_f32 / SciPy
 * -- without the device->host copy and sync of gfl_hungarian_assigner.py:143-151.  Asynchronous on `stream`.
 * d_assigned_gt int64 [P, rows] (1-based matched column or 0); d_status int32 [P] (DSKD_OK / DSKD_EINFEASIBLE for a
 * matrix with NaN / -inf or no feasible matching; may be NULL).  max(rows, max_cols) <= 1024. */
int dskd_lsap_batch_device(const float* d_cost, int32_t num_problems, int32_t N, int32_t rows, int32_t ld,
                           const int32_t* d_gt_start, int32_t max_cols, int64_t* d_assigned_gt, int32_t* d_status,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-scale deformable attention: the sampling core of the detector that hosts the losses (next-row 1 harness).
 * Replaces mmcv's CUDA op `MultiScaleDeformableAttention` that the reference's transformer imports
 * (mmdet/models/utils/transformer.py:23; mmcv-full pinned by requirements/mminstall.txt:1, not vendored):
 *   out[n,q,m,:] = sum_{l,p} attn[n,q,m,l,p] * bilinear(value_l[n,:,m,:], loc[n,q,m,l,p,:])
 * with grid_sample(align_corners=False, zeros) sampling, loc = (x, y) in [0,1] per level.
 * value [N,S,M,D], loc [N,Lq,M,L,P,2], attn [N,Lq,M,L,P], out / grad_out [N,Lq,M,D]; fp32, contiguous; `levels` is a
 * HOST array (H, W, cell_offset = first token of the level); L*P <= 16.  The backward zero-fills d_grad_value itself.
 * ------------------------------------------------------------------------------------------- */
int dskd_msda_forward(const float* d_value, const DskdLevel* levels, int32_t num_levels, const float* d_loc,
                      const float* d_attn, int32_t N, int64_t S, int32_t M, int32_t D, int64_t Lq, int32_t P,
                      float* d_out, void* stream);
int dskd_msda_backward(const float* d_value, const DskdLevel* levels, int32_t num_levels, const float* d_loc,
                       const float* d_attn, const float* d_grad_out, int32_t N, int64_t S, int32_t M, int32_t D,
                       int64_t Lq, int32_t P, float* d_grad_value, float* d_grad_loc, float* d_grad_attn,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Registry loss modules (mse_loss.py, kd_loss.py, utils.py) -- generic tensors
 * ------------------------------------------------------------------------------------------- */
/* elementwise (pred-target)^2 * weight (weight may be NULL); d_elem (n) optional; d_sum[1] (double,
 * caller zero-fills) optional; grads (n each, optional) = +-2*(pred-target)*weight*grad_scale. */
int dskd_mse_elementwise(const float* d_pred, const float* d_target, const float* d_weight, int64_t n,
                         float grad_scale, float* d_elem, double* d_sum, float* d_grad_pred,
                         float* d_grad_target, void* stream);

/* Same contract for the other elementwise criteria the head builds by config: kind 0 = MSE, 1 = smooth L1 with
 * threshold beta (`loss_ld_bbox=dict(type='SmoothL1Loss')`, chaosuan_..._40_...py:118; smooth_l1_loss.py:12-35),
 * 2 = L1 (smooth_l1_loss.py:38-56). */
int dskd_elementwise_loss(int32_t kind, float beta, const float* d_pred, const float* d_target, const float* d_weight,
                          int64_t n, float grad_scale, float* d_elem, double* d_sum, float* d_grad_pred,
                          float* d_grad_target, void* stream);

/* knowledge_distillation_kl_div_loss on [outer, D, inner] with softmax over D (dim=1 of the
 * reference's [N,D] or [C,H,W] inputs): d_rowloss [outer*inner] = T^2/D * sum_d t(log t - logp);
 * optional d_row_weight[outer*inner] scales loss & grad; d_grad_pred (same shape as pred, optional)
 * = grad_scale * w * (T/D)(softmax(pred/T) - t). */
int dskd_kd_kl_rows(const float* d_pred, const float* d_soft, int64_t outer, int32_t D, int64_t inner,
                    float temperature, const float* d_row_weight, float grad_scale, float* d_rowloss,
                    double* d_sum, float* d_grad_pred, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CUDA-graph launch policy (no reference counterpart: the reference runs eagerly).  `graph` is a cudaGraph_t holding
 * a captured loss step (e.g. torch.cuda.CUDAGraph(keep_graph=True).raw_cuda_graph()).  Every kernel node with fewer
 * than big_grid_ctas CTAs gets the highest launch priority, the others the lowest, and the graph is instantiated with
 * cudaGraphInstantiateFlagUseNodePriority so that the latency-bound chains (mask build, BCDD, the prototype
 * all-reduce) are dispatched ahead of the remaining CTAs of the HBM-bound streaming kernel instead of behind them.
 * *exec_out receives a cudaGraphExec_t owned by the caller (dskd_graph_exec_destroy).
 * ------------------------------------------------------------------------------------------- */
int dskd_graph_instantiate_prioritized(void* graph, int32_t big_grid_ctas, void** exec_out, int32_t* num_small,
                                       int32_t* num_big);
int dskd_graph_launch(void* exec, void* stream);
int dskd_graph_exec_destroy(void* exec);

/* ---------------------------------------------------------------------------------------------
 * The exchange step of SURVEY.md section 8e over NVLink / NVSwitch peer memory (no reference counterpart: the reference's
 * prototypes are rank-local, gfl_deformable_detr_head_il.py:531-551; north_star asks for global ones).  One process per
 * GPU: every rank allocates a "symmetric" buffer of dskd_peer_buffer_floats(table_floats) floats, ZERO-FILLED, exports
 * it with dskd_ipc_export (CUDA IPC handle of the allocation that holds the pointer + the pointer's byte offset in it),
 * the handles travel through the caller's own channel (torch.distributed), and every rank maps the others with
 * dskd_ipc_open (+ offset).  dskd_peer_allreduce then replaces `d_table` ([table_floats] fp32, a multiple of 4, 16-byte
 * aligned) by its sum over ranks in ONE kernel launch: publish into the own buffer, raise a flag in every peer's buffer,
 * wait for all flags, add the peers' tables in rank order (the result is bit-identical on every rank).  `bufs` is a HOST
 * array of `world` device pointers (entry `rank` = the own buffer).  Collective: every rank must call it the same number
 * of times; asynchronous on `stream`, capturable in a CUDA graph; a peer that does not arrive within about twenty seconds
 * makes the kernel trap.
 * ------------------------------------------------------------------------------------------- */
#define DSKD_PEER_MAX_WORLD 16
#define DSKD_PEER_CTRL_FLOATS 64 /* control words in front of the two table slots of a symmetric buffer */
#define DSKD_IPC_HANDLE_BYTES 64
int dskd_ipc_export(const void* d_ptr, void* handle_out /* DSKD_IPC_HANDLE_BYTES */, int64_t* offset_out);
int dskd_ipc_open(const void* handle, void** base_out);
int dskd_ipc_close(void* base);
int64_t dskd_peer_buffer_floats(int64_t table_floats);
int dskd_peer_allreduce(float* d_table, int64_t table_floats, void* const* bufs, int32_t world, int32_t rank,
                        void* stream);

/* x[i] *= *d_factor for i<n, skipped entirely when *d_factor == 1.0f (read on the device: no host
 * sync).  Used by autograd backward to apply grad_output to gradients staged by the fused kernels. */
int dskd_scale_inplace(float* d_x, int64_t n, const float* d_factor, void* stream);

/* sum of doubles -> float (deterministic order), helper for partial-sum buffers. */
int dskd_f64_to_f32(const double* d_in, float* d_out, int32_t n, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSKD_B200_H_ */
