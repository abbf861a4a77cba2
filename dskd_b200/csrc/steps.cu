// Composite entry points: the whole DSG-FD (or BCDD tail) of one batch as a fixed launch sequence, so
// the Python module pays one FFI call per loss instead of one per kernel.
#include "common.cuh"

using namespace dskd;

namespace dskd {
int launch_prepare_v1(const DskdDsgfdStepArgs* a, int* owner, float* rows, float* energy, int64_t* ids, double* acc,
                      unsigned* counter, cudaStream_t st);
int launch_rows_finish_final(const DskdDsgfdStepArgs* a, const float* rows, const float* energy, const int64_t* ids,
                             double* acc, unsigned* counter, float* grad_hs, cudaStream_t st);
int launch_v1_loss_final(const DskdDsgfdStepArgs* a, const double* acc, const unsigned* counter, cudaStream_t st);
}  // namespace dskd

namespace {
// Timing events survive CUDA-graph capture as external event-record nodes.
cudaError_t record_event(void* ev, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaError_t e = cudaStreamIsCapturing(st, &cs);
  if (e != cudaSuccess) return e;
  return cudaEventRecordWithFlags(static_cast<cudaEvent_t>(ev), st,
                                  cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}
constexpr int64_t kAlign = 256;
inline int64_t align_up(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Workspace {
  int64_t cell_map, rows, energy, ids, acc, kl_redo, kl_redo_bytes, total;
  Workspace(int32_t N, int64_t cells, int32_t P, int32_t C, int64_t kl_bytes = 0) {
    int64_t off = 0;
    cell_map = off; off += align_up((int64_t)N * cells * 4);
    rows = off;     off += align_up((int64_t)std::max(P, 1) * C * 4);
    energy = off;   off += align_up((int64_t)std::max(P, 1) * C * 4);
    acc = off;      off += align_up(16);      // right after energy: one memset clears both; [8] = finish counter, [12] = matched student queries
    ids = off;      off += align_up((int64_t)std::max(P, 1) * 8);
    kl_redo = off;  kl_redo_bytes = kl_bytes; off += align_up(kl_bytes);  // redo list of the KL kernels
    total = off;
  }
};
}  // namespace

// Without the level table the KL kernels' workspace (one word per column tile, image and channel) is bounded through
// sum_l ceil(W_l / 32) <= cells / 32 + levels; dskd_dsgfd_step_workspace_bytes_for() gives the exact figure.
extern "C" int64_t dskd_dsgfd_step_workspace_bytes(int32_t N, int64_t cells_per_image, int32_t num_pairs, int32_t C) {
  if (N < 0 || cells_per_image < 0 || num_pairs < 0 || C <= 0) return -1;
  const int64_t kl = std::max<int64_t>(16, 4 * (cells_per_image / 32 + DSKD_MAX_LEVELS) * N * (int64_t)C);
  return Workspace(N, cells_per_image, num_pairs, C, kl).total;
}

extern "C" int64_t dskd_dsgfd_step_workspace_bytes_for(const DskdDsgfdStepArgs* a) {
  if (a == nullptr || a->N < 0 || a->cells_per_image < 0 || a->num_pairs < 0 || a->C <= 0) return -1;
  int64_t kl = 0;
  if (a->criterion == DSKD_CRIT_KL) {
    kl = dskd_dsgfd_kl_workspace_bytes(a->N, a->num_levels, a->levels, a->C);
    if (kl < 0) return -1;
  }
  return Workspace(a->N, a->cells_per_image, a->num_pairs, a->C, kl).total;
}

extern "C" int dskd_dsgfd_step(const DskdDsgfdStepArgs* a, void* stream) {
  DSKD_REQUIRE(a != nullptr, "dskd_dsgfd_step: null args");
  DSKD_REQUIRE(a->criterion == DSKD_CRIT_MSE || a->criterion == DSKD_CRIT_KL, "dskd_dsgfd_step: bad criterion %d", a->criterion);
  DSKD_REQUIRE(a->mask_mode >= DSKD_MODE_DECODE_V1 && a->mask_mode <= DSKD_MODE_FG_BK, "dskd_dsgfd_step: bad mask_mode %d", a->mask_mode);
  DSKD_REQUIRE(a->num_levels > 0 && a->num_levels <= DSKD_MAX_LEVELS && a->N >= 0 && a->C > 0 && a->num_pairs >= 0,
               "dskd_dsgfd_step: bad sizes");
  DSKD_REQUIRE(a->d_loss != nullptr, "dskd_dsgfd_step: d_loss is null");
  DSKD_REQUIRE(a->criterion == DSKD_CRIT_MSE || a->layout == DSKD_LAYOUT_NCHW || a->mask_mode >= DSKD_MODE_SG_OUT,
               "dskd_dsgfd_step: KL on [S,N,C] memory takes the per-cell masks (sg_out / fg_only) only");
  const int64_t kl_ws = a->criterion == DSKD_CRIT_KL ? dskd_dsgfd_kl_workspace_bytes(a->N, a->num_levels, a->levels, a->C) : 0;
  const Workspace ws(a->N, a->cells_per_image, a->num_pairs, a->C, kl_ws);
  DSKD_REQUIRE(a->d_workspace != nullptr && a->workspace_bytes >= ws.total &&
                   (reinterpret_cast<uintptr_t>(a->d_workspace) % kAlign) == 0,
               "dskd_dsgfd_step: workspace must be %lld bytes, 256-byte aligned", (long long)ws.total);
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(a->d_workspace);
  void* cell_map = base + ws.cell_map;
  float* rows = reinterpret_cast<float*>(base + ws.rows);
  float* energy = reinterpret_cast<float*>(base + ws.energy);
  double* acc = reinterpret_cast<double*>(base + ws.acc);
  int64_t* ids = reinterpret_cast<int64_t*>(base + ws.ids);
  const bool row_mode = a->mask_mode <= DSKD_MODE_DECODE_V2;
  const bool v1 = a->mask_mode == DSKD_MODE_DECODE_V1;
  const int P = a->num_pairs;
  if (a->N == 0) {  // empty batch: zero loss, zero gradients, nothing matched
    DSKD_CUDA_OK(cudaMemsetAsync(a->d_loss, 0, sizeof(float), st));
    if (a->d_matched_count) DSKD_CUDA_OK(cudaMemsetAsync(a->d_matched_count, 0, sizeof(int32_t), st));
    if (a->d_grad_hs_student)
      DSKD_CUDA_OK(cudaMemsetAsync(a->d_grad_hs_student, 0, sizeof(float) * (size_t)a->num_query_rows * a->C, st));
    return DSKD_OK;
  }
  const bool want_hs = v1 && a->d_grad_hs_student != nullptr;
  unsigned* counter = reinterpret_cast<unsigned*>(base + ws.acc + 8);
  // decode_v1 with at least one pair: matched ids, mask rows, owner raster and every clear in ONE launch
  const bool fused_v1 = v1 && P > 0;
  int rc;
  if (fused_v1) {
    DSKD_REQUIRE(a->d_hs_teacher && a->d_teacher_keepid && a->d_hs_student && a->d_student_labels && a->d_prev_mask &&
                     a->d_boxes && a->d_box_start && a->d_img_hw,
                 "dskd_dsgfd_step: decode_v1 needs embeddings, keep-ids, labels and boxes");
    rc = launch_prepare_v1(a, static_cast<int*>(cell_map), rows, energy, ids, acc, counter, st);
    if (rc) return rc;
  } else {
    // energy (or KL grad_rows) + loss accumulator + counter: one clear
    DSKD_CUDA_OK(cudaMemsetAsync(energy, 0, (size_t)(ws.acc + 16 - ws.energy), st));
    if (a->d_grad_hs_student != nullptr)
      DSKD_CUDA_OK(cudaMemsetAsync(a->d_grad_hs_student, 0, sizeof(float) * (size_t)a->num_query_rows * a->C, st));
  }
  if (fused_v1) {
    // nothing else to prepare
  } else if (row_mode) {
    DSKD_REQUIRE(a->d_hs_teacher && (P == 0 || a->d_teacher_keepid), "dskd_dsgfd_step: decode_* needs teacher embeddings / keep-ids");
    if (v1) {
      DSKD_REQUIRE(a->d_hs_student && a->d_student_labels && a->d_prev_mask, "dskd_dsgfd_step: decode_v1 needs the student side");
      rc = dskd_select_prev_queries(a->d_student_labels, a->num_query_rows, a->d_prev_mask, a->num_classes, P, ids,
                                    a->d_matched_count, stream);
      if (rc) return rc;
    }
    rc = dskd_mask_rows(v1 ? DSKD_MASK_DECODE_V1 : DSKD_MASK_DECODE_V2, a->d_hs_teacher, a->d_hs_student,
                        a->d_teacher_keepid, ids, P, a->C, rows, stream);
    if (rc) return rc;
    rc = dskd_raster_cells(DSKD_RASTER_OWNER_EXCL, a->d_boxes, a->d_box_start, nullptr, nullptr, a->d_img_hw, a->N,
                           a->max_boxes_per_image, a->levels, a->num_levels, a->cells_per_image, cell_map, stream);
    if (rc) return rc;
  } else {
    const int mode = a->mask_mode == DSKD_MODE_SG_OUT ? DSKD_RASTER_BINARY_INCL
                     : (a->mask_mode == DSKD_MODE_FG_ONLY ? DSKD_RASTER_AREA_INCL : DSKD_RASTER_AREA_FGBK);
    rc = dskd_raster_cells(mode, a->d_boxes, a->d_box_start, a->d_gt_boxes, a->d_gt_start, a->d_img_hw, a->N,
                           a->max_boxes_per_image, a->levels, a->num_levels, a->cells_per_image, cell_map, stream);
    if (rc) return rc;
  }
  if (a->ev_kernel_begin) DSKD_CUDA_OK(record_event(a->ev_kernel_begin, st));
  if (a->criterion == DSKD_CRIT_MSE) {
    DskdDsgfdMseArgs m;
    memset(&m, 0, sizeof(m));
    m.layout = a->layout; m.num_levels = a->num_levels; m.N = a->N; m.C = a->C;
    m.cells_per_image = a->cells_per_image;
    for (int l = 0; l < a->num_levels; ++l) {
      m.levels[l] = a->levels[l];
      m.d_student[l] = a->d_student[l]; m.d_teacher[l] = a->d_teacher[l]; m.d_grad_student[l] = a->d_grad_student[l];
      m.scale[l] = a->scale[l];
    }
    if (row_mode) {
      m.d_owner = static_cast<const int32_t*>(cell_map); m.d_rows = rows; m.d_energy = energy; m.num_pairs = P;
    } else {
      m.d_cell_weight = static_cast<const float*>(cell_map); m.d_loss = acc;
    }
    rc = dskd_dsgfd_mse_fwd_bwd(&m, stream);
    if (rc) return rc;
  } else {
    DskdDsgfdKlArgs k;
    memset(&k, 0, sizeof(k));
    k.num_levels = a->num_levels; k.N = a->N; k.C = a->C; k.temperature = a->temperature;
    k.cells_per_image = a->cells_per_image;
    for (int l = 0; l < a->num_levels; ++l) {
      k.levels[l] = a->levels[l];
      k.d_student[l] = a->d_student[l]; k.d_teacher[l] = a->d_teacher[l];
      k.scale[l] = a->scale[l];
    }
    k.d_loss = acc;
    k.layout = a->layout;
    k.d_workspace = base + ws.kl_redo;
    k.workspace_bytes = ws.kl_redo_bytes;
    if (row_mode) {
      k.d_owner = static_cast<const int32_t*>(cell_map); k.d_rows = rows; k.num_pairs = P;
      k.d_grad_rows = want_hs ? energy : nullptr;
    } else {
      k.d_cell_weight = static_cast<const float*>(cell_map);
    }
    rc = dskd_dsgfd_kl_fwd_bwd(&k, stream);
    if (rc) return rc;
  }
  if (a->ev_kernel_end) DSKD_CUDA_OK(record_event(a->ev_kernel_end, st));
  if (fused_v1 && a->criterion == DSKD_CRIT_MSE)  // row finish whose last CTA also writes the loss
    return launch_rows_finish_final(a, rows, energy, ids, acc, counter, want_hs ? a->d_grad_hs_student : nullptr, st);
  if (row_mode) {
    rc = dskd_dsgfd_rows_finish(a->criterion, a->d_hs_teacher, a->d_hs_student, a->d_teacher_keepid, ids, rows, energy, P,
                                a->C, acc, want_hs ? a->d_grad_hs_student : nullptr, stream);
    if (rc) return rc;
  }
  if (fused_v1) return launch_v1_loss_final(a, acc, counter, st);  // NaN when teacher detections went unpaired
  return dskd_f64_to_f32(acc, a->d_loss, 1, 1.0f, stream);
}
