// Hungarian cost matrices for every (decoder layer, image) problem in one launch.
// Reference: gfl_hungarian_assigner.py:120-140; match_cost.py:34-51 (BBoxL1Cost, box_format='xywh'),
// :215-230 (QualityFocalLossCost), :460-476 (IoUCost 'giou'); iou2d_calculator.py:213-260 (eps = 1e-6
// clamps on union and enclosing area); head_il.py:54-59,1427-1432 (Integral_average box decode);
// transforms.py:245-270.  Same fp32 operation order as the reference (IEEE division, no fast-math).
#include "common.cuh"

namespace dskd {

constexpr int kQTile = 32;  // queries per CTA

struct Box4 { float x1, y1, x2, y2; };

__device__ __forceinline__ float iou_or_giou(const Box4& a, const Box4& b, bool giou) {
  const float eps = 1e-6f;
  const float area1 = (a.x2 - a.x1) * (a.y2 - a.y1);
  const float area2 = (b.x2 - b.x1) * (b.y2 - b.y1);
  const float w = fmaxf(fminf(a.x2, b.x2) - fmaxf(a.x1, b.x1), 0.f);
  const float h = fmaxf(fminf(a.y2, b.y2) - fmaxf(a.y1, b.y1), 0.f);
  const float overlap = w * h;
  const float uni = fmaxf(area1 + area2 - overlap, eps);
  const float iou = __fdiv_rn(overlap, uni);
  if (!giou) return iou;
  const float ew = fmaxf(fmaxf(a.x2, b.x2) - fminf(a.x1, b.x1), 0.f);
  const float eh = fmaxf(fmaxf(a.y2, b.y2) - fminf(a.y1, b.y1), 0.f);
  const float earea = fmaxf(ew * eh, eps);
  return iou - __fdiv_rn(earea - uni, earea);
}

// grid (ceil(Q / kQTile), num_problems); dynamic smem: max_gt * 13 floats + max_gt ints.
__global__ void __launch_bounds__(128) cost_matrix_kernel(const float* __restrict__ cls, const float* __restrict__ box,
                                                          int N, int Q, int num_classes, int reg_max,
                                                          const float* __restrict__ gt_boxes,
                                                          const int64_t* __restrict__ gt_labels,
                                                          const int* __restrict__ gt_start,
                                                          const int* __restrict__ img_hw, int max_gt, float w_cls,
                                                          float w_reg, float w_iou, float* __restrict__ cost) {
  extern __shared__ float sm[];
  __shared__ float expect[kQTile][4];
  __shared__ float pred[kQTile][4];  // cx, cy, w, h (normalised)
  const int p = blockIdx.y, img = p % N;
  const int q0 = blockIdx.x * kQTile;
  const int g0 = gt_start[img], G = gt_start[img + 1] - g0;
  if (G == 0) return;
  const float img_h = (float)img_hw[2 * img], img_w = (float)img_hw[2 * img + 1];
  float* gt_px = sm;                   // [G][4]
  float* gt_nx = sm + 4 * max_gt;      // [G][4] normalised xyxy
  float* gt_cw = sm + 8 * max_gt;      // [G][4] normalised cxcywh
  int* gt_lab = reinterpret_cast<int*>(sm + 12 * max_gt);
  const int bins = reg_max + 1, box_ch = (reg_max > 0) ? 2 + 4 * bins : 4;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float* b = gt_boxes + (int64_t)(g0 + g) * 4;
    const float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
    gt_px[4 * g + 0] = x1; gt_px[4 * g + 1] = y1; gt_px[4 * g + 2] = x2; gt_px[4 * g + 3] = y2;
    const float nx1 = __fdiv_rn(x1, img_w), ny1 = __fdiv_rn(y1, img_h), nx2 = __fdiv_rn(x2, img_w), ny2 = __fdiv_rn(y2, img_h);
    gt_nx[4 * g + 0] = nx1; gt_nx[4 * g + 1] = ny1; gt_nx[4 * g + 2] = nx2; gt_nx[4 * g + 3] = ny2;
    gt_cw[4 * g + 0] = __fdiv_rn(nx1 + nx2, 2.f);
    gt_cw[4 * g + 1] = __fdiv_rn(ny1 + ny2, 2.f);
    gt_cw[4 * g + 2] = nx2 - nx1;
    gt_cw[4 * g + 3] = ny2 - ny1;
    gt_lab[g] = (int)gt_labels[g0 + g];
  }
  // Integral_average: each (reg_max+1)-bin group normalised by its sum, expectation over b/reg_max/2
  {
    const int ql = threadIdx.x >> 2, grp = threadIdx.x & 3;
    const int q = q0 + ql;
    if (q < Q && reg_max > 0) {
      const float* x = box + ((int64_t)p * Q + q) * box_ch + 2 + grp * bins;
      float s = 0.f;
      for (int b = 0; b < bins; ++b) s += x[b];
      float e = 0.f;
      for (int b = 0; b < bins; ++b) {
        const float space = __fdiv_rn(__fdiv_rn((float)b, (float)reg_max), 2.f);
        e += __fmul_rn(__fdiv_rn(x[b], s), space);
      }
      expect[ql][grp] = e;
    }
  }
  __syncthreads();
  if (threadIdx.x < kQTile && q0 + threadIdx.x < Q) {
    const int ql = threadIdx.x;
    const float* x = box + ((int64_t)p * Q + q0 + ql) * box_ch;
    pred[ql][0] = x[0];
    pred[ql][1] = x[1];
    pred[ql][2] = (reg_max > 0) ? expect[ql][0] + expect[ql][1] : x[2];
    pred[ql][3] = (reg_max > 0) ? expect[ql][2] + expect[ql][3] : x[3];
  }
  __syncthreads();
  float* out = cost + (int64_t)p * Q * max_gt;
  for (int idx = threadIdx.x; idx < kQTile * G; idx += blockDim.x) {
    const int ql = idx / G, g = idx - ql * G;
    const int q = q0 + ql;
    if (q >= Q) break;
    const float cx = pred[ql][0], cy = pred[ql][1], w = pred[ql][2], h = pred[ql][3];
    // BBoxL1Cost('xywh'): cdist(pred_cxcywh, gt_cxcywh, p=1) * w_reg
    const float l1 = fabsf(cx - gt_cw[4 * g]) + fabsf(cy - gt_cw[4 * g + 1]) + fabsf(w - gt_cw[4 * g + 2]) +
                     fabsf(h - gt_cw[4 * g + 3]);
    const float reg = l1 * w_reg;
    Box4 pn{cx - 0.5f * w, cy - 0.5f * h, cx + 0.5f * w, cy + 0.5f * h};
    Box4 pp{pn.x1 * img_w, pn.y1 * img_h, pn.x2 * img_w, pn.y2 * img_h};
    Box4 gp{gt_px[4 * g], gt_px[4 * g + 1], gt_px[4 * g + 2], gt_px[4 * g + 3]};
    Box4 gn{gt_nx[4 * g], gt_nx[4 * g + 1], gt_nx[4 * g + 2], gt_nx[4 * g + 3]};
    const float iou_c = -iou_or_giou(pp, gp, true) * w_iou;
    // QualityFocalLossCost: BCE-with-logits(logit, IoU) * |IoU - sigmoid(logit)|^2 * w_cls
    const float score = iou_or_giou(pn, gn, false);
    const float x = cls[((int64_t)p * Q + q) * num_classes + gt_lab[g]];
    const float sig = __fdiv_rn(1.f, 1.f + expf(-x));
    const float log_sig = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
    const float bce = (1.f - score) * x - log_sig;
    const float sf = fabsf(score - sig);
    const float clsc = bce * (sf * sf) * w_cls;
    out[(int64_t)q * max_gt + g] = clsc + reg + iou_c;
  }
}

// One thread per (problem, query).
__global__ void __launch_bounds__(256) assign_targets_kernel(const int64_t* __restrict__ assigned, int P, int N, int Q,
                                                             int num_classes, const float* __restrict__ gt_boxes,
                                                             const int64_t* __restrict__ gt_labels,
                                                             const int* __restrict__ gt_start,
                                                             const int* __restrict__ img_hw,
                                                             const uint8_t* __restrict__ prev_mask,
                                                             int64_t* __restrict__ labels, float* __restrict__ bt,
                                                             float* __restrict__ bw, float* __restrict__ teacher_only) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)P * Q) return;
  const int p = (int)(t / Q), img = p % N;
  const int64_t a = assigned[t];
  int64_t lab = num_classes;
  float4 tgt = make_float4(0.f, 0.f, 0.f, 0.f);
  const float wgt = a > 0 ? 1.f : 0.f;
  if (a > 0) {
    const int g = gt_start[img] + (int)a - 1;
    lab = gt_labels[g];
    const float img_h = (float)img_hw[2 * img], img_w = (float)img_hw[2 * img + 1];
    const float x1 = __fdiv_rn(gt_boxes[4 * g], img_w), y1 = __fdiv_rn(gt_boxes[4 * g + 1], img_h);
    const float x2 = __fdiv_rn(gt_boxes[4 * g + 2], img_w), y2 = __fdiv_rn(gt_boxes[4 * g + 3], img_h);
    tgt = make_float4(__fdiv_rn(x1 + x2, 2.f), __fdiv_rn(y1 + y2, 2.f), x2 - x1, y2 - y1);
  }
  if (labels) labels[t] = lab;
  if (bt) reinterpret_cast<float4*>(bt)[t] = tgt;
  if (bw) reinterpret_cast<float4*>(bw)[t] = make_float4(wgt, wgt, wgt, wgt);
  if (teacher_only) teacher_only[t] = (prev_mask && lab >= 0 && lab < num_classes && prev_mask[lab]) ? 1.f : 0.f;
}

// Single CTA ordered compaction (n is a few thousand): ids of the queries whose label is a previous-task
// label.  Each of the 32 warps compacts a contiguous 256-label segment with ballots, a shuffle scan orders the
// segments, then every warp copies its hits to their final slots: 2 block barriers per 8192 labels.
constexpr int kSelSeg = 256;
constexpr int kSelChunk = 32 * kSelSeg;

__global__ void __launch_bounds__(1024) select_prev_kernel(const int64_t* __restrict__ labels, int n,
                                                           const uint8_t* __restrict__ prev_mask, int num_classes,
                                                           int max_out, int64_t* __restrict__ ids, int* __restrict__ count) {
  __shared__ int list[kSelChunk];
  __shared__ int warp_off[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int running = 0;
  for (int base = 0; base < n; base += kSelChunk) {
    int cnt = 0;
    const int seg0 = base + warp * kSelSeg;
    for (int it = 0; it < kSelSeg; it += 32) {
      const int q = seg0 + it + lane;
      bool hit = false;
      if (q < n) {
        const int64_t lab = labels[q];
        hit = lab >= 0 && lab < num_classes && prev_mask[lab] != 0;
      }
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (hit) list[warp * kSelSeg + cnt + __popc(b & ((1u << lane) - 1u))] = q;
      cnt += __popc(b);
    }
    if (lane == 0) warp_off[warp + 1] = cnt;
    __syncthreads();
    if (warp == 0) {                       // inclusive scan of the 32 segment counts
      int v = warp_off[lane + 1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      warp_off[lane + 1] = v;
      if (lane == 0) warp_off[0] = 0;
    }
    __syncthreads();
    const int off = running + warp_off[warp];
    for (int e = lane; e < cnt; e += 32)
      if (off + e < max_out) ids[off + e] = list[warp * kSelSeg + e];
    running += warp_off[32];
    __syncthreads();
  }
  for (int k = running + threadIdx.x; k < max_out; k += blockDim.x) ids[k] = 0;
  if (threadIdx.x == 0 && count) count[0] = running;
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_assign_targets(const int64_t* d_assigned_gt, int32_t num_problems, int32_t N, int32_t Q,
                                   int32_t num_classes, const float* d_gt_boxes, const int64_t* d_gt_labels,
                                   const int32_t* d_gt_start, const int32_t* d_img_hw, const uint8_t* d_prev_mask,
                                   int64_t* d_labels, float* d_bbox_targets, float* d_bbox_weights,
                                   float* d_teacher_only, void* stream) {
  DSKD_REQUIRE(num_problems >= 0 && N > 0 && Q > 0 && num_classes > 0, "dskd_assign_targets: bad sizes");
  if (num_problems == 0) return DSKD_OK;
  DSKD_REQUIRE(d_assigned_gt && d_gt_start && d_img_hw, "dskd_assign_targets: null pointer");
  DSKD_REQUIRE((!d_bbox_targets || aligned16(d_bbox_targets)) && (!d_bbox_weights || aligned16(d_bbox_weights)),
               "dskd_assign_targets: outputs must be 16-byte aligned");
  const int64_t total = (int64_t)num_problems * Q;
  assign_targets_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      d_assigned_gt, num_problems, N, Q, num_classes, d_gt_boxes, d_gt_labels, d_gt_start, d_img_hw, d_prev_mask,
      d_labels, d_bbox_targets, d_bbox_weights, d_teacher_only);
  DSKD_LAUNCH_OK("assign_targets_kernel");
  return DSKD_OK;
}

extern "C" int dskd_select_prev_queries(const int64_t* d_labels, int32_t n, const uint8_t* d_prev_mask,
                                        int32_t num_classes, int32_t max_out, int64_t* d_ids, int32_t* d_count,
                                        void* stream) {
  DSKD_REQUIRE(n >= 0 && max_out >= 0 && num_classes > 0, "dskd_select_prev_queries: bad sizes");
  DSKD_REQUIRE((n == 0 || d_labels) && d_prev_mask && (max_out == 0 || d_ids), "dskd_select_prev_queries: null pointer");
  select_prev_kernel<<<1, 1024, 0, as_stream(stream)>>>(d_labels, n, d_prev_mask, num_classes, max_out, d_ids, d_count);
  DSKD_LAUNCH_OK("select_prev_kernel");
  return DSKD_OK;
}

extern "C" int dskd_cost_matrix(const float* d_cls, const float* d_box, int32_t num_problems, int32_t N, int32_t Q,
                                int32_t num_classes, int32_t reg_max, const float* d_gt_boxes,
                                const int64_t* d_gt_labels, const int32_t* d_gt_start, const int32_t* d_img_hw,
                                int32_t max_gt, float w_cls, float w_reg, float w_iou, float* d_cost, void* stream) {
  DSKD_REQUIRE(num_problems >= 0 && N > 0 && Q > 0 && num_classes > 0 && reg_max >= 0 && max_gt >= 0,
               "dskd_cost_matrix: bad sizes");
  if (num_problems == 0 || max_gt == 0) return DSKD_OK;
  DSKD_REQUIRE(d_cls && d_box && d_gt_boxes && d_gt_labels && d_gt_start && d_img_hw && d_cost,
               "dskd_cost_matrix: null pointer");
  const size_t smem = (size_t)max_gt * 13 * sizeof(float);
  DSKD_REQUIRE(smem <= 160 * 1024, "dskd_cost_matrix: max_gt %d too large", max_gt);
  if (smem > 48 * 1024)
    DSKD_CUDA_OK(cudaFuncSetAttribute(cost_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(Q, kQTile), (unsigned)num_problems);
  cost_matrix_kernel<<<grid, 128, smem, as_stream(stream)>>>(d_cls, d_box, N, Q, num_classes, reg_max, d_gt_boxes,
                                                            d_gt_labels, d_gt_start, d_img_hw, max_gt, w_cls, w_reg,
                                                            w_iou, d_cost);
  DSKD_LAUNCH_OK("cost_matrix_kernel");
  return DSKD_OK;
}
