// Teacher keep-ids: what the frozen teacher's last decoder layer contributes to the
// distillation path, without per-image Python or `nonzero()` / `sort()` host round trips.
// Reference: gfl_deformable_detr_head_il.py:1622-1668 (`_get_bboxes_single`, sigmoid branch, need_logits),
// core/utils/misc.py:143-152 (`filter_scores_and_topk`: scores > thr, sort descending, keep <= topk
// (query, class) pairs -- a query may appear more than once), head_il.py:42-59 (`Integral_average`),
// transforms.py:245-256, detectors/deformable_detr_il.py:138-151 (keep-id flattening q + Q*i).
//
// One CTA per image.  Every (query, class) score is turned into a 64-bit key
//   (float bits of sigmoid(logit)) << 32 | ~flat_index
// so "larger key" == "higher score, ties to the lower flat index" (a stable descending sort of the
// row-major `nonzero()` order).  An 8-pass radix select finds the K-th largest valid key, the <= K winners
// are gathered and bitonic-sorted in shared memory, and each winner decodes its own box.
#include "common.cuh"

namespace dskd {

constexpr int kTeacherThreads = 1024;
constexpr int kTeacherMaxKeep = 1024;  // max_per_img supported by the shared-memory sort

__device__ __forceinline__ unsigned long long score_key(float logit, unsigned flat, float thr, bool& valid) {
  const float s = __fdiv_rn(1.f, 1.f + expf(-logit));
  valid = s > thr;
  return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xffffffffu - flat);
}

__global__ void __launch_bounds__(kTeacherThreads) teacher_decode_kernel(
    const float* __restrict__ cls, const float* __restrict__ box, int Q, int num_classes, int reg_max,
    const int* __restrict__ img_hw, float thr, int max_keep, int* __restrict__ count, float* __restrict__ out_boxes,
    float* __restrict__ out_scores, int64_t* __restrict__ out_labels, int64_t* __restrict__ out_keepid,
    float* __restrict__ out_logits) {
  __shared__ unsigned hist[256];
  __shared__ unsigned long long keys[kTeacherMaxKeep];
  __shared__ unsigned long long sel_prefix;
  __shared__ int sel_k, n_valid, n_out;
  const int img = blockIdx.x, tid = threadIdx.x;
  const int total = Q * num_classes;
  const float* __restrict__ x = cls + (int64_t)img * total;

  // ---- number of valid scores
  if (tid == 0) n_valid = 0;
  __syncthreads();
  {
    int c = 0;
    for (int f = tid; f < total; f += kTeacherThreads) {
      bool v;
      score_key(x[f], (unsigned)f, thr, v);
      c += v ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((tid & 31) == 0 && c) atomicAdd(&n_valid, c);
  }
  __syncthreads();
  const int K = min(max_keep, n_valid);
  if (tid == 0) count[img] = K;
  // padding of the fixed-capacity outputs
  for (int r = K + tid; r < max_keep; r += kTeacherThreads) {
    const int64_t o = (int64_t)img * max_keep + r;
    if (out_scores) out_scores[o] = 0.f;
    if (out_labels) out_labels[o] = -1;
    if (out_keepid) out_keepid[o] = -1;
    if (out_boxes) reinterpret_cast<float4*>(out_boxes)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (K == 0) return;

  // ---- radix select: the K-th largest key
  if (tid == 0) { sel_prefix = 0ull; sel_k = K; }
  for (int byte = 7; byte >= 0; --byte) {
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    const unsigned long long prefix = sel_prefix;
    const unsigned long long himask = (byte == 7) ? 0ull : (~0ull << (8 * (byte + 1)));
    for (int f = tid; f < total; f += kTeacherThreads) {
      bool v;
      const unsigned long long k = score_key(x[f], (unsigned)f, thr, v);
      if (v && (k & himask) == prefix) atomicAdd(&hist[(unsigned)(k >> (8 * byte)) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int need = sel_k, b = 255;
      for (; b > 0; --b) {
        if ((int)hist[b] >= need) break;
        need -= (int)hist[b];
      }
      sel_prefix = prefix | ((unsigned long long)b << (8 * byte));
      sel_k = need;
    }
    __syncthreads();
  }
  const unsigned long long kth = sel_prefix;  // keys are distinct: exactly K valid keys are >= kth

  // ---- gather the winners, sort descending (bitonic over the next power of two)
  if (tid == 0) n_out = 0;
  int P2 = 1;
  while (P2 < K) P2 <<= 1;
  for (int r = tid; r < P2; r += kTeacherThreads) keys[r] = 0ull;
  __syncthreads();
  for (int f = tid; f < total; f += kTeacherThreads) {
    bool v;
    const unsigned long long k = score_key(x[f], (unsigned)f, thr, v);
    if (v && k >= kth) keys[atomicAdd(&n_out, 1)] = k;
  }
  __syncthreads();
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < P2 / 2; t += kTeacherThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }

  // ---- decode: one winner per thread
  const float img_h = (float)img_hw[2 * img], img_w = (float)img_hw[2 * img + 1];
  const int bins = reg_max + 1, box_ch = (reg_max > 0) ? 2 + 4 * bins : 4;
  for (int r = tid; r < K; r += kTeacherThreads) {
    const unsigned long long k = keys[r];
    const unsigned flat = 0xffffffffu - (unsigned)(k & 0xffffffffull);
    const int q = (int)(flat / (unsigned)num_classes), lab = (int)(flat % (unsigned)num_classes);
    const int64_t o = (int64_t)img * max_keep + r;
    if (out_scores) out_scores[o] = __uint_as_float((unsigned)(k >> 32));
    if (out_labels) out_labels[o] = lab;
    if (out_keepid) out_keepid[o] = (int64_t)q + (int64_t)Q * img;
    if (out_boxes) {
      const float* b = box + ((int64_t)img * Q + q) * box_ch;
      float wh[4];
      if (reg_max > 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // Integral_average, head_il.py:54-59
          const float* p = b + 2 + g * bins;
          float s = 0.f;
          for (int j = 0; j < bins; ++j) s += p[j];
          float e = 0.f;
          for (int j = 0; j < bins; ++j)
            e += __fmul_rn(__fdiv_rn(p[j], s), __fdiv_rn(__fdiv_rn((float)j, (float)reg_max), 2.f));
          wh[g] = e;
        }
      }
      const float cx = b[0], cy = b[1];
      const float w = (reg_max > 0) ? wh[0] + wh[1] : b[2], h = (reg_max > 0) ? wh[2] + wh[3] : b[3];
      float x1 = __fmul_rn(cx - 0.5f * w, img_w), y1 = __fmul_rn(cy - 0.5f * h, img_h);
      float x2 = __fmul_rn(cx + 0.5f * w, img_w), y2 = __fmul_rn(cy + 0.5f * h, img_h);
      x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);
      y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
      reinterpret_cast<float4*>(out_boxes)[o] = make_float4(x1, y1, x2, y2);
    }
  }
  if (out_logits) {  // `det_logits = cls_score.sigmoid()[bbox_index]`  (:1636)
    for (int e = tid; e < K * num_classes; e += kTeacherThreads) {
      const int r = e / num_classes, c = e - r * num_classes;
      const unsigned flat = 0xffffffffu - (unsigned)(keys[r] & 0xffffffffull);
      const int q = (int)(flat / (unsigned)num_classes);
      out_logits[((int64_t)img * max_keep + r) * num_classes + c] =
          __fdiv_rn(1.f, 1.f + expf(-x[(int64_t)q * num_classes + c]));
    }
  }
}

// Ragged -> concatenated: start[i] = sum_{j<i} count[j] and the first count[i] entries of every image copied
// behind each other (the `torch.cat` of deformable_detr_il.py:151 / head_il.py:462-465).  Single CTA.
__global__ void __launch_bounds__(1024) teacher_compact_kernel(const int* __restrict__ count, int N, int max_keep,
                                                               const float* __restrict__ boxes,
                                                               const float* __restrict__ scores,
                                                               const int64_t* __restrict__ labels,
                                                               const int64_t* __restrict__ keepid,
                                                               int* __restrict__ start, float* __restrict__ c_boxes,
                                                               float* __restrict__ c_scores, int64_t* __restrict__ c_labels,
                                                               int64_t* __restrict__ c_keepid) {
  extern __shared__ int s_start[];  // N + 1
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < N; ++i) { s_start[i] = acc; acc += count[i]; }
    s_start[N] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= N; i += blockDim.x) start[i] = s_start[i];
  for (int e = threadIdx.x; e < N * max_keep; e += blockDim.x) {
    const int i = e / max_keep, r = e - i * max_keep;
    if (r >= s_start[i + 1] - s_start[i]) continue;
    const int o = s_start[i] + r;
    if (c_boxes) reinterpret_cast<float4*>(c_boxes)[o] = reinterpret_cast<const float4*>(boxes)[e];
    if (c_scores) c_scores[o] = scores[e];
    if (c_labels) c_labels[o] = labels[e];
    if (c_keepid) c_keepid[o] = keepid[e];
  }
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_teacher_decode(const float* d_cls, const float* d_box, int32_t N, int32_t Q, int32_t num_classes,
                                   int32_t reg_max, const int32_t* d_img_hw, float score_thr, int32_t max_per_img,
                                   int32_t* d_count, float* d_bboxes, float* d_scores, int64_t* d_labels,
                                   int64_t* d_keepid, float* d_logits, void* stream) {
  DSKD_REQUIRE(N >= 0 && Q > 0 && num_classes > 0 && reg_max >= 0, "dskd_teacher_decode: bad sizes");
  DSKD_REQUIRE(max_per_img > 0 && max_per_img <= kTeacherMaxKeep, "dskd_teacher_decode: max_per_img must be in 1..%d",
               kTeacherMaxKeep);
  DSKD_REQUIRE((int64_t)Q * num_classes < (1ll << 31), "dskd_teacher_decode: Q * num_classes too large");
  if (N == 0) return DSKD_OK;
  DSKD_REQUIRE(d_cls && d_img_hw && d_count && (d_bboxes == nullptr || d_box), "dskd_teacher_decode: null pointer");
  DSKD_REQUIRE(d_bboxes == nullptr || aligned16(d_bboxes), "dskd_teacher_decode: d_bboxes must be 16-byte aligned");
  teacher_decode_kernel<<<N, kTeacherThreads, 0, as_stream(stream)>>>(d_cls, d_box, Q, num_classes, reg_max, d_img_hw,
                                                                     score_thr, max_per_img, d_count, d_bboxes, d_scores,
                                                                     d_labels, d_keepid, d_logits);
  DSKD_LAUNCH_OK("teacher_decode_kernel");
  return DSKD_OK;
}

extern "C" int dskd_teacher_compact(const int32_t* d_count, int32_t N, int32_t max_per_img, const float* d_bboxes,
                                    const float* d_scores, const int64_t* d_labels, const int64_t* d_keepid,
                                    int32_t* d_start, float* d_cat_bboxes, float* d_cat_scores, int64_t* d_cat_labels,
                                    int64_t* d_cat_keepid, void* stream) {
  DSKD_REQUIRE(N >= 0 && max_per_img > 0 && N <= 8192, "dskd_teacher_compact: bad sizes");
  DSKD_REQUIRE(d_start && (N == 0 || d_count), "dskd_teacher_compact: null pointer");
  DSKD_REQUIRE((!d_cat_bboxes || (d_bboxes && aligned16(d_bboxes) && aligned16(d_cat_bboxes))) &&
                   (!d_cat_scores || d_scores) && (!d_cat_labels || d_labels) && (!d_cat_keepid || d_keepid),
               "dskd_teacher_compact: an output was requested without its (16-byte aligned) input");
  teacher_compact_kernel<<<1, 1024, (N + 1) * sizeof(int), as_stream(stream)>>>(
      d_count, N, max_per_img, d_bboxes, d_scores, d_labels, d_keepid, d_start, d_cat_bboxes, d_cat_scores, d_cat_labels,
      d_cat_keepid);
  DSKD_LAUNCH_OK("teacher_compact_kernel");
  return DSKD_OK;
}
