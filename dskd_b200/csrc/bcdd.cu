// Between-Class Distance Distillation.
// Reference: gfl_deformable_detr_head_il.py:525-552 (prototype sums / counts, Python loop with one
// `+=` per query) and :1197-1222 (correlation_mat: 2*L^2 torch.dist calls + MSELoss / L).
// The whole block is < 3 MFLOP and < 1 MB: latency-bound, so it is three small launches (prototypes,
// distance+loss+gradient, scatter) with deterministic summation order; the distance uses the direct
// difference in fp32 like torch.dist (a Gram-trick on tensor cores loses the diagonal).
#include "common.cuh"

namespace dskd {

// grid (num_classes, 2): side 0 = teacher, 1 = student.  The rows of one class are gathered in ascending
// index order (each warp compacts a contiguous segment of the label array with ballots, segments are then
// walked in order), so every class sums its rows in the reference's Python-loop order: bit-exact sums.
constexpr int kProtoSeg = 512;                 // labels per warp per chunk
constexpr int kProtoChunk = 8 * kProtoSeg;     // labels per CTA per chunk

__global__ void __launch_bounds__(256) bcdd_proto_kernel(const float* __restrict__ hs_s,
                                                         const int64_t* __restrict__ s_labels, int n_s,
                                                         const float* __restrict__ hs_t,
                                                         const int64_t* __restrict__ t_keep,
                                                         const int64_t* __restrict__ t_labels, int n_t,
                                                         const uint8_t* __restrict__ prev_mask, int C,
                                                         float* __restrict__ proto, int num_classes) {
  __shared__ int list[kProtoChunk];
  __shared__ int warp_cnt[8];
  const int cls = blockIdx.x, side = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* out = proto + ((int64_t)side * num_classes + cls) * (C + 1);
  const int n = side ? n_s : n_t;
  const int64_t* labels = side ? s_labels : t_labels;
  const float* hs = side ? hs_s : hs_t;
  const bool cls_on = side ? (prev_mask[cls] != 0) : true;
  constexpr int kMaxPerThread = 4;  // supports C <= 1024
  float acc[kMaxPerThread] = {0.f, 0.f, 0.f, 0.f};
  int count = 0;
  for (int base = 0; base < n && cls_on; base += kProtoChunk) {
    int cnt = 0;
    const int seg0 = base + warp * kProtoSeg;
    int64_t lab[kProtoSeg / 32];  // all of the lane's labels first: independent loads, one memory latency
#pragma unroll
    for (int u = 0; u < kProtoSeg / 32; ++u) {
      const int q = seg0 + u * 32 + lane;
      lab[u] = (q < n) ? __ldg(labels + q) : -1;
    }
#pragma unroll
    for (int u = 0; u < kProtoSeg / 32; ++u) {
      const int it = u * 32;
      const int q = seg0 + it + lane;
      const bool hit = lab[u] == (int64_t)cls;
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (hit) list[warp * kProtoSeg + cnt + __popc(b & ((1u << lane) - 1u))] = q;
      cnt += __popc(b);
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    for (int w = 0; w < 8; ++w) {
      const int m = warp_cnt[w];
      // four rows in flight at a time, added in ascending index order (the reference's summation order)
      for (int e0 = 0; e0 < m; e0 += 4) {
        const float* rowp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = list[w * kProtoSeg + min(e0 + u, m - 1)];
          rowp[u] = hs + (side ? (int64_t)q : __ldg(t_keep + q)) * (int64_t)C;
        }
#pragma unroll
        for (int k = 0; k < kMaxPerThread; ++k) {
          const int c = threadIdx.x + k * 256;
          if (c < C) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(rowp[u] + c);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (e0 + u < m) acc[k] += v[u];
          }
        }
      }
      count += m;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kMaxPerThread; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < C) out[c] = acc[k];
  }
  if (threadIdx.x == 0) out[C] = (float)count;
}

// normalised prototype element (head_il.py:1198-1206): both sides divide only where the TEACHER count
// is non-zero; the student divides by its own count (0/0 = NaN reproduces the reference's edge).
__device__ __forceinline__ float proto_elem(const float* proto, int num_classes, int C, int side, int k, int c) {
  const float* t_row = proto + (int64_t)k * (C + 1);
  const float* row = proto + ((int64_t)side * num_classes + k) * (C + 1);
  const float n_t = t_row[C];
  return (n_t != 0.f) ? __fdiv_rn(row[c], row[C]) : row[c];
}

// grid L: CTA k computes row k of both distance matrices, then d loss / d (student sum row k).
// STAGED: every CTA first normalises all 2*L prototypes into shared memory (2*L*C floats: 82 KB at L=40,
// 144 KB at L=70) so the L x C inner loops never leave the SM; otherwise they are re-derived from global.
template <bool STAGED>
__global__ void __launch_bounds__(256) bcdd_distance_kernel(const float* __restrict__ proto, int num_classes,
                                                            int C, int L, float gcoef, float grad_scale,
                                                            float* __restrict__ dist,
                                                            float* __restrict__ grad_proto_s) {
  extern __shared__ float sm[];
  float* coef = sm;                              // [L]
  float* ck_t = sm + L;                          // [C] (not STAGED) | all teacher rows [L][C] (STAGED)
  float* ck_s = STAGED ? sm + L + L * C : sm + L + C;
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (STAGED) {
    // counts first (coef doubles as scratch: teacher count, then the student divisor goes to registers per row)
    float* cnt_t = coef;                         // [L], overwritten by the coefficients after the distances
    for (int j = threadIdx.x; j < L; j += blockDim.x) cnt_t[j] = proto[(int64_t)j * (C + 1) + C];
    __syncthreads();
    // one warp per prototype row: coalesced row loads, one IEEE division per element, four rows in flight
    for (int j = warp; j < L; j += nw) {
      const float n_t = cnt_t[j];
      const float* t_row = proto + (int64_t)j * (C + 1);
      const float* s_row = proto + ((int64_t)num_classes + j) * (C + 1);
      const float n_s = s_row[C];
      for (int c = lane; c < C; c += 32) {
        const float a = t_row[c], b = s_row[c];
        ck_t[j * C + c] = (n_t != 0.f) ? __fdiv_rn(a, n_t) : a;
        ck_s[j * C + c] = (n_t != 0.f) ? __fdiv_rn(b, n_s) : b;
      }
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      ck_t[c] = proto_elem(proto, num_classes, C, 0, k, c);
      ck_s[c] = proto_elem(proto, num_classes, C, 1, k, c);
    }
  }
  __syncthreads();
  const float* mine_t = STAGED ? ck_t + k * C : ck_t;
  const float* mine_s = STAGED ? ck_s + k * C : ck_s;
  for (int j = warp; j < L; j += nw) {
    float st = 0.f, ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float ot = STAGED ? ck_t[j * C + c] : proto_elem(proto, num_classes, C, 0, j, c);
      const float os = STAGED ? ck_s[j * C + c] : proto_elem(proto, num_classes, C, 1, j, c);
      const float dt = mine_t[c] - ot;
      const float ds = mine_s[c] - os;
      st = fmaf(dt, dt, st);
      ss = fmaf(ds, ds, ss);
    }
    st = warp_sum(st);
    ss = warp_sum(ss);
    if (lane == 0) {
      const float d_t = sqrtf(st), d_s = sqrtf(ss);
      dist[(int64_t)k * L + j] = d_t;
      dist[(int64_t)L * L + (int64_t)k * L + j] = d_s;
      // d loss / d D_S[k,j] = -gcoef * (D_T - D_S); row k and column k both carry c_k: factor 2.
      // torch.dist backward yields 0 where the distance is 0 (diagonal, coincident prototypes).
      coef[j] = (d_s > 0.f) ? (2.f * -gcoef * (d_t - d_s) / d_s) : ((d_s == 0.f) ? 0.f : d_s /*NaN*/);
    }
  }
  if (grad_proto_s == nullptr) return;
  __syncthreads();
  const float n_t = proto[(int64_t)k * (C + 1) + C];
  const float n_s = proto[((int64_t)num_classes + k) * (C + 1) + C];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g = 0.f;
    const float ck = mine_s[c];
    for (int j = 0; j < L; ++j) {
      const float os = STAGED ? ck_s[j * C + c] : proto_elem(proto, num_classes, C, 1, j, c);
      g = fmaf(coef[j], ck - os, g);
    }
    if (n_t != 0.f) g = __fdiv_rn(g, n_s);
    grad_proto_s[(int64_t)k * (C + 1) + c] = g * grad_scale;
  }
}

// The whole tail of BCDD after the prototypes (and their optional all-reduce) in ONE launch, grid L: CTA k
//   1. normalises and stages all 2 L prototypes (as bcdd_distance_kernel<true>),
//   2. computes row k of both distance matrices and d loss / d (student sum row k),
//   3. scatters that row to every query labelled k (head_il.py:533-539 backward) and zero-fills its share of the queries
//      without a previous label -- no second kernel needs the other rows,
//   4. the LAST CTA to finish adds up (D_T - D_S)^2 over the finished matrices in a fixed order (deterministic loss).
// `counter` is a zero-initialised word the last CTA resets.
__global__ void __launch_bounds__(1024) bcdd_tail_kernel(const float* __restrict__ proto, int num_classes, int C, int L,
                                                        float gcoef, float factor, float grad_scale,
                                                        float* __restrict__ dist, float* __restrict__ loss,
                                                        float* __restrict__ grad_proto_s,
                                                        const int64_t* __restrict__ labels, int n_rows,
                                                        const uint8_t* __restrict__ prev_mask,
                                                        float* __restrict__ grad_hs, unsigned* __restrict__ counter) {
  extern __shared__ float sm[];
  __shared__ double red[32];
  __shared__ bool last_s;
  __shared__ uint8_t prev_s[1024];  // previous-class flags (num_classes <= 1024, checked by the host)
  for (int i = threadIdx.x; i < num_classes; i += blockDim.x) prev_s[i] = prev_mask[i];
  float* coef = sm;               // [L]
  float* cnt_s = sm + L;          // [L] student counts
  float* ck_t = sm + 2 * L;       // [L][C]
  float* ck_s = ck_t + L * C;     // [L][C]
  float* grow = ck_s + L * C;     // [C] gradient row k
  const int k = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  {
    // counts first, then the rows in batches of 8: a thread owns a channel and has 16 independent loads in flight
    // reciprocal counts (1 where the teacher saw no query of the class: both rows stay un-normalised, head_il.py:1205).
    // One IEEE division per class and side; the elements are then scaled by it (<= 1 ulp from the reference's division
    // per element: 2 L C divisions per CTA were most of this kernel's instructions)
    float* cnt_t = coef;  // overwritten by the coefficients after the distances
    for (int j = tid; j < L; j += blockDim.x) {
      const float n_t = __ldg(proto + (int64_t)j * (C + 1) + C);
      const float n_s = __ldg(proto + ((int64_t)num_classes + j) * (C + 1) + C);
      cnt_t[j] = (n_t != 0.f) ? __fdiv_rn(1.f, n_t) : 1.f;
      cnt_s[j] = (n_t != 0.f) ? __fdiv_rn(1.f, n_s) : 1.f;   // n_s == 0: inf, and 0 * inf = NaN like the reference's 0 / 0
    }
    __syncthreads();
    const float* __restrict__ pt = proto;
    const float* __restrict__ ps = proto + (int64_t)num_classes * (C + 1);
    // work item = (batch of 8 rows, channel)
    const int items = ((L + 7) / 8) * C;
    for (int it = tid; it < items; it += blockDim.x) {
      const int c = it % C, j0 = (it / C) * 8;
      {
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = min(j0 + u, L - 1);
          a[u] = __ldg(pt + (int64_t)j * (C + 1) + c);
          b[u] = __ldg(ps + (int64_t)j * (C + 1) + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u;
          if (j < L) {
            ck_t[j * C + c] = a[u] * cnt_t[j];
            ck_s[j * C + c] = b[u] * cnt_s[j];
          }
        }
      }
    }
  }
  __syncthreads();
  const float* mine_t = ck_t + k * C;
  const float* mine_s = ck_s + k * C;
  for (int j = warp; j < L; j += nw) {
    float st = 0.f, ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float dt = mine_t[c] - ck_t[j * C + c];
      const float ds = mine_s[c] - ck_s[j * C + c];
      st = fmaf(dt, dt, st);
      ss = fmaf(ds, ds, ss);
    }
    st = warp_sum(st);
    ss = warp_sum(ss);
    if (lane == 0) {
      const float d_t = sqrtf(st), d_s = sqrtf(ss);
      dist[(int64_t)k * L + j] = d_t;
      dist[(int64_t)L * L + (int64_t)k * L + j] = d_s;
      coef[j] = (d_s > 0.f) ? (2.f * -gcoef * (d_t - d_s) / d_s) : ((d_s == 0.f) ? 0.f : d_s /*NaN*/);
    }
  }
  __syncthreads();
  if (grad_hs != nullptr || grad_proto_s != nullptr) {
    const float n_t = proto[(int64_t)k * (C + 1) + C];
    const float n_s = proto[((int64_t)num_classes + k) * (C + 1) + C];
    for (int c = tid; c < C; c += blockDim.x) {
      float g = 0.f;
      const float ck = mine_s[c];
      for (int j = 0; j < L; ++j) g = fmaf(coef[j], ck - ck_s[j * C + c], g);
      if (n_t != 0.f) g = __fdiv_rn(g, n_s);
      g *= grad_scale;
      grow[c] = g;
      if (grad_proto_s != nullptr) grad_proto_s[(int64_t)k * (C + 1) + c] = g;
    }
    __syncthreads();
  }
  if (grad_hs != nullptr) {
    // queries labelled k take the row; the queries without a previous label below L are zero-filled by CTA (q mod L).
    // Eight labels per thread are fetched before any of them is looked at (one memory latency per 2048 queries).
    const bool k_on = prev_mask[k] != 0;
    constexpr int kBatch = 8;
    for (int q0 = 0; q0 < n_rows; q0 += kBatch * (int)blockDim.x) {
      int64_t lab[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int q = q0 + u * blockDim.x + tid;
        lab[u] = (q < n_rows) ? __ldg(labels + q) : -2;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int q = q0 + u * blockDim.x + tid;
        int kind = 0;  // 1: my row, 2: zero row
        if (q < n_rows) {
          const bool prev = lab[u] >= 0 && lab[u] < L && lab[u] < num_classes && prev_s[lab[u]] != 0;
          if (prev) kind = (lab[u] == k && k_on) ? 1 : 0;
          else kind = (q % L == k) ? 2 : 0;
        }
        unsigned mine = __ballot_sync(0xffffffffu, kind == 1), zero = __ballot_sync(0xffffffffu, kind == 2);
        const int qw = q0 + u * blockDim.x + warp * 32;
        while (mine | zero) {
          const int b = __ffs(mine | zero) - 1;
          const bool z = (zero >> b) & 1u;
          float* dst = grad_hs + (int64_t)(qw + b) * C;
          for (int c = lane; c < C; c += 32) dst[c] = z ? 0.f : grow[c];
          mine &= ~(1u << b);
          zero &= ~(1u << b);
        }
      }
    }
  }
  if (counter == nullptr) return;
  // ---- the last CTA reduces the loss over the finished matrices
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  double acc = 0.0;
  const int n = L * L;
  for (int i = tid; i < n; i += blockDim.x) {
    const double d = (double)__ldcg(dist + i) - (double)__ldcg(dist + n + i);
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (tid == 0) {
    loss[0] = (float)(acc * (double)factor);
    *counter = 0u;
  }
}

// loss = loss_weight * reduce((D_T - D_S)^2) / L, fixed summation order, double accumulation.
__global__ void __launch_bounds__(256) bcdd_loss_kernel(const float* __restrict__ dist, int L, float factor,
                                                        float* __restrict__ loss) {
  __shared__ double red[32];
  double acc = 0.0;
  const int n = L * L;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)dist[i] - (double)dist[n + i];
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = (float)(acc * (double)factor);
}

__global__ void __launch_bounds__(256) bcdd_scatter_kernel(const float* __restrict__ grad_proto_s,
                                                           const int64_t* __restrict__ labels, int n,
                                                           const uint8_t* __restrict__ prev_mask,
                                                           int num_classes, int C, float* __restrict__ grad_hs) {
  const int q = blockIdx.x;
  const int64_t lab = labels[q];
  const bool on = lab >= 0 && lab < num_classes && prev_mask[lab] != 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    grad_hs[(int64_t)q * C + c] = on ? grad_proto_s[lab * (int64_t)(C + 1) + c] : 0.f;
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_bcdd_prototypes(const float* d_hs_student, const int64_t* d_student_labels,
                                    int32_t num_student_rows, const float* d_hs_teacher,
                                    const int64_t* d_teacher_keepid, const int64_t* d_teacher_labels,
                                    int32_t num_teacher, const uint8_t* d_prev_mask, int32_t num_classes,
                                    int32_t C, float* d_proto, void* stream) {
  DSKD_REQUIRE(num_classes > 0 && C > 0 && C <= 1024 && num_student_rows >= 0 && num_teacher >= 0,
               "dskd_bcdd_prototypes: bad sizes (C must be <= 1024)");
  DSKD_REQUIRE(d_proto && d_prev_mask, "dskd_bcdd_prototypes: null pointer");
  DSKD_REQUIRE(num_student_rows == 0 || (d_hs_student && d_student_labels), "dskd_bcdd_prototypes: null student input");
  DSKD_REQUIRE(num_teacher == 0 || (d_hs_teacher && d_teacher_keepid && d_teacher_labels),
               "dskd_bcdd_prototypes: null teacher input");
  bcdd_proto_kernel<<<dim3(num_classes, 2), 256, 0, as_stream(stream)>>>(
      d_hs_student, d_student_labels, num_student_rows, d_hs_teacher, d_teacher_keepid, d_teacher_labels,
      num_teacher, d_prev_mask, C, d_proto, num_classes);
  DSKD_LAUNCH_OK("bcdd_proto_kernel");
  return DSKD_OK;
}

extern "C" int dskd_bcdd_distance_loss(const float* d_proto, int32_t num_classes, int32_t C, int32_t L,
                                       int32_t reduction, float loss_weight, float grad_scale, float* d_dist,
                                       float* d_loss, float* d_grad_proto_student, void* stream) {
  DSKD_REQUIRE(d_proto && d_dist && d_loss, "dskd_bcdd_distance_loss: null pointer");
  DSKD_REQUIRE(L > 0 && L <= num_classes && C > 0, "dskd_bcdd_distance_loss: need 0 < L <= num_classes");
  DSKD_REQUIRE(reduction == 1 || reduction == 2, "dskd_bcdd_distance_loss: reduction must be 1 (mean) or 2 (sum)");
  cudaStream_t st = as_stream(stream);
  // loss = w * red * sum((D_T-D_S)^2) / L, red = 1/L^2 (mean) or 1 (sum)
  const double red = (reduction == 1) ? 1.0 / ((double)L * (double)L) : 1.0;
  const float factor = (float)((double)loss_weight * red / (double)L);
  const float gcoef = 2.f * factor;  // d loss / d D_S = -gcoef * (D_T - D_S)
  if (d_grad_proto_student != nullptr)
    DSKD_CUDA_OK(cudaMemsetAsync(d_grad_proto_student, 0, sizeof(float) * (size_t)num_classes * (C + 1), st));
  const size_t staged = sizeof(float) * (2 * (size_t)L * C + L);
  if (staged <= 200 * 1024) {
    if (staged > 48 * 1024)
      DSKD_CUDA_OK(cudaFuncSetAttribute(bcdd_distance_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
    bcdd_distance_kernel<true><<<L, 256, staged, st>>>(d_proto, num_classes, C, L, gcoef, grad_scale, d_dist,
                                                       d_grad_proto_student);
  } else {
    const size_t smem = sizeof(float) * (2 * (size_t)C + L);
    bcdd_distance_kernel<false><<<L, 256, smem, st>>>(d_proto, num_classes, C, L, gcoef, grad_scale, d_dist,
                                                      d_grad_proto_student);
  }
  DSKD_LAUNCH_OK("bcdd_distance_kernel");
  bcdd_loss_kernel<<<1, 256, 0, st>>>(d_dist, L, factor, d_loss);
  DSKD_LAUNCH_OK("bcdd_loss_kernel");
  return DSKD_OK;
}

extern "C" int dskd_bcdd_scatter_grad(const float* d_grad_proto_student, const int64_t* d_student_labels,
                                      int32_t num_student_rows, const uint8_t* d_prev_mask, int32_t num_classes,
                                      int32_t C, float* d_grad_hs_student, void* stream) {
  DSKD_REQUIRE(num_student_rows >= 0 && C > 0 && num_classes > 0, "dskd_bcdd_scatter_grad: bad sizes");
  if (num_student_rows == 0) return DSKD_OK;
  DSKD_REQUIRE(d_grad_proto_student && d_student_labels && d_prev_mask && d_grad_hs_student,
               "dskd_bcdd_scatter_grad: null pointer");
  bcdd_scatter_kernel<<<num_student_rows, 256, 0, as_stream(stream)>>>(
      d_grad_proto_student, d_student_labels, num_student_rows, d_prev_mask, num_classes, C, d_grad_hs_student);
  DSKD_LAUNCH_OK("bcdd_scatter_kernel");
  return DSKD_OK;
}

// dskd_bcdd_distance_loss + dskd_bcdd_scatter_grad as one launch when the staged prototypes fit shared memory (L <= 90 at
// C = 256); the count column of the last row of the [num_classes, C+1] gradient workspace doubles as the finish counter
// (it is zero by contract and the last CTA leaves it zero).
extern "C" int dskd_bcdd_loss_and_grad(const float* d_proto, int32_t num_classes, int32_t C, int32_t L, int32_t reduction,
                                       float loss_weight, float grad_scale, const int64_t* d_student_labels,
                                       int32_t num_student_rows, const uint8_t* d_prev_mask, float* d_dist, float* d_loss,
                                       float* d_grad_proto_student, float* d_grad_hs_student, void* stream) {
  DSKD_REQUIRE((d_grad_hs_student == nullptr) || d_grad_proto_student, "dskd_bcdd_loss_and_grad: gradient needs the prototype workspace");
  const size_t staged = sizeof(float) * (2 * (size_t)L * C + 2 * (size_t)L + C);
  const bool fused = d_grad_hs_student != nullptr && L > 0 && L <= num_classes && num_classes <= 1024 && C > 0 &&
                     staged <= 200 * 1024 &&
                     (reduction == 1 || reduction == 2) && num_student_rows > 0;
  if (!fused) {
    int rc = dskd_bcdd_distance_loss(d_proto, num_classes, C, L, reduction, loss_weight, grad_scale, d_dist, d_loss,
                                     d_grad_hs_student ? d_grad_proto_student : nullptr, stream);
    if (rc || d_grad_hs_student == nullptr) return rc;
    return dskd_bcdd_scatter_grad(d_grad_proto_student, d_student_labels, num_student_rows, d_prev_mask, num_classes, C,
                                  d_grad_hs_student, stream);
  }
  DSKD_REQUIRE(d_proto && d_dist && d_loss && d_student_labels && d_prev_mask, "dskd_bcdd_loss_and_grad: null pointer");
  cudaStream_t st = as_stream(stream);
  const double red = (reduction == 1) ? 1.0 / ((double)L * (double)L) : 1.0;
  const float factor = (float)((double)loss_weight * red / (double)L);
  const float gcoef = 2.f * factor;
  DSKD_CUDA_OK(cudaMemsetAsync(d_grad_proto_student, 0, sizeof(float) * (size_t)num_classes * (C + 1), st));
  unsigned* counter = reinterpret_cast<unsigned*>(d_grad_proto_student + (size_t)num_classes * (C + 1) - 1);
  if (staged > 48 * 1024)
    DSKD_CUDA_OK(cudaFuncSetAttribute(bcdd_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
  bcdd_tail_kernel<<<L, 1024, staged, st>>>(d_proto, num_classes, C, L, gcoef, factor, grad_scale, d_dist, d_loss,
                                           d_grad_proto_student, d_student_labels, num_student_rows, d_prev_mask,
                                           d_grad_hs_student, counter);
  DSKD_LAUNCH_OK("bcdd_tail_kernel");
  return DSKD_OK;
}
