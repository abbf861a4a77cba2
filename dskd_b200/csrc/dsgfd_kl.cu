// DSG-FD with the shipped criterion: KnowledgeDistillationKLDivLoss(T=2, reduction='sum') applied to
// [C,H,W] maps, i.e. softmax over H (kd_loss.py:28-34 with dim=1 == H), target = student*mask
// (detached), pred = teacher*mask.  Reference: gfl_deformable_detr_head_il.py:707-718 + kd_loss.py:12-43.
//
// Cells outside every box have logit 0 for both softmaxes and need no feature bytes; strips without any owned cell
// contribute exactly 0 and are skipped.  No gradient reaches the student features (target is detached): the only
// gradient is d loss / d rows.  Algorithmic bytes: read S + read T = 8 B per element (45.51 MB per 800x1333 image).
#include "common.cuh"

namespace dskd {

// Strip kernel.  A CTA owns one strip = (level, image, kKlCols consecutive w) x all H rows and a chunk of channels; each
// of its warps walks kKlChan channels one after the other; the 32 lanes are kKlCols columns x kKlPhases row phases (a
// lane takes every kKlPhases-th row of its column; narrow strips keep the shared-memory footprint per warp small enough
// for 32 resident warps per SM).  Once per CTA every column is cut into SEGMENTS, the maximal runs of rows with one
// owning box; all per-channel work is a loop over the segments of the lane's column, so only rows inside boxes are
// loaded or visited (the others have logit 0 in both softmaxes and enter in closed form), the mask value is a
// per-segment constant and the inner loops are branch-free.  Per channel the warp
//   0. cp.async's the segment rows of the student / teacher strip into ITS shared-memory buffers (every feature byte
//      is read from HBM exactly once),
//   A. turns them in place into the logits x = feature * mask / T and takes the column maxima,
//   B. sums e^(x - max) for both softmaxes and sum e^(xs - max) (xs - xt) -- the KL of a column needs nothing else:
//        KL = sum_h t_h (xs_h - xt_h) - (lse_s - lse_t),
//   C. (only when the mask rows need a gradient) accumulates T_h * (p_h - t_h) over each segment and issues one
//      red.global per (column, segment, channel).
constexpr int kKlChan = 4;        // channels per warp (sequential)
constexpr int kKlCols = 8;        // columns (w) per strip: a warp covers kKlCols columns x kKlPhases interleaved row phases
constexpr int kKlPhases = 32 / kKlCols;
constexpr int kKlMaxWarps = 32;   // warps per CTA
constexpr int kKlMaxH = 800;      // rows per level the shared-memory strips can hold (one warp per CTA at the limit)
constexpr size_t kKlSmemBudget = 220 * 1024;
__host__ __device__ inline size_t kl_header_bytes(int max_h) {
  // owner strip + first row / end row / owner of every segment: 4 x int32 [max_h][kKlCols]; 2 x int32 [kKlCols] counters
  return ((size_t)max_h * kKlCols * 4 * 4 + (size_t)kKlCols * 4 * 2 + 127) / 128 * 128;
}

struct KlParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* student[DSKD_MAX_LEVELS];
  const float* teacher[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  int block_start[DSKD_MAX_LEVELS + 1];
  int wtiles[DSKD_MAX_LEVELS];
  int num_levels, N, C;
  int warps, max_h;               // warps per CTA, max H over the levels (sizes the shared-memory strips)
  float temperature, inv_temperature;
  int64_t cells_per_image;
  const int* owner;
  const float* rows;
  float* grad_rows;
  const float* cell_weight;
  double* loss;
};

__device__ __forceinline__ void cp_async_f32(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// POW2: T is a power of two, so (feature * mask) / T == feature * (mask / T) bit for bit and the division is hoisted.
template <bool CELL, bool POW2>
__global__ void __launch_bounds__(32 * kKlMaxWarps) dsgfd_kl_kernel(const __grid_constant__ KlParams prm) {
  extern __shared__ __align__(16) unsigned char kl_smem[];
  __shared__ double red[32];
  __shared__ int strip_any, max_seg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wl = lane % kKlCols, ph = lane / kKlCols;  // column inside the strip, row phase (rows ph, ph + kKlPhases, ...)
  int lvl = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && (int)blockIdx.x >= prm.block_start[k]) lvl = k;
  const int H = prm.levels[lvl].H, W = prm.levels[lvl].W, C = prm.C;
  const int HW = H * W;
  const int ch_per_cta = prm.warps * kKlChan;
  const int nchunks = (C + ch_per_cta - 1) / ch_per_cta;
  int idx = blockIdx.x - prm.block_start[lvl];
  const int chunk = idx % nchunks;     // channel chunks of one strip are neighbours: the owner strip stays in L2 / L1
  idx /= nchunks;
  const int wt = idx % prm.wtiles[lvl];
  const int img = idx / prm.wtiles[lvl];
  const int w = wt * kKlCols + wl;
  const float Temp = prm.temperature;
  const float scale = prm.scale[lvl];
  const int64_t strip_base = (int64_t)img * prm.cells_per_image + prm.levels[lvl].cell_offset + wt * kKlCols;
  const int64_t cell_base = strip_base + wl;

  // shared memory header (kl_header_bytes): the segment tables of the strip; then per warp the xs / xt strips
  const int mh = prm.max_h;
  int* own_s = reinterpret_cast<int*>(kl_smem);     // [h][col] owner strip (setup only)
  int* seg_h0 = own_s + (size_t)mh * kKlCols;       // [col][k] first row of the k-th segment of the column
  int* seg_h1 = seg_h0 + (size_t)mh * kKlCols;      // [col][k] one past its last row
  int* seg_own = seg_h1 + (size_t)mh * kKlCols;     // [col][k] its owner (pair index; 0 in cell-mask mode)
  int* nseg_s = seg_own + (size_t)mh * kKlCols;     // [col] segments of the column
  int* nown_s = nseg_s + kKlCols;                   // [col] owned rows of the column
  float* xs_s = reinterpret_cast<float*>(kl_smem + kl_header_bytes(mh)) + (size_t)warp * 2 * mh * kKlCols;
  float* xt_s = xs_s + (size_t)mh * kKlCols;

  // ---- once per CTA: the owner strip and, per column, its SEGMENTS (maximal runs of rows with one owner).  Only the
  // rows inside segments are ever loaded or visited: the others have logit 0 in both softmaxes and enter in closed form.
  if (threadIdx.x == 0) { strip_any = 0; max_seg = 0; }
  __syncthreads();
  {
    bool any = false;
    for (int i = threadIdx.x; i < H * kKlCols; i += blockDim.x) {
      const int h = i / kKlCols, c = i % kKlCols;
      int o = -1;
      if (wt * kKlCols + c < W) {
        if (CELL) o = (__ldg(prm.cell_weight + strip_base + c + (int64_t)h * W) != 0.f) ? 0 : -1;
        else o = __ldg(prm.owner + strip_base + c + (int64_t)h * W);
      }
      own_s[i] = o;
      any |= o >= 0;
    }
    if (__any_sync(0xffffffffu, any) && lane == 0) strip_any = 1;
  }
  __syncthreads();
  if (!strip_any) return;  // no box touches this strip: every column's KL is exactly 0
  if (threadIdx.x < kKlCols) {  // one thread per column walks its H rows
    const int c = threadIdx.x;
    int n = 0, prev = -1, owned = 0;
    for (int h = 0; h < H; ++h) {
      const int o = own_s[h * kKlCols + c];
      if (o != prev) {
        if (prev >= 0) seg_h1[c * mh + n - 1] = h;
        if (o >= 0) { seg_h0[c * mh + n] = h; seg_own[c * mh + n] = o; ++n; }
      }
      owned += o >= 0 ? 1 : 0;
      prev = o;
    }
    if (prev >= 0) seg_h1[c * mh + n - 1] = H;
    nseg_s[c] = n;
    nown_s[c] = owned;
    atomicMax(&max_seg, n);
  }
  __syncthreads();

  const float kLog2e = 1.4426950408889634f;
  const float gcoef = scale * Temp / (float)H;  // d loss / d pred = scale * (T/H) * (p - t)
  const bool want_grad = !CELL && prm.grad_rows != nullptr;
  constexpr int kRow = kKlCols;                  // floats per strip row
  const int my_nseg = nseg_s[wl], my_owned = nown_s[wl], nsweep = max_seg;  // the sweep loops are warp-uniform
  const int* my_h0 = seg_h0 + wl * mh;
  const int* my_h1 = seg_h1 + wl * mh;
  const int* my_own = seg_own + wl * mh;
  const uint32_t xs_a = (uint32_t)__cvta_generic_to_shared(xs_s) + (uint32_t)wl * 4u;
  const uint32_t xt_a = (uint32_t)__cvta_generic_to_shared(xt_s) + (uint32_t)wl * 4u;
  float* xs_c = xs_s + wl;                       // xs_c[h * kRow]
  float* xt_c = xt_s + wl;
  auto phase_max = [](float v) {
#pragma unroll
    for (int o = kKlCols; o < 32; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  };
  auto phase_sum = [](float v) {
#pragma unroll
    for (int o = kKlCols; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  // rows of segment k that belong to this lane: hb, hb + kKlPhases, ... < h1 (an empty range when k >= my_nseg)
  auto seg_range = [&](int k, int& hb, int& h1) {
    if (k < my_nseg) {
      const int h0 = my_h0[k];
      h1 = my_h1[k];
      hb = h0 + ((ph - h0) & (kKlPhases - 1));
    } else {
      hb = 0;
      h1 = 0;
    }
  };
  double kl_total = 0.0;

  for (int kc = 0; kc < kKlChan; ++kc) {
    const int c = (chunk * prm.warps + warp) * kKlChan + kc;
    if (c >= C) break;
    const float* __restrict__ S = prm.student[lvl] + ((int64_t)img * C + c) * HW + w;
    const float* __restrict__ T = prm.teacher[lvl] + ((int64_t)img * C + c) * HW + w;
    auto seg_mask = [&](int k) -> float {  // mask value of segment k (divided by T when that is exact)
      const float m = __ldg(prm.rows + (int64_t)my_own[k] * C + c);
      return POW2 ? m * prm.inv_temperature : m;
    };
    // ---- 0: the rows inside segments -> shared memory, asynchronously (every feature byte leaves HBM once)
    for (int k = 0; k < nsweep; ++k) {
      int h, h1;
      seg_range(k, h, h1);
      for (; h < h1; h += kKlPhases) {
        cp_async_f32(xs_a + (uint32_t)(h * kRow) * 4u, S + (unsigned)(h * W));
        cp_async_f32(xt_a + (uint32_t)(h * kRow) * 4u, T + (unsigned)(h * W));
      }
    }
    cp_async_wait_all();
    __syncwarp();
    // ---- A: logits in place, maxima over the owned rows
    float ms = -INFINITY, mt = -INFINITY;
    for (int k = 0; k < nsweep; ++k) {
      int h, h1;
      seg_range(k, h, h1);
      if (h >= h1) continue;
      float m = CELL ? 0.f : seg_mask(k);
      for (; h < h1; h += kKlPhases) {
        if (CELL) {
          m = __ldg(prm.cell_weight + cell_base + (int64_t)h * W);
          if (POW2) m *= prm.inv_temperature;
        }
        float x = xs_c[h * kRow] * m, y = xt_c[h * kRow] * m;
        if (!POW2) { x = __fdiv_rn(x, Temp); y = __fdiv_rn(y, Temp); }
        xs_c[h * kRow] = x;
        xt_c[h * kRow] = y;
        ms = fmaxf(ms, x);
        mt = fmaxf(mt, y);
      }
    }
    ms = phase_max(ms);
    mt = phase_max(mt);
    if (my_owned < H) { ms = fmaxf(ms, 0.f); mt = fmaxf(mt, 0.f); }  // rows outside boxes: logit 0 in both softmaxes
    // ---- B: softmax sums and the t-weighted logit difference over the owned rows
    const float nms = -ms * kLog2e, nmt = -mt * kLog2e;
    float ss0 = 0.f, ss1 = 0.f, st0 = 0.f, st1 = 0.f, ws0 = 0.f, ws1 = 0.f;
    for (int k = 0; k < nsweep; ++k) {
      int h, h1;
      seg_range(k, h, h1);
      for (; h + kKlPhases < h1; h += 2 * kKlPhases) {
        const float a0 = xs_c[h * kRow], b0 = xt_c[h * kRow];
        const float a1 = xs_c[(h + kKlPhases) * kRow], b1 = xt_c[(h + kKlPhases) * kRow];
        const float e0 = fast_ex2(fmaf(a0, kLog2e, nms)), e1 = fast_ex2(fmaf(a1, kLog2e, nms));
        ss0 += e0; ss1 += e1;
        st0 += fast_ex2(fmaf(b0, kLog2e, nmt)); st1 += fast_ex2(fmaf(b1, kLog2e, nmt));
        ws0 = fmaf(e0, a0 - b0, ws0); ws1 = fmaf(e1, a1 - b1, ws1);
      }
      if (h < h1) {
        const float a0 = xs_c[h * kRow], b0 = xt_c[h * kRow];
        const float e0 = fast_ex2(fmaf(a0, kLog2e, nms));
        ss0 += e0;
        st0 += fast_ex2(fmaf(b0, kLog2e, nmt));
        ws0 = fmaf(e0, a0 - b0, ws0);
      }
    }
    // the H - owned rows outside boxes add e^(0 - max) to each sum and nothing to the weighted difference
    const float rest = (float)(H - my_owned);
    const float sum_s = fmaf(rest, fast_ex2(nms), phase_sum(ss0 + ss1));
    const float sum_t = fmaf(rest, fast_ex2(nmt), phase_sum(st0 + st1));
    const float wsum = phase_sum(ws0 + ws1);
    // KL of the column = sum_h t_h (xs - xt) - (lse_s - lse_t): a second-order quantity (both log-softmaxes sit near
    // -log H), so the log-sum-exp difference is taken in double.  One lane per column keeps it.
    if (ph == 0) {
      const double dl = ((double)ms - (double)mt) + log((double)sum_s / (double)sum_t);
      kl_total += (double)wsum / (double)sum_s - dl;
    }

    // ---- C: d loss / d mask rows: sum over a segment of T_h (p_h - t_h) = (T/mask) sum xt_h (p_h - t_h); the row phases
    // of a column are combined by shuffles and one lane issues one red.global per (column, segment, channel)
    if (want_grad) {
      const float rs = __fdividef(1.f, sum_s), rt = __fdividef(1.f, sum_t);
      for (int k = 0; k < nsweep; ++k) {
        int h, h1;
        seg_range(k, h, h1);
        float acc0 = 0.f, acc1 = 0.f;
        for (; h + kKlPhases < h1; h += 2 * kKlPhases) {
          const float b0 = xt_c[h * kRow], a0 = xs_c[h * kRow], b1 = xt_c[(h + kKlPhases) * kRow], a1 = xs_c[(h + kKlPhases) * kRow];
          const float p0 = fast_ex2(fmaf(b0, kLog2e, nmt)) * rt, t0 = fast_ex2(fmaf(a0, kLog2e, nms)) * rs;
          const float p1 = fast_ex2(fmaf(b1, kLog2e, nmt)) * rt, t1 = fast_ex2(fmaf(a1, kLog2e, nms)) * rs;
          acc0 = fmaf(b0, p0 - t0, acc0);
          acc1 = fmaf(b1, p1 - t1, acc1);
        }
        if (h < h1) {
          const float b0 = xt_c[h * kRow], a0 = xs_c[h * kRow];
          acc0 = fmaf(b0, fast_ex2(fmaf(b0, kLog2e, nmt)) * rt - fast_ex2(fmaf(a0, kLog2e, nms)) * rs, acc0);
        }
        float v = phase_sum(acc0 + acc1);
        if (ph == 0 && k < my_nseg) {
          const float m = seg_mask(k);
          float div = POW2 ? m : m / Temp;
          if (m == 0.f) {
            // the mask value underflowed to 0: the logits carry no trace of the teacher feature, re-read it
            v = 0.f;
            div = 1.f;
            for (int r = my_h0[k]; r < my_h1[k]; ++r) {
              const float pp = fast_ex2(fmaf(xt_c[r * kRow], kLog2e, nmt)) * rt;
              const float tt = fast_ex2(fmaf(xs_c[r * kRow], kLog2e, nms)) * rs;
              v = fmaf(ld_stream_f1(T + (int64_t)r * W), pp - tt, v);
            }
          }
          atomicAdd(prm.grad_rows + (int64_t)my_own[k] * C + c, __fdividef(gcoef * v, div));
        }
      }
    }
    __syncwarp();
  }
  // loss = scale * T^2 / H * sum over columns of sum_h t (log t - log p)
  double tot = block_sum(kl_total, red);
  if (threadIdx.x == 0 && tot != 0.0)
    atomicAdd(prm.loss, tot * (double)scale * (double)Temp * (double)Temp / (double)H);
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_dsgfd_kl_fwd_bwd(const DskdDsgfdKlArgs* a, void* stream) {
  DSKD_REQUIRE(a != nullptr, "dskd_dsgfd_kl_fwd_bwd: null args");
  DSKD_REQUIRE(a->num_levels > 0 && a->num_levels <= DSKD_MAX_LEVELS && a->N >= 0 && a->C > 0, "dsgfd_kl: bad sizes");
  DSKD_REQUIRE(a->temperature >= 1.f, "dsgfd_kl: T must be >= 1 (kd_loss.py:58)");
  const bool cell = a->d_cell_weight != nullptr;
  DSKD_REQUIRE(cell != (a->d_owner != nullptr), "dsgfd_kl: exactly one of d_owner / d_cell_weight must be set");
  DSKD_REQUIRE(a->d_loss != nullptr, "dsgfd_kl: d_loss is null");
  DSKD_REQUIRE(cell || a->num_pairs == 0 || a->d_rows != nullptr, "dsgfd_kl: d_rows is null");
  if (a->N == 0) return DSKD_OK;
  KlParams prm;
  prm.num_levels = a->num_levels;
  prm.N = a->N;
  prm.C = a->C;
  prm.temperature = a->temperature;
  prm.inv_temperature = 1.f / a->temperature;
  prm.cells_per_image = a->cells_per_image;
  prm.owner = a->d_owner;
  prm.rows = a->d_rows;
  prm.grad_rows = a->d_grad_rows;
  prm.cell_weight = a->d_cell_weight;
  prm.loss = a->d_loss;
  int64_t cells = 0;
  int max_h = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    DSKD_REQUIRE(a->levels[l].H > 0 && a->levels[l].W > 0 && a->levels[l].cell_offset == cells,
                 "dsgfd_kl: level %d is not densely packed", l);
    DSKD_REQUIRE(a->d_student[l] && a->d_teacher[l], "dsgfd_kl: null feature pointer at level %d", l);
    cells += (int64_t)a->levels[l].H * a->levels[l].W;
    max_h = std::max(max_h, a->levels[l].H);
  }
  DSKD_REQUIRE(cells == a->cells_per_image, "dsgfd_kl: cells_per_image mismatch");
  DSKD_REQUIRE(max_h <= kKlMaxH, "dsgfd_kl: H (%d) above the supported %d", max_h, kKlMaxH);
  // shared memory: owner strip (128 B per row) + two logit strips (256 B per row) per warp
  int warps = (int)std::min<int64_t>(kKlMaxWarps, ((int64_t)kKlSmemBudget - (int64_t)kl_header_bytes(max_h)) / (8ll * kKlCols * max_h));
  DSKD_REQUIRE(warps >= 1, "dsgfd_kl: H (%d) does not fit the shared-memory strips", max_h);
  warps = std::min(warps, std::max(1, a->C / kKlChan));
  prm.warps = warps;
  prm.max_h = max_h;
  const int ch_per_cta = warps * kKlChan;
  const int nchunks = (a->C + ch_per_cta - 1) / ch_per_cta;
  int blocks = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    prm.levels[l] = a->levels[l];
    prm.student[l] = a->d_student[l];
    prm.teacher[l] = a->d_teacher[l];
    prm.scale[l] = a->scale[l];
    prm.wtiles[l] = (a->levels[l].W + kKlCols - 1) / kKlCols;
    prm.block_start[l] = blocks;
    blocks += prm.wtiles[l] * nchunks * a->N;
  }
  prm.block_start[a->num_levels] = blocks;
  const size_t smem = kl_header_bytes(max_h) + 8ull * kKlCols * max_h * warps;
  int texp = 0;
  const bool pow2 = frexpf(a->temperature, &texp) == 0.5f;  // T = 2^k: the division by T is an exact scaling
  cudaStream_t st = as_stream(stream);
#define DSKD_KL_LAUNCH(CELLV, P2V)                                                                                  \
  do {                                                                                                              \
    DSKD_CUDA_OK(cudaFuncSetAttribute(dsgfd_kl_kernel<CELLV, P2V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    dsgfd_kl_kernel<CELLV, P2V><<<blocks, 32 * warps, smem, st>>>(prm);                                             \
  } while (0)
  if (cell) { if (pow2) DSKD_KL_LAUNCH(true, true); else DSKD_KL_LAUNCH(true, false); }
  else { if (pow2) DSKD_KL_LAUNCH(false, true); else DSKD_KL_LAUNCH(false, false); }
#undef DSKD_KL_LAUNCH
  DSKD_LAUNCH_OK("dsgfd_kl_kernel");
  return DSKD_OK;
}
