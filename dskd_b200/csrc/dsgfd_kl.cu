// DSG-FD with the shipped criterion: KnowledgeDistillationKLDivLoss(T=2, reduction='sum') applied to
// [C,H,W] maps, i.e. softmax over H (kd_loss.py:28-34 with dim=1 == H), target = student*mask
// (detached), pred = teacher*mask.  Reference: gfl_deformable_detr_head_il.py:707-718 + kd_loss.py:12-43.
//
// Cells outside every box have logit 0 for both softmaxes and need no feature bytes; strips without any owned cell
// contribute exactly 0 and are skipped.  No gradient reaches the student features (target is detached): the only
// gradient is d loss / d rows.  Algorithmic bytes: read S + read T = 8 B per element (45.51 MB per 800x1333 image).
#include <type_traits>

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


// ------------------------------------------------------------------------------------------------------------------
// Column kernel (box masks, H <= 400): the logits of a column never leave the register file.
// A warp owns 32 consecutive columns (w) of kColRows consecutive rows of one channel plane: every feature load is one
// coalesced 128 B row piece, each lane keeps its kColRows student / teacher values in registers through all three
// sweeps (logits + maxima, softmax sums, gradient), and the loops over rows are fully unrolled, so there is no address
// arithmetic, no shared-memory staging of features and no per-segment loop.  Taller levels are cut into `parts` row
// parts held by `parts` warps of the CTA; they exchange (local max, local sums) once per channel through shared memory
// and rescale.  What varies per row -- the owning box -- is turned ONCE per CTA into a byte offset into a small
// shared-memory table of the mask rows of the tile's boxes (entry 0 = the zero row for cells outside boxes), so a row
// costs one LDS for its mask value and cells outside boxes need no branch: their logit is feature * 0.
// Row parts without any box skip their loads and arithmetic (closed form), tiles without any box exit.
constexpr int kColRows = 25;        // rows per lane (100 / 50 / 25 rows of the COCO pyramid = 4 / 2 / 1 parts)
constexpr int kColBlk = 5;          // rows per skippable block
constexpr int kColBlocks = (kColRows + kColBlk - 1) / kColBlk;
constexpr int kColTableCap = 3072;  // floats of staged mask rows per CTA
constexpr int kColMaxPairs = kColTableCap - 3;
constexpr int kColChunk = 32;       // channels per CTA

struct KlColParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* student[DSKD_MAX_LEVELS];
  const float* teacher[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  int block_start[DSKD_MAX_LEVELS + 1];
  int wtiles[DSKD_MAX_LEVELS];
  int parts[DSKD_MAX_LEVELS];  // row parts (warps per column tile)
  int rpp[DSKD_MAX_LEVELS];    // rows per part
  int num_levels, N, C;
  int chunk;                   // channels per CTA
  float temperature, inv_temperature;
  int64_t cells_per_image;
  const int* owner;
  const float* cell_weight;    // per-cell mask weights instead of owners (CELL kernels)
  const float* rows;
  float* grad_rows;
  double* loss;
};

// bar.sync on a compile-time barrier id (a run-time id makes ptxas reserve all 16 barriers for every CTA)
template <int MAXG>
__device__ __forceinline__ void group_barrier(int group, int threads) {
#define DSKD_BAR_CASE(G)                                                               \
  case G:                                                                              \
    if (G < MAXG) asm volatile("bar.sync %0, %1;" ::"n"(G + 1), "r"(threads) : "memory"); \
    break;
  switch (group) {
    DSKD_BAR_CASE(0) DSKD_BAR_CASE(1) DSKD_BAR_CASE(2) DSKD_BAR_CASE(3)
    DSKD_BAR_CASE(4) DSKD_BAR_CASE(5) DSKD_BAR_CASE(6) DSKD_BAR_CASE(7)
    default: break;
  }
#undef DSKD_BAR_CASE
}

// Registers over occupancy: 3 CTAs of 4 warps per SM (168 registers, no spills of the row arrays) beat 4, 5 and 6 CTAs
// (128 / 96 / 80 registers, the row arrays partly in local memory): 0.33 / 0.39 / 0.40 / 0.42 ms on the COCO batch.
// CELL: per-cell mask weights (sg_out / fg_only) instead of box owners: the lane keeps the weights of its rows in
// registers, there is no mask table and no gradient.
template <int MAXW, bool POW2, bool CELL>
__global__ void __launch_bounds__(32 * MAXW, MAXW == 4 ? 3 : 1) dsgfd_kl_col_kernel(const __grid_constant__ KlColParams prm) {
  __shared__ float rows_s[kColTableCap];   // [channel of the sub-chunk][1 + owner - omin] mask values (/T when exact)
  __shared__ float ex_s[2][MAXW][5][32];   // per warp: local max_s, max_t, sum_s, sum_t, weighted sum (double-buffered)
  __shared__ double red[32];
  __shared__ int orange_s[2];
  __shared__ int zero_s[kColChunk];     // per staged channel: some mask value is exactly 0 (underflow)
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kPacked = (kColRows + 1) / 2;
  constexpr float kExcluded = -1e30f;      // logit of a row that does not exist (partial parts): e^x == 0, x - y == 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int lvl = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && (int)blockIdx.x >= prm.block_start[k]) lvl = k;
  const int H = prm.levels[lvl].H, W = prm.levels[lvl].W, C = prm.C;
  const int HW = H * W;
  const int nchunks = (C + prm.chunk - 1) / prm.chunk;
  int idx = blockIdx.x - prm.block_start[lvl];
  const int chunk = idx % nchunks;  // the channel chunks of one tile are neighbours: its owner lines stay in L2
  idx /= nchunks;
  const int wt = idx % prm.wtiles[lvl];
  const int img = idx / prm.wtiles[lvl];
  const int parts = prm.parts[lvl], rpp = prm.rpp[lvl];
  const int groups = MAXW / parts;            // warps of one group hold the row parts of the same channel
  const int part = warp % parts, group = warp / parts;
  const bool active = group < groups;
  const int w = wt * 32 + lane;
  const bool col_ok = w < W;
  const int wc = min(w, W - 1);               // lanes past the last column load it again and are ignored
  const int row0 = part * rpp;
  const int nrows = active ? min(rpp, H - row0) : 0;
  const float Temp = prm.temperature;
  const float scale = prm.scale[lvl];
  const unsigned uW = (unsigned)W;

  // ---- once per CTA: the owners of this lane's rows -> byte offsets into the staged mask table, flush rows
  if (tid == 0) { orange_s[0] = 0x7fffffff; orange_s[1] = -1; }
  __syncthreads();
  const int64_t cell0 = (int64_t)img * prm.cells_per_image + prm.levels[lvl].cell_offset + (int64_t)row0 * W + wc;
  unsigned moffp[kPacked];  // two 16-bit byte offsets per register
  float mw[kColRows];       // CELL: the mask weights of this lane's rows (/T when exact)
  unsigned fb = 0;          // rows after which the accumulated gradient of a run is flushed
  bool part_any;
  unsigned rowany;          // rows with a box in some column of the warp
  int omin, omax;
  {
    int own[kColRows];
    int lo = 0x7fffffff, hi = -1;
#pragma unroll
    for (int r = 0; r < kColRows; ++r) {
      int o = -1;
      if (CELL) {
        const float wv = (col_ok && r < nrows) ? __ldg(prm.cell_weight + cell0 + r * uW) : 0.f;
        mw[r] = POW2 ? wv * prm.inv_temperature : wv;
        o = wv != 0.f ? 0 : -1;
      } else if (col_ok && r < nrows) {
        o = __ldg(prm.owner + cell0 + r * uW);
      }
      own[r] = o;
      if (o >= 0) { lo = min(lo, o); hi = max(hi, o); }
    }
    lo = __reduce_min_sync(kFull, lo);
    hi = __reduce_max_sync(kFull, hi);
    part_any = hi >= 0;  // some cell of this warp's rows lies inside a box
    {
      unsigned ob = 0;
#pragma unroll
      for (int r = 0; r < kColRows; ++r) ob |= own[r] >= 0 ? 1u << r : 0u;
      rowany = __reduce_or_sync(kFull, ob);
    }
    if (lane == 0 && part_any) { atomicMin(&orange_s[0], lo); atomicMax(&orange_s[1], hi); }
    __syncthreads();
    omin = orange_s[0];
    omax = orange_s[1];
    if (omax < 0) return;  // no box touches this tile: every column's KL is exactly 0
#pragma unroll
    for (int r = 0; r < kColRows; ++r) {
      const int nx = (r + 1 < kColRows) ? own[r + 1] : -1;
      if (!CELL && own[r] >= 0 && nx != own[r]) fb |= 1u << r;
    }
#pragma unroll
    for (int k = 0; k < kPacked; ++k) {
      const int o0 = own[2 * k], o1 = (2 * k + 1 < kColRows) ? own[2 * k + 1] : -1;
      const unsigned b0 = o0 >= 0 ? (unsigned)(o0 - omin + 1) * 4u : 0u;
      const unsigned b1 = o1 >= 0 ? (unsigned)(o1 - omin + 1) * 4u : 0u;
      moffp[k] = b0 | (b1 << 16);
    }
  }
  const unsigned anyfb = __reduce_or_sync(kFull, fb);
  // blocks of kColBlk rows without any box in the warp's 32 columns are skipped (closed form: logit 0)
  unsigned blk_on = 0;
  int nskip_i = 0;
#pragma unroll
  for (int b = 0; b < kColBlocks; ++b) {
    const unsigned bm = ((1u << kColBlk) - 1u) << (b * kColBlk);
    if (rowany & bm) blk_on |= 1u << b;
    else nskip_i += max(0, min(nrows - b * kColBlk, kColBlk));
  }
  const float nskip = (float)nskip_i;
  const int nown = omax - omin + 1;
  const int nstride = (nown + 1) | 1;  // odd: the staging writes of one owner spread over the banks
  const int chs = min(prm.chunk, kColTableCap / nstride);  // channels staged at a time (>= 1: host bounds num_pairs)

  const float kLog2e = 1.4426950408889634f;
  const float gcoef = scale * Temp / (float)H;  // d loss / d pred = scale * (T/H) * (p - t)
  const bool want_grad = !CELL && prm.grad_rows != nullptr;
  const int c_begin = chunk * prm.chunk, c_end = min(C, c_begin + prm.chunk);
  double kl_total = 0.0;
  int par = 0;
#define DSKD_MOFF(r) (((r) & 1) ? (moffp[(r) >> 1] >> 16) : (moffp[(r) >> 1] & 0xffffu))
#define DSKD_MTAB(off) (*reinterpret_cast<const float*>(reinterpret_cast<const char*>(mtab) + (off)))

  // one channel of this warp's rows; FULL: all kColRows rows exist
  auto channel = [&](auto full_tag, const int c, const float* mtab, const bool zero_any) {
    constexpr bool FULL = decltype(full_tag)::value;
    float s[kColRows], t[kColRows];
    float ml_s = 0.f, ml_t = 0.f, ss = (float)nrows, st = (float)nrows, ws = 0.f;
    const int64_t plane = ((int64_t)img * C + c) * HW + (int64_t)row0 * W + wc;
#define DSKD_FOR_ROWS_ON(body)                                   \
  _Pragma("unroll") for (int b_ = 0; b_ < kColBlocks; ++b_) {    \
    if (blk_on & (1u << b_)) {                                   \
      _Pragma("unroll") for (int k_ = 0; k_ < kColBlk; ++k_) {   \
        const int r = b_ * kColBlk + k_;                         \
        if (r < kColRows) { body }                               \
      }                                                          \
    }                                                            \
  }
    if (part_any) {
      const float* __restrict__ Sp = prm.student[lvl] + plane;
      const float* __restrict__ Tp = prm.teacher[lvl] + plane;
      const uint64_t pitch = (uint64_t)uW * 4u;
#pragma unroll
      for (int b_ = 0; b_ < kColBlocks; ++b_) {
        if (blk_on & (1u << b_)) {
          // byte addresses advanced by one row pitch: two 64-bit adds per row instead of re-deriving base + r * W
          const unsigned r0 = FULL ? (unsigned)(b_ * kColBlk) : (unsigned)min(b_ * kColBlk, nrows - 1);
          uint64_t sa = reinterpret_cast<uint64_t>(Sp + r0 * uW), ta = reinterpret_cast<uint64_t>(Tp + r0 * uW);
#pragma unroll
          for (int k_ = 0; k_ < kColBlk; ++k_) {
            const int r = b_ * kColBlk + k_;
            if (r < kColRows) {
              s[r] = ld_stream_f1(reinterpret_cast<const float*>(sa));
              t[r] = ld_stream_f1(reinterpret_cast<const float*>(ta));
              if (FULL || r + 1 < nrows) { sa += pitch; ta += pitch; }  // rows past the part load its last row again
            }
          }
        }
      }
      // A: logits, maxima of this part (skipped rows: logit 0)
      ml_s = nskip_i > 0 ? 0.f : -INFINITY;
      ml_t = ml_s;
      DSKD_FOR_ROWS_ON(
        const float m = CELL ? mw[r] : DSKD_MTAB(DSKD_MOFF(r));
        float x = s[r] * m; float y = t[r] * m;
        if (!POW2) { x = __fdiv_rn(x, Temp); y = __fdiv_rn(y, Temp); }
        if (!FULL) { x = r < nrows ? x : kExcluded; y = r < nrows ? y : kExcluded; }
        s[r] = x;
        t[r] = y;
        ml_s = fmaxf(ml_s, x);
        ml_t = fmaxf(ml_t, y);
      )
      // B: sums against the part's own maxima
      const float nms = -ml_s * kLog2e, nmt = -ml_t * kLog2e;
      float ss0 = 0.f, ss1 = 0.f, st0 = 0.f, st1 = 0.f, ws0 = 0.f, ws1 = 0.f;
      DSKD_FOR_ROWS_ON(
        const float e = fast_ex2(fmaf(s[r], kLog2e, nms));
        const float f = fast_ex2(fmaf(t[r], kLog2e, nmt));
        if (r & 1) { ss1 += e; st1 += f; ws1 = fmaf(e, s[r] - t[r], ws1); }
        else { ss0 += e; st0 += f; ws0 = fmaf(e, s[r] - t[r], ws0); }
        s[r] = e;  // the student logit is not needed again: keep e^(xs - local max) for the gradient sweep
      )
      ss = fmaf(nskip, fast_ex2(nms), ss0 + ss1);
      st = fmaf(nskip, fast_ex2(nmt), st0 + st1);
      ws = ws0 + ws1;
    }
    // combine the row parts of the column: sum_p e^(max_p - max) * sum_p
    float Ms = ml_s, Mt = ml_t, sum_s = ss, sum_t = st, wsum = ws;
    if (parts > 1) {
      float(*ex)[5][32] = ex_s[par];
      ex[warp][0][lane] = ml_s;
      ex[warp][1][lane] = ml_t;
      ex[warp][2][lane] = ss;
      ex[warp][3][lane] = st;
      ex[warp][4][lane] = ws;
      group_barrier<MAXW / 2>(group, 32 * parts);
      const int w0 = group * parts;
#pragma unroll 1
      for (int p = 0; p < parts; ++p) {
        Ms = fmaxf(Ms, ex[w0 + p][0][lane]);
        Mt = fmaxf(Mt, ex[w0 + p][1][lane]);
      }
      sum_s = 0.f;
      sum_t = 0.f;
      wsum = 0.f;
#pragma unroll 1
      for (int p = 0; p < parts; ++p) {
        const float fs = fast_ex2((ex[w0 + p][0][lane] - Ms) * kLog2e);
        const float ft = fast_ex2((ex[w0 + p][1][lane] - Mt) * kLog2e);
        sum_s = fmaf(ex[w0 + p][2][lane], fs, sum_s);
        sum_t = fmaf(ex[w0 + p][3][lane], ft, sum_t);
        wsum = fmaf(ex[w0 + p][4][lane], fs, wsum);
      }
      par ^= 1;
    }
    // KL of the column = sum_h t_h (xs - xt) - (lse_s - lse_t): second order, so the difference is taken in double
    if (part == 0 && col_ok) {
      const double dl = ((double)Ms - (double)Mt) + log((double)sum_s / (double)sum_t);
      kl_total += (double)wsum / (double)sum_s - dl;
    }
    // C: d loss / d mask rows: sum over a run of rows with one owner of T_h (p_h - t_h) = (T/mask) sum xt_h (p_h - t_h)
    if (want_grad && part_any) {
      const float rs = __fdividef(1.f, sum_s), rt = __fdividef(1.f, sum_t);
      const float nms = -Ms * kLog2e, nmt = -Mt * kLog2e;
      // s[r] holds e^(xs - local max) since sweep B: t_h = s[r] * e^(local max - max) / sum_s
      const float cs = fast_ex2(fmaf(ml_s, kLog2e, nms)) * rs;
      float* __restrict__ grow = prm.grad_rows + (int64_t)(omin - 1) * C + c;
      // first every row's term T_h (p_h - t_h) * mask / T, branch-free (the exponentials of all rows overlap) ...
      DSKD_FOR_ROWS_ON(
        const float pt = fast_ex2(fmaf(t[r], kLog2e, nmt)) * rt;
        s[r] = t[r] * fmaf(-s[r], cs, pt);
      )
      // ... then the running sum down the rows; a run of one owner ends where the owner of the next row differs
      float acc = 0.f;
      DSKD_FOR_ROWS_ON(
        acc += s[r];
        if (anyfb & (1u << r)) {  // warp-uniform: most rows end no run in any lane
          unsigned fbv;           // (opaque copy: keeps the uniform test from being folded into the per-lane one)
          asm volatile("mov.u32 %0, %1;" : "=r"(fbv) : "r"(fb));
          if (fbv & (1u << r)) {
            const unsigned off = DSKD_MOFF(r);
            const float m = DSKD_MTAB(off);
            // m == 0 (underflow): the run is handled below from the raw teacher feature
            if (m != 0.f) atomicAdd(grow + (off >> 2) * (unsigned)C, __fdividef(gcoef * acc, POW2 ? m : m / Temp));
            acc = 0.f;
          }
        }
      )
      if (zero_any) {
        // some mask value of this channel underflowed to 0: the logits of such a run are all 0, p - t is one
        // constant, and the registers carry no trace of the teacher feature -- walk the rows again
        const float d0 = fast_ex2(nmt) * rt - fast_ex2(nms) * rs;
        const float* __restrict__ Tp = prm.teacher[lvl] + plane;
        float tsum = 0.f;
        for (int r = 0; r < nrows; ++r) {
          const int o = col_ok ? __ldg(prm.owner + cell0 + r * uW) : -1;
          const int nx = (col_ok && r + 1 < nrows) ? __ldg(prm.owner + cell0 + (r + 1) * uW) : -1;
          const bool zero = o >= 0 && mtab[o - omin + 1] == 0.f;
          if (zero) tsum += ld_stream_f1(Tp + r * uW);
          if (nx != o) {
            if (zero) atomicAdd(prm.grad_rows + (int64_t)o * C + c, gcoef * tsum * d0);
            tsum = 0.f;
          }
        }
      }
    }
  };

  for (int sc = c_begin; sc < c_end; sc += chs) {
    const int nch = min(chs, c_end - sc);
    if (!CELL) {
      __syncthreads();  // the readers of the previous sub-chunk are done
      for (int cc = tid; cc < nch; cc += 32 * MAXW) {
        rows_s[cc * nstride] = 0.f;  // cells outside boxes
        zero_s[cc] = 0;
      }
      __syncthreads();
      for (int i = tid; i < nown * nch; i += 32 * MAXW) {
        const int o = i / nch, cc = i - o * nch;
        float m = __ldg(prm.rows + (int64_t)(omin + o) * C + sc + cc);
        if (POW2) m *= prm.inv_temperature;
        rows_s[cc * nstride + o + 1] = m;
        if (m == 0.f) zero_s[cc] = 1;
      }
      __syncthreads();
    }
    if (active) {
      for (int cc = group; cc < nch; cc += groups) {
        const bool zero_any = !CELL && zero_s[cc] != 0;
        if (nrows == kColRows) channel(std::true_type{}, sc + cc, rows_s + cc * nstride, zero_any);
        else channel(std::false_type{}, sc + cc, rows_s + cc * nstride, zero_any);
      }
    }
  }
#undef DSKD_MOFF
#undef DSKD_MTAB
#undef DSKD_FOR_ROWS_ON
  // loss = scale * T^2 / H * sum over columns of sum_h t (log t - log p)
  double tot = block_sum(kl_total, red);
  if (tid == 0 && tot != 0.0) atomicAdd(prm.loss, tot * (double)scale * (double)Temp * (double)Temp / (double)H);
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
  cudaStream_t st = as_stream(stream);
  int texp = 0;
  const bool pow2 = frexpf(a->temperature, &texp) == 0.5f;  // T = 2^k: the division by T is an exact scaling

  // box masks on levels of at most 16 x kColRows rows: the register-resident column kernel
  if (max_h <= 16 * kColRows && (cell || a->num_pairs <= kColMaxPairs)) {
    KlColParams cp;
    cp.num_levels = a->num_levels;
    cp.N = a->N;
    cp.C = a->C;
    cp.temperature = a->temperature;
    cp.inv_temperature = 1.f / a->temperature;
    cp.cells_per_image = a->cells_per_image;
    cp.owner = a->d_owner;
    cp.cell_weight = a->d_cell_weight;
    cp.rows = a->d_rows;
    cp.grad_rows = a->d_grad_rows;
    cp.loss = a->d_loss;
    cp.chunk = kColChunk;
    const int max_parts = (max_h + kColRows - 1) / kColRows;
    const int maxw = max_parts <= 4 ? 4 : (max_parts <= 8 ? 8 : 16);
    const int nchunks = (a->C + cp.chunk - 1) / cp.chunk;
    int blocks = 0;
    for (int l = 0; l < a->num_levels; ++l) {
      cp.levels[l] = a->levels[l];
      cp.student[l] = a->d_student[l];
      cp.teacher[l] = a->d_teacher[l];
      cp.scale[l] = a->scale[l];
      cp.wtiles[l] = (a->levels[l].W + 31) / 32;
      cp.parts[l] = (a->levels[l].H + kColRows - 1) / kColRows;
      cp.rpp[l] = (a->levels[l].H + cp.parts[l] - 1) / cp.parts[l];
      cp.block_start[l] = blocks;
      blocks += cp.wtiles[l] * nchunks * a->N;
    }
    cp.block_start[a->num_levels] = blocks;
#define DSKD_COL_LAUNCH(MW)                                                                        \
  do {                                                                                             \
    if (cell) {                                                                                    \
      if (pow2) dsgfd_kl_col_kernel<MW, true, true><<<blocks, 32 * MW, 0, st>>>(cp);               \
      else dsgfd_kl_col_kernel<MW, false, true><<<blocks, 32 * MW, 0, st>>>(cp);                   \
    } else {                                                                                       \
      if (pow2) dsgfd_kl_col_kernel<MW, true, false><<<blocks, 32 * MW, 0, st>>>(cp);              \
      else dsgfd_kl_col_kernel<MW, false, false><<<blocks, 32 * MW, 0, st>>>(cp);                  \
    }                                                                                              \
  } while (0)
    if (maxw == 4) DSKD_COL_LAUNCH(4);
    else if (maxw == 8) DSKD_COL_LAUNCH(8);
    else DSKD_COL_LAUNCH(16);
#undef DSKD_COL_LAUNCH
    DSKD_LAUNCH_OK("dsgfd_kl_col_kernel");
    return DSKD_OK;
  }
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
