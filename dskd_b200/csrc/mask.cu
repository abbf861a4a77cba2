// Mask construction for DSG-FD: per-pair channel rows (softmax of |hs_T - hs_S|), their backward,
// and the cell rasters (last-writer-wins owner map, binary / area cell weights).
// Reference: gfl_deformable_detr_head_il.py:685-706 (decode_v1), :742-756 (decode_v2),
// :883-914 (sg_out), :1107-1122 (fg_only); gfl_deformable_detr_head_il_fg_bk.py:548-566 (fg_bk).
#include "common.cuh"

namespace dskd {

// One CTA per (teacher, student) query pair.  C floats of dynamic shared memory.
template <int MODE>
__global__ void __launch_bounds__(128) mask_rows_kernel(const float* __restrict__ hs_t,
                                                        const float* __restrict__ hs_s,
                                                        const int64_t* __restrict__ id_soft,
                                                        const int64_t* __restrict__ id_pred, int C,
                                                        float* __restrict__ rows) {
  extern __shared__ float a[];
  __shared__ float red[32];
  __shared__ float bcast;
  const int p = blockIdx.x;
  const float* t = hs_t + id_soft[p] * (int64_t)C;
  const float* s = (MODE == DSKD_MASK_DECODE_V1) ? hs_s + id_pred[p] * (int64_t)C : nullptr;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = (MODE == DSKD_MASK_DECODE_V1) ? fabsf(t[c] - s[c]) : t[c];
    a[c] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float e = expf(a[c] - mx);
    a[c] = e;
    sum += e;
  }
  __syncthreads();
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) bcast = sum;
  __syncthreads();
  sum = bcast;
  for (int c = threadIdx.x; c < C; c += blockDim.x) rows[(int64_t)p * C + c] = a[c] / sum;
}

// d a = A * (g - <g, A>),  d hs_S = -sign(hs_T - hs_S) * d a   (softmax, abs, and the minus of T - S).
__global__ void __launch_bounds__(128) mask_rows_bwd_kernel(const float* __restrict__ hs_t,
                                                            const float* __restrict__ hs_s,
                                                            const int64_t* __restrict__ id_soft,
                                                            const int64_t* __restrict__ id_pred,
                                                            const float* __restrict__ rows,
                                                            const float* __restrict__ grad_rows, int C,
                                                            float* __restrict__ grad_hs_s) {
  __shared__ float red[32];
  __shared__ float bcast;
  const int p = blockIdx.x;
  const float* A = rows + (int64_t)p * C;
  const float* g = grad_rows + (int64_t)p * C;
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dot += A[c] * g[c];
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) bcast = dot;
  __syncthreads();
  dot = bcast;
  const int64_t qs = id_pred[p];
  const float* t = hs_t + id_soft[p] * (int64_t)C;
  const float* s = hs_s + qs * (int64_t)C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float da = A[c] * (g[c] - dot);
    const float delta = t[c] - s[c];
    const float sgn = (delta > 0.f) ? 1.f : ((delta < 0.f) ? -1.f : 0.f);
    atomicAdd(grad_hs_s + qs * (int64_t)C + c, -sgn * da);
  }
}

// Row-mask finish, one CTA per pair p:  (mse) loss += sum_c rows^2 * energy, g = 2 * rows * energy;
// (kl) g = grad_rows as accumulated by the KL kernel.  Then, if requested, the softmax / abs backward of
// mask_rows_bwd_kernel into grad_hs_s.  Replaces the single-CTA dsgfd_mse_finish + mask_rows_bwd pair.
template <bool MSE>
__global__ void __launch_bounds__(128) rows_finish_kernel(const float* __restrict__ hs_t, const float* __restrict__ hs_s,
                                                          const int64_t* __restrict__ id_soft,
                                                          const int64_t* __restrict__ id_pred,
                                                          const float* __restrict__ rows, const float* __restrict__ eg,
                                                          int C, double* __restrict__ loss_acc,
                                                          float* __restrict__ grad_hs_s) {
  __shared__ float red[32];
  __shared__ double dred[32];
  __shared__ float bcast;
  const int p = blockIdx.x;
  const float* A = rows + (int64_t)p * C;
  const float* E = eg + (int64_t)p * C;
  float dot = 0.f;
  double part = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = A[c], e = E[c];
    const float g = MSE ? 2.f * a * e : e;
    if (MSE) part += (double)a * (double)a * (double)e;
    dot += a * g;
  }
  if (MSE) {
    part = block_sum(part, dred);
    if (threadIdx.x == 0 && part != 0.0) atomicAdd(loss_acc, part);
  }
  if (grad_hs_s == nullptr) return;
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) bcast = dot;
  __syncthreads();
  dot = bcast;
  const int64_t qs = id_pred[p];
  const float* t = hs_t + id_soft[p] * (int64_t)C;
  const float* sv = hs_s + qs * (int64_t)C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = A[c], e = E[c];
    const float g = MSE ? 2.f * a * e : e;
    const float da = a * (g - dot);
    const float delta = t[c] - sv[c];
    const float sgn = (delta > 0.f) ? 1.f : ((delta < 0.f) ? -1.f : 0.f);
    atomicAdd(grad_hs_s + qs * (int64_t)C + c, -sgn * da);
  }
}

struct RasterParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  int num_levels;
};

struct Rect {
  int wmin, wmax, hmin, hmax;  // as computed by the reference (before inclusive/exclusive use)
  float area;
};

// head_il.py:688-696: x / img_w * W in fp32 (IEEE division, no contraction), floor / ceil, .int()
__device__ __forceinline__ Rect make_rect(const float* b, int img_h, int img_w, int H, int W, bool swap_scale) {
  const float sx = swap_scale ? (float)H : (float)W;
  const float sy = swap_scale ? (float)W : (float)H;
  Rect r;
  r.wmin = (int)floorf(__fmul_rn(__fdiv_rn(b[0], (float)img_w), sx));
  r.wmax = (int)ceilf(__fmul_rn(__fdiv_rn(b[2], (float)img_w), sx));
  r.hmin = (int)floorf(__fmul_rn(__fdiv_rn(b[1], (float)img_h), sy));
  r.hmax = (int)ceilf(__fmul_rn(__fdiv_rn(b[3], (float)img_h), sy));
  // area = 1.0 / (hmax + 1 - hmin) / (wmax + 1 - wmin)   (head_il.py:1113-1114)
  r.area = __fdiv_rn(__fdiv_rn(1.0f, (float)(r.hmax + 1 - r.hmin)), (float)(r.wmax + 1 - r.wmin));
  return r;
}

// grid (cell tiles, N); dynamic smem: (boxes_i + gts_i) * num_levels Rects.
template <int MODE>
__global__ void __launch_bounds__(256) raster_kernel(const float* __restrict__ boxes,
                                                     const int* __restrict__ box_start,
                                                     const float* __restrict__ gt_boxes,
                                                     const int* __restrict__ gt_start,
                                                     const int* __restrict__ img_hw, RasterParams prm,
                                                     int64_t cells_per_image, void* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Rect* rects = reinterpret_cast<Rect*>(smem_raw);
  const int i = blockIdx.y;
  const int b0 = box_start[i], nb = box_start[i + 1] - b0;
  const int g0 = (MODE == DSKD_RASTER_BINARY_INCL) ? gt_start[i] : 0;
  const int ng = (MODE == DSKD_RASTER_BINARY_INCL) ? gt_start[i + 1] - g0 : 0;
  const int img_h = img_hw[2 * i], img_w = img_hw[2 * i + 1];
  const int per_level = nb + ng;
  for (int k = threadIdx.x; k < per_level * prm.num_levels; k += blockDim.x) {
    const int l = k / per_level, j = k - l * per_level;
    const float* b = (j < nb) ? boxes + (int64_t)(b0 + j) * 4 : gt_boxes + (int64_t)(g0 + j - nb) * 4;
    rects[k] = make_rect(b, img_h, img_w, prm.levels[l].H, prm.levels[l].W, MODE == DSKD_RASTER_AREA_FGBK);
  }
  __syncthreads();
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= cells_per_image) return;
  int l = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && cell >= prm.levels[k].cell_offset) l = k;
  const int W = prm.levels[l].W;
  const int local = (int)(cell - prm.levels[l].cell_offset);
  const int h = local / W, w = local - h * W;
  const Rect* R = rects + l * per_level;
  if (MODE == DSKD_RASTER_OWNER_EXCL) {
    int owner = -1;
    for (int j = nb - 1; j >= 0; --j) {
      const Rect r = R[j];
      if (h >= r.hmin && h < r.hmax && w >= r.wmin && w < r.wmax) { owner = b0 + j; break; }
    }
    reinterpret_cast<int*>(out)[(int64_t)i * cells_per_image + cell] = owner;
  } else if (MODE == DSKD_RASTER_BINARY_INCL) {
    float m = 0.f;
    for (int j = 0; j < nb; ++j) {
      const Rect r = R[j];
      if (h >= r.hmin && h <= r.hmax && w >= r.wmin && w <= r.wmax) { m = 1.f; break; }
    }
    for (int j = nb; j < per_level && m != 0.f; ++j) {
      const Rect r = R[j];
      if (h >= r.hmin && h <= r.hmax && w >= r.wmin && w <= r.wmax) m = 0.f;
    }
    reinterpret_cast<float*>(out)[(int64_t)i * cells_per_image + cell] = m;  // sqrt(0/1) == itself
  } else {
    float m = 0.f;
    for (int j = 0; j < nb; ++j) {
      const Rect r = R[j];
      if (h >= r.hmin && h <= r.hmax && w >= r.wmin && w <= r.wmax) m = fmaxf(m, r.area);
    }
    reinterpret_cast<float*>(out)[(int64_t)i * cells_per_image + cell] = sqrtf(m);
  }
}


// ------------------------------------------------------------------------------------------------
// decode_v1 in three launches instead of eight.  The step is latency-bound outside its streaming kernel: every
// dependent launch of a captured graph costs 2-3 us, and matched ids -> mask rows -> raster (+ two memsets) and
// row finish -> loss conversion were eight of them.
//
// prepare_v1_kernel, one 1-D grid with three CTA roles:
//   [0, num_pairs)              pair p: finds ITS matched student query -- the p-th query in ascending order whose
//                               label is a previous-task label (head_il.py:1453-1455, :672) -- by a block-wide rank
//                               search over the labels, writes ids[p], the mask row softmax_c|hs_T - hs_S|, and clears
//                               its row of the energy table
//   [.., + raster_ctas)         owner raster of one 256-cell tile of one image (as raster_kernel<OWNER_EXCL>)
//   [.., + zero_ctas)           clears grad_hs_student, the loss accumulator and the finish counter
// ------------------------------------------------------------------------------------------------
constexpr int kPrepTile = 256;   // cells per raster CTA of prepare_v1_kernel (1024 = 4 per thread measured slower: fewer warps to hide latency)
struct PrepareV1Params {
  RasterParams raster;
  int raster_tiles, raster_ctas, num_pairs, zero_ctas;
  int N, C, num_rows, num_classes;
  int64_t cells_per_image;
  const float* boxes; const int* box_start; const int* img_hw;
  const float* hs_t; const float* hs_s; const int64_t* keepid; const int64_t* labels; const uint8_t* prev_mask;
  int* owner; float* rows; float* energy; int64_t* ids; int* matched_count;
  float* grad_hs; int64_t grad_hs_floats; double* acc; unsigned* counter;
};

__global__ void __launch_bounds__(256, 8) prepare_v1_kernel(const __grid_constant__ PrepareV1Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[32];
  __shared__ int s_scan[8];
  __shared__ int s_id, s_total;
  __shared__ float bcast;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // block order = start order: the pair CTAs carry the longest chain of dependent loads (labels -> prev_mask -> ids ->
  // embeddings -> softmax) and go first, the short raster CTAs fill in behind them, the clears come last
  int b = (int)blockIdx.x - p.num_pairs;
  if (b >= 0 && b < p.raster_ctas) {
    // ---- owner raster: last box (highest pair index) whose half-open rectangle holds the cell (head_il.py:706)
    Rect* rects = reinterpret_cast<Rect*>(smem_raw);
    const int i = b / p.raster_tiles, tile = b - i * p.raster_tiles;
    const int b0 = p.box_start[i], nb = p.box_start[i + 1] - b0;
    const int img_h = p.img_hw[2 * i], img_w = p.img_hw[2 * i + 1];
    for (int k = tid; k < nb * p.raster.num_levels; k += blockDim.x) {
      const int l = k / nb, j = k - l * nb;
      rects[k] = make_rect(p.boxes + (int64_t)(b0 + j) * 4, img_h, img_w, p.raster.levels[l].H, p.raster.levels[l].W, false);
    }
    __syncthreads();
    const int cells = (int)p.cells_per_image;  // < 2^31 (checked at launch): 32-bit cell arithmetic
    const int cell0 = tile * kPrepTile;
    const int cell_last = min(cell0 + kPrepTile, cells) - 1;
    auto level_of = [&](int c) {
      int l = 0;
#pragma unroll
      for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
        if (k < p.raster.num_levels && c >= (int)p.raster.levels[k].cell_offset) l = k;
      return l;
    };
    // A tile is kPrepTile consecutive cells -- a few rows of one level.  Only the boxes that reach those rows can own any
    // of its cells: one ballot per warp keeps them as a bit mask and the per-cell search walks the set bits from the top
    // (the kernel was issue-bound on the 40-box loop of every cell).  Tiles that straddle two levels take every box.
    __shared__ unsigned cand[8];
    const int l_first = level_of(cell0), l_last = level_of(cell_last);
    {
      bool keep = false;
      if (tid < nb) {
        keep = true;
        if (l_first == l_last) {
          const int off = (int)p.raster.levels[l_first].cell_offset;
          const int Wl = p.raster.levels[l_first].W;
          const int h0 = (cell0 - off) / Wl, h1 = (cell_last - off) / Wl;
          const Rect r = rects[l_first * nb + tid];
          keep = r.hmin <= h1 && r.hmax > h0 && r.wmin < r.wmax;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) cand[warp] = m;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPrepTile / 256; ++u) {
      const int cell = cell0 + u * 256 + tid;
      if (cell > cell_last) break;
      const int l = (l_first == l_last) ? l_first : level_of(cell);
      const int W = p.raster.levels[l].W;
      const int local = cell - (int)p.raster.levels[l].cell_offset;
      const int h = local / W, w = local - h * W;
      const Rect* R = rects + l * nb;
      int owner = -1;
      if (nb <= 256) {
        for (int wv = (nb - 1) >> 5; wv >= 0 && owner < 0; --wv) {
          unsigned m = cand[wv];
          while (m) {
            const int j = (wv << 5) + 31 - __clz(m);
            m &= ~(1u << (j & 31));
            const Rect r = R[j];
            if (h >= r.hmin && h < r.hmax && w >= r.wmin && w < r.wmax) { owner = b0 + j; break; }
          }
        }
      } else {
        for (int j = nb - 1; j >= 0; --j) {
          const Rect r = R[j];
          if (h >= r.hmin && h < r.hmax && w >= r.wmin && w < r.wmax) { owner = b0 + j; break; }
        }
      }
      p.owner[(int64_t)i * cells + cell] = owner;
    }
    return;
  }
  if (b < 0) {
    b += p.num_pairs;
    // ---- pair b: rank search for the b-th previous-labelled query.  The labels are taken 256 at a time in query order
    // (thread = query, coalesced); all chunks' loads are in flight together (a thread walking its own contiguous slice
    // paid two dependent L2 round trips per label: 10 of this kernel's 15 us).  Per chunk and warp the ballot of the
    // hits and its population go to shared memory; warp 0 then scans the counts and picks the b-th set bit.
    const int n = p.num_rows;
    const int chunks = (n + 255) >> 8;
    int* cnt = reinterpret_cast<int*>(smem_raw + (size_t)p.C * sizeof(float));  // [chunks][8] hits per (chunk, warp)
    unsigned* bal = reinterpret_cast<unsigned*>(cnt + chunks * 8);               // [chunks][8] their ballots
    constexpr int kBatch = 10;  // chunks whose label loads are in flight together (registers: 2 per chunk)
    for (int k0 = 0; k0 < chunks; k0 += kBatch) {
      int64_t lab[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int q = ((k0 + u) << 8) + tid;
        lab[u] = (k0 + u < chunks && q < n) ? __ldg(p.labels + q) : -1;
      }
      unsigned hits = 0;
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
        if (lab[u] >= 0 && lab[u] < p.num_classes && __ldg(p.prev_mask + lab[u]) != 0) hits |= 1u << u;
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned m = __ballot_sync(0xffffffffu, (hits >> u) & 1u);
        if (lane == 0 && k0 + u < chunks) { cnt[(k0 + u) * 8 + warp] = __popc(m); bal[(k0 + u) * 8 + warp] = m; }
      }
    }
    if (tid == 0) s_id = 0;
    __syncthreads();
    if (warp == 0) {
      // entries in (chunk, warp) order == query order; lane takes a contiguous run of them
      const int entries = chunks * 8, per = (entries + 31) >> 5;
      const int lo = min(entries, lane * per), hi = min(entries, lo + per);
      int c = 0;
      for (int e = lo; e < hi; ++e) c += cnt[e];
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int before = incl - c;
      if (b >= before && b < incl) {
        for (int e = lo; e < hi; ++e) {
          if (b < before + cnt[e]) {
            s_id = (e << 5) + (int)__fns(bal[e], 0, b - before + 1);  // e * 32 = first query of this (chunk, warp)
            break;
          }
          before += cnt[e];
        }
      }
      if (lane == 31) s_total = incl;
    }
    __syncthreads();
    // fewer than b + 1 queries carry a previous label: the reference raises IndexError at :705.  Without a host sync
    // the failure is made loud on the device instead: the pair's mask row and the step's loss become NaN.
    const bool unmatched = b >= s_total;
    const int id_pred = s_id;
    if (tid == 0) {
      p.ids[b] = id_pred;
      if (b == 0) {
        reinterpret_cast<int*>(p.counter)[1] = s_total;  // read back by the epilogue of the step
        if (p.matched_count) *p.matched_count = s_total;
      }
    }
    // ---- mask row b = softmax_c |hs_T[keepid[b]] - hs_S[id_pred]| (head_il.py:705-706); energy row cleared
    float* a = reinterpret_cast<float*>(smem_raw);
    const int C = p.C;
    const float* t = p.hs_t + p.keepid[b] * (int64_t)C;
    const float* sv = p.hs_s + (int64_t)id_pred * C;
    float mx = -INFINITY;
    for (int ch = tid; ch < C; ch += blockDim.x) {
      const float v = fabsf(t[ch] - sv[ch]);
      a[ch] = v;
      mx = fmaxf(mx, v);
      p.energy[(int64_t)b * C + ch] = 0.f;
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
      float m = red[0];
      for (int wv = 1; wv < ((int)blockDim.x >> 5); ++wv) m = fmaxf(m, red[wv]);
      bcast = m;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int ch = tid; ch < C; ch += blockDim.x) {
      const float e = expf(a[ch] - mx);
      a[ch] = e;
      sum += e;
    }
    __syncthreads();
    sum = block_sum(sum, red);
    if (tid == 0) bcast = sum;
    __syncthreads();
    sum = bcast;
    for (int ch = tid; ch < C; ch += blockDim.x) p.rows[(int64_t)b * C + ch] = unmatched ? __int_as_float(0x7fc00000) : a[ch] / sum;
    return;
  }
  b -= p.raster_ctas;
  // ---- clears
  if (b == 0 && tid == 0) { *p.acc = 0.0; *p.counter = 0u; }
  if (p.grad_hs != nullptr) {
    const int64_t n4 = p.grad_hs_floats >> 2;
    float4* g4 = reinterpret_cast<float4*>(p.grad_hs);
    for (int64_t i = (int64_t)b * blockDim.x + tid; i < n4; i += (int64_t)p.zero_ctas * blockDim.x)
      g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b == 0 && tid < (int)(p.grad_hs_floats & 3)) p.grad_hs[(n4 << 2) + tid] = 0.f;
  }
}

// rows_finish_kernel<true> whose last CTA also converts the accumulated loss: loss[0] = (float) sum.
__global__ void __launch_bounds__(128) rows_finish_final_kernel(const float* __restrict__ hs_t, const float* __restrict__ hs_s,
                                                                const int64_t* __restrict__ id_soft,
                                                                const int64_t* __restrict__ id_pred,
                                                                const float* __restrict__ rows, const float* __restrict__ eg,
                                                                int C, double* __restrict__ loss_acc, unsigned* __restrict__ counter,
                                                                float* __restrict__ loss, float* __restrict__ grad_hs_s) {
  // counter[1]: matched student queries (prepare_v1_kernel); fewer than pairs => NaN loss (see there)
  __shared__ float red[32];
  __shared__ double dred[32];
  __shared__ float bcast;
  const int p = blockIdx.x;
  const float* A = rows + (int64_t)p * C;
  const float* E = eg + (int64_t)p * C;
  float dot = 0.f;
  double part = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = A[c], e = E[c];
    part += (double)a * (double)a * (double)e;
    dot += a * (2.f * a * e);
  }
  part = block_sum(part, dred);
  if (grad_hs_s != nullptr) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) bcast = dot;
    __syncthreads();
    dot = bcast;
    const int64_t qs = id_pred[p];
    const float* t = hs_t + id_soft[p] * (int64_t)C;
    const float* sv = hs_s + qs * (int64_t)C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float a = A[c], e = E[c];
      const float da = a * (2.f * a * e - dot);
      const float delta = t[c] - sv[c];
      const float sgn = (delta > 0.f) ? 1.f : ((delta < 0.f) ? -1.f : 0.f);
      atomicAdd(grad_hs_s + qs * (int64_t)C + c, -sgn * da);
    }
  }
  if (threadIdx.x == 0) {
    if (part != 0.0) atomicAdd(loss_acc, part);
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {  // last pair: every partial sum is in
      __threadfence();
      const bool short_of_queries = reinterpret_cast<const volatile int*>(counter)[1] < (int)gridDim.x;
      loss[0] = short_of_queries ? __int_as_float(0x7fc00000) : (float)(*reinterpret_cast<volatile double*>(loss_acc));
    }
  }
}

// loss[0] = (float) acc for the KL criterion of the fused decode_v1 step, NaN when pairs went unmatched
__global__ void v1_loss_final_kernel(const double* __restrict__ acc, const int* __restrict__ matched, int num_pairs,
                                     float* __restrict__ loss) {
  loss[0] = *matched < num_pairs ? __int_as_float(0x7fc00000) : (float)acc[0];
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_mask_rows(int32_t mode, const float* d_hs_teacher, const float* d_hs_student,
                              const int64_t* d_id_soft, const int64_t* d_id_pred, int32_t num_pairs,
                              int32_t C, float* d_rows, void* stream) {
  DSKD_REQUIRE(mode == DSKD_MASK_DECODE_V1 || mode == DSKD_MASK_DECODE_V2, "dskd_mask_rows: bad mode %d", mode);
  DSKD_REQUIRE(num_pairs >= 0 && C > 0 && C <= 8192, "dskd_mask_rows: bad sizes pairs=%d C=%d", num_pairs, C);
  if (num_pairs == 0) return DSKD_OK;
  DSKD_REQUIRE(d_hs_teacher && d_id_soft && d_rows, "dskd_mask_rows: null pointer");
  DSKD_REQUIRE(mode == DSKD_MASK_DECODE_V2 || (d_hs_student && d_id_pred), "dskd_mask_rows: decode_v1 needs the student side");
  const size_t smem = (size_t)C * sizeof(float);
  if (mode == DSKD_MASK_DECODE_V1)
    mask_rows_kernel<DSKD_MASK_DECODE_V1><<<num_pairs, 128, smem, as_stream(stream)>>>(
        d_hs_teacher, d_hs_student, d_id_soft, d_id_pred, C, d_rows);
  else
    mask_rows_kernel<DSKD_MASK_DECODE_V2><<<num_pairs, 128, smem, as_stream(stream)>>>(
        d_hs_teacher, nullptr, d_id_soft, nullptr, C, d_rows);
  DSKD_LAUNCH_OK("mask_rows_kernel");
  return DSKD_OK;
}

extern "C" int dskd_mask_rows_bwd(const float* d_hs_teacher, const float* d_hs_student,
                                  const int64_t* d_id_soft, const int64_t* d_id_pred, const float* d_rows,
                                  const float* d_grad_rows, int32_t num_pairs, int32_t C,
                                  float* d_grad_hs_student, void* stream) {
  DSKD_REQUIRE(num_pairs >= 0 && C > 0, "dskd_mask_rows_bwd: bad sizes");
  if (num_pairs == 0) return DSKD_OK;
  DSKD_REQUIRE(d_hs_teacher && d_hs_student && d_id_soft && d_id_pred && d_rows && d_grad_rows && d_grad_hs_student,
               "dskd_mask_rows_bwd: null pointer");
  mask_rows_bwd_kernel<<<num_pairs, 128, 0, as_stream(stream)>>>(d_hs_teacher, d_hs_student, d_id_soft, d_id_pred,
                                                                 d_rows, d_grad_rows, C, d_grad_hs_student);
  DSKD_LAUNCH_OK("mask_rows_bwd_kernel");
  return DSKD_OK;
}

extern "C" int dskd_dsgfd_rows_finish(int32_t criterion, const float* d_hs_teacher, const float* d_hs_student,
                                      const int64_t* d_id_soft, const int64_t* d_id_pred, const float* d_rows,
                                      const float* d_energy_or_grad_rows, int32_t num_pairs, int32_t C,
                                      double* d_loss_acc, float* d_grad_hs_student, void* stream) {
  DSKD_REQUIRE(criterion == 0 || criterion == 1, "dskd_dsgfd_rows_finish: criterion must be 0 (mse) or 1 (kl)");
  DSKD_REQUIRE(num_pairs >= 0 && C > 0, "dskd_dsgfd_rows_finish: bad sizes");
  if (num_pairs == 0) return DSKD_OK;
  DSKD_REQUIRE(d_rows && d_energy_or_grad_rows && (criterion == 1 || d_loss_acc), "dskd_dsgfd_rows_finish: null pointer");
  DSKD_REQUIRE(d_grad_hs_student == nullptr || (d_hs_teacher && d_hs_student && d_id_soft && d_id_pred),
               "dskd_dsgfd_rows_finish: the embedding gradient needs hs / ids");
  if (criterion == 1 && d_grad_hs_student == nullptr) return DSKD_OK;
  if (criterion == 0)
    rows_finish_kernel<true><<<num_pairs, 128, 0, as_stream(stream)>>>(d_hs_teacher, d_hs_student, d_id_soft, d_id_pred,
                                                                       d_rows, d_energy_or_grad_rows, C, d_loss_acc,
                                                                       d_grad_hs_student);
  else
    rows_finish_kernel<false><<<num_pairs, 128, 0, as_stream(stream)>>>(d_hs_teacher, d_hs_student, d_id_soft, d_id_pred,
                                                                        d_rows, d_energy_or_grad_rows, C, d_loss_acc,
                                                                        d_grad_hs_student);
  DSKD_LAUNCH_OK("rows_finish_kernel");
  return DSKD_OK;
}

extern "C" int dskd_raster_cells(int32_t mode, const float* d_boxes, const int32_t* d_box_start,
                                 const float* d_gt_boxes, const int32_t* d_gt_start, const int32_t* d_img_hw,
                                 int32_t N, int32_t max_boxes_per_image, const DskdLevel* levels,
                                 int32_t num_levels, int64_t cells_per_image, void* d_out, void* stream) {
  DSKD_REQUIRE(mode >= DSKD_RASTER_OWNER_EXCL && mode <= DSKD_RASTER_AREA_FGBK, "dskd_raster_cells: bad mode %d", mode);
  DSKD_REQUIRE(N >= 0 && num_levels > 0 && num_levels <= DSKD_MAX_LEVELS && cells_per_image > 0 && levels,
               "dskd_raster_cells: bad sizes");
  if (N == 0) return DSKD_OK;
  DSKD_REQUIRE(d_box_start && d_img_hw && d_out && (d_boxes || max_boxes_per_image == 0), "dskd_raster_cells: null pointer");
  DSKD_REQUIRE(mode != DSKD_RASTER_BINARY_INCL || d_gt_start, "dskd_raster_cells: sg_out needs GT boxes");
  RasterParams prm;
  prm.num_levels = num_levels;
  int64_t cells = 0;
  for (int l = 0; l < num_levels; ++l) {
    prm.levels[l] = levels[l];
    DSKD_REQUIRE(levels[l].H > 0 && levels[l].W > 0 && levels[l].cell_offset == cells,
                 "dskd_raster_cells: level %d is not densely packed", l);
    cells += (int64_t)levels[l].H * levels[l].W;
  }
  DSKD_REQUIRE(cells == cells_per_image, "dskd_raster_cells: cells_per_image %lld != sum H*W %lld",
               (long long)cells_per_image, (long long)cells);
  const size_t smem = (size_t)max_boxes_per_image * num_levels * sizeof(Rect);
  DSKD_REQUIRE(smem <= 200 * 1024, "dskd_raster_cells: too many boxes per image (%d)", max_boxes_per_image);
  dim3 grid((unsigned)ceil_div(cells_per_image, 256), (unsigned)N);
  cudaStream_t st = as_stream(stream);
#define DSKD_RASTER_LAUNCH(M)                                                                              \
  do {                                                                                                     \
    if (smem > 48 * 1024)                                                                                  \
      DSKD_CUDA_OK(cudaFuncSetAttribute(raster_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    raster_kernel<M><<<grid, 256, smem, st>>>(d_boxes, d_box_start, d_gt_boxes, d_gt_start, d_img_hw, prm,   \
                                              cells_per_image, d_out);                                      \
  } while (0)
  switch (mode) {
    case DSKD_RASTER_OWNER_EXCL: DSKD_RASTER_LAUNCH(DSKD_RASTER_OWNER_EXCL); break;
    case DSKD_RASTER_BINARY_INCL: DSKD_RASTER_LAUNCH(DSKD_RASTER_BINARY_INCL); break;
    case DSKD_RASTER_AREA_INCL: DSKD_RASTER_LAUNCH(DSKD_RASTER_AREA_INCL); break;
    default: DSKD_RASTER_LAUNCH(DSKD_RASTER_AREA_FGBK); break;
  }
#undef DSKD_RASTER_LAUNCH
  DSKD_LAUNCH_OK("raster_kernel");
  return DSKD_OK;
}


// Internal to the library (used by dskd_dsgfd_step): the fused decode_v1 prologue / epilogue.
namespace dskd {
int launch_prepare_v1(const DskdDsgfdStepArgs* a, int* owner, float* rows, float* energy, int64_t* ids, double* acc,
                      unsigned* counter, cudaStream_t st) {
  PrepareV1Params p;
  memset(&p, 0, sizeof(p));
  p.raster.num_levels = a->num_levels;
  for (int l = 0; l < a->num_levels; ++l) p.raster.levels[l] = a->levels[l];
  p.raster_tiles = (int)ceil_div(a->cells_per_image, kPrepTile);
  p.raster_ctas = p.raster_tiles * a->N;
  p.num_pairs = a->num_pairs;
  p.N = a->N; p.C = a->C; p.num_rows = a->num_query_rows; p.num_classes = a->num_classes;
  p.cells_per_image = a->cells_per_image;
  p.boxes = a->d_boxes; p.box_start = a->d_box_start; p.img_hw = a->d_img_hw;
  p.hs_t = a->d_hs_teacher; p.hs_s = a->d_hs_student; p.keepid = a->d_teacher_keepid;
  p.labels = a->d_student_labels; p.prev_mask = a->d_prev_mask;
  p.owner = owner; p.rows = rows; p.energy = energy; p.ids = ids; p.matched_count = a->d_matched_count;
  p.grad_hs = a->d_grad_hs_student;
  p.grad_hs_floats = a->d_grad_hs_student ? (int64_t)a->num_query_rows * a->C : 0;
  p.acc = acc; p.counter = counter;
  p.zero_ctas = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(p.grad_hs_floats, 256 * 16), 2 * kNumSMs));
  // raster role: the image's rectangles; pair role: one mask row + the (chunk, warp) hit counts and ballots of the rank search
  const size_t search = (size_t)ceil_div(p.num_rows, 256) * 8 * 2 * sizeof(int);
  const size_t smem = std::max((size_t)a->max_boxes_per_image * a->num_levels * sizeof(Rect), (size_t)a->C * sizeof(float) + search);
  DSKD_REQUIRE(smem <= 200 * 1024, "dskd_dsgfd_step: too many boxes per image (%d) or query rows (%d)", a->max_boxes_per_image, p.num_rows);
  DSKD_REQUIRE(a->cells_per_image < (1ll << 31) - 256, "dskd_dsgfd_step: cells_per_image too large");
  DSKD_REQUIRE(a->d_grad_hs_student == nullptr || aligned16(a->d_grad_hs_student), "dskd_dsgfd_step: d_grad_hs_student must be 16-byte aligned");
  if (smem > 48 * 1024)
    DSKD_CUDA_OK(cudaFuncSetAttribute(prepare_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prepare_v1_kernel<<<p.raster_ctas + p.num_pairs + p.zero_ctas, 256, smem, st>>>(p);
  DSKD_LAUNCH_OK("prepare_v1_kernel");
  return DSKD_OK;
}

int launch_rows_finish_final(const DskdDsgfdStepArgs* a, const float* rows, const float* energy, const int64_t* ids,
                             double* acc, unsigned* counter, float* grad_hs, cudaStream_t st) {
  rows_finish_final_kernel<<<a->num_pairs, 128, 0, st>>>(a->d_hs_teacher, a->d_hs_student, a->d_teacher_keepid, ids, rows,
                                                         energy, a->C, acc, counter, a->d_loss, grad_hs);
  DSKD_LAUNCH_OK("rows_finish_final_kernel");
  return DSKD_OK;
}

int launch_v1_loss_final(const DskdDsgfdStepArgs* a, const double* acc, const unsigned* counter, cudaStream_t st) {
  v1_loss_final_kernel<<<1, 1, 0, st>>>(acc, reinterpret_cast<const int*>(counter) + 1, a->num_pairs, a->d_loss);
  DSKD_LAUNCH_OK("v1_loss_final_kernel");
  return DSKD_OK;
}
}  // namespace dskd
