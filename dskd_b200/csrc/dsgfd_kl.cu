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
// for 16 resident warps per SM).  Per channel the warp
//   0. cp.async's the owned cells of the student / teacher strip [H][32] into ITS shared-memory buffers (every
//      feature byte is read from HBM exactly once; ~200 independent 128-byte copies in flight per warp),
//   A. turns them in place into the logits x = feature * mask / T (0 outside boxes) and takes the column maxima,
//   B. sums e^(x - max) for both softmaxes and sum e^(xs - max) (xs - xt) -- the KL of a column needs nothing else:
//        KL = sum_h t_h (xs_h - xt_h) - (lse_s - lse_t),
//   C. (only when the mask rows need a gradient) accumulates T_h * (p_h - t_h) per box segment of each column and
//      issues one red.global per (column, segment, channel).
// The owner strip and the segment number of every cell are computed once per CTA and shared by all its channels, so
// the gradient sweep carries no ownership logic.
constexpr int kKlChan = 4;        // channels per warp (sequential)
constexpr int kKlCols = 8;        // columns (w) per strip: a warp covers kKlCols columns x kKlPhases interleaved row phases
constexpr int kKlPhases = 32 / kKlCols;
constexpr int kKlMaxWarps = 32;   // warps per CTA
constexpr int kKlMaxH = 800;      // rows per level the shared-memory strips can hold (one warp per CTA at the limit)
constexpr size_t kKlSmemBudget = 220 * 1024;
constexpr int kKlSegGroup = 4;    // box segments of a column accumulated per pass of the gradient sweep
__host__ __device__ inline size_t kl_header_bytes(int max_h) {
  // owner strip int32 [max_h][kKlCols] | segment owners int32 [kKlCols][max_h] | segments per column int32 [kKlCols]
  // | segment slot of every cell uint16 [max_h][kKlCols]
  return ((size_t)max_h * kKlCols * 4 * 2 + (size_t)kKlCols * 4 + (size_t)max_h * kKlCols * 2 + 127) / 128 * 128;
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
  const bool col_ok = w < W;
  const float Temp = prm.temperature;
  const float scale = prm.scale[lvl];
  const int64_t strip_base = (int64_t)img * prm.cells_per_image + prm.levels[lvl].cell_offset + wt * kKlCols;
  const int64_t cell_base = strip_base + wl;

  // shared memory: header (owner strip, segment tables; kl_header_bytes) | per warp: xs, xt [max_h][kKlCols]
  int* own_s = reinterpret_cast<int*>(kl_smem);
  int* seg_owner = own_s + (size_t)prm.max_h * kKlCols;          // [col][k]: owner of the k-th box segment of the column
  int* nseg_s = seg_owner + (size_t)prm.max_h * kKlCols;         // [col]
  unsigned short* slot_s = reinterpret_cast<unsigned short*>(nseg_s + kKlCols);   // [h][col]: segment index of the cell
  float* xs_s = reinterpret_cast<float*>(kl_smem + kl_header_bytes(prm.max_h)) + (size_t)warp * 2 * prm.max_h * kKlCols;
  float* xt_s = xs_s + (size_t)prm.max_h * kKlCols;

  // ---- once per CTA: owner strip; any owned cell at all?  For the gradient: every column's cells are numbered by
  // the box segment (maximal run of rows with one owner) they belong to, so that the per-channel gradient sweep is a
  // plain loop over rows with kKlSegGroup accumulators and no ownership logic.
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
  const bool want_grad = !CELL && prm.grad_rows != nullptr;
  if (want_grad) {
    if (threadIdx.x < kKlCols) {  // one thread per column walks its H rows
      const int c = threadIdx.x;
      int n = 0, prev = -1;
      for (int h = 0; h < H; ++h) {
        const int o = own_s[h * kKlCols + c];
        if (o >= 0 && o != prev) seg_owner[c * prm.max_h + n++] = o;
        slot_s[h * kKlCols + c] = (unsigned short)(o >= 0 ? n - 1 : 0xffff);
        prev = o;
      }
      nseg_s[c] = n;
      atomicMax(&max_seg, n);
    }
    __syncthreads();
  }

  const float kLog2e = 1.4426950408889634f;
  const float gcoef = scale * Temp / (float)H;  // d loss / d pred = scale * (T/H) * (p - t)
  double kl_total = 0.0;
  constexpr int kRow = kKlCols;                  // floats per strip row
  constexpr int kStep = kKlPhases * kKlCols;     // floats between two rows of the same lane
  const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(xs_s) + (uint32_t)(ph * kRow + wl) * 4u;
  const uint32_t xt_addr = (uint32_t)__cvta_generic_to_shared(xt_s) + (uint32_t)(ph * kRow + wl) * 4u;
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

  for (int k = 0; k < kKlChan; ++k) {
    const int c = (chunk * prm.warps + warp) * kKlChan + k;
    if (c >= C) break;
    const float* __restrict__ S = prm.student[lvl] + ((int64_t)img * C + c) * HW + w;
    const float* __restrict__ T = prm.teacher[lvl] + ((int64_t)img * C + c) * HW + w;
    auto mask_of = [&](int o, int h) -> float {  // mask value of the cell (divided by T when that is exact)
      if (o < 0) return 0.f;
      const float m = CELL ? __ldg(prm.cell_weight + cell_base + (int64_t)h * W) : __ldg(prm.rows + (int64_t)o * C + c);
      return POW2 ? m * prm.inv_temperature : m;
    };
    // Lane-private views: row r of this lane is row ph + r * kKlPhases of the strip; consecutive rows of a lane are
    // kStep floats apart, so the unrolled loops below address shared memory with immediate offsets.
    const int nrl = (H - ph + kKlPhases - 1) / kKlPhases;
    const int* oq0 = own_s + ph * kRow + wl;
    float* xq0 = xs_s + ph * kRow + wl;
    float* tq0 = xt_s + ph * kRow + wl;
    // ---- 0: owned cells -> shared memory asynchronously (every feature byte leaves HBM once)
    {
      const int* oq = oq0;
      uint32_t xa = xs_addr, ta = xt_addr;
      unsigned goff = (unsigned)ph * (unsigned)W;
      const unsigned gstep = (unsigned)kKlPhases * (unsigned)W;
      int r = 0;
      for (; r + 4 <= nrl; r += 4, oq += 4 * kStep, xa += 16u * kStep, ta += 16u * kStep, goff += 4u * gstep) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (oq[q * kStep] >= 0) {
            cp_async_f32(xa + 4u * q * kStep, S + (goff + q * gstep));
            cp_async_f32(ta + 4u * q * kStep, T + (goff + q * gstep));
          }
        }
      }
      for (; r < nrl; ++r, oq += kStep, xa += 4u * kStep, ta += 4u * kStep, goff += gstep) {
        if (oq[0] >= 0) {
          cp_async_f32(xa, S + goff);
          cp_async_f32(ta, T + goff);
        }
      }
    }
    cp_async_wait_all();
    __syncwarp();
    // ---- A: logits in place (exactly 0 outside boxes; those cells were never loaded), column maxima
    float ms = -INFINITY, mt = -INFINITY;
    {
      int prev = -2;
      float m = 0.f;
      const int* oq = oq0;
      float* xq = xq0;
      float* tq = tq0;
      auto row = [&](int q, int h) {
        const int o = oq[q * kStep];
        if (CELL ? (o >= 0) : (o != prev)) { m = mask_of(o, h); prev = o; }  // the mask is constant along a run of rows
        float x = 0.f, y = 0.f;
        if (o >= 0) {
          x = xq[q * kStep] * m;
          y = tq[q * kStep] * m;
          if (!POW2) { x = __fdiv_rn(x, Temp); y = __fdiv_rn(y, Temp); }
        }
        xq[q * kStep] = x;
        tq[q * kStep] = y;
        ms = fmaxf(ms, x);
        mt = fmaxf(mt, y);
      };
      int r = 0;
      for (; r + 4 <= nrl; r += 4, oq += 4 * kStep, xq += 4 * kStep, tq += 4 * kStep) {
#pragma unroll
        for (int q = 0; q < 4; ++q) row(q, ph + (r + q) * kKlPhases);
      }
      for (; r < nrl; ++r, oq += kStep, xq += kStep, tq += kStep) row(0, ph + r * kKlPhases);
      ms = phase_max(ms);
      mt = phase_max(mt);
    }
    // ---- B: softmax sums and the t-weighted logit difference (four interleaved accumulator sets)
    const float nms = -ms * kLog2e, nmt = -mt * kLog2e;
    float ss[4] = {0.f, 0.f, 0.f, 0.f}, st4[4] = {0.f, 0.f, 0.f, 0.f}, ws[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const float* xq = xq0;
      const float* tq = tq0;
      int r = 0;
      for (; r + 4 <= nrl; r += 4, xq += 4 * kStep, tq += 4 * kStep) {
        float a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { a[q] = xq[q * kStep]; b[q] = tq[q * kStep]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float ea = fast_ex2(fmaf(a[q], kLog2e, nms));
          ss[q] += ea;
          st4[q] += fast_ex2(fmaf(b[q], kLog2e, nmt));
          ws[q] = fmaf(ea, a[q] - b[q], ws[q]);
        }
      }
      for (; r < nrl; ++r, xq += kStep, tq += kStep) {
        const float a = xq[0], b = tq[0];
        const float ea = fast_ex2(fmaf(a, kLog2e, nms));
        ss[0] += ea;
        st4[0] += fast_ex2(fmaf(b, kLog2e, nmt));
        ws[0] = fmaf(ea, a - b, ws[0]);
      }
    }
    const float sum_s = phase_sum((ss[0] + ss[2]) + (ss[1] + ss[3])), sum_t = phase_sum((st4[0] + st4[2]) + (st4[1] + st4[3]));
    const float wsum = phase_sum((ws[0] + ws[2]) + (ws[1] + ws[3]));
    // KL of the column = sum_h t_h (xs - xt) - (lse_s - lse_t): a second-order quantity (both log-softmaxes sit near
    // -log H), so the log-sum-exp difference is taken in double.  One lane per column keeps it.
    if (ph == 0) {
      const double dl = ((double)ms - (double)mt) + log((double)sum_s / (double)sum_t);
      kl_total += (double)wsum / (double)sum_s - dl;
    }

    // ---- C: d loss / d mask rows.  sum over a box segment of T_h (p_h - t_h) = (T/mask) sum xt_h (p_h - t_h); the cells
    // carry their segment number, kKlSegGroup segments per column are accumulated per pass (one pass unless a column
    // crosses more boxes), then the row phases are combined and one lane per column issues one red.global per segment.
    if (want_grad) {
      const float rs = __fdividef(1.f, sum_s), rt = __fdividef(1.f, sum_t);
      const unsigned short* sq0 = slot_s + ph * kRow + wl;
      const int my_nseg = nseg_s[wl];
      for (int base = 0; base < max_seg; base += kKlSegGroup) {
        float acc[kKlSegGroup];
#pragma unroll
        for (int k = 0; k < kKlSegGroup; ++k) acc[k] = 0.f;
        const float* xq = xq0;
        const float* tq = tq0;
        const unsigned short* sq = sq0;
        auto row = [&](int q) {
          const float a = xq[q * kStep], b = tq[q * kStep];
          const int sl = (int)sq[q * kStep] - base;
          const float v = b * (fast_ex2(fmaf(b, kLog2e, nmt)) * rt - fast_ex2(fmaf(a, kLog2e, nms)) * rs);
#pragma unroll
          for (int k = 0; k < kKlSegGroup; ++k) acc[k] += (sl == k) ? v : 0.f;
        };
        int r = 0;
        for (; r + 4 <= nrl; r += 4, xq += 4 * kStep, tq += 4 * kStep, sq += 4 * kStep) {
#pragma unroll
          for (int q = 0; q < 4; ++q) row(q);
        }
        for (; r < nrl; ++r, xq += kStep, tq += kStep, sq += kStep) row(0);
#pragma unroll
        for (int k = 0; k < kKlSegGroup; ++k) acc[k] = phase_sum(acc[k]);
        if (ph == 0) {
#pragma unroll
          for (int k = 0; k < kKlSegGroup; ++k) {
            if (base + k >= my_nseg) continue;
            const int who = seg_owner[wl * prm.max_h + base + k];
            const float m = mask_of(who, 0);
            float v = acc[k];
            float div = POW2 ? m : m / Temp;
            if (m == 0.f) {
              // the mask value underflowed to 0: the logits carry no trace of the teacher feature, re-read it
              v = 0.f;
              div = 1.f;
              for (int h = 0; h < H; ++h) {
                if ((int)slot_s[h * kRow + wl] != base + k) continue;
                const float pp = fast_ex2(fmaf(xt_s[h * kRow + wl], kLog2e, nmt)) * rt;
                const float tt = fast_ex2(fmaf(xs_s[h * kRow + wl], kLog2e, nms)) * rs;
                v = fmaf(ld_stream_f1(T + (int64_t)h * W), pp - tt, v);
              }
            }
            atomicAdd(prm.grad_rows + (int64_t)who * C + c, __fdividef(gcoef * v, div));
          }
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
