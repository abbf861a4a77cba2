// DSG-FD masked MSE, fused forward + backward, streaming (HBM-bound).
// Reference arithmetic: gfl_deformable_detr_head_il.py:707-718 with loss_fg_feature = MSELoss('sum')
// (mse_loss.py:9-57): per (level, image)  sum_c,h,w (T*M - S*M)^2,  M = mask row of the owning box.
//
// One launch covers every level and image.  Per element the kernel reads S and T once (128-bit,
// L1-bypassing loads) ONLY where a box owns the cell, writes dS once (128-bit store, zeros outside
// boxes) and folds (T-S)^2 into energy[pair, channel] with warp-shuffle reductions followed by one
// coalesced 32-lane red.global per 32 channels.  Algorithmic bytes: 3 * 4 B per element
// (DESIGN.md section "Kernels"), i.e. 68.27 MB per 800x1333 image.
#include "common.cuh"

namespace dskd {

constexpr int kMaxOwners = 16;  // distinct boxes per warp run handled by the shared-memory fast path
constexpr int kChanChunk = 32;  // channels per CTA (NCHW kernel) = lanes of the final red.global

struct MseParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* student[DSKD_MAX_LEVELS];
  const float* teacher[DSKD_MAX_LEVELS];
  float* grad[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  int block_start[DSKD_MAX_LEVELS + 1];  // first CTA of each level (NCHW kernel)
  int groups[DSKD_MAX_LEVELS];           // cell groups per plane
  int vec4[DSKD_MAX_LEVELS];             // 1: plane size % 4 == 0 and 16 B aligned -> 128-bit path
  int num_levels, N, C;
  int64_t cells_per_image;
  const int* owner;
  const float* rows;
  float* energy;
  const float* cell_weight;
  double* loss;
};

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = ld_stream_f4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const { st_stream_f4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = ld_stream_f1(p); }
  __device__ __forceinline__ void store(float* p) const { st_stream_f1(p, v[0]); }
};

// ------------------------------------------------------------------------------------------------
// NCHW: a warp owns 32*VEC consecutive cells of one (level, image) plane and walks kChanChunk
// channels; the 8 warps of a CTA cover 8 adjacent cell runs so each channel step touches one
// contiguous 4 KB (VEC=4) piece of the plane.
// ------------------------------------------------------------------------------------------------
struct NchwSmem {
  float rows[8][kMaxOwners][kChanChunk + 4];  // per warp: mask-row slice of each box it touches (+4: bank spread)
  int owner[8][kMaxOwners];
  double red[32];
};

template <int VEC, bool CELL>
__device__ __forceinline__ void nchw_tile(const MseParams& prm, const int lvl, NchwSmem& sm) {
  constexpr int U = 4;  // channels in flight per thread: 2*U independent 128-bit loads
  static_assert(U == 4 && kChanChunk == 32, "the transposed reduction below is written for 4 x 8 channels");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = prm.levels[lvl].H * prm.levels[lvl].W;
  const int nchunks = prm.C / kChanChunk;
  int idx = blockIdx.x - prm.block_start[lvl];
  const int group = idx % prm.groups[lvl];
  idx /= prm.groups[lvl];
  const int chunk = idx % nchunks;
  const int img = idx / nchunks;
  const int cell0 = (group * 8 + warp) * (32 * VEC) + lane * VEC;
  const float scale = prm.scale[lvl];
  const float* __restrict__ S = prm.student[lvl];
  const float* __restrict__ T = prm.teacher[lvl];
  float* __restrict__ G = prm.grad[lvl];
  const int c0 = chunk * kChanChunk;
  const int64_t plane0 = ((int64_t)img * prm.C + c0) * HW + cell0;
  const bool in_range = cell0 < HW;  // VEC == 4 requires HW % 4 == 0, so a vector never straddles the end

  // ---- per-cell mask source
  int own[VEC];
  float wgt[VEC];
  int lmax = -1;
  {
    const int64_t cbase = (int64_t)img * prm.cells_per_image + prm.levels[lvl].cell_offset + cell0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      own[k] = -1;
      wgt[k] = 0.f;
      if (in_range) {
        if (CELL) {
          wgt[k] = __ldg(prm.cell_weight + cbase + k);
          own[k] = (wgt[k] != 0.f) ? 0 : -1;
        } else {
          own[k] = __ldg(prm.owner + cbase + k);
        }
      }
      lmax = max(lmax, own[k]);
    }
  }
  const bool mine = lmax >= 0;  // this lane has at least one masked-in cell: it must read S and T

  // ---- row-mask mode: warp-uniform list of the distinct boxes owning these 32*VEC cells (descending
  // pair index), their mask-row slices staged in shared memory, and each cell's slot in that list
  int D = 0;
  bool overflow = false;
  int slot[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) slot[k] = -1;
  if (!CELL) {
    int cur = __reduce_max_sync(0xffffffffu, lmax);
    while (cur >= 0 && D < kMaxOwners) {
      if (lane == 0) sm.owner[warp][D] = cur;
      sm.rows[warp][D][lane] = __ldg(prm.rows + (int64_t)cur * prm.C + c0 + lane);  // coalesced 128 B
      int nxt = -1;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (own[k] == cur) slot[k] = D;
        if (own[k] < cur) nxt = max(nxt, own[k]);
      }
      ++D;
      cur = __reduce_max_sync(0xffffffffu, nxt);
    }
    overflow = cur >= 0;  // more than kMaxOwners boxes inside one warp's run of cells: per-cell atomics
    __syncwarp();
  }
  const bool warp_active = CELL ? (__any_sync(0xffffffffu, mine) != 0) : (D > 0);
  float loss_acc = 0.f;

  if (!warp_active) {
    // nothing owned in these 32*VEC cells: the gradient is zero and no feature byte is needed
    if (G != nullptr && in_range) {
      Vec<VEC> z;
#pragma unroll
      for (int k = 0; k < VEC; ++k) z.v[k] = 0.f;
#pragma unroll 8
      for (int cc = 0; cc < kChanChunk; ++cc) z.store(G + plane0 + (int64_t)cc * HW);
    }
  } else {
    const bool hi16 = lane & 16, hi8 = lane & 8;
    // Software pipeline: the loads of step cb+U are issued as soon as step cb's features have been consumed,
    // so they are in flight while the per-box energy reduction of step cb runs.
    Vec<VEC> s[U], t[U];
    auto issue_loads = [&](int cb) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (mine) {
          s[u].load(S + plane0 + (int64_t)(cb + u) * HW);
          t[u].load(T + plane0 + (int64_t)(cb + u) * HW);
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) s[u].v[k] = t[u].v[k] = 0.f;
        }
      }
    };
    issue_loads(0);
    for (int cb = 0; cb < kChanChunk; cb += U) {
      // mask value of every cell for the U channels of this step: one 128-bit shared load per cell
      float m[VEC][U];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (CELL) {
#pragma unroll
          for (int u = 0; u < U; ++u) m[k][u] = wgt[k];
        } else if (overflow) {
#pragma unroll
          for (int u = 0; u < U; ++u)
            m[k][u] = own[k] >= 0 ? __ldg(prm.rows + (int64_t)own[k] * prm.C + c0 + cb + u) : 0.f;
        } else {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (slot[k] >= 0) v = *reinterpret_cast<const float4*>(&sm.rows[warp][slot[k]][cb]);
          m[k][0] = v.x; m[k][1] = v.y; m[k][2] = v.z; m[k][3] = v.w;
        }
      }
      float dsq[VEC][U];  // (T - S)^2, kept for the per-box energy sums
#pragma unroll
      for (int u = 0; u < U; ++u) {
        Vec<VEC> g;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float dd = t[u].v[k] - s[u].v[k];
          const float m2 = m[k][u] * m[k][u];
          dsq[k][u] = dd * dd;
          g.v[k] = -2.f * scale * m2 * dd;
          if (CELL) loss_acc = fmaf(m2, dsq[k][u], loss_acc);
        }
        if (G != nullptr && in_range) g.store(G + plane0 + (int64_t)(cb + u) * HW);
      }
      if (cb + U < kChanChunk) issue_loads(cb + U);
      if (!CELL) {
        if (!overflow) {
          // energy[box, channel] += sum over this warp's cells owned by the box.  Transposed warp reduction:
          // 6 shuffles give the totals of 4 channels (lanes 8u..8u+7 hold channel cb+u), then one 4-lane red.
          for (int d = 0; d < D; ++d) {
            float e[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              e[u] = 0.f;
#pragma unroll
              for (int k = 0; k < VEC; ++k) e[u] += (slot[k] == d) ? dsq[k][u] : 0.f;
            }
            float x0 = hi16 ? e[2] : e[0], y0 = hi16 ? e[0] : e[2];
            float x1 = hi16 ? e[3] : e[1], y1 = hi16 ? e[1] : e[3];
            x0 += __shfl_xor_sync(0xffffffffu, y0, 16);
            x1 += __shfl_xor_sync(0xffffffffu, y1, 16);
            float z = hi8 ? x1 : x0;
            const float w = hi8 ? x0 : x1;
            z += __shfl_xor_sync(0xffffffffu, w, 8);
            z += __shfl_xor_sync(0xffffffffu, z, 4);
            z += __shfl_xor_sync(0xffffffffu, z, 2);
            z += __shfl_xor_sync(0xffffffffu, z, 1);
            if ((lane & 7) == 0)
              atomicAdd(prm.energy + (int64_t)sm.owner[warp][d] * prm.C + c0 + cb + (lane >> 3), scale * z);
          }
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k)
            if (own[k] >= 0) {
#pragma unroll
              for (int u = 0; u < U; ++u)
                atomicAdd(prm.energy + (int64_t)own[k] * prm.C + c0 + cb + u, scale * dsq[k][u]);
            }
        }
      }
    }
  }
  if (CELL) {
    double tot = block_sum((double)loss_acc * (double)scale, sm.red);
    if (threadIdx.x == 0 && tot != 0.0) atomicAdd(prm.loss, tot);
  }
}

// One launch for every level: levels whose plane size is a multiple of 4 floats take the 128-bit path,
// the others (25x42 and 13x21 at 800x1333: 6 % of the bytes) the 32-bit path; the choice is CTA-uniform.
template <bool CELL>
__global__ void __launch_bounds__(256) dsgfd_mse_nchw_kernel(const __grid_constant__ MseParams prm) {
  __shared__ NchwSmem sm;
  int lvl = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && (int)blockIdx.x >= prm.block_start[k]) lvl = k;
  if (prm.vec4[lvl]) nchw_tile<4, CELL>(prm, lvl, sm);
  else nchw_tile<1, CELL>(prm, lvl, sm);
}

// ------------------------------------------------------------------------------------------------
// [S,N,C] encoder memory: channels are contiguous, so a warp reads one token row (C floats) per
// step with float4 lanes along C and keeps per-lane energy accumulators that are flushed with a
// coalesced red.global whenever the owning box changes along the token run.
// ------------------------------------------------------------------------------------------------
constexpr int kSncBatch = 4;     // tokens whose loads a warp issues together
__device__ __forceinline__ void red_add_f4(float* p, float a, float b, float c, float d) {  // 16-byte aligned p
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int NC, bool CELL>  // NC = ceil(C / 128): float4 slots per lane; tokens_per_warp <= 32
__global__ void __launch_bounds__(256) dsgfd_mse_snc_kernel(const __grid_constant__ MseParams prm, int tokens_per_warp) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const int C = prm.C, N = prm.N;
  const int64_t S_total = prm.cells_per_image;
  const int64_t t_begin = ((int64_t)blockIdx.x * 8 + warp) * tokens_per_warp;
  const int64_t t_end = min(t_begin + tokens_per_warp, S_total);
  const float* __restrict__ S = prm.student[0];
  const float* __restrict__ T = prm.teacher[0];
  float* __restrict__ G = prm.grad[0];
  float acc[NC][4];
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
  int cur_owner = -1;
  float loss_acc = 0.f;

  auto flush = [&](int owner) {  // mid-run: the owner changed under this warp
    if (owner < 0) return;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        red_add_f4(prm.energy + (int64_t)owner * C + c, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
      }
    }
  };

  // Per-token metadata first, one coalesced load for the whole token run: lane i <-> token t_begin + i (owner or cell
  // weight, level scale).  The tokens are then taken kSncBatch at a time: every feature load of a batch is issued
  // before the first is used (one token is only 2 x C x 4 bytes; token by token the warp had 2 KB in flight and the
  // kernel ran at 3.3 TB/s with every cell covered).
  const int n_tok = (int)max((int64_t)0, t_end - t_begin);
  int own_l = -1;
  float w_l = 0.f, sc_l = 0.f;
  if (lane < n_tok) {
    const int64_t t = t_begin + lane;
    int lvl = 0;
#pragma unroll
    for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
      if (k < prm.num_levels && t >= prm.levels[k].cell_offset) lvl = k;
    sc_l = prm.scale[lvl];
    if (CELL) {
      w_l = __ldg(prm.cell_weight + (int64_t)img * S_total + t);
      own_l = (w_l != 0.f) ? 0 : -1;
    } else {
      own_l = __ldg(prm.owner + (int64_t)img * S_total + t);
    }
  }
  for (int i0 = 0; i0 < n_tok; i0 += kSncBatch) {
    int own[kSncBatch];
    float4 s[kSncBatch][NC], tt[kSncBatch][NC];
#pragma unroll
    for (int r = 0; r < kSncBatch; ++r) {
      own[r] = __shfl_sync(0xffffffffu, own_l, (i0 + r) & 31);
      if (i0 + r >= n_tok) own[r] = -2;  // past the end of the run
      const int64_t row = ((t_begin + i0 + r) * N + img) * (int64_t)C;
      if (own[r] >= 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int c = (j * 32 + lane) * 4;
          if (c < C) {
            s[r][j] = ld_stream_f4(S + row + c);
            tt[r][j] = ld_stream_f4(T + row + c);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kSncBatch; ++r) {
      if (own[r] == -2) break;
      const int64_t row = ((t_begin + i0 + r) * N + img) * (int64_t)C;
      const int owner = own[r];
      if (!CELL && owner != cur_owner) {
        flush(cur_owner);
        cur_owner = owner;
      }
      if (owner < 0) {
        if (G != nullptr) {
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const int c = (j * 32 + lane) * 4;
            if (c < C) st_stream_f4(G + row + c, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        continue;
      }
      const float scale = __shfl_sync(0xffffffffu, sc_l, (i0 + r) & 31);
      const float w = __shfl_sync(0xffffffffu, w_l, (i0 + r) & 31);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int c = (j * 32 + lane) * 4;
        if (c < C) {
          const float4 a = CELL ? make_float4(w, w, w, w)
                                : __ldg(reinterpret_cast<const float4*>(prm.rows + (int64_t)owner * C + c));
          const float d0 = tt[r][j].x - s[r][j].x, d1 = tt[r][j].y - s[r][j].y, d2 = tt[r][j].z - s[r][j].z,
                      d3 = tt[r][j].w - s[r][j].w;
          const float m0 = a.x * a.x, m1 = a.y * a.y, m2 = a.z * a.z, m3 = a.w * a.w;
          if (CELL) {
            loss_acc += scale * (m0 * d0 * d0 + m1 * d1 * d1 + m2 * d2 * d2 + m3 * d3 * d3);
          } else {
            acc[j][0] = fmaf(scale * d0, d0, acc[j][0]);
            acc[j][1] = fmaf(scale * d1, d1, acc[j][1]);
            acc[j][2] = fmaf(scale * d2, d2, acc[j][2]);
            acc[j][3] = fmaf(scale * d3, d3, acc[j][3]);
          }
          if (G != nullptr) {
            const float k2 = -2.f * scale;
            st_stream_f4(G + row + c, make_float4(k2 * m0 * d0, k2 * m1 * d1, k2 * m2 * d2, k2 * m3 * d3));
          }
        }
      }
    }
  }
  if (!CELL) {
    // End of the run: the warps of a CTA usually finish inside the same box.  They add up in shared memory and the
    // first warp of every distinct owner issues the atomics (with one box over the whole image every warp of the grid
    // used to hit the same C addresses: 450 us instead of ~200).
    __shared__ __align__(16) float fin[8][NC * 128];
    __shared__ int fin_owner[8];
#pragma unroll
    for (int j = 0; j < NC; ++j)
      *reinterpret_cast<float4*>(&fin[warp][(j * 32 + lane) * 4]) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    if (lane == 0) fin_owner[warp] = cur_owner;
    __syncthreads();
    bool first = cur_owner >= 0;
    for (int v = 0; v < warp; ++v) first = first && fin_owner[v] != cur_owner;
    if (first) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int c = (j * 32 + lane) * 4;
        if (c < C) {
          float4 sum = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
          for (int v = warp + 1; v < 8; ++v)
            if (fin_owner[v] == cur_owner) {
              const float4 o = *reinterpret_cast<const float4*>(&fin[v][c]);
              sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
            }
          red_add_f4(prm.energy + (int64_t)cur_owner * C + c, sum.x, sum.y, sum.z, sum.w);
        }
      }
    }
  }
  if (CELL) {
    double tot = block_sum((double)loss_acc, red);
    if (threadIdx.x == 0 && tot != 0.0) atomicAdd(prm.loss, tot);
  }
}

// loss = sum rows^2 * energy (double accumulation, fixed order); grad_rows = 2 * rows * energy.
__global__ void __launch_bounds__(1024) dsgfd_mse_finish_kernel(const float* __restrict__ rows,
                                                               const float* energy, int64_t n,
                                                               float* __restrict__ loss, float* grad_rows) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = rows[i], e = energy[i];
    acc += (double)a * (double)a * (double)e;
    if (grad_rows != nullptr) grad_rows[i] = 2.f * a * e;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = (float)acc;
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_dsgfd_mse_fwd_bwd(const DskdDsgfdMseArgs* a, void* stream) {
  DSKD_REQUIRE(a != nullptr, "dskd_dsgfd_mse_fwd_bwd: null args");
  DSKD_REQUIRE(a->layout == DSKD_LAYOUT_NCHW || a->layout == DSKD_LAYOUT_SNC, "dsgfd_mse: bad layout %d", a->layout);
  DSKD_REQUIRE(a->num_levels > 0 && a->num_levels <= DSKD_MAX_LEVELS && a->N >= 0 && a->C > 0, "dsgfd_mse: bad sizes");
  const bool cell = a->d_cell_weight != nullptr;
  DSKD_REQUIRE(cell != (a->d_owner != nullptr), "dsgfd_mse: exactly one of d_owner / d_cell_weight must be set");
  DSKD_REQUIRE(cell ? a->d_loss != nullptr : (a->num_pairs == 0 || (a->d_rows && a->d_energy)),
               "dsgfd_mse: missing %s", cell ? "d_loss" : "d_rows / d_energy");
  if (a->N == 0) return DSKD_OK;
  MseParams prm;
  prm.num_levels = a->num_levels;
  prm.N = a->N;
  prm.C = a->C;
  prm.cells_per_image = a->cells_per_image;
  prm.owner = a->d_owner;
  prm.rows = a->d_rows;
  prm.energy = a->d_energy;
  prm.cell_weight = a->d_cell_weight;
  prm.loss = a->d_loss;
  int64_t cells = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    prm.levels[l] = a->levels[l];
    prm.scale[l] = a->scale[l];
    DSKD_REQUIRE(a->levels[l].H > 0 && a->levels[l].W > 0 && a->levels[l].cell_offset == cells,
                 "dsgfd_mse: level %d is not densely packed", l);
    cells += (int64_t)a->levels[l].H * a->levels[l].W;
  }
  DSKD_REQUIRE(cells == a->cells_per_image, "dsgfd_mse: cells_per_image mismatch");
  cudaStream_t st = as_stream(stream);

  if (a->layout == DSKD_LAYOUT_NCHW) {
    DSKD_REQUIRE(a->C % kChanChunk == 0, "dsgfd_mse: C (%d) must be a multiple of %d for the NCHW layout", a->C, kChanChunk);
    int blocks = 0;
    for (int l = 0; l < a->num_levels; ++l) {
      DSKD_REQUIRE(a->d_student[l] && a->d_teacher[l], "dsgfd_mse: null feature pointer at level %d", l);
      prm.student[l] = a->d_student[l];
      prm.teacher[l] = a->d_teacher[l];
      prm.grad[l] = a->d_grad_student[l];
      const int64_t HW = (int64_t)a->levels[l].H * a->levels[l].W;
      const bool v4 = (HW % 4 == 0) && aligned16(a->d_student[l]) && aligned16(a->d_teacher[l]) &&
                      (a->d_grad_student[l] == nullptr || aligned16(a->d_grad_student[l]));
      prm.vec4[l] = v4 ? 1 : 0;
      prm.groups[l] = (int)ceil_div(HW, 8 * 32 * (v4 ? 4 : 1));
      prm.block_start[l] = blocks;
      blocks += prm.groups[l] * (a->C / kChanChunk) * a->N;
    }
    prm.block_start[a->num_levels] = blocks;
    if (cell) dsgfd_mse_nchw_kernel<true><<<blocks, 256, 0, st>>>(prm);
    else dsgfd_mse_nchw_kernel<false><<<blocks, 256, 0, st>>>(prm);
    DSKD_LAUNCH_OK("dsgfd_mse_nchw_kernel");
    return DSKD_OK;
  }

  // SNC
  DSKD_REQUIRE(a->d_student[0] && a->d_teacher[0], "dsgfd_mse: null memory pointer");
  DSKD_REQUIRE(a->C % 4 == 0 && a->C <= 512, "dsgfd_mse: C (%d) must be a multiple of 4 and <= 512 for the SNC layout", a->C);
  DSKD_REQUIRE(aligned16(a->d_student[0]) && aligned16(a->d_teacher[0]) &&
                   (a->d_grad_student[0] == nullptr || aligned16(a->d_grad_student[0])) &&
                   (a->d_rows == nullptr || aligned16(a->d_rows)) && (cell || aligned16(a->d_energy)),
               "dsgfd_mse: SNC tensors must be 16-byte aligned");
  prm.student[0] = a->d_student[0];
  prm.teacher[0] = a->d_teacher[0];
  prm.grad[0] = a->d_grad_student[0];
  const int tokens_per_warp = 16;
  dim3 grid((unsigned)ceil_div(a->cells_per_image, 8 * tokens_per_warp), (unsigned)a->N);
  const int nc = (a->C + 127) / 128;
#define DSKD_SNC_LAUNCH(NCV)                                                                     \
  do {                                                                                           \
    if (cell) dsgfd_mse_snc_kernel<NCV, true><<<grid, 256, 0, st>>>(prm, tokens_per_warp);        \
    else dsgfd_mse_snc_kernel<NCV, false><<<grid, 256, 0, st>>>(prm, tokens_per_warp);            \
  } while (0)
  switch (nc) {
    case 1: DSKD_SNC_LAUNCH(1); break;
    case 2: DSKD_SNC_LAUNCH(2); break;
    case 3: DSKD_SNC_LAUNCH(3); break;
    default: DSKD_SNC_LAUNCH(4); break;
  }
#undef DSKD_SNC_LAUNCH
  DSKD_LAUNCH_OK("dsgfd_mse_snc_kernel");
  return DSKD_OK;
}

extern "C" int dskd_dsgfd_mse_finish(const float* d_rows, const float* d_energy, int32_t num_pairs, int32_t C,
                                     float* d_loss, float* d_grad_rows, void* stream) {
  DSKD_REQUIRE(d_loss != nullptr && num_pairs >= 0 && C > 0, "dskd_dsgfd_mse_finish: bad arguments");
  DSKD_REQUIRE(num_pairs == 0 || (d_rows && d_energy), "dskd_dsgfd_mse_finish: null pointer");
  dsgfd_mse_finish_kernel<<<1, 1024, 0, as_stream(stream)>>>(d_rows, d_energy, (int64_t)num_pairs * C, d_loss, d_grad_rows);
  DSKD_LAUNCH_OK("dsgfd_mse_finish_kernel");
  return DSKD_OK;
}
