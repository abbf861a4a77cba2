// DSG-FD with the shipped criterion: KnowledgeDistillationKLDivLoss(T=2, reduction='sum') applied to
// [C,H,W] maps, i.e. softmax over H (kd_loss.py:28-34 with dim=1 == H), target = student*mask
// (detached), pred = teacher*mask.  Reference: gfl_deformable_detr_head_il.py:707-718 + kd_loss.py:12-43.
//
// Column kernel: lane = one (channel, w) column, walked over H twice (online softmax statistics,
// then the KL terms and d loss / d mask).  Cells outside every box have logit 0 for both softmaxes
// and need no feature bytes; columns without any owned cell contribute exactly 0 and are skipped.
// No gradient reaches the student features (target is detached): the only gradient is d loss / d rows.
#include "common.cuh"

namespace dskd {

constexpr int kKlChan = 4;          // channels walked together by one warp (8 independent loads per row)
constexpr int kKlWarps = 4;         // warps per CTA -> 16 channels per CTA

struct KlParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* student[DSKD_MAX_LEVELS];
  const float* teacher[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  int block_start[DSKD_MAX_LEVELS + 1];
  int wtiles[DSKD_MAX_LEVELS];
  int num_levels, N, C;
  float temperature;
  int64_t cells_per_image;
  const int* owner;
  const float* rows;
  float* grad_rows;
  const float* cell_weight;
  double* loss;
};

struct OnlineLse {  // running max / sum of exp for a softmax over H
  float mx, sum;
  __device__ __forceinline__ void init() { mx = -INFINITY; sum = 0.f; }
  __device__ __forceinline__ void push(float x) {
    if (x > mx) {
      sum = sum * expf(mx - x) + 1.f;
      mx = x;
    } else {
      sum += expf(x - mx);
    }
  }
  // log-sum-exp in double: the KL below is a second-order quantity (sum_h t_h (log t_h - log p_h) with both
  // log-softmaxes ~ -log H), so an fp32 logf here (abs. error ~3e-7) would be ~1 % of a column's KL.
  __device__ __forceinline__ double lse() const { return (double)mx + log((double)sum); }
};

template <bool CELL>
__global__ void __launch_bounds__(32 * kKlWarps) dsgfd_kl_kernel(const __grid_constant__ KlParams prm) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lvl = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && (int)blockIdx.x >= prm.block_start[k]) lvl = k;
  const int H = prm.levels[lvl].H, W = prm.levels[lvl].W, C = prm.C;
  const int HW = H * W;
  const int nchunks = C / (kKlChan * kKlWarps);
  int idx = blockIdx.x - prm.block_start[lvl];
  const int wt = idx % prm.wtiles[lvl];
  idx /= prm.wtiles[lvl];
  const int chunk = idx % nchunks;
  const int img = idx / nchunks;
  const int w = wt * 32 + lane;
  const bool col_ok = w < W;
  const int c0 = (chunk * kKlWarps + warp) * kKlChan;
  const float Temp = prm.temperature;
  const float scale = prm.scale[lvl];
  const float* __restrict__ S = prm.student[lvl] + ((int64_t)img * C + c0) * HW + w;
  const float* __restrict__ T = prm.teacher[lvl] + ((int64_t)img * C + c0) * HW + w;
  const int64_t cell_base = (int64_t)img * prm.cells_per_image + prm.levels[lvl].cell_offset + w;

  auto cell_owner = [&](int h) -> int {
    if (!col_ok) return -1;
    if (CELL) return (__ldg(prm.cell_weight + cell_base + (int64_t)h * W) != 0.f) ? 0 : -1;
    return __ldg(prm.owner + cell_base + (int64_t)h * W);
  };
  auto mask_value = [&](int owner, int h, int k) -> float {
    if (CELL) return __ldg(prm.cell_weight + cell_base + (int64_t)h * W);
    return __ldg(prm.rows + (int64_t)owner * C + c0 + k);
  };

  // pass 0: does this column meet any box?  (owners are shared by every channel)
  bool any = false;
  for (int h = 0; h < H; ++h) any |= cell_owner(h) >= 0;
  double kl_total = 0.0;
  if (__any_sync(0xffffffffu, any)) {
    // pass 1: softmax statistics over H for target (student*mask/T) and pred (teacher*mask/T)
    OnlineLse ls[kKlChan], lt[kKlChan];
#pragma unroll
    for (int k = 0; k < kKlChan; ++k) { ls[k].init(); lt[k].init(); }
    if (any) {
      for (int h = 0; h < H; ++h) {
        const int o = cell_owner(h);
        float xs[kKlChan], xt[kKlChan];
#pragma unroll
        for (int k = 0; k < kKlChan; ++k) xs[k] = xt[k] = 0.f;
        if (o >= 0) {
#pragma unroll
          for (int k = 0; k < kKlChan; ++k) {
            const float m = mask_value(o, h, k);
            xs[k] = __fdiv_rn(ld_stream_f1(S + (int64_t)k * HW + (int64_t)h * W) * m, Temp);
            xt[k] = __fdiv_rn(ld_stream_f1(T + (int64_t)k * HW + (int64_t)h * W) * m, Temp);
          }
        }
#pragma unroll
        for (int k = 0; k < kKlChan; ++k) { ls[k].push(xs[k]); lt[k].push(xt[k]); }
      }
    }
    // pass 2: KL terms and d loss / d mask, accumulated per owning box along the column
    float lse_s[kKlChan], lse_t[kKlChan], acc[kKlChan];
    double dl[kKlChan], kl[kKlChan];  // dl = lse_s - lse_t
#pragma unroll
    for (int k = 0; k < kKlChan; ++k) {
      const double a = any ? ls[k].lse() : 0.0, b = any ? lt[k].lse() : 0.0;
      lse_s[k] = (float)a;
      lse_t[k] = (float)b;
      dl[k] = a - b;
      kl[k] = 0.0;
      acc[k] = 0.f;
    }
    const float gcoef = scale * Temp / (float)H;  // d loss / d pred = scale * (T/H) * (p - t)
    int cur = -1;
    for (int h = 0; h <= H; ++h) {
      const int o = (any && h < H) ? cell_owner(h) : -1;
      // flush the per-lane accumulators of the box that just ended (warp-cooperative reduction)
      const bool need = (o != cur) && (cur >= 0);
      unsigned pending = __ballot_sync(0xffffffffu, need);
      if (!CELL && prm.grad_rows != nullptr) {
        while (pending) {
          const int leader = __ffs(pending) - 1;
          const int who = __shfl_sync(0xffffffffu, cur, leader);
          const unsigned same = __ballot_sync(0xffffffffu, need && cur == who);
#pragma unroll
          for (int k = 0; k < kKlChan; ++k) {
            float v = (need && cur == who) ? acc[k] : 0.f;
            v = warp_sum(v);
            if (lane == leader) atomicAdd(prm.grad_rows + (int64_t)who * C + c0 + k, v);
          }
          pending &= ~same;
        }
      }
      if (need) {
#pragma unroll
        for (int k = 0; k < kKlChan; ++k) acc[k] = 0.f;
      }
      cur = o;
      if (h == H || !any) continue;
      float xs[kKlChan], xt[kKlChan], tf[kKlChan];
#pragma unroll
      for (int k = 0; k < kKlChan; ++k) xs[k] = xt[k] = tf[k] = 0.f;
      if (o >= 0) {
#pragma unroll
        for (int k = 0; k < kKlChan; ++k) {
          const float m = mask_value(o, h, k);
          tf[k] = ld_stream_f1(T + (int64_t)k * HW + (int64_t)h * W);
          xs[k] = __fdiv_rn(ld_stream_f1(S + (int64_t)k * HW + (int64_t)h * W) * m, Temp);
          xt[k] = __fdiv_rn(tf[k] * m, Temp);
        }
      }
#pragma unroll
      for (int k = 0; k < kKlChan; ++k) {
        const float log_t = xs[k] - lse_s[k];
        const float log_p = xt[k] - lse_t[k];
        const float t = expf(log_t);
        // log t - log p = (xs - xt) - (lse_s - lse_t), the difference of the small parts taken first
        kl[k] += (double)t * ((double)(xs[k] - xt[k]) - dl[k]);
        if (o >= 0) acc[k] = fmaf(tf[k], gcoef * (expf(log_p) - t), acc[k]);
      }
    }
    if (any) {
#pragma unroll
      for (int k = 0; k < kKlChan; ++k) kl_total += kl[k];
    }
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
  DSKD_REQUIRE(a->C % (kKlChan * kKlWarps) == 0, "dsgfd_kl: C (%d) must be a multiple of %d", a->C, kKlChan * kKlWarps);
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
  prm.cells_per_image = a->cells_per_image;
  prm.owner = a->d_owner;
  prm.rows = a->d_rows;
  prm.grad_rows = a->d_grad_rows;
  prm.cell_weight = a->d_cell_weight;
  prm.loss = a->d_loss;
  int64_t cells = 0;
  int blocks = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    DSKD_REQUIRE(a->levels[l].H > 0 && a->levels[l].W > 0 && a->levels[l].cell_offset == cells,
                 "dsgfd_kl: level %d is not densely packed", l);
    DSKD_REQUIRE(a->d_student[l] && a->d_teacher[l], "dsgfd_kl: null feature pointer at level %d", l);
    cells += (int64_t)a->levels[l].H * a->levels[l].W;
    prm.levels[l] = a->levels[l];
    prm.student[l] = a->d_student[l];
    prm.teacher[l] = a->d_teacher[l];
    prm.scale[l] = a->scale[l];
    prm.wtiles[l] = (a->levels[l].W + 31) / 32;
    prm.block_start[l] = blocks;
    blocks += prm.wtiles[l] * (a->C / (kKlChan * kKlWarps)) * a->N;
  }
  prm.block_start[a->num_levels] = blocks;
  DSKD_REQUIRE(cells == a->cells_per_image, "dsgfd_kl: cells_per_image mismatch");
  if (cell) dsgfd_kl_kernel<true><<<blocks, 32 * kKlWarps, 0, as_stream(stream)>>>(prm);
  else dsgfd_kl_kernel<false><<<blocks, 32 * kKlWarps, 0, as_stream(stream)>>>(prm);
  DSKD_LAUNCH_OK("dsgfd_kl_kernel");
  return DSKD_OK;
}
