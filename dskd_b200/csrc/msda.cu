// Multi-scale deformable attention, forward and backward -- the sampling core of every encoder / decoder layer of the
// detector that hosts the distillation losses.  The reference calls mmcv's CUDA op
// (`mmcv.ops.multi_scale_deform_attn.MultiScaleDeformableAttention`, imported at mmdet/models/utils/transformer.py:23,
// mmcv-full pinned by requirements/mminstall.txt:1; not vendored).  Semantics = its published definition, which is
// `F.grid_sample(value_l, 2 * loc - 1, bilinear, zeros, align_corners=False)` per level, weighted by the attention
// weights and summed over levels and points (mmcv `multi_scale_deformable_attn_pytorch`); a torch restatement of that
// definition lives with the test infrastructure and is what tests/test_gpu_msda.py compares the kernels against.
//
// One warp per (image, query, head), a CTA = 8 consecutive queries of one head; the lanes are the head's channels (D = 32 for 256 / 8), so every bilinear tap is
// one coalesced 128 B load (forward) or one coalesced 128 B red.global (backward) and the value tensor of an image
// (22.8 MB at 800x1333) stays in L2.  The 32 sampling coordinates and 16 weights of the warp's query arrive as one
// coalesced load each (lane = coordinate); lane k prepares point k (tap offset, fractions, in-map bits) once and
// the warp takes it by shuffles instead of repeating the coordinate arithmetic in 32 lanes.  Backward: the per-point sums over channels
// (d attention weight, d x, d y: 48 values per warp) are reduced with a transposing butterfly (16 shuffles per 16
// values instead of 80) and leave as coalesced stores.
#include "common.cuh"

namespace dskd {

constexpr int kMsdaMaxPoints = 16;  // levels * points per head (4 x 4 in Deformable-DETR)
constexpr int kMsdaWarps = 8;

struct MsdaParams {
  int H[DSKD_MAX_LEVELS], W[DSKD_MAX_LEVELS];
  int64_t start[DSKD_MAX_LEVELS];
  int num_levels, P, LP;
  int N, M, D;
  int64_t S, Lq;
  const float* value;     // [N, S, M, D]
  const float* loc;       // [N, Lq, M, L, P, 2]  (x, y) in [0, 1]
  const float* attn;      // [N, Lq, M, L, P]
  float* out;             // [N, Lq, M, D]
  const float* grad_out;  // [N, Lq, M, D]
  float* grad_value;      // [N, S, M, D] (zero-filled before the launch)
  float* grad_loc;
  float* grad_attn;
};

// sum of v[i] over the 32 lanes for 16 values at once: afterwards v[0] of lane l is the total of value l >> 1
__device__ __forceinline__ void warp_reduce_scatter16(float (&v)[kMsdaMaxPoints], const int lane) {
#pragma unroll
  for (int half = kMsdaMaxPoints / 2, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool BWD>
__global__ void __launch_bounds__(32 * kMsdaWarps) msda_kernel(const __grid_constant__ MsdaParams prm) {
  __shared__ int sH[DSKD_MAX_LEVELS], sW[DSKD_MAX_LEVELS], sStart[DSKD_MAX_LEVELS];
  if (threadIdx.x < DSKD_MAX_LEVELS) {
    const int l = min((int)threadIdx.x, prm.num_levels - 1);
    sH[threadIdx.x] = prm.H[l];
    sW[threadIdx.x] = prm.W[l];
    sStart[threadIdx.x] = (int)prm.start[l];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // the warps of a CTA are consecutive queries of ONE head: in the encoder these are neighbouring cells whose sampling
  // footprints overlap, so most of their taps hit L1
  const int64_t qblocks = (prm.Lq + kMsdaWarps - 1) / kMsdaWarps;
  const int64_t q = (blockIdx.x % qblocks) * kMsdaWarps + (threadIdx.x >> 5);
  if (q >= prm.Lq) return;
  const int nm = (int)(blockIdx.x / qblocks);
  const int m = nm % prm.M, n = nm / prm.M;
  const int LP = prm.LP, D = prm.D, M = prm.M;
  const int64_t wid = ((int64_t)n * prm.Lq + q) * M + m;  // (n, q, m) flattened
  // the query's sampling coordinates (lane = coordinate) and weights (lane = point)
  const float locv = lane < 2 * LP ? __ldg(prm.loc + wid * (2 * LP) + lane) : 0.f;
  const float attv = lane < LP ? __ldg(prm.attn + wid * LP + lane) : 0.f;
  // lane k < LP prepares point k once for the whole warp: cell offset of the top-left tap, the fractions, which of the
  // four taps lie inside the map (grid_sample(align_corners=False, zeros): pixel = loc * size - 0.5)
  int p_off = 0, p_w = 1;
  unsigned p_mask = 0;
  float p_fx = 0.f, p_fy = 0.f;
  {
    const float px = __shfl_sync(0xffffffffu, locv, (2 * lane) & 31);
    const float py = __shfl_sync(0xffffffffu, locv, (2 * lane + 1) & 31);
    if (lane < LP) {
      const int l = lane / prm.P;
      const int H = sH[l], W = sW[l];
      const float ix = px * (float)W - 0.5f, iy = py * (float)H - 0.5f;
      // far outside (or not finite): nothing to sample, and the int conversion stays defined
      if (ix > -1.f && iy > -1.f && ix < (float)W && iy < (float)H) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        p_fx = ix - fx0;
        p_fy = iy - fy0;
        const bool xl = x0 >= 0, xh = x0 + 1 < W, yl = y0 >= 0, yh = y0 + 1 < H;
        p_mask = (yl && xl ? 1u : 0u) | (yl && xh ? 2u : 0u) | (yh && xl ? 4u : 0u) | (yh && xh ? 8u : 0u);
        p_off = (sStart[l] + y0 * W + x0) * (M * D) * 4;  // BYTES inside one image's value tensor: below 2^32 (host check)
      }
      p_w = W * (M * D) * 4;
    }
  }
  const float* __restrict__ vbase = prm.value + ((int64_t)n * prm.S * M + m) * D;
  float* __restrict__ gvbase = BWD ? prm.grad_value + ((int64_t)n * prm.S * M + m) * D : nullptr;
  const unsigned tokb = (unsigned)(M * D) * 4u;  // bytes between tokens of one head
  const char* __restrict__ vb = reinterpret_cast<const char*>(vbase);
  char* __restrict__ gb = reinterpret_cast<char*>(gvbase);
  float ga[kMsdaMaxPoints], gx[kMsdaMaxPoints], gy[kMsdaMaxPoints];
  if (BWD) {
#pragma unroll
    for (int k = 0; k < kMsdaMaxPoints; ++k) { ga[k] = 0.f; gx[k] = 0.f; gy[k] = 0.f; }
  }
  for (int d0 = 0; d0 < D; d0 += 32) {  // D = 32: one pass
    const int d = d0 + lane;
    const bool act = d < D;
    const int dd = act ? d : D - 1;
    const float go = (BWD && act) ? __ldg(prm.grad_out + wid * D + dd) : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kMsdaMaxPoints; ++k) {
      if (k < LP) {  // warp-uniform
        const unsigned mask = __shfl_sync(0xffffffffu, p_mask, k);
        if (mask == 0u) continue;  // the point lies outside its map
        const float a = __shfl_sync(0xffffffffu, attv, k);
        const float fx = __shfl_sync(0xffffffffu, p_fx, k), fy = __shfl_sync(0xffffffffu, p_fy, k);
        // 32-bit element offsets from the slice base: one 64-bit add per tap
        // 32-bit byte offsets from the slice base (zero-extended: a tap outside the map is never dereferenced)
        const unsigned off = (unsigned)__shfl_sync(0xffffffffu, p_off, k) + (unsigned)dd * 4u;
        const unsigned wtok = (unsigned)__shfl_sync(0xffffffffu, p_w, k);
#define DSKD_TAP(o) (*reinterpret_cast<const float*>(vb + (size_t)(o)))
        const float v00 = (mask & 1u) ? __ldg(&DSKD_TAP(off)) : 0.f;
        const float v01 = (mask & 2u) ? __ldg(&DSKD_TAP(off + tokb)) : 0.f;
        const float v10 = (mask & 4u) ? __ldg(&DSKD_TAP(off + wtok)) : 0.f;
        const float v11 = (mask & 8u) ? __ldg(&DSKD_TAP(off + wtok + tokb)) : 0.f;
#undef DSKD_TAP
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
        const float sampled = w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11;
        if (!BWD) {
          acc = fmaf(a, sampled, acc);
        } else {
          ga[k] = fmaf(go, sampled, ga[k]);
          const float ag = a * go;
          gx[k] = fmaf(ag, (1.f - fy) * (v01 - v00) + fy * (v11 - v10), gx[k]);
          gy[k] = fmaf(ag, (1.f - fx) * (v10 - v00) + fx * (v11 - v01), gy[k]);
          if (act) {
#define DSKD_GTAP(o) reinterpret_cast<float*>(gb + (size_t)(o))
            if (mask & 1u) atomicAdd(DSKD_GTAP(off), ag * w00);
            if (mask & 2u) atomicAdd(DSKD_GTAP(off + tokb), ag * w01);
            if (mask & 4u) atomicAdd(DSKD_GTAP(off + wtok), ag * w10);
            if (mask & 8u) atomicAdd(DSKD_GTAP(off + wtok + tokb), ag * w11);
#undef DSKD_GTAP
          }
        }
      }
    }
    if (!BWD && act) prm.out[wid * D + d] = acc;
  }
  if (BWD) {
    warp_reduce_scatter16(ga, lane);
    warp_reduce_scatter16(gx, lane);
    warp_reduce_scatter16(gy, lane);
    const int k = lane >> 1;  // the point whose totals this lane holds
    if (k < LP) {
      const int l = k / prm.P;
      if ((lane & 1) == 0) prm.grad_attn[wid * LP + k] = ga[0];
      // d pixel / d loc = size; lane even writes x, lane odd writes y: one coalesced 128 B store
      prm.grad_loc[wid * (2 * LP) + lane] = (lane & 1) ? gy[0] * (float)sH[l] : gx[0] * (float)sW[l];
    }
  }
}

static int msda_fill(MsdaParams& p, const char* who, const float* d_value, const DskdLevel* levels, int32_t num_levels,
                     const float* d_loc, const float* d_attn, int32_t N, int64_t S, int32_t M, int32_t D, int64_t Lq,
                     int32_t P) {
  DSKD_REQUIRE(levels != nullptr && num_levels > 0 && num_levels <= DSKD_MAX_LEVELS, "%s: bad level table", who);
  DSKD_REQUIRE(N >= 0 && S > 0 && M > 0 && D > 0 && Lq >= 0 && P > 0, "%s: bad sizes", who);
  DSKD_REQUIRE(num_levels * P <= kMsdaMaxPoints, "%s: levels * points (%d) above the supported %d", who, num_levels * P,
               kMsdaMaxPoints);
  DSKD_REQUIRE(d_value && d_loc && d_attn, "%s: null pointer", who);
  int64_t cells = 0;
  for (int l = 0; l < num_levels; ++l) {
    DSKD_REQUIRE(levels[l].H > 0 && levels[l].W > 0 && levels[l].cell_offset == cells, "%s: level %d is not densely packed", who, l);
    p.H[l] = levels[l].H;
    p.W[l] = levels[l].W;
    p.start[l] = cells;
    cells += (int64_t)levels[l].H * levels[l].W;
  }
  DSKD_REQUIRE(S * M * D < (1ll << 29), "%s: one image's value tensor (S * M * D) must stay below 2^29 elements", who);
  DSKD_REQUIRE(cells == S, "%s: the levels hold %lld tokens, S is %lld", who, (long long)cells, (long long)S);
  p.num_levels = num_levels;
  p.P = P;
  p.LP = num_levels * P;
  p.N = N; p.M = M; p.D = D; p.S = S; p.Lq = Lq;
  p.value = d_value; p.loc = d_loc; p.attn = d_attn;
  p.out = nullptr; p.grad_out = nullptr; p.grad_value = nullptr; p.grad_loc = nullptr; p.grad_attn = nullptr;
  return DSKD_OK;
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_msda_forward(const float* d_value, const DskdLevel* levels, int32_t num_levels, const float* d_loc,
                                 const float* d_attn, int32_t N, int64_t S, int32_t M, int32_t D, int64_t Lq, int32_t P,
                                 float* d_out, void* stream) {
  MsdaParams p;
  const int rc = msda_fill(p, "dskd_msda_forward", d_value, levels, num_levels, d_loc, d_attn, N, S, M, D, Lq, P);
  if (rc != DSKD_OK) return rc;
  DSKD_REQUIRE(d_out != nullptr, "dskd_msda_forward: d_out is null");
  const int64_t warps = (int64_t)N * Lq * M;
  if (warps == 0) return DSKD_OK;
  p.out = d_out;
  msda_kernel<false><<<(unsigned)(ceil_div(Lq, kMsdaWarps) * N * M), 32 * kMsdaWarps, 0, as_stream(stream)>>>(p);
  DSKD_LAUNCH_OK("msda_kernel<fwd>");
  return DSKD_OK;
}

extern "C" int dskd_msda_backward(const float* d_value, const DskdLevel* levels, int32_t num_levels, const float* d_loc,
                                  const float* d_attn, const float* d_grad_out, int32_t N, int64_t S, int32_t M, int32_t D,
                                  int64_t Lq, int32_t P, float* d_grad_value, float* d_grad_loc, float* d_grad_attn,
                                  void* stream) {
  MsdaParams p;
  const int rc = msda_fill(p, "dskd_msda_backward", d_value, levels, num_levels, d_loc, d_attn, N, S, M, D, Lq, P);
  if (rc != DSKD_OK) return rc;
  DSKD_REQUIRE(d_grad_out && d_grad_value && d_grad_loc && d_grad_attn, "dskd_msda_backward: null pointer");
  cudaStream_t st = as_stream(stream);
  DSKD_CUDA_OK(cudaMemsetAsync(d_grad_value, 0, sizeof(float) * (size_t)N * S * M * D, st));
  const int64_t warps = (int64_t)N * Lq * M;
  if (warps == 0) return DSKD_OK;
  p.grad_out = d_grad_out;
  p.grad_value = d_grad_value;
  p.grad_loc = d_grad_loc;
  p.grad_attn = d_grad_attn;
  msda_kernel<true><<<(unsigned)(ceil_div(Lq, kMsdaWarps) * N * M), 32 * kMsdaWarps, 0, st>>>(p);
  DSKD_LAUNCH_OK("msda_kernel<bwd>");
  return DSKD_OK;
}
