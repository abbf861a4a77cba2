// Registry loss modules on generic tensors (the reduction boundary of the hot path).
// Reference: mmdet/models/losses/mse_loss.py:9-57, kd_loss.py:12-94, utils.py:30-59.
#include "common.cuh"

namespace dskd {

// KIND 0: (pred-target)^2 (mse_loss.py:9-12); 1: smooth L1 with threshold beta (smooth_l1_loss.py:12-35);
// 2: |pred-target| (smooth_l1_loss.py:38-56).
template <int KIND>
__global__ void __launch_bounds__(256) mse_elementwise_kernel(const float* __restrict__ pred,
                                                              const float* __restrict__ target,
                                                              const float* __restrict__ weight, int64_t n, float beta,
                                                              float grad_scale, float* __restrict__ elem,
                                                              double* __restrict__ sum, float* __restrict__ gp,
                                                              float* __restrict__ gt) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - target[i];
    const float w = weight ? weight[i] : 1.f;
    float l, dl;  // loss and d loss / d pred before the weight
    if (KIND == 0) {
      l = d * d;
      dl = 2.f * d;
    } else {
      const float a = fabsf(d), sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      if (KIND == 1 && a < beta) {
        l = __fdiv_rn(0.5f * a * a, beta);
        dl = __fdiv_rn(d, beta);
      } else {
        l = (KIND == 1) ? a - 0.5f * beta : a;
        dl = sgn;
      }
    }
    const float e = l * w;
    if (elem) elem[i] = e;
    acc += (double)e;
    const float g = dl * w * grad_scale;
    if (gp) gp[i] = g;
    if (gt) gt[i] = -g;
  }
  if (sum != nullptr) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(sum, acc);
  }
}

// One thread per (outer, inner) column, softmax over D with stride `inner` (coalesced along inner).
// For inner == 1 a warp owns a row and its lanes stride over D.
template <bool WARP_ROW>
__global__ void __launch_bounds__(256) kd_kl_rows_kernel(const float* __restrict__ pred, const float* __restrict__ soft,
                                                         int64_t outer, int D, int64_t inner, float Temp,
                                                         const float* __restrict__ row_weight, float grad_scale,
                                                         float* __restrict__ rowloss, double* __restrict__ sum,
                                                         float* __restrict__ gp) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int64_t rows = outer * inner;
  double acc = 0.0;
  const int64_t step = WARP_ROW ? ((int64_t)gridDim.x * blockDim.x) >> 5 : (int64_t)gridDim.x * blockDim.x;
  const int64_t first = WARP_ROW ? (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5
                                 : blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  // every lane of a warp runs the same number of iterations in WARP_ROW mode (row index is warp-uniform)
  for (int64_t r = first; r < rows; r += step) {
    const int64_t o = r / inner, in = r - o * inner;
    const float* p = pred + o * D * inner + in;
    const float* s = soft + o * D * inner + in;
    const int d0 = WARP_ROW ? lane : 0, dstep = WARP_ROW ? 32 : 1;
    float mp = -INFINITY, ms = -INFINITY;
    for (int d = d0; d < D; d += dstep) {
      mp = fmaxf(mp, __fdiv_rn(p[(int64_t)d * inner], Temp));
      ms = fmaxf(ms, __fdiv_rn(s[(int64_t)d * inner], Temp));
    }
    if (WARP_ROW) { mp = warp_max(mp); ms = warp_max(ms); }
    float sp = 0.f, ss = 0.f;
    for (int d = d0; d < D; d += dstep) {
      sp += expf(__fdiv_rn(p[(int64_t)d * inner], Temp) - mp);
      ss += expf(__fdiv_rn(s[(int64_t)d * inner], Temp) - ms);
    }
    if (WARP_ROW) { sp = warp_sum(sp); ss = warp_sum(ss); }
    // double log-sum-exp and difference: KL is second order in (pred - soft), see dsgfd_kl.cu
    const double lse_pd = (double)mp + log((double)sp), lse_sd = (double)ms + log((double)ss);
    const float lse_p = (float)lse_pd, lse_s = (float)lse_sd;
    const double dl = lse_sd - lse_pd;
    const float w = row_weight ? row_weight[r] : 1.f;
    const float gc = grad_scale * w * Temp / (float)D;
    double kl = 0.0;
    for (int d = d0; d < D; d += dstep) {
      const float xp = __fdiv_rn(p[(int64_t)d * inner], Temp), xs = __fdiv_rn(s[(int64_t)d * inner], Temp);
      const float log_p = xp - lse_p;
      const float log_t = xs - lse_s;
      const float t = expf(log_t);
      kl += (double)t * ((double)(xs - xp) - dl);
      if (gp) gp[o * D * inner + (int64_t)d * inner + in] = gc * (expf(log_p) - t);
    }
    if (WARP_ROW) kl = warp_sum(kl);
    const float rl = (float)(kl * (double)(Temp * Temp) / (double)D) * w;  // .mean(1) * T^2, then elementwise weight
    if (!WARP_ROW || lane == 0) {
      if (rowloss) rowloss[r] = rl;
      acc += (double)rl;
    }
  }
  if (sum != nullptr) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && acc != 0.0) atomicAdd(sum, acc);
  }
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_elementwise_loss(int32_t kind, float beta, const float* d_pred, const float* d_target,
                                     const float* d_weight, int64_t n, float grad_scale, float* d_elem, double* d_sum,
                                     float* d_grad_pred, float* d_grad_target, void* stream) {
  DSKD_REQUIRE(n >= 0, "dskd_elementwise_loss: negative size");
  DSKD_REQUIRE(kind >= 0 && kind <= 2, "dskd_elementwise_loss: kind must be 0 (mse), 1 (smooth L1) or 2 (L1)");
  DSKD_REQUIRE(kind != 1 || beta > 0.f, "dskd_elementwise_loss: smooth L1 needs beta > 0 (smooth_l1_loss.py:24)");
  if (n == 0) return DSKD_OK;
  DSKD_REQUIRE(d_pred && d_target, "dskd_elementwise_loss: null pointer");
  const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  cudaStream_t st = as_stream(stream);
  if (kind == 0)
    mse_elementwise_kernel<0><<<grid, 256, 0, st>>>(d_pred, d_target, d_weight, n, beta, grad_scale, d_elem, d_sum, d_grad_pred, d_grad_target);
  else if (kind == 1)
    mse_elementwise_kernel<1><<<grid, 256, 0, st>>>(d_pred, d_target, d_weight, n, beta, grad_scale, d_elem, d_sum, d_grad_pred, d_grad_target);
  else
    mse_elementwise_kernel<2><<<grid, 256, 0, st>>>(d_pred, d_target, d_weight, n, beta, grad_scale, d_elem, d_sum, d_grad_pred, d_grad_target);
  DSKD_LAUNCH_OK("mse_elementwise_kernel");
  return DSKD_OK;
}

extern "C" int dskd_mse_elementwise(const float* d_pred, const float* d_target, const float* d_weight, int64_t n,
                                    float grad_scale, float* d_elem, double* d_sum, float* d_grad_pred,
                                    float* d_grad_target, void* stream) {
  return dskd_elementwise_loss(0, 0.f, d_pred, d_target, d_weight, n, grad_scale, d_elem, d_sum, d_grad_pred, d_grad_target, stream);
}

extern "C" int dskd_kd_kl_rows(const float* d_pred, const float* d_soft, int64_t outer, int32_t D, int64_t inner,
                               float temperature, const float* d_row_weight, float grad_scale, float* d_rowloss,
                               double* d_sum, float* d_grad_pred, void* stream) {
  DSKD_REQUIRE(outer >= 0 && D > 0 && inner > 0, "dskd_kd_kl_rows: bad sizes");
  DSKD_REQUIRE(temperature >= 1.f, "dskd_kd_kl_rows: T must be >= 1 (kd_loss.py:58)");
  if (outer == 0) return DSKD_OK;
  DSKD_REQUIRE(d_pred && d_soft, "dskd_kd_kl_rows: null pointer");
  const int64_t rows = outer * inner;
  if (inner == 1) {
    const int grid = (int)std::min<int64_t>(ceil_div(rows, 8), (int64_t)kNumSMs * 8);
    kd_kl_rows_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(d_pred, d_soft, outer, D, inner, temperature,
                                                                 d_row_weight, grad_scale, d_rowloss, d_sum,
                                                                 d_grad_pred);
  } else {
    const int grid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)kNumSMs * 8);
    kd_kl_rows_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(d_pred, d_soft, outer, D, inner, temperature,
                                                                  d_row_weight, grad_scale, d_rowloss, d_sum,
                                                                  d_grad_pred);
  }
  DSKD_LAUNCH_OK("kd_kl_rows_kernel");
  return DSKD_OK;
}
