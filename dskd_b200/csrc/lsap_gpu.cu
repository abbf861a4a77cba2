// Batched rectangular linear sum assignment ON THE DEVICE: removes the last
// device->host sync of the assignment path (gfl_hungarian_assigner.py:143-151 `.cpu()` + SciPy per (layer, image)).
//
// Same published algorithm as lsap.cpp -- the shortest-augmenting-path Jonker-Volgenant variant of D. F. Crouse,
// "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016, which is what
// scipy.optimize.linear_sum_assignment runs -- with the same float64 arithmetic in the same operation order and the
// same tie-breaking, so the indices are identical to SciPy's:
//   * rows > cols is solved on the transpose;
//   * candidate columns are scanned in the order of the `remaining` list (initialised descending, swap-removal);
//     among equal reduced costs an unassigned column wins over an assigned one, the LAST unassigned one in scan order
//     wins among unassigned, the FIRST one among assigned.
// One CTA per problem, one thread per column of the (transposed) problem: the per-column state (dual v, shortest path
// cost, scanned flag, position in `remaining`) lives in registers, the cost matrix is staged in shared memory with
// the tree row as the slow index, and every Dijkstra step is one block-wide lexicographic arg-min.
#include <math_constants.h>

#include "common.cuh"

namespace dskd {

struct LsapKey {
  double val;
  int assigned;  // 0: unassigned column (preferred on ties)
  int pos;       // position in `remaining`
  int col;
};

__device__ __forceinline__ bool lsap_better(const LsapKey& a, const LsapKey& b) {
  if (a.val != b.val) return a.val < b.val;
  if (a.assigned != b.assigned) return a.assigned < b.assigned;
  return a.assigned ? (a.pos < b.pos) : (a.pos > b.pos);
}

__device__ __forceinline__ LsapKey lsap_shfl(const LsapKey& k, int off) {
  LsapKey o;
  o.val = __shfl_xor_sync(0xffffffffu, k.val, off);
  o.assigned = __shfl_xor_sync(0xffffffffu, k.assigned, off);
  o.pos = __shfl_xor_sync(0xffffffffu, k.pos, off);
  o.col = __shfl_xor_sync(0xffffffffu, k.col, off);
  return o;
}

// dynamic shared memory: u[nr] (double) | red_val[32] (double) | ints: path[nc] row4col[nc] col4row[nr] remaining[nc]
// red_i[32*3] | cost_t[nr*nc] floats (optional)
__global__ void __launch_bounds__(1024) lsap_batch_kernel(const float* __restrict__ cost_all, int N, int rows, int ld,
                                                          const int* __restrict__ gt_start, int64_t* __restrict__ assigned_all,
                                                          int* __restrict__ status, int stage_cost) {
  extern __shared__ __align__(16) unsigned char lsap_smem[];
  __shared__ int s_i, s_sink, s_nrem, s_moved_col, s_moved_pos, s_winner, s_bad;
  __shared__ double s_min;
  const int p = blockIdx.x, img = p % N;
  const int cols = gt_start[img + 1] - gt_start[img];
  int64_t* out = assigned_all + (int64_t)p * rows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (cols <= 0 || rows <= 0) {
    for (int r = tid; r < rows; r += blockDim.x) out[r] = 0;
    if (tid == 0 && status) status[p] = DSKD_OK;
    return;
  }
  const float* __restrict__ cost = cost_all + (int64_t)p * rows * ld;   // [rows][ld], cols valid
  const bool transpose = cols < rows;
  const int nr = transpose ? cols : rows, nc = transpose ? rows : cols;  // nr <= nc; thread j owns column j < nc
  double* u = reinterpret_cast<double*>(lsap_smem);
  double* red_val = u + nr;
  int* path = reinterpret_cast<int*>(red_val + 32);
  int* row4col = path + nc;
  int* col4row = row4col + nc;
  int* remaining = col4row + nr;
  int* red_i = remaining + nc;
  float* cost_t = reinterpret_cast<float*>(red_i + 96);                  // [nr][nc]: tree row slow, column fast
  // c(i, j) of the problem being solved (i < nr, j < nc)
  auto cost_at = [&](int i, int j) -> double {
    if (stage_cost) return (double)cost_t[i * nc + j];
    return transpose ? (double)cost[(int64_t)j * ld + i] : (double)cost[(int64_t)i * ld + j];
  };
  if (tid == 0) s_bad = 0;
  __syncthreads();
  {
    bool bad = false;
    for (int e = tid; e < rows * cols; e += blockDim.x) {
      const int r = e / cols, c = e - r * cols;
      const float x = cost[(int64_t)r * ld + c];
      bad |= isnan(x) || (isinf(x) && x < 0.f);
      if (stage_cost) cost_t[transpose ? (c * nc + r) : (r * nc + c)] = x;
    }
    if (bad) s_bad = 1;
  }
  for (int i = tid; i < nr; i += blockDim.x) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = tid; j < nc; j += blockDim.x) { row4col[j] = -1; path[j] = -1; }
  __syncthreads();
  if (s_bad) {  // NaN or -inf entries: SciPy raises; report and assign nothing
    for (int r = tid; r < rows; r += blockDim.x) out[r] = 0;
    if (tid == 0 && status) status[p] = DSKD_EINFEASIBLE;
    return;
  }
  const int j = tid;
  const bool col_on = j < nc;
  double v = 0.0;
  for (int cur = 0; cur < nr; ++cur) {
    double shortest = CUDART_INF;
    bool scanned = false;
    int pos = nc - 1 - j;                         // remaining[it] = nc - it - 1
    if (col_on) remaining[nc - 1 - j] = j;
    if (tid == 0) { s_i = cur; s_sink = -1; s_nrem = nc; s_min = 0.0; }
    __syncthreads();
    while (true) {
      const int i = s_i;
      const double min_val = s_min;
      LsapKey k;
      k.val = CUDART_INF; k.assigned = 1; k.pos = 0x7fffffff; k.col = -1;
      if (col_on && !scanned) {
        const double r = min_val + cost_at(i, j) - u[i] - v;
        if (r < shortest) { shortest = r; path[j] = i; }
        k.val = shortest; k.assigned = row4col[j] != -1 ? 1 : 0; k.pos = pos; k.col = j;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const LsapKey o = lsap_shfl(k, off);
        if (lsap_better(o, k)) k = o;
      }
      if (lane == 0) { red_val[warp] = k.val; red_i[warp * 3] = k.assigned; red_i[warp * 3 + 1] = k.pos; red_i[warp * 3 + 2] = k.col; }
      __syncthreads();
      if (warp == 0) {
        LsapKey b;
        b.val = CUDART_INF; b.assigned = 1; b.pos = 0x7fffffff; b.col = -1;
        if (lane < nwarps) { b.val = red_val[lane]; b.assigned = red_i[lane * 3]; b.pos = red_i[lane * 3 + 1]; b.col = red_i[lane * 3 + 2]; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const LsapKey o = lsap_shfl(b, off);
          if (lsap_better(o, b)) b = o;
        }
        if (lane == 0) {
          s_min = b.val;
          s_winner = b.col;
          if (b.col < 0 || b.val == CUDART_INF) {
            s_sink = -2;  // infeasible
          } else {
            if (row4col[b.col] == -1) s_sink = b.col;
            else s_i = row4col[b.col];
            const int last = remaining[s_nrem - 1];   // swap-removal: the last candidate takes the winner's slot
            remaining[b.pos] = last;
            s_moved_col = last;
            s_moved_pos = b.pos;
            s_nrem -= 1;
          }
        }
      }
      __syncthreads();
      const int sink = s_sink;
      if (sink == -2) break;
      if (j == s_winner) scanned = true;
      else if (j == s_moved_col) pos = s_moved_pos;
      if (sink >= 0) break;
    }
    if (s_sink == -2) {
      __syncthreads();
      for (int r = tid; r < rows; r += blockDim.x) out[r] = 0;
      if (tid == 0 && status) status[p] = DSKD_EINFEASIBLE;
      return;
    }
    // dual update (before the assignment changes): every scanned column but the sink was assigned; its row is in the tree
    const double min_val = s_min;
    if (col_on && scanned) {
      const double d = min_val - shortest;
      const int r4 = row4col[j];
      if (r4 != -1) u[r4] += d;
      v -= d;
    }
    if (tid == 0) u[cur] += min_val;
    __syncthreads();
    if (tid == 0) {  // augment along the alternating path
      int jj = s_sink;
      while (true) {
        const int i = path[jj];
        row4col[jj] = i;
        const int t = col4row[i];
        col4row[i] = jj;
        jj = t;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }
  // 1-based matched GT per query row of the ORIGINAL problem (gfl_hungarian_assigner.py:153-158)
  if (transpose) {
    if (col_on) out[j] = (int64_t)(row4col[j] + 1);      // column of the transposed problem = query
  } else {
    for (int r = tid; r < rows; r += blockDim.x) out[r] = (int64_t)(col4row[r] + 1);
  }
  if (tid == 0 && status) status[p] = DSKD_OK;
}

}  // namespace dskd

using namespace dskd;

extern "C" int dskd_lsap_batch_device(const float* d_cost, int32_t num_problems, int32_t N, int32_t rows, int32_t ld,
                                      const int32_t* d_gt_start, int32_t max_cols, int64_t* d_assigned_gt,
                                      int32_t* d_status, void* stream) {
  DSKD_REQUIRE(num_problems >= 0 && N > 0 && rows >= 0 && ld >= 0 && max_cols >= 0, "dskd_lsap_batch_device: bad sizes");
  if (num_problems == 0) return DSKD_OK;
  DSKD_REQUIRE(d_gt_start && d_assigned_gt && (d_cost || max_cols == 0), "dskd_lsap_batch_device: null pointer");
  DSKD_REQUIRE(max_cols <= ld || max_cols == 0, "dskd_lsap_batch_device: max_cols (%d) exceeds the row stride (%d)", max_cols, ld);
  const int nc_max = std::max(rows, max_cols), nr_max = std::max(1, std::min(rows, max_cols));
  DSKD_REQUIRE(nc_max <= 1024, "dskd_lsap_batch_device: max(rows, cols) = %d exceeds 1024 (use dskd_lsap_batch_f32)", nc_max);
  const int threads = std::max(32, (nc_max + 31) / 32 * 32);
  const size_t fixed = sizeof(double) * (nr_max + 32) + sizeof(int) * (3 * (size_t)nc_max + nr_max + 96);
  const size_t staged = fixed + sizeof(float) * (size_t)nr_max * nc_max;
  const int stage = staged <= 200 * 1024 ? 1 : 0;
  const size_t smem = stage ? staged : fixed;
  if (smem > 48 * 1024)
    DSKD_CUDA_OK(cudaFuncSetAttribute(lsap_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lsap_batch_kernel<<<num_problems, threads, smem, as_stream(stream)>>>(d_cost, N, rows, ld, d_gt_start, d_assigned_gt,
                                                                      d_status, stage);
  DSKD_LAUNCH_OK("lsap_batch_kernel");
  return DSKD_OK;
}
