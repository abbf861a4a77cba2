// Rectangular linear sum assignment, host side, float64.
//
// Replaces the call `scipy.optimize.linear_sum_assignment(cost)` at
// mmdet/core/bbox/assigners/gfl_hungarian_assigner.py:147 (SciPy itself is not vendored by the
// reference; 1.18.1 is what is installed next to it here).  SciPy's solver is the shortest
// augmenting path variant of Jonker-Volgenant described by D. F. Crouse, "On implementing 2D
// rectangular assignment algorithms", IEEE TAES 52(4), 2016.  This file re-implements that published
// algorithm with the same tie-breaking rules (unassigned column preferred among equal reduced costs,
// candidate columns scanned in descending-index-initialised order, rows > cols solved on the
// transpose) so the returned indices are identical to SciPy's; tests/test_lsap.py checks that on
// thousands of random, tied and degenerate matrices.
//
// Attribution: the operation order and variable roles (u, v, path, col4row, row4col, SR, SC, remaining) follow
// SciPy's `rectangular_lsap` (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp, BSD 3-Clause License,
// Copyright (c) 2019, PM Larsen; SciPy Developers) -- required for index-identical results; no SciPy source is
// included or linked.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

#include "../../include/dskd_b200.h"

namespace dskd {
void set_error(const char* fmt, ...);

namespace {

struct Workspace {
  std::vector<double> u, v, shortest, cost_t;
  std::vector<int64_t> path, col4row, row4col, remaining;
  std::vector<char> SR, SC;
};

// Grow one alternating tree from row `i0`; returns the sink column or -1 when infeasible.
int64_t augmenting_path(int64_t nc, const double* cost, Workspace& w, int64_t i0, double* p_min_val) {
  double min_val = 0.0;
  int64_t num_remaining = nc;
  for (int64_t it = 0; it < nc; ++it) w.remaining[it] = nc - it - 1;
  std::fill(w.SR.begin(), w.SR.end(), 0);
  std::fill(w.SC.begin(), w.SC.end(), 0);
  std::fill(w.shortest.begin(), w.shortest.end(), std::numeric_limits<double>::infinity());
  int64_t sink = -1;
  int64_t i = i0;
  while (sink == -1) {
    int64_t index = -1;
    double lowest = std::numeric_limits<double>::infinity();
    w.SR[i] = 1;
    const double* row = cost + i * nc;
    const double ui = w.u[i];
    for (int64_t it = 0; it < num_remaining; ++it) {
      const int64_t j = w.remaining[it];
      const double r = min_val + row[j] - ui - w.v[j];
      if (r < w.shortest[j]) {
        w.path[j] = i;
        w.shortest[j] = r;
      }
      // among equal candidates prefer one that is a fresh sink
      if (w.shortest[j] < lowest || (w.shortest[j] == lowest && w.row4col[j] == -1)) {
        lowest = w.shortest[j];
        index = it;
      }
    }
    min_val = lowest;
    if (min_val == std::numeric_limits<double>::infinity()) return -1;
    const int64_t j = w.remaining[index];
    if (w.row4col[j] == -1) sink = j;
    else i = w.row4col[j];
    w.SC[j] = 1;
    w.remaining[index] = w.remaining[--num_remaining];
  }
  *p_min_val = min_val;
  return sink;
}

// cost: row-major [nr, nc]; a/b receive min(nr, nc) pairs sorted by a.
int solve(int64_t nr, int64_t nc, const double* cost_in, Workspace& w, int64_t* a, int64_t* b) {
  if (nr == 0 || nc == 0) return DSKD_OK;
  const bool transpose = nc < nr;
  const double* cost = cost_in;
  if (transpose) {
    w.cost_t.resize((size_t)nr * nc);
    for (int64_t i = 0; i < nr; ++i)
      for (int64_t j = 0; j < nc; ++j) w.cost_t[j * nr + i] = cost_in[i * nc + j];
    std::swap(nr, nc);
    cost = w.cost_t.data();
  }
  for (int64_t k = 0; k < nr * nc; ++k)
    if (std::isnan(cost[k]) || cost[k] == -std::numeric_limits<double>::infinity()) return DSKD_EINFEASIBLE;
  w.u.assign(nr, 0.0);
  w.v.assign(nc, 0.0);
  w.shortest.assign(nc, 0.0);
  w.path.assign(nc, -1);
  w.col4row.assign(nr, -1);
  w.row4col.assign(nc, -1);
  w.SR.assign(nr, 0);
  w.SC.assign(nc, 0);
  w.remaining.assign(nc, 0);
  for (int64_t cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    const int64_t sink = augmenting_path(nc, cost, w, cur, &min_val);
    if (sink < 0) return DSKD_EINFEASIBLE;
    // dual update
    w.u[cur] += min_val;
    for (int64_t i = 0; i < nr; ++i)
      if (w.SR[i] && i != cur) w.u[i] += min_val - w.shortest[w.col4row[i]];
    for (int64_t j = 0; j < nc; ++j)
      if (w.SC[j]) w.v[j] -= min_val - w.shortest[j];
    // augment along the alternating path
    int64_t j = sink;
    while (true) {
      const int64_t i = w.path[j];
      w.row4col[j] = i;
      std::swap(w.col4row[i], j);
      if (i == cur) break;
    }
  }
  if (transpose) {
    std::vector<int64_t> order(nr);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(),
              [&](int64_t x, int64_t y) { return w.col4row[x] < w.col4row[y]; });
    for (int64_t k = 0; k < nr; ++k) {
      a[k] = w.col4row[order[k]];
      b[k] = order[k];
    }
  } else {
    for (int64_t i = 0; i < nr; ++i) {
      a[i] = i;
      b[i] = w.col4row[i];
    }
  }
  return DSKD_OK;
}

}  // namespace
}  // namespace dskd

extern "C" int dskd_lsap_f64(const double* h_cost, int32_t rows, int32_t cols, int64_t* h_row_ind,
                             int64_t* h_col_ind) {
  if (rows < 0 || cols < 0 || ((rows > 0 && cols > 0) && (!h_cost || !h_row_ind || !h_col_ind))) {
    dskd::set_error("dskd_lsap_f64: bad arguments");
    return DSKD_EINVAL;
  }
  dskd::Workspace w;
  const int rc = dskd::solve(rows, cols, h_cost, w, h_row_ind, h_col_ind);
  if (rc == DSKD_EINFEASIBLE) dskd::set_error("dskd_lsap_f64: cost matrix is infeasible or has NaN / -inf entries");
  return rc;
}

extern "C" int dskd_lsap_batch_f32(const float* h_cost, int32_t num_problems, int32_t rows, int32_t ld,
                                   const int32_t* h_cols, int64_t* h_assigned_gt, int32_t num_threads) {
  if (num_problems < 0 || rows < 0 || ld < 0 || (num_problems > 0 && (!h_cols || !h_assigned_gt || (!h_cost && ld > 0)))) {
    dskd::set_error("dskd_lsap_batch_f32: bad arguments");
    return DSKD_EINVAL;
  }
  if (num_problems == 0) return DSKD_OK;
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, (int)num_problems));
  std::atomic<int> next{0};
  std::atomic<int> status{DSKD_OK};
  auto worker = [&]() {
    dskd::Workspace w;
    std::vector<double> cost;
    std::vector<int64_t> ri, ci;
    for (;;) {
      const int p = next.fetch_add(1);
      if (p >= num_problems) break;
      int64_t* out = h_assigned_gt + (int64_t)p * rows;
      std::fill(out, out + rows, (int64_t)0);
      const int cols = h_cols[p];
      if (cols <= 0 || rows == 0) continue;
      if (cols > ld) { status.store(DSKD_EINVAL); continue; }
      // SciPy converts the fp32 cost to float64 before solving (gfl_hungarian_assigner.py:143-147)
      cost.resize((size_t)rows * cols);
      const float* src = h_cost + (int64_t)p * rows * ld;
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) cost[(size_t)r * cols + c] = (double)src[(int64_t)r * ld + c];
      const int k = std::min(rows, cols);
      ri.resize(k);
      ci.resize(k);
      const int rc = dskd::solve(rows, cols, cost.data(), w, ri.data(), ci.data());
      if (rc != DSKD_OK) { status.store(rc); continue; }
      for (int t = 0; t < k; ++t) out[ri[t]] = ci[t] + 1;
    }
  };
  if (nt == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  const int rc = status.load();
  if (rc != DSKD_OK) dskd::set_error("dskd_lsap_batch_f32: a problem was infeasible / malformed (status %d)", rc);
  return rc;
}
