// og_tree.hpp -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
//
// CPU restatement of kd_tree.ml and interpolate_pdf.ml.  The reference keeps
// OCaml lists of objects in every Cell (kd_tree.ml:173); here a node is the
// range perm[begin, end) of one index array that is stably partitioned in
// place, which yields the same lists in the same order.  Nodes are numbered
// breadth first (children adjacent: left, left + 1), the numbering the GPU
// build uses, so the flat arrays can be compared element by element.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

#include "og_rng.hpp"

namespace og {

// Pervasives.compare on floats (kd_tree.ml:88-91): -0.0 = +0.0; NaN inputs are
// rejected up front (documented deviation).
inline int fcompare(double a, double b) { return a < b ? -1 : (a > b ? 1 : 0); }

struct Tree {
  int D = 0;
  int64_t N = 0;
  int min_split = 2;
  std::vector<double> pts;        // [N][D]
  std::vector<double> low, high;  // root box (caller supplied)
  std::vector<int32_t> perm;
  std::vector<int32_t> begin, end, dim, left;
  std::vector<double> split;
  int nlevels = 0;

  const double *pt(int64_t i) const { return &pts[(size_t)i * D]; }
  int64_t nnodes() const { return (int64_t)begin.size(); }

  int32_t add_node(int32_t b, int32_t e) {
    begin.push_back(b); end.push_back(e); dim.push_back(-1); left.push_back(-1);
    split.push_back(0.0);
    return (int32_t)begin.size() - 1;
  }

  // kd_tree.ml:155-175, one node.  Returns true if the node was split.
  bool split_node(int32_t id) {
    int32_t b = begin[id], e = end[id];
    int32_t n = e - b;
    if (n <= 1) return false;                       // :157-158
    if (n < min_split) return false;                // truncation (not in the reference)
    {                                               // :159-160 all coordinates equal
      const double *x0 = pt(perm[b]);
      bool all_eq = true;
      for (int32_t k = b + 1; k < e && all_eq; ++k) {
        const double *y = pt(perm[k]);
        for (int d = 0; d < D; ++d) if (fcompare(x0[d], y[d]) != 0) { all_eq = false; break; }
      }
      if (all_eq) return false;
    }
    int32_t i = n / 2;                              // :162-163
    // bounds_of_objects :96-110
    std::vector<double> l(pt(perm[b]), pt(perm[b]) + D), h(l);
    for (int32_t k = b + 1; k < e; ++k) {
      const double *c = pt(perm[k]);
      for (int d = 0; d < D; ++d) { if (c[d] < l[d]) l[d] = c[d]; if (c[d] > h[d]) h[d] = c[d]; }
    }
    // longest_dim :120-130 (first strictly largest spread)
    int sd = -1; double dx_max = -std::numeric_limits<double>::infinity();
    for (int d = 0; d < D; ++d) { double dx = h[d] - l[d]; if (dx > dx_max) { sd = d; dx_max = dx; } }
    // find_ith :69-86 -- randomized quickselect; its RESULT is the i-th order
    // statistic, so the Random.int draws are not emulated.
    std::vector<double> keys(n);
    for (int32_t k = 0; k < n; ++k) keys[k] = pt(perm[b + k])[sd];
    std::nth_element(keys.begin(), keys.begin() + i, keys.end());
    double pvt = keys[i];
    // List.partition (<= pvt) :168, order preserved
    auto mid = std::stable_partition(perm.begin() + b, perm.begin() + e,
                                     [&](int32_t o) { return fcompare(pt(o)[sd], pvt) <= 0; });
    int32_t nl = (int32_t)(mid - (perm.begin() + b));
    if (nl == n) {                                  // adjust_for_empty_split :150-152
      double mx = pt(perm[b])[sd];
      for (int32_t k = b + 1; k < e; ++k) if (fcompare(pt(perm[k])[sd], mx) > 0) mx = pt(perm[k])[sd];
      mid = std::stable_partition(perm.begin() + b, perm.begin() + e,
                                  [&](int32_t o) { return fcompare(pt(o)[sd], mx) < 0; });
      nl = (int32_t)(mid - (perm.begin() + b));
    }
    if (nl == 0 || nl == n) throw std::runtime_error("kd_tree: empty side after adjust");
    // find_max lte / find_min gt :170-171
    double lt_bound = pt(perm[b])[sd];
    for (int32_t k = b + 1; k < b + nl; ++k) if (fcompare(pt(perm[k])[sd], lt_bound) > 0) lt_bound = pt(perm[k])[sd];
    double gt_bound = pt(perm[b + nl])[sd];
    for (int32_t k = b + nl + 1; k < e; ++k) if (fcompare(pt(perm[k])[sd], gt_bound) < 0) gt_bound = pt(perm[k])[sd];
    double x = 0.5 * (lt_bound + gt_bound);         // split_bounds :113
    dim[id] = sd; split[id] = x;
    int32_t L = add_node(b, b + nl);
    add_node(b + nl, e);
    left[id] = L;
    return true;
  }

  void build() {
    perm.resize(N);
    for (int64_t i = 0; i < N; ++i) perm[i] = (int32_t)i;
    begin.clear(); end.clear(); dim.clear(); left.clear(); split.clear();
    nlevels = 0;
    if (N == 0) return;                             // :157 Empty
    add_node(0, (int32_t)N);
    int32_t lvl_b = 0, lvl_e = 1;
    while (lvl_b < lvl_e) {
      ++nlevels;
      for (int32_t id = lvl_b; id < lvl_e; ++id) split_node(id);
      lvl_b = lvl_e; lvl_e = (int32_t)begin.size();
    }
  }

  // kd_tree.ml:177-182
  static double bounds_volume(const double *lo, const double *hi, int D) {
    double v = 1.0;
    for (int i = 0; i < D; ++i) v = v * (hi[i] - lo[i]);
    return v + 0.0;
  }

  // interpolate_pdf.ml:88-109 (find_cell) and :121-133 (high level: stop at
  // the first cell with <= nstop objects).  lo/hi receive the cell's box.
  // Returns -1 where the reference raises (empty tree / Empty child reached).
  int32_t find_cell(const double *q, int nstop, double *lo, double *hi) const {
    if (N == 0) return -1;
    for (int d = 0; d < D; ++d) { lo[d] = low[d]; hi[d] = high[d]; }
    int32_t id = 0;
    for (;;) {
      if (nstop > 0) {
        if (end[id] - begin[id] <= nstop) return id;
        if (left[id] < 0) return -1;                // :124 Failure "encountered empty tree"
      } else if (left[id] < 0) return id;           // :103
      int sd = dim[id]; double s = split[id];
      // in_tree pt left = in_bounds pt low high' with high'.(dim) = s  (:88-99)
      double save = hi[sd]; hi[sd] = s;
      int i = 0;
      while (i < D && q[i] >= lo[i] && q[i] <= hi[i]) ++i;
      if (i == D) { id = left[id]; }
      else { hi[sd] = save; lo[sd] = s; id = left[id] + 1; }
    }
  }

  // interpolate_pdf.ml:135-142,144-159
  double jump_prob(const double *q, int nstop, int32_t *node_out) const {
    std::vector<double> lo(D), hi(D);
    int32_t id = find_cell(q, nstop, lo.data(), hi.data());
    if (node_out) *node_out = id;
    if (id < 0) return std::numeric_limits<double>::quiet_NaN();
    double nobjs = (double)(end[id] - begin[id]);
    double v = bounds_volume(lo.data(), hi.data(), D);
    if (nstop > 0) return nobjs / ((double)N * v);  // :154 (ncell / (npts * v))
    return nobjs / (v * (double)N);                 // :142 (nobjs / (v * n))
  }

  // interpolate_pdf.ml:114-119,121-133 with random_in_volume :80-86
  bool draw(Rng &r, int nstop, double *out) const {
    std::vector<double> lo(D), hi(D);
    int64_t k = (int64_t)r.below((uint64_t)N);
    int32_t id = find_cell(pt(k), nstop, lo.data(), hi.data());
    if (id < 0) return false;
    for (int i = 0; i < D; ++i) out[i] = lo[i] + (hi[i] - lo[i]) * r.uniform();
    return true;
  }
};

}  // namespace og
