// ellipse.cu -- Ellipse (ellipse.ml:36-173): enclosing ellipses of point clouds and the recursive ellipse tree.
//
// SURVEY.md 8f rank 4 ("next"): the ellipsoidal decomposition a nested sampler would use beyond MCMC inside the
// constraint.  The reference builds the tree by recursion over OCaml arrays with one LAPACK eigen-decomposition per
// node (Lacaml syevr, ellipse.ml:58-61).  Here the tree is built level by level with the rows kept in node order:
//
//   mean      Ellipse.center (:36-46)          per-node column sums: (node, slice) CTAs stream their rows through
//   sigma     Ellipse.sigma2 (:48-66)           shared memory, partials combined in a fixed order (deterministic)
//   eigen     Ellipse.eigensystem (:58-61)      cyclic Jacobi, one warp per node, matrix in shared memory; eigenvalues
//                                               ascending like LAPACK, eigenvector sign fixed (largest component > 0)
//   range     max_elliptical_range (:75-81)     thread per point, exact maximum -> rescale_ellipse (:83-86)
//   split     widest_dimension (:134-143) and the stable partition  coord.(split) < center.(split)  (:151-156):
//             flags -> one exclusive scan over all positions -> rows, ids and node numbers scattered
//   circle    ellipse_circumcircle / union_circumcircles (:112-132,:159-167) bottom-up, one thread per node
//
// Quirks kept because the code says so: `split` is the index of the widest AXIS used as a COORDINATE index (:151-156);
// the circumcircle "radius" is the largest scaled eigenvalue (:131-132) and a union's radius is r1 + r12 + r2 (:122).
// A node whose points all fall on one side of its centre recurses forever in the reference (Stack_overflow); here it
// is MG_EFAIL.  Sums are taken per slice and combined in slice order, not in one left fold with a division per term
// as :41-44,:56-64 do: centre and covariance agree with the reference's arithmetic to ~1e-15 relative, stated in the
// tests; the tree's structure depends on comparisons with the centre only.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

struct mg_ellipse_tree {
  mg_ctx *ctx = nullptr;
  int64_t N = 0; int32_t D = 0; int64_t nnodes = 0; int32_t nlevels = 0;
  // device arrays, [nnodes] unless noted
  int32_t *left = nullptr, *right = nullptr, *begin = nullptr, *end = nullptr, *perm = nullptr /* [N] */;
  double *center = nullptr /* [nnodes][D] */, *axes = nullptr /* [nnodes][D] */, *ori = nullptr /* [nnodes][D][D] */;
  double *cc_center = nullptr /* [nnodes][D] */, *cc_radius = nullptr;
};

namespace mg {

constexpr int EL_TB = 256;        // threads per CTA
constexpr int EL_CH = 128;        // rows per shared-memory chunk
constexpr int EL_DMAX = 32;
constexpr int EL_ACC = 3;         // accumulators per thread: ceil(D (D + 1) / 2 / 256) at D = 32
constexpr int EL_SMAX = 592;      // slices per node (four CTAs per SM when a level has one node)
constexpr int EL_SMALL = 1024;    // a node of at most this many doubles (points x D) is finished by ONE warp (el_small_kernel)
constexpr int EL_SWARPS = 4;      // warps (nodes) per CTA of that kernel

struct ElNodes {                  // growing node table (device)
  int32_t *left, *right, *begin, *end;
  double *center, *axes, *ori, *evals;
  int32_t *split;
};

__device__ __forceinline__ double neg_inf_d() { return __longlong_as_double(0xFFF0000000000000ll); }
__device__ __forceinline__ int el_pairs(int D) { return D * (D + 1) / 2; }

// slice s of S of the rows [b, e)
__device__ __forceinline__ void el_slice(int b, int e, int s, int S, int *sb, int *se) {
  const int64_t n = e - b;
  *sb = b + (int)(n * s / S);
  *se = b + (int)(n * (s + 1) / S);
}

__device__ __forceinline__ bool el_is_small(const ElNodes &nd, int node, int D) { return (long long)(nd.end[node] - nd.begin[node]) * D <= EL_SMALL; }

// ---- Ellipse.center (:36-46): per-(node, slice) column sums --------------------------------------------------------------
__global__ void __launch_bounds__(EL_TB)
el_mean_partial_kernel(const double *__restrict__ rows, int D, ElNodes nd, int lb, int S, double *__restrict__ part) {
  extern __shared__ double el_sm[];                 // [EL_CH][D] then [G][D]
  const int node = lb + blockIdx.x, s = blockIdx.y;
  if (el_is_small(nd, node, D)) return;             // finished by el_small_kernel
  int sb, se;
  el_slice(nd.begin[node], nd.end[node], s, S, &sb, &se);
  const int G = EL_TB / D;                          // row groups: thread (g, j) sums rows g, g + G, ... of column j
  const int g = threadIdx.x / D, j = threadIdx.x - g * D;
  const bool worker = g < G;
  double acc = 0.0;
  for (int c0 = sb; c0 < se; c0 += EL_CH) {
    const int cnt = min(EL_CH, se - c0);
    __syncthreads();
    const double *src = rows + (int64_t)c0 * D;
    for (int k = threadIdx.x; k < cnt * D; k += EL_TB) el_sm[k] = src[k];
    __syncthreads();
    if (worker)
      for (int i = g; i < cnt; i += G) acc = acc + el_sm[i * D + j];
  }
  double *red = el_sm + EL_CH * D;
  __syncthreads();
  if (worker) red[g * D + j] = acc;
  __syncthreads();
  if (threadIdx.x < D) {
    double t = 0.0;
    for (int gg = 0; gg < G; ++gg) t = t + red[gg * D + threadIdx.x];
    part[((int64_t)blockIdx.x * S + s) * D + threadIdx.x] = t;
  }
}

__global__ void el_mean_finish_kernel(int D, ElNodes nd, int lb, int nn, int S, const double *__restrict__ part) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nn * D) return;
  const int r = k / D, j = k - r * D, node = lb + r;
  if (el_is_small(nd, node, D)) return;
  double t = 0.0;
  for (int s = 0; s < S; ++s) t = t + part[((int64_t)r * S + s) * D + j];
  nd.center[(int64_t)node * D + j] = t / (double)(nd.end[node] - nd.begin[node]);
}

// ---- Ellipse.sigma2 (:48-66): per-(node, slice) sums of dx_j dx_k, j <= k ---------------------------------------------
__global__ void __launch_bounds__(EL_TB)
el_cov_partial_kernel(const double *__restrict__ rows, int D, ElNodes nd, int lb, int S, double *__restrict__ part) {
  extern __shared__ double el_sm[];                 // [EL_CH][D] centred rows, then the reduction scratch [G][P]
  __shared__ double mu[EL_DMAX];
  __shared__ unsigned char pj[EL_DMAX * (EL_DMAX + 1) / 2], pk[EL_DMAX * (EL_DMAX + 1) / 2];
  const int node = lb + blockIdx.x, s = blockIdx.y;
  if (el_is_small(nd, node, D)) return;
  const int P = el_pairs(D);
  if (threadIdx.x < D) mu[threadIdx.x] = nd.center[(int64_t)node * D + threadIdx.x];
  for (int q = threadIdx.x; q < P; q += EL_TB) {     // pair q -> (j, k), row-major upper triangle
    int j = 0, rem = q;
    while (rem >= D - j) { rem -= D - j; ++j; }
    pj[q] = (unsigned char)j; pk[q] = (unsigned char)(j + rem);
  }
  int sb, se;
  el_slice(nd.begin[node], nd.end[node], s, S, &sb, &se);
  const int G = P >= EL_TB ? 1 : EL_TB / P;          // row groups when the pairs do not fill the CTA
  double acc[EL_ACC];
#pragma unroll
  for (int a = 0; a < EL_ACC; ++a) acc[a] = 0.0;
  for (int c0 = sb; c0 < se; c0 += EL_CH) {
    const int cnt = min(EL_CH, se - c0);
    __syncthreads();
    const double *src = rows + (int64_t)c0 * D;
    for (int k = threadIdx.x; k < cnt * D; k += EL_TB) { const int j = k % D; el_sm[k] = src[k] - mu[j]; }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < EL_ACC; ++a) {
      const int w = threadIdx.x + a * EL_TB;
      if (w >= P * G) break;
      const int g = w / P, q = w - g * P;
      const int j = pj[q], k = pk[q];
      double t = acc[a];
      for (int i = g; i < cnt; i += G) t = t + el_sm[i * D + j] * el_sm[i * D + k];
      acc[a] = t;
    }
  }
  double *red = el_sm + EL_CH * D;
  __syncthreads();
#pragma unroll
  for (int a = 0; a < EL_ACC; ++a) { const int w = threadIdx.x + a * EL_TB; if (w < P * G) red[w] = acc[a]; }
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += EL_TB) {
    double t = 0.0;
    for (int g = 0; g < G; ++g) t = t + red[g * P + q];
    part[((int64_t)blockIdx.x * S + s) * P + q] = t;
  }
}

// cyclic Jacobi on the symmetric matrix A (leading dimension LD) by one warp; V receives the eigenvectors (columns)
__device__ __forceinline__ void el_jacobi_warp(double *A, double *V, int D, int LD, int lane) {
  for (int q = lane; q < D * D; q += 32) { const int i = q / D, j = q - i * D; V[i * LD + j] = (i == j) ? 1.0 : 0.0; }
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int q = lane; q < D * D; q += 32) { const int i = q / D, j = q - i * D; const double a = A[i * LD + j]; if (i == j) dg += a * a; else off += a * a; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
    if (off <= 1e-33 * dg || off == 0.0) break;
    for (int p = 0; p < D - 1; ++p)
      for (int q = p + 1; q < D; ++q) {
        const double apq = A[p * LD + q];
        if (apq == 0.0) continue;                   // warp-uniform: every lane reads the same word
        const double app = A[p * LD + p], aqq = A[q * LD + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        __syncwarp();
        const int k = lane;
        double akp = 0.0, akq = 0.0, vkp = 0.0, vkq = 0.0;
        if (k < D) { akp = A[k * LD + p]; akq = A[k * LD + q]; vkp = V[k * LD + p]; vkq = V[k * LD + q]; }
        __syncwarp();
        if (k < D) {
          V[k * LD + p] = c * vkp - sn * vkq; V[k * LD + q] = sn * vkp + c * vkq;
          if (k != p && k != q) {
            const double np_ = c * akp - sn * akq, nq_ = sn * akp + c * akq;
            A[k * LD + p] = np_; A[p * LD + k] = np_; A[k * LD + q] = nq_; A[q * LD + k] = nq_;
          }
        }
        if (lane == 0) { A[p * LD + p] = app - t * apq; A[q * LD + q] = aqq + t * apq; A[p * LD + q] = 0.0; A[q * LD + p] = 0.0; }
        __syncwarp();
      }
  }
}
// ascending order (ties by index), sign: the component of largest magnitude (first on ties) is positive; written to the
// node table and, when given, to a shared-memory copy (ev_s[D], ori_s[D][D])
__device__ __forceinline__ void el_eigen_store(const double *A, const double *V, double *w, int D, int LD, int lane, const ElNodes &nd,
                                               int node, double *ev_s, double *ori_s) {
  if (lane < D) w[lane] = A[lane * LD + lane];
  __syncwarp();
  if (lane < D) {
    const double me = w[lane];
    int rank = 0;
    for (int i = 0; i < D; ++i) { const double o = w[i]; rank += (o < me || (o == me && i < lane)) ? 1 : 0; }
    double big = 0.0; double sgn = 1.0;
    for (int i = 0; i < D; ++i) { const double v = V[i * LD + lane]; if (fabs(v) > big) { big = fabs(v); sgn = v < 0.0 ? -1.0 : 1.0; } }
    nd.evals[(int64_t)node * D + rank] = me;
    if (ev_s) ev_s[rank] = me;
    for (int i = 0; i < D; ++i) {
      const double v = sgn * V[i * LD + lane];
      nd.ori[((int64_t)node * D + i) * D + rank] = v;
      if (ori_s) ori_s[i * D + rank] = v;
    }
  }
  __syncwarp();
}

// ---- Ellipse.eigensystem (:58-61): cyclic Jacobi, one warp per node -------------------------------------------------------
// Output like LAPACK's syevr through Lacaml: eigenvalues ascending, ori[i][j] = component i of eigenvector j.
__global__ void __launch_bounds__(32)
el_eigen_kernel(int D, ElNodes nd, int lb, int S, const double *__restrict__ part) {
  extern __shared__ double el_sm[];                 // A[D][D+1], V[D][D+1], w[D]
  const int r = blockIdx.x, node = lb + r, lane = threadIdx.x;
  if (el_is_small(nd, node, D)) return;
  const int P = el_pairs(D), LD = D + 1;
  double *A = el_sm, *V = A + D * LD, *w = V + D * LD;
  const double nf = (double)(nd.end[node] - nd.begin[node]);
  for (int q = lane; q < P; q += 32) {
    int j = 0, rem = q;
    while (rem >= D - j) { rem -= D - j; ++j; }
    const int k = j + rem;
    double t = 0.0;
    for (int s = 0; s < S; ++s) t = t + part[((int64_t)r * S + s) * P + q];
    t = t / nf;
    A[j * LD + k] = t; A[k * LD + j] = t;
  }
  el_jacobi_warp(A, V, D, LD, lane);
  el_eigen_store(A, V, w, D, LD, lane, nd, node, nullptr, nullptr);
}

// ---- Ellipse.elliptical_range (:63-73) -------------------------------------------------------------------------------------
__device__ __forceinline__ double el_range(const double *pt, const double *c, const double *a, const double *ori, int D) {
  double r = 0.0;
  for (int j = 0; j < D; ++j) {
    double d = 0.0;
    for (int i = 0; i < D; ++i) d = d + (pt[i] - c[i]) * ori[i * D + j];
    r = r + d * d / a[j];
  }
  return r + 0.0;
}

// max_elliptical_range (:75-81) with the unscaled ellipse (axes = eigenvalues): per-(node, slice) maxima
__global__ void __launch_bounds__(EL_TB)
el_range_partial_kernel(const double *__restrict__ rows, int D, ElNodes nd, int lb, int S, double *__restrict__ part) {
  extern __shared__ double el_sm[];                 // c[D], a[D], ori[D][D], then the chunk [EL_CH][D]
  __shared__ double wmax[EL_TB / 32];
  const int node = lb + blockIdx.x, s = blockIdx.y;
  if (el_is_small(nd, node, D)) return;
  double *c = el_sm, *a = c + D, *ori = a + D, *chunk = ori + D * D;
  for (int k = threadIdx.x; k < D; k += EL_TB) { c[k] = nd.center[(int64_t)node * D + k]; a[k] = nd.evals[(int64_t)node * D + k]; }
  for (int k = threadIdx.x; k < D * D; k += EL_TB) ori[k] = nd.ori[(int64_t)node * D * D + k];
  int sb, se;
  el_slice(nd.begin[node], nd.end[node], s, S, &sb, &se);
  double m = neg_inf_d();
  for (int c0 = sb; c0 < se; c0 += EL_CH) {
    const int cnt = min(EL_CH, se - c0);
    __syncthreads();
    const double *src = rows + (int64_t)c0 * D;
    for (int k = threadIdx.x; k < cnt * D; k += EL_TB) chunk[k] = src[k];
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += EL_TB) { const double r = el_range(chunk + i * D, c, a, ori, D); m = r > m ? r : m; }   // max: NaN never wins (:79)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < EL_TB / 32; ++k) m = wmax[k] > m ? wmax[k] : m;
    part[(int64_t)blockIdx.x * S + s] = m;
  }
}

// rescale_ellipse (:83-86) and widest_dimension (:134-143)
__global__ void el_rescale_kernel(int D, ElNodes nd, int lb, int nn, int S, const double *__restrict__ part, double dim_sf) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nn) return;
  const int node = lb + r;
  if (el_is_small(nd, node, D)) return;
  double rmax = neg_inf_d();
  for (int s = 0; s < S; ++s) { const double x = part[(int64_t)r * S + s]; rmax = x > rmax ? x : rmax; }
  int imax = -1; double amax = neg_inf_d();
  for (int j = 0; j < D; ++j) {
    const double a = nd.evals[(int64_t)node * D + j] * dim_sf * rmax;
    nd.axes[(int64_t)node * D + j] = a;
    if (a > amax) { amax = a; imax = j; }
  }
  nd.split[node] = imax;
}

// ---- small nodes: enclosing_ellipse (:98-103) of a node of at most EL_SMALL doubles by ONE warp -----------------------------
// The deep levels of the tree hold most of its nodes (D + 1 .. a few dozen points each); a CTA per (node, slice) and
// five launches per node are wasteful there.  One warp stages the node's rows in shared memory and does centre,
// covariance, eigen-system, largest elliptical range and the rescaling in one go (same formulas; sums in a fixed order).
__global__ void __launch_bounds__(32 * EL_SWARPS)
el_small_kernel(const double *__restrict__ rows, int D, ElNodes nd, int lb, int nn, double dim_sf) {
  extern __shared__ double el_sm[];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int r = blockIdx.x * EL_SWARPS + wrp;
  if (r >= nn) return;
  const int node = lb + r;
  const int b = nd.begin[node], n = nd.end[node] - b;
  if ((long long)n * D > EL_SMALL) return;          // the (node, slice) kernels take it
  const int LD = D + 1, P = el_pairs(D);
  double *base = el_sm + (size_t)wrp * (EL_SMALL + 2 * D * LD + 4 * D + D * D + 64);
  double *pt = base, *A = pt + EL_SMALL, *V = A + D * LD, *w = V + D * LD, *mu = w + D, *ev = mu + D, *ori = ev + D, *red = ori + D * D;
  const double *src = rows + (int64_t)b * D;
  for (int k = lane; k < n * D; k += 32) pt[k] = src[k];
  __syncwarp();
  const double nf = (double)n;
  // centre (:36-46): lane (g, j) sums rows g, g + G, ... of column j; groups combined in order
  {
    const int G = 32 / D;                           // D <= 32: at least one group
    const int g = lane / D, j = lane - g * D;
    const bool worker = g < G;
    double acc = 0.0;
    if (worker) for (int i = g; i < n; i += G) acc = acc + pt[i * D + j];
    if (worker) red[g * D + j] = acc;
    __syncwarp();
    if (lane < D) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t = t + red[gg * D + lane];
      t = t / nf;
      mu[lane] = t; nd.center[(int64_t)node * D + lane] = t;
    }
    __syncwarp();
  }
  for (int k = lane; k < n * D; k += 32) pt[k] = pt[k] - mu[k % D];    // centred rows
  __syncwarp();
  // covariance (:48-66): one pair (j <= k) per lane and round, rows in order
  for (int q = lane; q < P; q += 32) {
    int j = 0, rem = q;
    while (rem >= D - j) { rem -= D - j; ++j; }
    const int k = j + rem;
    double t = 0.0;
    for (int i = 0; i < n; ++i) t = t + pt[i * D + j] * pt[i * D + k];
    t = t / nf;
    A[j * LD + k] = t; A[k * LD + j] = t;
  }
  __syncwarp();
  el_jacobi_warp(A, V, D, LD, lane);
  el_eigen_store(A, V, w, D, LD, lane, nd, node, ev, ori);
  // largest elliptical range (:75-81) of the centred rows against the unscaled ellipse, then rescale_ellipse (:83-86)
  double m = neg_inf_d();
  for (int i = lane; i < n; i += 32) {
    double rr = 0.0;
    for (int j = 0; j < D; ++j) {
      double d = 0.0;
      for (int k = 0; k < D; ++k) d = d + pt[i * D + k] * ori[k * D + j];
      rr = rr + d * d / ev[j];
    }
    rr = rr + 0.0;
    m = rr > m ? rr : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
  if (lane == 0) {
    int imax = -1; double amax = neg_inf_d();
    for (int j = 0; j < D; ++j) {
      const double a = ev[j] * dim_sf * m;
      nd.axes[(int64_t)node * D + j] = a;
      if (a > amax) { amax = a; imax = j; }
    }
    nd.split[node] = imax;
  }
}

// ---- the partition of a level (:151-156) ---------------------------------------------------------------------------------------
__global__ void el_flag_kernel(const double *__restrict__ rows, int64_t N, int D, ElNodes nd, const int32_t *__restrict__ seg,
                               int lb, int le, int32_t *__restrict__ flag) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > N) return;
  int f = 0;
  if (p < N) {
    const int node = seg[p];
    if (node >= lb && node < le) {
      const int sp = nd.split[node];
      if (sp >= 0) f = rows[p * D + sp] < nd.center[(int64_t)node * D + sp] ? 1 : 0;
    }
  }
  flag[p] = f;                                       // flag[N] = 0: the scan's last entry is the total
}

// children of the level's nodes: a side with fewer than D + 1 points is Empty (:157-158)
__global__ void el_children_count_kernel(int D, ElNodes nd, int lb, int nn, const int32_t *__restrict__ lscan,
                                         int32_t *__restrict__ has, int *__restrict__ stuck) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nn) return;
  const int node = lb + r;
  const int b = nd.begin[node], e = nd.end[node];
  const int nL = lscan[e] - lscan[b], nR = (e - b) - nL;
  if (nL == e - b || nR == e - b) *stuck = 1;        // the reference would recurse on the same points forever
  has[2 * r] = nL >= D + 1 ? 1 : 0;
  has[2 * r + 1] = nR >= D + 1 ? 1 : 0;
  const int big = (nL >= D + 1 && nL > nR) ? nL : (nR >= D + 1 ? nR : (nL >= D + 1 ? nL : 0));
  if (big > 0) atomicMax(stuck + 1, big);            // the largest node of the next level
}
__global__ void el_children_make_kernel(ElNodes nd, int lb, int nn, int next_base, const int32_t *__restrict__ lscan,
                                        const int32_t *__restrict__ has, const int32_t *__restrict__ hscan) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nn) return;
  const int node = lb + r;
  const int b = nd.begin[node], e = nd.end[node];
  const int nL = lscan[e] - lscan[b];
  int l = -1, rt = -1;
  if (has[2 * r]) { l = next_base + hscan[2 * r]; nd.begin[l] = b; nd.end[l] = b + nL; }
  if (has[2 * r + 1]) { rt = next_base + hscan[2 * r + 1]; nd.begin[rt] = b + nL; nd.end[rt] = e; }
  nd.left[node] = l; nd.right[node] = rt;
}

// rows, ids and node numbers to their places: stable on both sides (List.partition keeps the order, :145-148)
__global__ void __launch_bounds__(256)
el_scatter_kernel(const double *__restrict__ rows, double *__restrict__ rows_out, const int32_t *__restrict__ perm,
                  int32_t *__restrict__ perm_out, const int32_t *__restrict__ seg, int32_t *__restrict__ seg_out, int64_t N, int D,
                  ElNodes nd, int lb, int le, const int32_t *__restrict__ lscan, const int32_t *__restrict__ flag) {
  __shared__ int64_t dest_sh[256];
  const int64_t p0 = (int64_t)blockIdx.x * 256, p = p0 + threadIdx.x;
  int64_t dest = p;
  if (p < N) {
    const int node = seg[p];
    int child = -1;
    if (node >= lb && node < le) {
      const int b = nd.begin[node], e = nd.end[node];
      const int nL = lscan[e] - lscan[b], lr = lscan[p] - lscan[b];
      if (flag[p]) { dest = b + lr; child = nd.left[node]; }
      else { dest = (int64_t)b + nL + (p - b - lr); child = nd.right[node]; }
    }
    perm_out[dest] = perm[p];
    seg_out[dest] = child;
  }
  dest_sh[threadIdx.x] = dest;
  __syncthreads();
  const int cnt = (int)((N - p0 < 256) ? (N - p0) : 256);
  for (int k = threadIdx.x; k < cnt * D; k += 256) {           // coalesced reads of the tile's rows
    const int i = k / D, j = k - i * D;
    rows_out[dest_sh[i] * D + j] = rows[(p0 + i) * D + j];
  }
}

// ---- circumcircles (:106-132,:159-167), one level at a time from the deepest -----------------------------------------------
__device__ __forceinline__ double el_distance(const double *p1, const double *p2, int D) {
  double r = 0.0;
  for (int i = 0; i < D; ++i) { const double dx = p1[i] - p2[i]; r = r + dx * dx; }
  return sqrt(r);
}
// union_circumcircles cc1 cc2 -> (out, return radius); out may alias c1
__device__ __forceinline__ double el_union(const double *c1, double r1, const double *c2, double r2, double *out, int D) {
  const double r12 = el_distance(c1, c2, D);
  if (r12 + r2 < r1) { for (int i = 0; i < D; ++i) out[i] = c1[i]; return r1; }
  if (r12 + r1 < r2) { for (int i = 0; i < D; ++i) out[i] = c2[i]; return r2; }
  const double rnew = r1 + r12 + r2;
  const double mag = 0.5 * (r2 + r12 - r1) / r12;
  for (int i = 0; i < D; ++i) out[i] = c1[i] + mag * (c2[i] - c1[i]);
  return rnew;
}
__global__ void el_circum_kernel(int D, ElNodes nd, int lb, int nn, double *__restrict__ ccc, double *__restrict__ ccr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nn) return;
  const int node = lb + r;
  const double *ec = nd.center + (int64_t)node * D;
  double er = neg_inf_d();                                       // ellipse_circumcircle: fold max over the axes
  for (int j = 0; j < D; ++j) { const double a = nd.axes[(int64_t)node * D + j]; er = a > er ? a : er; }
  double *out = ccc + (int64_t)node * D;
  const int l = nd.left[node], rt = nd.right[node];
  double tmp[EL_DMAX];
  if (l < 0 && rt < 0) { for (int i = 0; i < D; ++i) out[i] = ec[i]; ccr[node] = er; return; }
  if (l >= 0 && rt >= 0) {
    const double ru = el_union(ccc + (int64_t)l * D, ccr[l], ccc + (int64_t)rt * D, ccr[rt], tmp, D);
    ccr[node] = el_union(tmp, ru, ec, er, out, D);
    return;
  }
  const int c = l >= 0 ? l : rt;
  ccr[node] = el_union(ccc + (int64_t)c * D, ccr[c], ec, er, out, D);
}

__global__ void el_iota_kernel(int32_t *perm, int32_t *seg, int64_t N) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < N) { perm[p] = (int32_t)p; seg[p] = 0; }
}

__global__ void el_range_points_kernel(const double *__restrict__ q, int64_t M, int D, const double *__restrict__ cao, double *__restrict__ out) {
  extern __shared__ double el_sm[];
  for (int k = threadIdx.x; k < 2 * D + D * D; k += blockDim.x) el_sm[k] = cao[k];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < M) out[p] = el_range(q + p * D, el_sm, el_sm + D, el_sm + 2 * D, D);
}

// the growing node table
struct ElTable {
  DevBuf<int32_t> left, right, begin, end, split;
  DevBuf<double> center, axes, ori, evals;
  int64_t cap = 0;
  ElNodes view() const { return ElNodes{left.get(), right.get(), begin.get(), end.get(), center.get(), axes.get(), ori.get(), evals.get(), split.get()}; }
};
template <class T>
static cudaError_t el_grow(DevBuf<T> &b, size_t old_n, size_t new_n, cudaStream_t s) {
  DevBuf<T> nb;
  cudaError_t e = nb.alloc(new_n, s);
  if (e != cudaSuccess) return e;
  if (old_n && b.get()) e = cudaMemcpyAsync(nb.get(), b.get(), old_n * sizeof(T), cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return e;
  std::swap(b.p, nb.p); std::swap(b.n, nb.n); std::swap(b.s, nb.s);
  return cudaSuccess;
}
static int el_reserve(mg_ctx *ctx, ElTable &t, int64_t used, int64_t want, int D, cudaStream_t s) {
  if (want <= t.cap) return MG_OK;
  const int64_t nc = std::max<int64_t>(want, t.cap * 2);
  MG_CUDA(ctx, el_grow(t.left, (size_t)used, (size_t)nc, s)); MG_CUDA(ctx, el_grow(t.right, (size_t)used, (size_t)nc, s));
  MG_CUDA(ctx, el_grow(t.begin, (size_t)used, (size_t)nc, s)); MG_CUDA(ctx, el_grow(t.end, (size_t)used, (size_t)nc, s));
  MG_CUDA(ctx, el_grow(t.split, (size_t)used, (size_t)nc, s));
  MG_CUDA(ctx, el_grow(t.center, (size_t)used * D, (size_t)nc * D, s)); MG_CUDA(ctx, el_grow(t.axes, (size_t)used * D, (size_t)nc * D, s));
  MG_CUDA(ctx, el_grow(t.evals, (size_t)used * D, (size_t)nc * D, s));
  MG_CUDA(ctx, el_grow(t.ori, (size_t)used * D * D, (size_t)nc * D * D, s));
  t.cap = nc;
  return MG_OK;
}

// enclosing_ellipse (:98-103) of every node of a level: centre, covariance, eigen-system, rescaling
static int el_level_ellipses(mg_ctx *ctx, const double *rows, int D, const ElTable &t, int lb, int nn, double sf, DevBuf<double> &part,
                             int64_t max_points) {
  cudaStream_t s = ctx->stream;
  const double dim_sf = pow(sf, 1.0 / (double)D);              // :85, host libm like the reference
  {
    const size_t sm_small = sizeof(double) * (size_t)EL_SWARPS * (EL_SMALL + 2 * D * (D + 1) + 4 * D + D * D + 64);
    if (sm_small > 48 * 1024) MG_CUDA(ctx, cudaFuncSetAttribute(el_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
    el_small_kernel<<<(unsigned)((nn + EL_SWARPS - 1) / EL_SWARPS), 32 * EL_SWARPS, sm_small, s>>>(rows, D, t.view(), lb, nn, dim_sf);
    MG_CHECK_LAUNCH(ctx);
    if (max_points * D <= EL_SMALL) return MG_OK;             // every node of the level was a small one
  }
  const int P = D * (D + 1) / 2;
  const int S = (int)std::max<int64_t>(1, std::min<int64_t>(EL_SMAX, (4 * (int64_t)ctx->sm_count + nn - 1) / nn));
  const size_t need = (size_t)nn * S * std::max(P, D);
  if (part.n < need) MG_CUDA(ctx, part.alloc(need, s));
  const ElNodes nd = t.view();
  const dim3 grid((unsigned)nn, (unsigned)S);
  const int Gm = EL_TB / D, Gc = P >= EL_TB ? 1 : EL_TB / P;
  const size_t sm_mean = sizeof(double) * ((size_t)EL_CH * D + (size_t)Gm * D);
  const size_t sm_cov = sizeof(double) * ((size_t)EL_CH * D + (size_t)Gc * P);
  const size_t sm_eig = sizeof(double) * ((size_t)2 * D * (D + 1) + D);
  const size_t sm_rng = sizeof(double) * ((size_t)2 * D + (size_t)D * D + (size_t)EL_CH * D);
  if (sm_cov > 48 * 1024) MG_CUDA(ctx, cudaFuncSetAttribute(el_cov_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cov));
  el_mean_partial_kernel<<<grid, EL_TB, sm_mean, s>>>(rows, D, nd, lb, S, part.get());
  MG_CHECK_LAUNCH(ctx);
  el_mean_finish_kernel<<<(unsigned)(((int64_t)nn * D + 255) / 256), 256, 0, s>>>(D, nd, lb, nn, S, part.get());
  MG_CHECK_LAUNCH(ctx);
  el_cov_partial_kernel<<<grid, EL_TB, sm_cov, s>>>(rows, D, nd, lb, S, part.get());
  MG_CHECK_LAUNCH(ctx);
  el_eigen_kernel<<<(unsigned)nn, 32, sm_eig, s>>>(D, nd, lb, S, part.get());
  MG_CHECK_LAUNCH(ctx);
  el_range_partial_kernel<<<grid, EL_TB, sm_rng, s>>>(rows, D, nd, lb, S, part.get());
  MG_CHECK_LAUNCH(ctx);
  el_rescale_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(D, nd, lb, nn, S, part.get(), dim_sf);
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

}  // namespace mg

using namespace mg;

static int el_check(mg_ctx *ctx, const void *pts, int64_t N, int32_t D, const char *who) {
  MG_REQUIRE(ctx, pts != nullptr, "%s: null points", who);
  MG_REQUIRE(ctx, D >= 1 && D <= EL_DMAX, "%s: dimension must be in 1..%d", who, EL_DMAX);
  MG_REQUIRE(ctx, N >= 1 && N < (1ll << 31) - 1, "%s: 1 <= N < 2^31 - 1 points", who);
  return MG_OK;
}

// Ellipse.enclosing_ellipse sf to_coord pts (ellipse.ml:98-103)
extern "C" int mg_ellipse_enclosing(mg_ctx *ctx, const double *pts, int64_t N, int32_t D, double sf, double *center, double *axes,
                                    double *orientation) {
  if (!ctx) return MG_EINVAL;
  int rc = el_check(ctx, pts, N, D, "enclosing_ellipse");
  if (rc) return rc;
  MG_REQUIRE(ctx, center && axes && orientation, "enclosing_ellipse: null output");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<double> rows, part;
  MG_CUDA(ctx, rows.alloc((size_t)N * D, s));
  MG_CUDA(ctx, cudaMemcpyAsync(rows.get(), pts, sizeof(double) * (size_t)N * D, cudaMemcpyHostToDevice, s));
  ElTable t;
  if ((rc = el_reserve(ctx, t, 0, 1, D, s))) return rc;
  const int32_t be[2] = {0, (int32_t)N};
  MG_CUDA(ctx, cudaMemcpyAsync(t.begin.get(), &be[0], sizeof(int32_t), cudaMemcpyHostToDevice, s));
  MG_CUDA(ctx, cudaMemcpyAsync(t.end.get(), &be[1], sizeof(int32_t), cudaMemcpyHostToDevice, s));
  if ((rc = el_level_ellipses(ctx, rows.get(), D, t, 0, 1, sf, part, N))) return rc;
  MG_CUDA(ctx, cudaMemcpyAsync(center, t.center.get(), sizeof(double) * D, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaMemcpyAsync(axes, t.axes.get(), sizeof(double) * D, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaMemcpyAsync(orientation, t.ori.get(), sizeof(double) * D * D, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  return MG_OK;
}

// Ellipse.elliptical_range ell pt (ellipse.ml:63-73) for M points
extern "C" int mg_ellipse_range(mg_ctx *ctx, const double *center, const double *axes, const double *orientation, int32_t D,
                                const double *q, int64_t M, double *out) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, center && axes && orientation && out && (q || M == 0), "elliptical_range: null argument");
  MG_REQUIRE(ctx, D >= 1 && D <= EL_DMAX && M >= 0, "elliptical_range: dimension must be in 1..%d", EL_DMAX);
  if (M == 0) return MG_OK;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  std::vector<double> h((size_t)2 * D + (size_t)D * D);
  std::copy(center, center + D, h.begin()); std::copy(axes, axes + D, h.begin() + D);
  std::copy(orientation, orientation + (size_t)D * D, h.begin() + 2 * D);
  DevBuf<double> cao, dq, dout;
  MG_CUDA(ctx, upload(cao, h.data(), h.size(), s));
  MG_CUDA(ctx, upload(dq, q, (size_t)M * D, s));
  MG_CUDA(ctx, dout.alloc((size_t)M, s));
  el_range_points_kernel<<<(unsigned)((M + 127) / 128), 128, sizeof(double) * h.size(), s>>>(dq.get(), M, D, cao.get(), dout.get());
  MG_CHECK_LAUNCH(ctx);
  MG_CUDA(ctx, cudaMemcpyAsync(out, dout.get(), sizeof(double) * (size_t)M, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  return MG_OK;
}

static int ellipse_tree_build(mg_ctx *ctx, const double *pts, bool on_device, int64_t N, int32_t D, double sf, mg_ellipse_tree **out) {
  int rc = el_check(ctx, pts, N, D, "ellipse_tree");
  if (rc) return rc;
  MG_REQUIRE(ctx, out != nullptr, "ellipse_tree: null output");
  MG_REQUIRE(ctx, N >= D + 1, "ellipse_tree: Assert_failure ellipse.ml:152 (fewer than ndim + 1 points)");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<double> rowsA, rowsB, part;
  DevBuf<int32_t> permA, permB, segA, segB, flag, lscan, tmp, has, hscan, htmp, htot;
  DevBuf<int> stuck;
  MG_CUDA(ctx, rowsA.alloc((size_t)N * D, s)); MG_CUDA(ctx, rowsB.alloc((size_t)N * D, s));
  MG_CUDA(ctx, cudaMemcpyAsync(rowsA.get(), pts, sizeof(double) * (size_t)N * D, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  MG_CUDA(ctx, permA.alloc((size_t)N, s)); MG_CUDA(ctx, permB.alloc((size_t)N, s));
  MG_CUDA(ctx, segA.alloc((size_t)N, s)); MG_CUDA(ctx, segB.alloc((size_t)N, s));
  MG_CUDA(ctx, flag.alloc((size_t)N + 1, s)); MG_CUDA(ctx, lscan.alloc((size_t)N + 1, s));
  MG_CUDA(ctx, tmp.alloc((size_t)scan_tmp_elems(N + 1, 1), s));
  MG_CUDA(ctx, htot.alloc(1, s)); MG_CUDA(ctx, stuck.alloc(2, s));   // [0] a node cannot be split, [1] largest node of the next level
  MG_CUDA(ctx, cudaMemsetAsync(stuck.get(), 0, 2 * sizeof(int), s));
  el_iota_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(permA.get(), segA.get(), N);
  MG_CHECK_LAUNCH(ctx);
  ElTable t;
  if ((rc = el_reserve(ctx, t, 0, std::max<int64_t>(64, 2 * N / (D + 1) + 64), D, s))) return rc;
  const int32_t be[2] = {0, (int32_t)N};
  MG_CUDA(ctx, cudaMemcpyAsync(t.begin.get(), &be[0], sizeof(int32_t), cudaMemcpyHostToDevice, s));
  MG_CUDA(ctx, cudaMemcpyAsync(t.end.get(), &be[1], sizeof(int32_t), cudaMemcpyHostToDevice, s));
  double *rin = rowsA.get(), *rout = rowsB.get();
  int32_t *pin = permA.get(), *pout = permB.get(), *sin = segA.get(), *sout = segB.get();
  std::vector<std::pair<int64_t, int64_t>> levels;       // [lb, le) of every level
  int64_t lb = 0, le = 1, level_max = N;
  while (le > lb) {
    const int64_t nn = le - lb;
    MG_REQUIRE(ctx, levels.size() < 4096, "ellipse_tree: more than 4096 levels");
    levels.push_back({lb, le});
    if ((rc = el_reserve(ctx, t, le, le + 2 * nn, D, s))) return rc;
    if ((rc = el_level_ellipses(ctx, rin, D, t, (int)lb, (int)nn, sf, part, level_max))) return rc;
    ElNodes nd = t.view();
    el_flag_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, s>>>(rin, N, D, nd, sin, (int)lb, (int)le, flag.get());
    MG_CHECK_LAUNCH(ctx);
    if ((rc = exclusive_scan_i32(ctx, flag.get(), lscan.get(), N + 1, 1, tmp.get(), nullptr))) return rc;
    if (has.n < (size_t)2 * nn) {
      MG_CUDA(ctx, has.alloc((size_t)2 * nn, s)); MG_CUDA(ctx, hscan.alloc((size_t)2 * nn, s));
      MG_CUDA(ctx, htmp.alloc((size_t)scan_tmp_elems(2 * nn, 1), s));
    }
    el_children_count_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(D, nd, (int)lb, (int)nn, lscan.get(), has.get(), stuck.get());
    MG_CHECK_LAUNCH(ctx);
    if ((rc = exclusive_scan_i32(ctx, has.get(), hscan.get(), 2 * nn, 1, htmp.get(), htot.get()))) return rc;
    el_children_make_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(nd, (int)lb, (int)nn, (int)le, lscan.get(), has.get(), hscan.get());
    MG_CHECK_LAUNCH(ctx);
    el_scatter_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(rin, rout, pin, pout, sin, sout, N, D, nd, (int)lb, (int)le, lscan.get(), flag.get());
    MG_CHECK_LAUNCH(ctx);
    int32_t nchild = 0; int h_st[2] = {0, 0};
    MG_CUDA(ctx, cudaMemcpyAsync(&nchild, htot.get(), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    MG_CUDA(ctx, cudaMemcpyAsync(h_st, stuck.get(), sizeof h_st, cudaMemcpyDeviceToHost, s));
    MG_CUDA(ctx, cudaMemsetAsync(stuck.get() + 1, 0, sizeof(int), s));
    MG_CUDA(ctx, cudaStreamSynchronize(s));
    const int h_stuck = h_st[0];
    level_max = h_st[1];
    if (h_stuck) return set_err(ctx, MG_EFAIL, "ellipse_tree: all points of a node lie on one side of its centre (the reference recurses forever, ellipse.ml:153-158)");
    std::swap(rin, rout); std::swap(pin, pout); std::swap(sin, sout);
    lb = le; le = le + nchild;
  }
  const int64_t nnodes = le;
  DevBuf<double> ccc, ccr;
  MG_CUDA(ctx, ccc.alloc((size_t)nnodes * D, s)); MG_CUDA(ctx, ccr.alloc((size_t)nnodes, s));
  mg_ellipse_tree *tr = new mg_ellipse_tree;
  tr->ctx = ctx; tr->N = N; tr->D = D; tr->nnodes = nnodes; tr->nlevels = (int32_t)levels.size();
  {
    const ElNodes nd = t.view();
    for (int L = (int)levels.size() - 1; L >= 0; --L) {
      const int64_t b = levels[L].first, n = levels[L].second - levels[L].first;
      el_circum_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(D, nd, (int)b, (int)n, ccc.get(), ccr.get());
      ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { delete tr; return set_err(ctx, MG_ECUDA, "cuda: %s (ellipse_tree circumcircles)", cudaGetErrorString(e)); }
  }
  // hand the arrays to the tree (ids in their final positions)
  DevBuf<int32_t> pfin;
  if (pin == permA.get()) std::swap(pfin.p, permA.p), std::swap(pfin.n, permA.n), pfin.s = s;
  else std::swap(pfin.p, permB.p), std::swap(pfin.n, permB.n), pfin.s = s;
  auto take_i = [](DevBuf<int32_t> &b) { int32_t *p = b.p; b.p = nullptr; b.n = 0; return p; };
  auto take_d = [](DevBuf<double> &b) { double *p = b.p; b.p = nullptr; b.n = 0; return p; };
  tr->left = take_i(t.left); tr->right = take_i(t.right); tr->begin = take_i(t.begin); tr->end = take_i(t.end);
  tr->perm = take_i(pfin);
  tr->center = take_d(t.center); tr->axes = take_d(t.axes); tr->ori = take_d(t.ori);
  tr->cc_center = take_d(ccc); tr->cc_radius = take_d(ccr);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { mg_ellipse_tree_destroy(tr); return set_err(ctx, MG_ECUDA, "cuda: %s (ellipse_tree)", cudaGetErrorString(e)); }
  *out = tr;
  return MG_OK;
}

// Ellipse.ellipse_tree sf to_coord pts (ellipse.ml:150-173)
extern "C" int mg_ellipse_tree_build(mg_ctx *ctx, const double *pts, int64_t N, int32_t D, double sf, mg_ellipse_tree **out) {
  if (!ctx) return MG_EINVAL;
  return ellipse_tree_build(ctx, pts, false, N, D, sf, out);
}
extern "C" int mg_ellipse_tree_build_dev(mg_ctx *ctx, const double *d_pts, int64_t N, int32_t D, double sf, mg_ellipse_tree **out) {
  if (!ctx) return MG_EINVAL;
  return ellipse_tree_build(ctx, d_pts, true, N, D, sf, out);
}

extern "C" void mg_ellipse_tree_destroy(mg_ellipse_tree *t) {
  if (!t) return;
  void *ps[] = {t->left, t->right, t->begin, t->end, t->perm, t->center, t->axes, t->ori, t->cc_center, t->cc_radius};
  for (void *p : ps) if (p) cudaFree(p);
  delete t;
}

extern "C" int mg_ellipse_tree_info(const mg_ellipse_tree *t, int64_t *npoints, int32_t *dim, int64_t *nnodes, int32_t *nlevels) {
  if (!t) return MG_EINVAL;
  if (npoints) *npoints = t->N;
  if (dim) *dim = t->D;
  if (nnodes) *nnodes = t->nnodes;
  if (nlevels) *nlevels = t->nlevels;
  return MG_OK;
}

extern "C" int mg_ellipse_tree_export(const mg_ellipse_tree *t, int32_t *left, int32_t *right, int32_t *begin, int32_t *end,
                                      int32_t *perm, double *center, double *axes, double *orientation, double *cc_center,
                                      double *cc_radius) {
  if (!t) return MG_EINVAL;
  mg_ctx *ctx = t->ctx;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t n = (size_t)t->nnodes, D = (size_t)t->D;
#define EL_OUT(dst, src, bytes) do { if (dst) MG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s)); } while (0)
  EL_OUT(left, t->left, n * 4); EL_OUT(right, t->right, n * 4); EL_OUT(begin, t->begin, n * 4); EL_OUT(end, t->end, n * 4);
  EL_OUT(perm, t->perm, (size_t)t->N * 4);
  EL_OUT(center, t->center, n * D * 8); EL_OUT(axes, t->axes, n * D * 8); EL_OUT(orientation, t->ori, n * D * D * 8);
  EL_OUT(cc_center, t->cc_center, n * D * 8); EL_OUT(cc_radius, t->cc_radius, n * 8);
#undef EL_OUT
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  return MG_OK;
}
