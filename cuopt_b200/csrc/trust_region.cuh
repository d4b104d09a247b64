// Trust-region restart of the Methodical1 preset on the device, written against the oracle restatement
// (oracle/pdlp_oracle.cpp: run_trust_region_restart / bound_optimal_objective, pinned to the reference's
// test_very_low_tolerance).  GPU acceptance: tests/test_methodical1.py (first run and green in round 2).
//
// Reference: restart_strategy/pdlp_restart_strategy.cu:278-364 (flow), :1034-1078 (bound_optimal_objective,
// compute_bound), :1291-1678 (trust-region solve), :1717-1900 (gradients, Lagrangian), utils.cuh:240-345.
//
// One call of bound_optimal_objective =
//   2 x k_spmv (A^T y, A x on the SCALED matrices)            [launched by the host code]
//   k_tr_prepare   element-wise over N = n + m: direction / threshold of every component of the joint problem,
//                  primal gradient, dual gradient, 5 reductions (Lagrangian terms, |objective|^2, the radius of
//                  the components that never hit a bound)
//   cub sort of the thresholds with an index permutation, k_tr_weights + two cub prefix sums over the sorted order
//   k_tr_bisect    ONE thread: the reference's median bisection, every partial radius a difference of prefix sums
//                  (the reference runs a cooperative kernel that re-reduces the active range per trial)
//   k_tr_bounds    element-wise: the trust-region point, <tr - x, g_p> and <tr - y, g_d> -> lower / upper bound
#pragma once

#include "pdlp_kernels.cuh"

namespace cuopt_b200 {

// scal layout (doubles): 0 c.x  1 x.A^Ty  2 y.subgradient  3 |objective|^2  4 radius^2 of never-fixed components
//                        5 target radius (in)  6 target threshold  7 lower bound (out)  8 upper bound (out)
constexpr int TR_SCALARS = 16;

struct tr_problem_t {
  int n, m;
  const double *px, *py;    // the point (center)
  const double *aty, *ax;   // A^T py, A px (scaled matrices)
  const double *c, *l, *u;  // scaled objective / variable bounds
  const double *lc, *uc;    // scaled constraint bounds
};

// One component of the joint problem [x; y] (utils.cuh:240-322).  Returns direction; threshold by reference.
__device__ __forceinline__ void tr_component(const tr_problem_t& P, int k, double tau, double sigma, double& center,
                                             double& obj, double& lo, double& up, double& w, double& grad, double& sub)
{
  if (k < P.n) {
    center = P.px[k];
    grad   = P.c[k] - P.aty[k];  // primal gradient c - A^T y
    obj    = grad;
    lo     = P.l[k];
    up     = P.u[k];
    w      = 1.0 / tau;
    sub    = 0.0;
  } else {
    const int i       = k - P.n;
    const double lower = P.lc[i], upper = P.uc[i], prod = P.ax[i], yi = P.py[i];
    double sc;  // compute_subgradient_kernel, pdlp_restart_strategy.cu:1746-1783
    if (yi < 0.0) sc = upper;
    else if (yi > 0.0) sc = lower;
    else if (!isfinite(upper) && !isfinite(lower)) sc = 0.0;
    else if (!isfinite(upper) && isfinite(lower)) sc = lower;
    else if (isfinite(upper) && !isfinite(lower)) sc = upper;
    else sc = prod < lower ? lower : (prod > upper ? upper : prod);
    sub    = sc;
    grad   = sc - prod;  // dual gradient
    obj    = -grad;
    center = yi;
    lo     = isfinite(upper) ? -CUDART_INF : 0.0;
    up     = isfinite(lower) ? CUDART_INF : 0.0;
    w      = 1.0 / sigma;
  }
}
__device__ __forceinline__ void tr_direction(double center, double obj, double lo, double up, double w, double& dir,
                                             double& thr)
{
  dir = 0.0;
  thr = 0.0;
  if (center >= up && obj <= 0.0) return;
  if (center <= lo && obj >= 0.0) return;
  if (obj == 0.0) {
    thr = CUDART_INF;
    return;
  }
  dir = -obj / w;
  if (dir > 0.0) thr = (up - center) / dir;
  else if (dir < 0.0) thr = (lo - center) / dir;
}

__global__ void __launch_bounds__(EW_THREADS) k_tr_prepare(pdhg_ctl_t* __restrict__ ctl,
                                                           tr_problem_t P,
                                                           double* __restrict__ dir_out,
                                                           double* __restrict__ thr_out,
                                                           double* __restrict__ grad_out,  // g_p (n) then g_d (m)
                                                           int* __restrict__ iota,
                                                           double* __restrict__ parts,  // 5 x gridDim.x
                                                           double* __restrict__ scal)
{
  __shared__ double red[32];
  const double tau = ctl->tau, sigma = ctl->sigma;
  const int N = P.n + P.m;
  double acc[5]    = {0, 0, 0, 0, 0};
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
    double center, obj, lo, up, w, grad, sub, dir, thr;
    tr_component(P, k, tau, sigma, center, obj, lo, up, w, grad, sub);
    tr_direction(center, obj, lo, up, w, dir, thr);
    dir_out[k]  = dir;
    thr_out[k]  = thr;
    grad_out[k] = grad;
    iota[k]     = k;
    if (k < P.n) {
      acc[0] += center * P.c[k];
      acc[1] += center * P.aty[k];
    } else {
      acc[2] += center * sub;
    }
    acc[3] += obj * obj;
    if (isinf(thr)) acc[4] += dir * dir * w;
  }
  if (!publish_and_elect<5>(acc, parts, &ctl->ticket[3], red)) return;
  for (int q = 0; q < 5; ++q) {
    const double t = gather_partials(parts + q * gridDim.x, gridDim.x, red);
    if (threadIdx.x == 0) scal[q] = t;
  }
}

// sorted position k -> A[k] = (thr * dir)^2 w (radius^2 once the component sits on its bound), B[k] = dir^2 w
__global__ void __launch_bounds__(EW_THREADS) k_tr_weights(const pdhg_ctl_t* __restrict__ ctl, int n, int N,
                                                           const double* __restrict__ thr_sorted,
                                                           const int* __restrict__ perm, const double* __restrict__ dir,
                                                           double* __restrict__ A, double* __restrict__ B)
{
  const double wp = 1.0 / ctl->tau, wd = 1.0 / ctl->sigma;
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
    const int q    = perm[k];
    const double d = dir[q], w = q < n ? wp : wd, t = thr_sorted[k];
    A[k]           = isinf(t) ? 0.0 : (t * d) * (t * d) * w;
    B[k]           = d * d * w;
  }
}

// solve_bound_constrained_trust_region_kernel (:1291-1358) + target_threshold_determination_kernel (:1081-1100)
__global__ void k_tr_bisect(int N, const double* __restrict__ thr, const double* __restrict__ PA,
                            const double* __restrict__ PB, double* __restrict__ scal)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto range_sum = [](const double* P, int a, int b) { return b > a ? P[b - 1] - (a > 0 ? P[a - 1] : 0.0) : 0.0; };
  auto first_ge  = [&](double t, int a, int b) {  // first index in [a, b) with thr >= t
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (thr[mid] >= t) b = mid; else a = mid + 1;
    }
    return a;
  };
  auto first_gt = [&](double t, int a, int b) {  // first index in [a, b) with thr > t
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (thr[mid] > t) b = mid; else a = mid + 1;
    }
    return a;
  };
  const double target = scal[5];
  double high_r2 = scal[4], low_r2 = 0.0;
  int low = 0, high = first_ge(CUDART_INF, 0, N);
  while (low != high) {
    const int size = high - low;
    const double t = (size & 1) == 0 ? 0.5 * (thr[low + size / 2 - 1] + thr[low + size / 2]) : thr[low + size / 2];
    const int p    = first_gt(t, low, high);
    const double test_r2 = range_sum(PA, low, p) + (t * t) * range_sum(PB, p, high);
    const bool too_high  = low_r2 + test_r2 + (t * t) * high_r2 >= target * target;
    if (too_high) {
      const int new_high = first_ge(t, low, high);
      high_r2 += range_sum(PB, new_high, high);
      high = new_high;
    } else {
      low_r2 += range_sum(PA, low, p);
      low = p;
    }
  }
  scal[6] = high_r2 <= 0.0 ? thr[N - 1] : sqrt((target * target - low_r2) / high_r2);
}

// trust-region point and the two bounds (compute_bound, :1053-1078); degenerate: 1 -> tr = center
__global__ void __launch_bounds__(EW_THREADS) k_tr_bounds(pdhg_ctl_t* __restrict__ ctl, tr_problem_t P,
                                                          const double* __restrict__ dir,
                                                          const double* __restrict__ grad, int degenerate,
                                                          double* __restrict__ parts,  // 2 x gridDim.x
                                                          double* __restrict__ scal)
{
  __shared__ double red[32];
  const double tau = ctl->tau, sigma = ctl->sigma, T = scal[6];
  const int N = P.n + P.m;
  double acc[2]    = {0.0, 0.0};
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
    double center, obj, lo, up, w, g, sub;
    tr_component(P, k, tau, sigma, center, obj, lo, up, w, g, sub);
    double tr = center;
    if (!degenerate) {
      const double d     = dir[k];
      const double moved = d == 0.0 ? center : center + T * d;  // a component that does not move keeps its value
      tr                 = fmin(fmax(moved, lo), up);
    }
    acc[k < P.n ? 0 : 1] += (tr - center) * grad[k];
  }
  if (!publish_and_elect<2>(acc, parts, &ctl->ticket[3], red)) return;
  const double dp = gather_partials(parts, gridDim.x, red);
  const double dd = gather_partials(parts + gridDim.x, gridDim.x, red);
  if (threadIdx.x != 0) return;
  const double lagrangian = scal[0] - scal[1] + scal[2];
  scal[7]                 = lagrangian + dp;
  scal[8]                 = lagrangian + dd;
}

}  // namespace cuopt_b200
