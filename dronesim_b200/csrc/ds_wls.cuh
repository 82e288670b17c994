// FP64 active-set weighted least squares: the slow path of the 6-DOF allocation.
//
// Restates dronesim/control/wls_alloc.py:125-350 (itself a transliteration of Paparazzi's
// wls_alloc.c) with the same control flow and the same integer working-set bookkeeping
// (free_index / free_index_lookup / W in {-1,0,1}).  The least-squares sub-problem, which the
// reference hands to np.linalg.lstsq (:252), is solved by Householder QR; A_free always has full
// column rank because of the identity block, so both give the unique minimiser.
//
// The fast path (first iteration feasible -> du = M nu, see ds_control.cuh) handles practically
// every call; this routine runs on the few lanes whose unconstrained step leaves the +-1.0
// feasibility slack (:264).  FP64 because the stacked matrix has rows of magnitude 1e10 on unit
// rows (cond ~ 2e4 for hexa_6DOF).
#pragma once
#include "ds_device.cuh"

#define WLS_NU 6
#define WLS_NC 12
#define WLS_FLT_EPSILON 1e-7   // wls_alloc.py:87
#define WLS_INFINITY 1e32      // wls_alloc.py:88

// least squares  min |A x - b|, A is (n_c x n) column-major with leading dimension WLS_NC
static __device__ __noinline__ void ds_lstsq_qr(double* A, double* b, int n_c, int n, double* x) {
  for (int k = 0; k < n; ++k) {
    double* ak = A + k * WLS_NC;
    double nrm = 0.0;
    for (int i = k; i < n_c; ++i) nrm += ak[i] * ak[i];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    double alpha = ak[k] > 0.0 ? -nrm : nrm;
    double vk = ak[k] - alpha;
    // v = (vk, ak[k+1..]) ; H = I - 2 v v^T / (v^T v)
    double vtv = vk * vk;
    for (int i = k + 1; i < n_c; ++i) vtv += ak[i] * ak[i];
    if (vtv == 0.0) continue;
    double beta = 2.0 / vtv;
    for (int j = k + 1; j < n; ++j) {
      double* aj = A + j * WLS_NC;
      double s = vk * aj[k];
      for (int i = k + 1; i < n_c; ++i) s += ak[i] * aj[i];
      s *= beta;
      aj[k] -= s * vk;
      for (int i = k + 1; i < n_c; ++i) aj[i] -= s * ak[i];
    }
    double s = vk * b[k];
    for (int i = k + 1; i < n_c; ++i) s += ak[i] * b[i];
    s *= beta;
    b[k] -= s * vk;
    for (int i = k + 1; i < n_c; ++i) b[i] -= s * ak[i];
    ak[k] = alpha;  // R diagonal; entries below are the Householder vector (no longer needed)
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[j * WLS_NC + k] * x[j];
    double r = A[k * WLS_NC + k];
    x[k] = (r != 0.0) ? s / r : 0.0;
  }
}

// returns the iteration count, or -iterations on non-convergence (reference returns None, :350)
// W_out (nullable): the final working set W in {-1, 0, +1} per actuator (wls_alloc.py:171, 284, 335-338)
static __device__ __noinline__ int ds_wls_alloc(const DsWlsDev* __restrict__ P, const double* v, const double* umin,
                                         const double* umax, double* u_out, int* W_out = nullptr) {
  const int n_u = P->n_u, n_v = P->n_v, n_c = n_u + n_v;
  double A[WLS_NC * WLS_NU];       // row-major [n_c][6]
  double A_free[WLS_NC * WLS_NU];  // row-major [n_c][6]
  double Aq[WLS_NC * WLS_NU];      // column-major scratch for QR
  double d[WLS_NC], dq[WLS_NC];
  double u[WLS_NU], u_opt[WLS_NU], p[WLS_NU], p_free[WLS_NU], W[WLS_NU], Lambda[WLS_NU];
  int free_index[WLS_NU], free_index_lookup[WLS_NU];
  int n_free = 0, free_chk = -1, iter = 0;
  int p_free_len = n_u;  // len(p_free): np.zeros(CA_N_U) initially (:156), then lstsq's n_free

  for (int i = 0; i < n_u; ++i) {  // :166-179
    u[i] = (umax[i] + umin[i]) * 0.5;
    W[i] = 0.0;
    p_free[i] = 0.0;
    free_index[i] = 0;
    free_index_lookup[i] = -1;
  }
  for (int i = 0; i < n_u; ++i) {  // :182-186
    if (W[i] == 0.0) { free_index_lookup[i] = n_free; free_index[n_free] = i; n_free++; }
  }
  for (int i = 0; i < n_v; ++i) {  // :190-203
    double bi = P->gamma * P->Wv[i] * v[i];
    d[i] = bi;
    for (int j = 0; j < n_u; ++j) {
      double a = P->gamma * P->Wv[i] * P->B[i * 6 + j];
      A[i * 6 + j] = a;
      d[i] -= a * u[j];
    }
  }
  for (int i = n_v; i < n_c; ++i) {  // :205-219  (Wu = 1, up = None)
    for (int j = 0; j < n_u; ++j) A[i * 6 + j] = 0.0;
    A[i * 6 + (i - n_v)] = 1.0;
    d[i] = 0.0 - 1.0 * u[i - n_v];
  }
  double alpha = WLS_INFINITY;  // NB: the reference leaves alpha unbound until the first infeasible pass
  int id_alpha = 0;

  while (iter < 100) {  // :222
    iter++;
    for (int i = 0; i < n_u; ++i) { p[i] = 0.0; u_opt[i] = u[i]; }
    if (free_chk != n_free) {  // :233-238
      for (int i = 0; i < n_c; ++i)
        for (int j = 0; j < n_free; ++j) A_free[i * 6 + j] = A[i * 6 + free_index[j]];
      free_chk = n_free;
    }
    if (n_free) {  // :243-252
      for (int j = 0; j < n_free; ++j)
        for (int i = 0; i < n_c; ++i) Aq[j * WLS_NC + i] = A_free[i * 6 + j];
      for (int i = 0; i < n_c; ++i) dq[i] = d[i];
      ds_lstsq_qr(Aq, dq, n_c, n_free, p_free);
      p_free_len = n_free;
    }
    for (int i = 0; i < n_free; ++i) {  // :257-259
      p[free_index[i]] = p_free[i];
      u_opt[free_index[i]] += p_free[i];
    }
    int n_infeasible = 0;  // :262-266
    for (int i = 0; i < n_u; ++i)
      if (u_opt[i] >= (umax[i] + 1.0) || u_opt[i] <= (umin[i] - 1.0)) n_infeasible++;
    if (n_infeasible == 0) {  // :269-298
      for (int i = 0; i < n_u; ++i) { u[i] = u_opt[i]; Lambda[i] = 0.0; }
      for (int i = 0; i < n_c; ++i) {
        for (int k = 0; k < n_free; ++k) d[i] -= A_free[i * 6 + k] * p_free[k];
        for (int k = 0; k < n_u; ++k) Lambda[k] += A[i * 6 + k] * d[i];
      }
      bool break_flag = true;
      for (int i = 0; i < n_u; ++i) {
        Lambda[i] *= W[i];
        if (Lambda[i] < -WLS_FLT_EPSILON) {
          break_flag = false;
          W[i] = 0.0;
          if (free_index_lookup[i] < 0) { free_index_lookup[i] = n_free; free_index[n_free] = i; n_free++; }
        }
      }
      if (break_flag) {
        for (int i = 0; i < n_u; ++i) u_out[i] = u[i];
        if (W_out) for (int i = 0; i < n_u; ++i) W_out[i] = (int)W[i];
        return iter;
      }
      // falls through with the previous alpha / id_alpha, as the reference does
    } else {  // :299-302
      alpha = WLS_INFINITY;
      id_alpha = 0;
    }
    for (int i = 0; i < n_free; ++i) {  // :305-317
      int id = free_index[i];
      double alpha_tmp;
      if (fabs(p[id]) > WLS_FLT_EPSILON) alpha_tmp = (p[id] < 0.0) ? (umin[id] - u[id]) / p[id] : (umax[id] - u[id]) / p[id];
      else alpha_tmp = WLS_INFINITY;
      if (alpha_tmp < alpha) { alpha = alpha_tmp; id_alpha = id; }
    }
    for (int i = 0; i < n_u; ++i) u[i] += alpha * p[i];  // :320-321
    {
      int k_len = n_free < p_free_len ? n_free : p_free_len;  // :325-327
      for (int i = 0; i < n_c; ++i)
        for (int k = 0; k < k_len; ++k) d[i] -= A_free[i * 6 + k] * alpha * p_free[k];
    }
    W[id_alpha] = (p[id_alpha] > 0.0) ? 1.0 : -1.0;  // :335-338
    n_free -= 1;                                     // :342-347
    {
      int lk = free_index_lookup[id_alpha];
      if (lk < 0) lk += n_u;  // numpy negative index wraps (only reachable on the stale-alpha path)
      if (n_free < 0) {
        if (W_out) for (int i = 0; i < n_u; ++i) W_out[i] = (int)W[i];
        return -iter;
      }
      free_index[lk] = free_index[n_free];
      int moved = free_index[lk];
      free_index_lookup[moved] = free_index_lookup[id_alpha];
      free_index_lookup[id_alpha] = -1;
    }
  }
  if (W_out) for (int i = 0; i < n_u; ++i) W_out[i] = (int)W[i];
  return -iter;
}
