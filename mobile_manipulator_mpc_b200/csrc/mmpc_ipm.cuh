// mmpc_ipm.cuh -- scalar pieces of the interior-point iteration shared by the lane-per-instance
// kernel (mmpc_lane.cuh) and the warp-cooperative kernel (mmpc_solver.cuh): IPOPT's scaled KKT
// error, the bound push of the starting point, and a cheap running log-barrier sum.
#pragma once
#include "mmpc_model.cuh"

namespace mmpc {

__device__ constexpr int POSE2X[6] = {0, 1, 2, 6, 7, 8};

struct KktParts {
  double e_stat, e_prim, c_hi, c_lo, sum_lam, sum_z;
  int n_z, n_eq;
};
__device__ __forceinline__ double kkt_error(const KktParts& k, double mu) {
  const double smax = 100.0;
  double sd = fmax(smax, (k.sum_lam + k.sum_z) / fmax(1.0, (double)(k.n_eq + k.n_z))) / smax;
  double sc = fmax(smax, k.sum_z / fmax(1.0, (double)k.n_z)) / smax;
  double ec = k.n_z ? fmax(fabs(k.c_hi - mu), fabs(k.c_lo - mu)) : 0.0;
  return fmax(fmax(k.e_stat / sd, k.e_prim), ec / sc);
}

__device__ __forceinline__ double push_in(double v, double lo, double hi) {
  const double k1 = 1e-2, k2 = 1e-2;
  bool fl = is_fin(lo), fh = is_fin(hi);
  if (fl && fh) {
    double pl = fmin(k1 * fmax(1.0, fabs(lo)), k2 * (hi - lo)), pu = fmin(k1 * fmax(1.0, fabs(hi)), k2 * (hi - lo));
    v = fmax(v, lo + pl); v = fmin(v, hi - pu);
  } else if (fl) v = fmax(v, lo + k1 * fmax(1.0, fabs(lo)));
  else if (fh) v = fmin(v, hi - k1 * fmax(1.0, fabs(hi)));
  return v;
}

// log of a running product without one log() per factor
struct LogProd {
  double prod, acc;
  __device__ __forceinline__ void init() { prod = 1.0; acc = 0.0; }
  __device__ __forceinline__ void mul(double v) {
    prod *= v;
    if (prod < 1e-120 || prod > 1e120) { acc += log(prod); prod = 1.0; }
  }
  __device__ __forceinline__ double value() const { return acc + log(prod); }
};

}  // namespace mmpc
