// mmpc_ipm.cuh -- scalar pieces of the interior-point iteration shared by the phase kernels: IPOPT's
// scaled KKT error, the bound push of the starting point, and a cheap running log-barrier sum.
#pragma once
#include "mmpc_model.cuh"

namespace mmpc {

__device__ constexpr int POSE2X[6] = {0, 1, 2, 6, 7, 8};

// angleDiff (controllers/mpc_wholebody_qref.py:92-117, controllers/mpc_base.py:59-84): a - b folded to the nearest
// representative; casadi fmod == C fmod (sign of the dividend)
__device__ inline double angle_diff(double a, double b) {
  const double PI = 3.14159265358979323846;
  a = fmod(a + PI, 2 * PI) - PI;
  b = fmod(b + PI, 2 * PI) - PI;
  const double d = a - b;
  if (a * b >= 0) return d;
  if (a > b) return d <= PI ? d : d - 2 * PI;
  return d > -PI ? d : d + 2 * PI;
}

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

// Reciprocal instead of an IEEE division (a double division is ~35 instructions with its slow-path
// check; one reciprocal feeds every quotient with the same denominator).
#if defined(MMPC_EMULATE_LANE)
__device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
#else
#ifdef MMPC_RCP_RN
__device__ __forceinline__ double rcp(double x) { return __drcp_rn(x); }
#else
// MUFU.RCP64H seed (relative error 2^-23) and two Newton steps: 1 + 4 instructions, no slow-path branch; within an ulp or two
// of 1/x for the normal, finite arguments the solver divides by (slacks, pivots, distances).  A zero, subnormal or non-finite
// argument gives a non-finite result, which is what those call sites test for anyway (pivot > 0, t > 0).
__device__ __forceinline__ double rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(y, fma(-x, y, 1.0), y);
  y = fma(y, fma(-x, y, 1.0), y);
  return y;
}
#endif
#endif
// 1/sqrt(x) for the squared distances of the circle and self-collision rows (x > 0, normal): MUFU.RSQ64H seed (relative
// error 2^-22) and two Newton steps, no special-case path.
#if defined(MMPC_EMULATE_LANE) || defined(MMPC_RCP_RN)
__device__ __forceinline__ double rsq(double x) { return rsqrt(x); }
#else
__device__ __forceinline__ double rsq(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = fma(y, fma(-h * y, y, 0.5), y);
  y = fma(y, fma(-h * y, y, 0.5), y);
  return y;
}
#endif
// IPOPT's multiplier safeguard  z <- max(min(z, kappa mu/d), mu/(kappa d)),  kappa = 1e10, id = 1/d
__device__ __forceinline__ double zclamp(double z, double mu, double id) { double r = mu * id; return fmax(fmin(z, 1e10 * r), 1e-10 * r); }
// Fraction to the boundary without a division per candidate: the smallest quotient num/den (num > 0,
// den > 0) is tracked as a pair and compared by cross-multiplication; den = 0 stands for +infinity.
struct MinRatio {
  double num, den;
  __device__ __forceinline__ void init() { num = 1.0; den = 0.0; }
  __device__ __forceinline__ void add(double n, double d) { if (n * den < num * d) { num = n; den = d; } }  // n/d < num/den
  __device__ __forceinline__ double value(double tau) const { return den > 0 ? fmin(1.0, tau * num / den) : 1.0; }
};

// log of a running product without one log() per factor.  The rare renormalisation calls log() OUT OF LINE: inlined, its
// ~60 instructions sat behind every one of the ~100 mul() sites of the trial kernel (31 % of the kernel's code, 6.7 % of its
// stall samples "no instruction": the taken branch around each copy runs past the fetched lines).
#if defined(MMPC_EMULATE) || defined(MMPC_EMULATE_LANE) || defined(MMPC_LOG_INLINE)
__device__ __forceinline__ double log_renorm(double p) { return log(p); }
#else
__device__ __noinline__ double log_renorm(double p) { return log(p); }
#endif
struct LogProd {
  double prod, acc;
  __device__ __forceinline__ void init() { prod = 1.0; acc = 0.0; }
  __device__ __forceinline__ void mul(double v) {
    prod *= v;
    if (prod < 1e-120 || prod > 1e120) { acc += log_renorm(prod); prod = 1.0; }
  }
  __device__ __forceinline__ double value() const { return acc + log(prod); }
};

}  // namespace mmpc
