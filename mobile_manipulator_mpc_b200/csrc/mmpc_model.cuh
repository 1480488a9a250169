// mmpc_model.cuh -- device-side robot model of the whole-body MPC hot path (sm_100a, FP64).
//
// What the reference evaluates symbolically through CasADi at every IPOPT iteration
// (controllers/mpc_wholebody_qref.py:177-270) is evaluated here analytically, per stage, by
// one lane of the warp that owns the instance:
//   dynamics   robot_models/base.py:19-26, robot_models/manipulator_3DoF.py:189-191
//   FK         robot_models/manipulator_3DoF.py:18-77 + robot_models/mobile_manipulator.py:28-55
//              in the compact theta-chain form: theta1=q1, theta2=q1-q2,
//              theta3=q1-q2-q3, segments v_s=(r,h) with dv/dtheta=(h,-r), d2v/dtheta2=-v.
//   rows       obsAvoid :49-54, self-collision :216-222, obsAvoidConvex :57-89
// Every body point / self-collision difference is  P(kappa; w) = (kappa*x + R cos(psi),
// kappa*y + R sin(psi), Z), R = sum_s w_s v_s.r + kappa*bx, Z = sum_s w_s v_s.h + kappa*bz,
// so one routine yields values, gradients and Hessians wrt the pose (x, y, psi, q1, q2, q3).
#pragma once
#include "mmpc_warp.cuh"
#include <math.h>

namespace mmpc {

constexpr int NX = 9, NU = 5, NP = 6;
// robot_models/manipulator_3DoF.py:18-22, robot_models/mobile_manipulator.py:14-15
constexpr double A2 = 0.316, A3 = 0.0825, A5 = 0.384, A6 = 0.088, A7 = 0.107;
constexpr double BX = -0.007, BZ = 0.606 + 0.333;

// (kappa, w1, w2, w3): body points mpc_wholebody_qref.py:216-217, self-collision differences :219-221
__device__ constexpr double BODY[6][4] = {{0.5, 0.5, 0, 0}, {1, 1, 0, 0}, {1, 1, 0.5, 0},
                                          {1, 1, 1, 0},     {1, 1, 1, 0.5}, {1, 1, 1, 1}};
__device__ constexpr double SELFD[4][4] = {{-1, -1, -1, -1}, {-0.5, -0.5, -1, -1}, {0, 0, -1, -1}, {0, 0, -0.5, -1}};

struct FK {
  double cp, sp;        // cos/sin psi
  double vr[3], vh[3];  // arm segments
};

__device__ __forceinline__ void fk_eval(double psi, double q1, double q2, double q3, FK& f) {
  double s1, c1, s2, c2, s3, c3;
  sincos(psi, &f.sp, &f.cp);
  sincos(q1, &s1, &c1);
  sincos(q1 - q2, &s2, &c2);
  sincos(q1 - q2 - q3, &s3, &c3);
  f.vr[0] = A2 * s1 + A3 * c1;  f.vh[0] = A2 * c1 - A3 * s1;
  f.vr[1] = -A3 * c2 + A5 * s2; f.vh[1] = A3 * s2 + A5 * c2;
  f.vr[2] = A6 * c3 - A7 * s3;  f.vh[2] = -A6 * s3 - A7 * c3;
}

// packed index of the symmetric 6x6 pose block, a <= b
__host__ __device__ constexpr int pidx(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }

struct Point {
  double kap, R, Z, P[3];
  double Rq[3], Zq[3];    // d/dq of R and Z
  double Rth[3], Zth[3];  // d/dtheta_s of R and Z
};

__device__ __forceinline__ void point_eval(double x, double y, const FK& f, const double (&kw)[4], Point& p) {
  p.kap = kw[0];
  double R = kw[0] * BX, Z = kw[0] * BZ;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    R = fma(kw[1 + s], f.vr[s], R);
    Z = fma(kw[1 + s], f.vh[s], Z);
    p.Rth[s] = kw[1 + s] * f.vh[s];
    p.Zth[s] = -kw[1 + s] * f.vr[s];
  }
  p.R = R; p.Z = Z;
  p.P[0] = fma(kw[0], x, R * f.cp);
  p.P[1] = fma(kw[0], y, R * f.sp);
  p.P[2] = Z;
  p.Rq[0] = p.Rth[0] + p.Rth[1] + p.Rth[2]; p.Rq[1] = -p.Rth[1] - p.Rth[2]; p.Rq[2] = -p.Rth[2];
  p.Zq[0] = p.Zth[0] + p.Zth[1] + p.Zth[2]; p.Zq[1] = -p.Zth[1] - p.Zth[2]; p.Zq[2] = -p.Zth[2];
}

// g = J^T n   (gradient of n . P wrt the pose)
__device__ __forceinline__ void point_grad(const FK& f, const Point& p, const double (&n)[3], double (&g)[NP]) {
  double npar = n[0] * f.cp + n[1] * f.sp, nperp = -n[0] * f.sp + n[1] * f.cp;
  g[0] = p.kap * n[0]; g[1] = p.kap * n[1]; g[2] = p.R * nperp;
#pragma unroll
  for (int j = 0; j < 3; ++j) g[3 + j] = p.Rq[j] * npar + p.Zq[j] * n[2];
}

// H[pidx] += coef * sum_c n_c Hess(P_c)
__device__ __forceinline__ void point_hess_acc(const FK& f, const Point& p, const double (&n)[3], double coef, double* H) {
  double npar = n[0] * f.cp + n[1] * f.sp, nperp = -n[0] * f.sp + n[1] * f.cp;
  H[pidx(2, 2)] -= coef * p.R * npar;
  double cn = coef * nperp;
#pragma unroll
  for (int j = 0; j < 3; ++j) H[pidx(2, 3 + j)] = fma(cn, p.Rq[j], H[pidx(2, 3 + j)]);
  double g0 = coef * (p.Zth[0] * npar - p.Rth[0] * n[2]);
  double g1 = coef * (p.Zth[1] * npar - p.Rth[1] * n[2]);
  double g2 = coef * (p.Zth[2] * npar - p.Rth[2] * n[2]);
  double g12 = g1 + g2;
  H[pidx(3, 3)] += g0 + g12; H[pidx(3, 4)] -= g12; H[pidx(3, 5)] -= g2;
  H[pidx(4, 4)] += g12;      H[pidx(4, 5)] += g2;  H[pidx(5, 5)] += g2;
}

// H[pidx] += coef * (J^T J)
__device__ __forceinline__ void point_jtj_acc(const FK& f, const Point& p, double coef, double* H) {
  double k = p.kap;
  H[pidx(0, 0)] += coef * k * k; H[pidx(1, 1)] += coef * k * k;
  double ck = coef * k;
  H[pidx(0, 2)] -= ck * p.R * f.sp; H[pidx(1, 2)] += ck * p.R * f.cp;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    H[pidx(0, 3 + j)] += ck * p.Rq[j] * f.cp;
    H[pidx(1, 3 + j)] += ck * p.Rq[j] * f.sp;
  }
  H[pidx(2, 2)] += coef * p.R * p.R;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) H[pidx(3 + i, 3 + j)] += coef * (p.Rq[i] * p.Rq[j] + p.Zq[i] * p.Zq[j]);
}

// f_kinematics: robot_models/mobile_manipulator.py:57-75
__device__ __forceinline__ void dyn_f(const double* x, const double* u, double dt, double cp, double sp, double* xn) {
  xn[0] = fma(dt, x[3], x[0]);
  xn[1] = fma(dt, x[4], x[1]);
  xn[2] = fma(dt, x[5], x[2]);
  xn[3] = x[3] + dt * (u[0] * cp - x[4] * x[5]);
  xn[4] = x[4] + dt * (u[0] * sp + x[3] * x[5]);
  xn[5] = fma(dt, u[1], x[5]);
  xn[6] = fma(u[2], dt, x[6]);
  xn[7] = fma(u[3], dt, x[7]);
  xn[8] = fma(u[4], dt, x[8]);
}

__device__ __forceinline__ bool is_fin(double v) { return v > -1e300 && v < 1e300; }

}  // namespace mmpc
