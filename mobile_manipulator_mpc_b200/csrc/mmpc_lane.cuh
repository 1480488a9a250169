// mmpc_lane.cuh -- lane-per-instance primal-dual interior-point solve of the whole-body MPC NLP
// (controllers/mpc_wholebody_qref.py:142-285), replacing opti.solve() (:315), i.e. the
// CasADi/IPOPT/MUMPS stack, for tens of thousands of independent instances at once.
//
// Mapping (B200): ONE THREAD owns one instance; a warp advances 32 instances in lock-step through
// the phases of an interior-point iteration; a persistent grid (resident warps = SMs x warps/SM)
// pulls instances from an atomic queue, and a lane that finishes picks up the next instance while
// its neighbours keep iterating.  There is no cross-lane communication in the solve: every phase
// is a sequential loop over the horizon executed by all lanes at full SIMT width (the
// warp-cooperative kernel in mmpc_solver.cuh uses 21 of 32 lanes in the stage-parallel phases and
// 5-10 in the Riccati sweep, and is latency bound).
//
// Data: the whole per-instance state (iterate, multipliers, stage QPs, Riccati factors, row
// slacks) lives in a lane-interleaved HBM workspace  ws[warp_slot][index][32 lanes]  so that every
// access of a warp is one fully coalesced 256-byte line; the Riccati recursion of a stage runs in
// registers (P 9x9 packed symmetric, A/B sparsity hard-coded).  Instance inputs are copied into
// the same layout once at init.
//
// Phases of one iteration (per lane, k = stage):
//   eval      k = N..0   commit of the previous step (lazy), dynamics, FK, all inequality rows with
//                        gradient and Hessian, barrier condensation, KKT error
//   riccati   k = N..0   backward recursion; the slack s_{k+1} is eliminated at stage k as a scalar
//                        control v_k (augmented form of oracle/mmpc_oracle.c), inertia check, delta_w
//   forward   k = 0..N   roll-out of the Newton step, costates, row/bound steps, fraction to the
//                        boundary, merit ingredients
//   trial                values-only evaluation for the filter line search
// Algorithm = oracle/mmpc_oracle.c (IPOPT-style: monotone mu, tau = max(0.99, 1-mu), bound push,
// gradient-based objective scaling, filter line search with slack reset).
#pragma once
#include <stdint.h>
#include "../../include/mmpc.h"
#include "mmpc_model.cuh"
#include "mmpc_warp.cuh"
#include "mmpc_ipm.cuh"

namespace mmpc {

#ifndef MMPC_LSTR
#define MMPC_LSTR 32  // lane stride of the interleaved workspace
#endif

struct LParams {
  MmpcConfig cfg;
  int B;
  const double *x_init, *x_ref, *u_ref, *u_last, *u_guess, *circles, *planes;
  const int32_t* n_pl_inst;
  const uint8_t* flags;
  double *U, *X, *s, *cost, *kkt;
  int32_t *iters, *status;
  double* ws;              // lane-interleaved workspace
  long long warp_stride;   // doubles per warp slot
  unsigned* counter;
  int R, STG, OG;          // rows per slack, doubles per stage, offset of the per-instance block
};

// ---- per-stage record (doubles) ------------------------------------------------------------------
constexpr int O_X = 0, O_U = 9, O_S = 14, O_LAM = 15, O_DX = 24, O_DU = 33, O_DS = 38, O_LAMN = 39;
constexpr int O_FK = 48, O_DFC = 56, O_ZXL = 65, O_ZXU = 74, O_ZUL = 83, O_ZUU = 88;
// stage QP: pose Hessian (21 packed) | velocity diagonal 3 | (dx,dpsi) (dy,dpsi) | control diagonal 5 |
// (psi,u0) | a = H[pose][s] 6 | c = H[s][s] | bv = H[pose][v] 6 | hvv | gA 16 | gB 16   (y = x9 s u5 v)
constexpr int O_HP = 93, O_HVD = 114, O_H35 = 117, O_H45 = 118, O_HUU = 119, O_HPU = 124, O_A = 125, O_C = 131,
              O_BV = 132, O_HVV = 138, O_GA = 139, O_GB = 155;
// Riccati: K 45 | kff 5 | w 14 | cv | l0 | Pxx 45 | pxx 9
constexpr int O_K = 171, O_KFF = 216, O_W = 221, O_CV = 235, O_L0 = 236, O_P = 237, O_PV = 282;
// inputs in lane layout
constexpr int O_XREF = 291, O_UREF = 300, O_ULAST = 305, O_ULO = 310, O_UHI = 315;
constexpr int O_ROW = 320;  // t[R] z[R] dt[R], then (moving obstacles) circles[3*nobs]
// per-instance block: filter 32 | planes 6*MAX | static circles 3*nobs
constexpr int G_FILT = 0, G_PL = 32, G_CIRC = 32 + 6 * MMPC_MAX_PLANES;
constexpr int GY_S = 9, GY_U = 10, GY_V = 15;

__host__ __device__ inline int lane_rows(const MmpcConfig& c) { return c.n_obs + 4 + (c.n_pl > 0 ? 6 : 0); }
__host__ __device__ inline int lane_stage_doubles(const MmpcConfig& c) {
  return O_ROW + 3 * lane_rows(c) + (c.obs_per_stage ? 3 * c.n_obs : 0);
}
__host__ __device__ inline long long lane_instance_doubles(const MmpcConfig& c) {
  return (long long)(c.N + 1) * lane_stage_doubles(c) + G_CIRC + (c.obs_per_stage ? 0 : 3 * c.n_obs);
}

__host__ __device__ constexpr int sidx(int i, int j) { return i <= j ? i * 9 - i * (i - 1) / 2 + (j - i) : j * 9 - j * (j - 1) / 2 + (i - j); }
struct ACoef { double dt, a32, a42, a34, a43, a35, a45, cp, sp; };

struct Lane {
  const LParams& P;
  const MmpcConfig& cfg;
  double* w;  // this lane's base inside the interleaved workspace
  int N, R, STG, OG, nobs, npl, b;
  double dt, os;  // objective scale (IPOPT nlp_scaling_max_gradient = 100)
  // interior-point state
  double mu, reg_last, theta_max, theta_min, E0;
  int nfilt, it;
  bool have_step;
  double alpha_p, alpha_d, mu_prev;
  // line-search ingredients of the current step
  double ls_ap, ls_ad, ls_gphi, ls_theta, ls_f, ls_logsum;

  __device__ Lane(const LParams& p) : P(p), cfg(p.cfg) {}

  __device__ __forceinline__ double& W(int k, int o) const { return w[((long long)k * STG + o) * MMPC_LSTR]; }
  __device__ __forceinline__ double& G(int o) const { return w[((long long)OG + o) * MMPC_LSTR]; }
  __device__ __forceinline__ int o_t(int r) const { return O_ROW + r; }
  __device__ __forceinline__ int o_z(int r) const { return O_ROW + R + r; }
  __device__ __forceinline__ int o_dt(int r) const { return O_ROW + 2 * R + r; }
  __device__ __forceinline__ double circ(int k, int i, int c) const {
    return cfg.obs_per_stage ? W(k, O_ROW + 3 * R + 3 * i + c) : G(G_CIRC + 3 * i + c);
  }

  __device__ void bind(double* lane_base, int b_) {
    w = lane_base; b = b_;
    N = cfg.N; R = P.R; STG = P.STG; OG = P.OG; nobs = cfg.n_obs; dt = cfg.dt;
    npl = P.n_pl_inst ? ldg(P.n_pl_inst + b) : cfg.n_pl;
  }

  // -max_j c[i][j] for body point i (:76-87); returns the arg-max plane
  __device__ __forceinline__ double plane_row(const Point& p, int& jbest) const {
    double cb = 0; jbest = 0;
    for (int j = 0; j < npl; ++j) {
      double n0 = G(G_PL + 6 * j + 3), n1 = G(G_PL + 6 * j + 4), n2 = G(G_PL + 6 * j + 5);
      double e = cfg.obstacle_expand_dist;
      double off = n0 * (G(G_PL + 6 * j + 0) - e * n0) + n1 * (G(G_PL + 6 * j + 1) - e * n1) + n2 * (G(G_PL + 6 * j + 2) - e * n2);
      double c = off - (n0 * p.P[0] + n1 * p.P[1] + n2 * p.P[2]);
      bool take = (j == 0) || (npl == 2 ? !(cb > c) : (c > cb));  // if_else(c0 > c1, c0, c1) :85 ; mmax :87
      if (take) { cb = c; jbest = j; }
    }
    return -cb;
  }

  // ------------------------------------------------------------------------------------------
  // init: copy the instance into the lane layout, reference initial guess (:302-304) + IPOPT
  // bound push; s lifted so every row starts strictly feasible; objective scaling.
  __device__ void init() {
    const double* xref = P.x_ref + (long long)b * (N + 1) * NX;
    const double* uref = P.u_ref + (long long)b * N * NU;
    const double* ulast = P.u_last + (long long)b * N * NU;
    for (int j = 0; j < cfg.n_pl; ++j)
      for (int c = 0; c < 6; ++c) G(G_PL + 6 * j + c) = ldg(P.planes + ((long long)b * cfg.n_pl + j) * 6 + c);
    if (!cfg.obs_per_stage)
      for (int i = 0; i < 3 * nobs; ++i) G(G_CIRC + i) = ldg(P.circles + (long long)b * 3 * nobs + i);
    double gmax = 0;
    for (int k = 0; k <= N; ++k) {
      double x[NX];
      if (cfg.obs_per_stage)
        for (int i = 0; i < 3 * nobs; ++i) W(k, O_ROW + 3 * R + i) = ldg(P.circles + ((long long)b * (N + 1) + k) * 3 * nobs + i);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = fmax(fmin(ldg(P.x_init + (long long)b * NX + i), cfg.xlim[1][i]), cfg.xlim[0][i]);  // :290-291
        if (k >= 1) v = push_in(v, cfg.xlim[0][i], cfg.xlim[1][i]);
        double xr = ldg(xref + k * NX + i);
        x[i] = v; W(k, O_X + i) = v; W(k, O_LAM + i) = 0; W(k, O_XREF + i) = xr;
        W(k, O_ZXL + i) = 1; W(k, O_ZXU + i) = 1;
        double Wx = (k < N ? cfg.Qd[i] : cfg.Pd[i]);
        if (k >= 1) gmax = fmax(gmax, fabs(2 * Wx * (v - xr)));
      }
      if (k < N) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double ul = ldg(ulast + k * NU + j), ur = ldg(uref + k * NU + j);
          double lo = fmax(cfg.ulim[0][j], ul + cfg.dulim[0][j]);  // mpc_wholebody_qref.py:203 and :205 merged
          double hi = fmin(cfg.ulim[1][j], ul + cfg.dulim[1][j]);
          double v = P.u_guess ? ldg(P.u_guess + ((long long)b * N + k) * NU + j) : ul;
          v = push_in(v, lo, hi);
          W(k, O_U + j) = v; W(k, O_UREF + j) = ur; W(k, O_ULAST + j) = ul; W(k, O_ULO + j) = lo; W(k, O_UHI + j) = hi;
          W(k, O_ZUL + j) = 1; W(k, O_ZUU + j) = 1;
          gmax = fmax(gmax, fabs(2 * cfg.Rd[j] * (v - ur) + 2 * cfg.Wd[j] * (v - ul)));
        }
      }
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
      double hmax = -1e300;
      for (int i = 0; i < nobs; ++i) {
        double ddx = x[0] - circ(k, i, 0), ddy = x[1] - circ(k, i, 1);
        double h = (circ(k, i, 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy);
        W(k, o_t(i)) = h; hmax = fmax(hmax, h);
      }
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        double h = cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]);
        W(k, o_t(nobs + m)) = h; hmax = fmax(hmax, h);
      }
      if (npl > 0) {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; double h = plane_row(p, jb);
          W(k, o_t(nobs + 4 + i)) = h; hmax = fmax(hmax, h);
        }
      }
      double s = fmax(0.0, hmax + 1e-2);
      W(k, O_S) = s;
      gmax = fmax(gmax, fabs(2 * cfg.S * s));
      for (int r = 0; r < R; ++r) { W(k, o_t(r)) = s - W(k, o_t(r)); W(k, o_z(r)) = 1.0; }
    }
    os = (gmax > 100.0) ? fmax(100.0 / gmax, 1e-8) : 1.0;
    have_step = false; alpha_p = alpha_d = 0; mu_prev = 0;
    mu = cfg.mu_init; reg_last = 0; theta_max = theta_min = -1; E0 = 1e300; nfilt = 0; it = 0;
  }

  // lazy multiplier update of a bound  v - lo >= 0  (sign = +1) or  hi - v >= 0  (sign = -1)
  __device__ __forceinline__ double box_z(double& zref, double v, double dv, double bound, double sgn) const {
    double z = zref;
    if (have_step) {
      double d_old = sgn * ((v - alpha_p * dv) - bound), d_new = sgn * (v - bound);
      double dz = mu_prev / d_old - z - sgn * (z / d_old) * dv;
      z += alpha_d * dz;
      z = fmax(fmin(z, 1e10 * mu_prev / d_new), mu_prev / (1e10 * d_new));
      zref = z;
    }
    return z;
  }

  struct RowAcc {
    double H[21], a[NP], gA[NP], gB[NP], st[NP];
    double csum, be0, be1, zrows, chi, clo, prim, sumz;
    int nz;
  };

  // bookkeeping of one slack row: lazy (t,z) update with slack reset
  __device__ __forceinline__ void row_state(int r, int k, double h, double s, double& z, double& it_, double& res, RowAcc& A) const {
    double t_old = W(k, o_t(r));
    z = W(k, o_z(r));
    double t = t_old;
    if (have_step) {
      double dtv = W(k, o_dt(r));
      double dz = (mu_prev - z * (t_old + dtv)) / t_old;
      z += alpha_d * dz;
      t = fmax(t_old + alpha_p * dtv, s - h);  // slack reset (Nocedal & Wright 19.30)
      z = fmax(fmin(z, 1e10 * mu_prev / t), mu_prev / (1e10 * t));
      W(k, o_t(r)) = t; W(k, o_z(r)) = z;
    }
    it_ = 1.0 / t;
    res = h - s + t;
    A.prim = fmax(A.prim, fabs(res));
    double zt = z * t;
    A.chi = fmax(A.chi, zt); A.clo = fmin(A.clo, zt);
    A.sumz += z; A.zrows += z; A.nz++;
    double sig = z * it_;
    A.csum += sig; A.be0 += sig * res; A.be1 += it_;
  }

  // ------------------------------------------------------------------------------------------
  // eval: commit of the pending step + full evaluation, backward sweep so that x_{k+1}, lam_{k+1}
  // are carried in registers.
  __device__ void eval(KktParts& kp) {
    RowAcc A;
    double e_stat = 0, e_prim = 0, c_hi = -1e300, c_lo = 1e300, sum_lam = 0, sum_z = 0;
    int n_z = 0, n_eq = 0;
    double xn1[NX], lam1[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { xn1[i] = 0; lam1[i] = 0; }
    for (int k = N; k >= 0; --k) {
      double x[NX], u[NU], dxo[NX], duo[NU], lam[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = W(k, O_X + i), d = 0;
        double l = (k >= 1) ? W(k, O_LAM + i) : 0.0;
        if (have_step) {
          d = W(k, O_DX + i); v = fma(alpha_p, d, v); W(k, O_X + i) = v;
          if (k >= 1) { l += alpha_p * (W(k, O_LAMN + i) - l); W(k, O_LAM + i) = l; }
        }
        x[i] = v; dxo[i] = d; lam[i] = l;
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double v = 0, d = 0;
        if (k < N) {
          v = W(k, O_U + j);
          if (have_step) { d = W(k, O_DU + j); v = fma(alpha_p, d, v); W(k, O_U + j) = v; }
        }
        u[j] = v; duo[j] = d;
      }
      double s = W(k, O_S);
      if (have_step) { s = fma(alpha_p, W(k, O_DS), s); W(k, O_S) = s; }
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
      W(k, O_FK + 0) = f.cp; W(k, O_FK + 1) = f.sp;
#pragma unroll
      for (int q = 0; q < 3; ++q) { W(k, O_FK + 2 + q) = f.vr[q]; W(k, O_FK + 5 + q) = f.vh[q]; }
      A.chi = -1e300; A.clo = 1e300; A.prim = 0; A.sumz = 0; A.zrows = 0; A.nz = 0; A.csum = A.be0 = A.be1 = 0;
#pragma unroll
      for (int e = 0; e < 21; ++e) A.H[e] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) A.a[a] = A.gA[a] = A.gB[a] = A.st[a] = 0;
      double es = 0;  // stationarity inf-norm of this stage
      double hpp = 0;
      // dynamics :180 -- defect and costate terms (A^T lam_{k+1}, B^T lam_{k+1})
      double stx[NX], stu[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) stx[i] = 0;
#pragma unroll
      for (int j = 0; j < NU; ++j) stu[j] = 0;
      if (k < N) {
        double xn[NX];
        dyn_f(x, u, dt, f.cp, f.sp, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double d = xn[i] - xn1[i];
          W(k, O_DFC + i) = d; A.prim = fmax(A.prim, fabs(d)); sum_lam += fabs(lam1[i]);
        }
        n_eq += NX;
        hpp = -dt * u[0] * (lam1[3] * f.cp + lam1[4] * f.sp);
        W(k, O_HPU) = dt * (-lam1[3] * f.sp + lam1[4] * f.cp);
        W(k, O_H45) = -dt * lam1[3];  // (dy,dpsi)
        W(k, O_H35) = dt * lam1[4];   // (dx,dpsi)
        stx[0] = lam1[0]; stx[1] = lam1[1];
        stx[2] = lam1[2] + dt * u[0] * (-f.sp * lam1[3] + f.cp * lam1[4]);
        stx[3] = dt * lam1[0] + lam1[3] + dt * x[5] * lam1[4];
        stx[4] = dt * lam1[1] - dt * x[5] * lam1[3] + lam1[4];
        stx[5] = dt * lam1[2] - dt * x[4] * lam1[3] + dt * x[3] * lam1[4] + lam1[5];
        stx[6] = lam1[6]; stx[7] = lam1[7]; stx[8] = lam1[8];
        stu[0] = dt * (f.cp * lam1[3] + f.sp * lam1[4]);
        stu[1] = dt * lam1[5];
        stu[2] = dt * lam1[6]; stu[3] = dt * lam1[7]; stu[4] = dt * lam1[8];
      } else {
        W(k, O_HPU) = 0; W(k, O_H45) = 0; W(k, O_H35) = 0;
      }
      // cost and boxes -- :192-205, :240-245.  Pose components seed the row accumulators, the
      // others are final here.
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]);
        double gr = 2 * Wx * (x[i] - W(k, O_XREF + i));
        double Hd = 2 * Wx, gA = gr, gB = 0, st = gr + stx[i] - lam[i];
        if (k >= 1) {
          double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
          if (is_fin(lo)) {
            double z = box_z(W(k, O_ZXL + i), x[i], dxo[i], lo, 1.0), d = x[i] - lo, id = 1.0 / d;
            Hd += z * id; gB -= id; st -= z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
          if (is_fin(hi)) {
            double z = box_z(W(k, O_ZXU + i), x[i], dxo[i], hi, -1.0), d = hi - x[i], id = 1.0 / d;
            Hd += z * id; gB += id; st += z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
        }
        if (i < 3 || i >= 6) {
          const int a = (i < 3) ? i : i - 3;
          A.H[pidx(a, a)] = Hd + (i == 2 ? hpp : 0.0); A.gA[a] = gA; A.gB[a] = gB; A.st[a] = st;
        } else {
          W(k, O_HVD + (i - 3)) = Hd; W(k, O_GA + i) = gA; W(k, O_GB + i) = gB;
          if (k >= 1) es = fmax(es, fabs(st));
        }
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double Hd = 0, gA = 0, gB = 0;
        if (k < N) {
          double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
          double gr = 2 * Rj * (u[j] - W(k, O_UREF + j)) + 2 * Wj * (u[j] - W(k, O_ULAST + j));
          Hd = 2 * Rj + 2 * Wj; gA = gr;
          double st = gr + stu[j];
          double lo = W(k, O_ULO + j), hi = W(k, O_UHI + j);
          if (is_fin(lo)) {
            double z = box_z(W(k, O_ZUL + j), u[j], duo[j], lo, 1.0), d = u[j] - lo, id = 1.0 / d;
            Hd += z * id; gB -= id; st -= z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
          if (is_fin(hi)) {
            double z = box_z(W(k, O_ZUU + j), u[j], duo[j], hi, -1.0), d = hi - u[j], id = 1.0 / d;
            Hd += z * id; gB += id; st += z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
          es = fmax(es, fabs(st));
        }
        W(k, O_HUU + j) = Hd; W(k, O_GA + GY_U + j) = gA; W(k, O_GB + GY_U + j) = gB;
      }
      // inequality rows with slack:  h(x_k) - s_k + t = 0
      for (int i = 0; i < nobs; ++i) {  // obsAvoid :49-54
        double ddx = x[0] - circ(k, i, 0), ddy = x[1] - circ(k, i, 1);
        double d2 = ddx * ddx + ddy * ddy, inv = rsqrt(d2), d = d2 * inv;
        double h = (circ(k, i, 2) + cfg.base_radius) - d;
        double z, it_, res; row_state(i, k, h, s, z, it_, res, A);
        double sig = z * it_, nx = ddx * inv, ny = ddy * inv, zd = z * inv;
        A.H[pidx(0, 0)] += sig * nx * nx - zd * (1 - nx * nx);
        A.H[pidx(0, 1)] += (sig + zd) * nx * ny;
        A.H[pidx(1, 1)] += sig * ny * ny - zd * (1 - ny * ny);
        double cb = sig * res;
        A.a[0] += sig * nx; A.a[1] += sig * ny; A.gA[0] -= cb * nx; A.gA[1] -= cb * ny;
        A.gB[0] -= it_ * nx; A.gB[1] -= it_ * ny; A.st[0] -= z * nx; A.st[1] -= z * ny;
      }
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {  // self collision :219-222
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        double d2 = p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2], inv = rsqrt(d2);
        double h = cfg.self_collision_radius - d2 * inv;
        double z, it_, res; row_state(nobs + m, k, h, s, z, it_, res, A);
        double sig = z * it_, zd = z * inv;
        double n[3] = {p.P[0] * inv, p.P[1] * inv, p.P[2] * inv}, g[NP];
        point_grad(f, p, n, g);  // grad h = -g
        double cgg = sig + zd;
#pragma unroll
        for (int a = 0; a < NP; ++a)
#pragma unroll
          for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += cgg * g[a] * g[c];
        point_jtj_acc(f, p, -zd, A.H);
        point_hess_acc(f, p, n, -z, A.H);
        double cb = sig * res;
#pragma unroll
        for (int a = 0; a < NP; ++a) { A.a[a] += sig * g[a]; A.gA[a] -= cb * g[a]; A.gB[a] -= it_ * g[a]; A.st[a] -= z * g[a]; }
      }
      if (npl > 0) {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {  // obsAvoidConvex :57-89 (proper row)
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; double h = plane_row(p, jb);
          double z, it_, res; row_state(nobs + 4 + i, k, h, s, z, it_, res, A);
          double sig = z * it_;
          double n[3] = {G(G_PL + 6 * jb + 3), G(G_PL + 6 * jb + 4), G(G_PL + 6 * jb + 5)}, g[NP];
          point_grad(f, p, n, g);  // grad c = +g, grad h = -g
          // h = -c:  grad h = -J^T n ... the row is  -max c <= s, c = off - n.P  =>  grad h = +g
#pragma unroll
          for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += sig * g[a] * g[c];
          point_hess_acc(f, p, n, z, A.H);
          double cb = sig * res;
#pragma unroll
          for (int a = 0; a < NP; ++a) { A.a[a] -= sig * g[a]; A.gA[a] += cb * g[a]; A.gB[a] += it_ * g[a]; A.st[a] += z * g[a]; }
        }
      }
      // slack column of the stage Hessian: H[s][s] = 2S + sum sigma, H[pose][s] = -sum sigma grad h
      double S2 = 2 * os * cfg.S;
      W(k, O_C) = S2 + A.csum;
      W(k, O_GA + GY_S) = S2 * s - A.be0;
      W(k, O_GB + GY_S) = -A.be1;
      W(k, O_HVV) = 0; W(k, O_GA + GY_V) = 0; W(k, O_GB + GY_V) = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) {
        W(k, O_A + a) = A.a[a]; W(k, O_BV + a) = 0;
        W(k, O_GA + POSE2X[a]) = A.gA[a]; W(k, O_GB + POSE2X[a]) = A.gB[a];
        if (k >= 1) es = fmax(es, fabs(A.st[a]));
      }
#pragma unroll
      for (int e = 0; e < 21; ++e) W(k, O_HP + e) = A.H[e];
      es = fmax(es, fabs(S2 * s - A.zrows));
      e_stat = fmax(e_stat, es); e_prim = fmax(e_prim, A.prim);
      c_hi = fmax(c_hi, A.chi); c_lo = fmin(c_lo, A.clo); sum_z += A.sumz; n_z += A.nz;
#pragma unroll
      for (int i = 0; i < NX; ++i) { xn1[i] = x[i]; lam1[i] = lam[i]; }
    }
    kp.e_stat = e_stat; kp.e_prim = e_prim; kp.c_hi = c_hi; kp.c_lo = c_lo;
    kp.sum_lam = sum_lam; kp.sum_z = sum_z; kp.n_z = n_z; kp.n_eq = n_eq;
    have_step = false;  // committed
  }

  __device__ __forceinline__ ACoef acoef(int k) const {
    ACoef c; c.dt = dt; c.cp = W(k, O_FK + 0); c.sp = W(k, O_FK + 1);
    double u0 = W(k, O_U + 0), x3 = W(k, O_X + 3), x4 = W(k, O_X + 4), x5 = W(k, O_X + 5);
    c.a32 = -dt * u0 * c.sp; c.a42 = dt * u0 * c.cp; c.a34 = -dt * x5; c.a43 = dt * x5; c.a35 = -dt * x4; c.a45 = dt * x3;
    return c;
  }
  // v <- A^T v  (in place on a 9-vector)
  __device__ __forceinline__ static void at_mul(double* v, const ACoef& c) {
    double v2 = v[2] + c.a32 * v[3] + c.a42 * v[4];
    double v3 = v[3] + c.dt * v[0] + c.a43 * v[4];
    double v4 = v[4] + c.dt * v[1] + c.a34 * v[3];
    double v5 = v[5] + c.dt * v[2] + c.a35 * v[3] + c.a45 * v[4];
    v[2] = v2; v[3] = v3; v[4] = v4; v[5] = v5;
  }
  // o <- B^T v
  __device__ __forceinline__ static void bt_mul(const double* v, double* o, const ACoef& c) {
    o[0] = c.dt * (c.cp * v[3] + c.sp * v[4]); o[1] = c.dt * v[5];
    o[2] = c.dt * v[6]; o[3] = c.dt * v[7]; o[4] = c.dt * v[8];
  }

  // ------------------------------------------------------------------------------------------
  // Riccati backward recursion in registers.  Stage k eliminates v_k = s_{k+1} (scalar pivot cv)
  // and u_k (5x5 LDL^T); reg = delta_w on the x and u diagonals.  Returns 0, or 1 on a
  // non-positive pivot (wrong inertia).
  __device__ int riccati(double reg) {
    double Pm[45], pv[NX], an[NP], cn, gsn;  // cost-to-go of stage k+1: Pxx, pxx, slack column a, c, g_s
    {
#pragma unroll
      for (int e = 0; e < 45; ++e) Pm[e] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int c = a; c < NP; ++c) Pm[sidx(POSE2X[a], POSE2X[c])] = W(N, O_HP + pidx(a, c));
#pragma unroll
      for (int q = 0; q < 3; ++q) Pm[sidx(3 + q, 3 + q)] = W(N, O_HVD + q);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        Pm[sidx(i, i)] += reg;
        pv[i] = W(N, O_GA + i) + mu * W(N, O_GB + i);
      }
#pragma unroll
      for (int a = 0; a < NP; ++a) an[a] = W(N, O_A + a);
      cn = W(N, O_C); gsn = W(N, O_GA + GY_S) + mu * W(N, O_GB + GY_S);
#pragma unroll
      for (int e = 0; e < 45; ++e) W(N, O_P + e) = Pm[e];
#pragma unroll
      for (int i = 0; i < NX; ++i) W(N, O_PV + i) = pv[i];
    }
    int bad = 0;
    for (int k = N - 1; k >= 0; --k) {
      ACoef c = acoef(k);
      double d[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) d[i] = W(k, O_DFC + i);
      // pd = p + P d ;  l0 = a.d + g_s(k+1) + g_v(k)
      double pd[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = pv[i];
#pragma unroll
        for (int q = 0; q < NX; ++q) v = fma(Pm[sidx(i, q)], d[q], v);
        pd[i] = v;
      }
      double l0 = gsn + W(k, O_GA + GY_V) + mu * W(k, O_GB + GY_V);
#pragma unroll
      for (int a = 0; a < NP; ++a) l0 = fma(an[a], d[POSE2X[a]], l0);
      // C = (P A)[:, 2..5]
      double C[NX][4];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double p0 = Pm[sidx(i, 0)], p1 = Pm[sidx(i, 1)], p2 = Pm[sidx(i, 2)], p3 = Pm[sidx(i, 3)], p4 = Pm[sidx(i, 4)], p5 = Pm[sidx(i, 5)];
        C[i][0] = p2 + c.a32 * p3 + c.a42 * p4;
        C[i][1] = p3 + c.dt * p0 + c.a43 * p4;
        C[i][2] = p4 + c.dt * p1 + c.a34 * p3;
        C[i][3] = p5 + c.dt * p2 + c.a35 * p3 + c.a45 * p4;
      }
#define PA_(l, j) (((j) >= 2 && (j) <= 5) ? C[l][(j) - 2] : Pm[sidx(l, j)])
      // Mux = B^T (P A) + H_ux ; Muu = B^T P B + H_uu + reg ; m_u
      double Mux[NU][NX], Muu[15], mvu[NU];
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        Mux[0][j] = c.dt * (c.cp * PA_(3, j) + c.sp * PA_(4, j));
        Mux[1][j] = c.dt * PA_(5, j);
        Mux[2][j] = c.dt * PA_(6, j); Mux[3][j] = c.dt * PA_(7, j); Mux[4][j] = c.dt * PA_(8, j);
      }
      Mux[0][2] += W(k, O_HPU);
      {
        const double dd = c.dt * c.dt;
        double q33 = Pm[sidx(3, 3)], q34 = Pm[sidx(3, 4)], q44 = Pm[sidx(4, 4)];
        Muu[0] = dd * (c.cp * (c.cp * q33 + c.sp * q34) + c.sp * (c.cp * q34 + c.sp * q44));
        Muu[1] = dd * (c.cp * Pm[sidx(3, 5)] + c.sp * Pm[sidx(4, 5)]);
        Muu[2] = dd * (c.cp * Pm[sidx(3, 6)] + c.sp * Pm[sidx(4, 6)]);
        Muu[3] = dd * (c.cp * Pm[sidx(3, 7)] + c.sp * Pm[sidx(4, 7)]);
        Muu[4] = dd * (c.cp * Pm[sidx(3, 8)] + c.sp * Pm[sidx(4, 8)]);
        // rows 1..4 <-> states 5..8
#pragma unroll
        for (int a = 1; a < NU; ++a)
#pragma unroll
          for (int b2 = a; b2 < NU; ++b2) Muu[a * 5 - a * (a - 1) / 2 + (b2 - a)] = dd * Pm[sidx(4 + a, 4 + b2)];
#pragma unroll
        for (int a = 0; a < NU; ++a) Muu[a * 5 - a * (a - 1) / 2] += W(k, O_HUU + a) + reg;
      }
      bt_mul(pd, mvu, c);
#pragma unroll
      for (int a = 0; a < NU; ++a) mvu[a] += W(k, O_GA + GY_U + a) + mu * W(k, O_GB + GY_U + a);
      // Mxx = A^T (P A) + H_xx + reg (upper triangle, packed)
      double Mxx[45], mvx[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = i; j < NX; ++j) {
          double v;
          if (i == 2) v = PA_(2, j) + c.a32 * PA_(3, j) + c.a42 * PA_(4, j);
          else if (i == 3) v = PA_(3, j) + c.dt * PA_(0, j) + c.a43 * PA_(4, j);
          else if (i == 4) v = PA_(4, j) + c.dt * PA_(1, j) + c.a34 * PA_(3, j);
          else if (i == 5) v = PA_(5, j) + c.dt * PA_(2, j) + c.a35 * PA_(3, j) + c.a45 * PA_(4, j);
          else v = PA_(i, j);
          Mxx[sidx(i, j)] = v;
        }
#undef PA_
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int c2 = a; c2 < NP; ++c2) Mxx[sidx(POSE2X[a], POSE2X[c2])] += W(k, O_HP + pidx(a, c2));
#pragma unroll
      for (int q = 0; q < 3; ++q) Mxx[sidx(3 + q, 3 + q)] += W(k, O_HVD + q);
      Mxx[sidx(3, 5)] += W(k, O_H35); Mxx[sidx(4, 5)] += W(k, O_H45);
#pragma unroll
      for (int i = 0; i < NX; ++i) { Mxx[sidx(i, i)] += reg; mvx[i] = pd[i]; }
      at_mul(mvx, c);
#pragma unroll
      for (int i = 0; i < NX; ++i) mvx[i] += W(k, O_GA + i) + mu * W(k, O_GB + i);
      // eliminate v_k = s_{k+1}:  w = [A^T a(k+1) + bv(k) ; B^T a(k+1)],  cv = hvv(k) + c(k+1)
      double wx[NX], wu[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) wx[i] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) wx[POSE2X[a]] = an[a];
      bt_mul(wx, wu, c);
      at_mul(wx, c);
#pragma unroll
      for (int a = 0; a < NP; ++a) wx[POSE2X[a]] += W(k, O_BV + a);
      double cv = W(k, O_HVV) + cn, icv = 1.0 / cv;
      bad |= !(cv > 1e-13);
      {
        double l0c = l0 * icv;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double wi = wx[i] * icv;
#pragma unroll
          for (int j = i; j < NX; ++j) Mxx[sidx(i, j)] = fma(-wi, wx[j], Mxx[sidx(i, j)]);
          mvx[i] = fma(-wx[i], l0c, mvx[i]);
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          double wa = wu[a] * icv;
#pragma unroll
          for (int j = 0; j < NX; ++j) Mux[a][j] = fma(-wa, wx[j], Mux[a][j]);
#pragma unroll
          for (int b2 = a; b2 < NU; ++b2) Muu[a * 5 - a * (a - 1) / 2 + (b2 - a)] = fma(-wa, wu[b2], Muu[a * 5 - a * (a - 1) / 2 + (b2 - a)]);
          mvu[a] = fma(-wu[a], l0c, mvu[a]);
        }
      }
      // LDL^T of Muu
      double L[NU][NU], D[NU], iD[NU];
#define MUU_(i, j) Muu[(j) * 5 - (j) * ((j) - 1) / 2 + ((i) - (j))]  // i >= j
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double dj = MUU_(j, j);
#pragma unroll
        for (int q = 0; q < j; ++q) dj -= L[j][q] * L[j][q] * D[q];
        bad |= !(dj > 1e-13);
        D[j] = dj; iD[j] = 1.0 / dj;
#pragma unroll
        for (int i = j + 1; i < NU; ++i) {
          double v = MUU_(i, j);
#pragma unroll
          for (int q = 0; q < j; ++q) v -= L[i][q] * L[j][q] * D[q];
          L[i][j] = v * iD[j];
        }
      }
#undef MUU_
      if (bad) return 1;
      // gains: K = -Muu^{-1} Mux (column by column), kff = -Muu^{-1} m_u
      double Kc[NU][NX], kff[NU];
#pragma unroll
      for (int j = 0; j <= NX; ++j) {
        double y[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) y[a] = (j < NX) ? Mux[a][j < NX ? j : 0] : mvu[a];
#pragma unroll
        for (int i = 1; i < NU; ++i)
#pragma unroll
          for (int q = 0; q < i; ++q) y[i] -= L[i][q] * y[q];
#pragma unroll
        for (int i = 0; i < NU; ++i) y[i] *= iD[i];
#pragma unroll
        for (int i = NU - 2; i >= 0; --i)
#pragma unroll
          for (int q = i + 1; q < NU; ++q) y[i] -= L[q][i] * y[q];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          if (j < NX) { Kc[a][j < NX ? j : 0] = -y[a]; W(k, O_K + a * NX + (j < NX ? j : 0)) = -y[a]; }
          else { kff[a] = -y[a]; W(k, O_KFF + a) = -y[a]; }
        }
      }
      // P_k = Mxx + Mux^T K ; p_k = m_x + Mux^T kff
#pragma unroll
      for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = i; j < NX; ++j) {
          double v = Mxx[sidx(i, j)];
#pragma unroll
          for (int a = 0; a < NU; ++a) v = fma(Mux[a][i], Kc[a][j], v);
          Pm[sidx(i, j)] = v; W(k, O_P + sidx(i, j)) = v;
        }
        double v = mvx[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) v = fma(Mux[a][i], kff[a], v);
        pv[i] = v; W(k, O_PV + i) = v;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) W(k, O_W + i) = wx[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) W(k, O_W + NX + a) = wu[a];
      W(k, O_CV) = cv; W(k, O_L0) = l0;
#pragma unroll
      for (int a = 0; a < NP; ++a) an[a] = W(k, O_A + a);
      cn = W(k, O_C); gsn = W(k, O_GA + GY_S) + mu * W(k, O_GB + GY_S);
    }
    if (!(cn > 1e-13)) return 1;
    return 0;
  }

  // ------------------------------------------------------------------------------------------
  // forward roll-out of the Newton step, new costates lam+ = P [dx; ds] + p, slack / multiplier
  // steps of every row and bound, fraction-to-boundary, merit ingredients.
  __device__ void forward(double tau) {
    double ap = 1.0, ad = 1.0, gphi = 0, theta = 0, fsum = 0;
    LogProd lp; lp.init();
    double dxv[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) dxv[i] = 0;
    double dsv = -(W(0, O_GA + GY_S) + mu * W(0, O_GB + GY_S)) / W(0, O_C);
    for (int k = 0; k <= N; ++k) {
      double x[NX], u[NU], duv[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = W(k, O_X + i); W(k, O_DX + i) = dxv[i]; }
      double s = W(k, O_S);
      W(k, O_DS) = dsv;
      double dsn = 0;
      if (k < N) {
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          double v = W(k, O_KFF + a);
#pragma unroll
          for (int j = 0; j < NX; ++j) v = fma(W(k, O_K + a * NX + j), dxv[j], v);
          duv[a] = v; W(k, O_DU + a) = v; u[a] = W(k, O_U + a);
        }
        double l = W(k, O_L0);
#pragma unroll
        for (int i = 0; i < NX; ++i) l = fma(W(k, O_W + i), dxv[i], l);
#pragma unroll
        for (int a = 0; a < NU; ++a) l = fma(W(k, O_W + NX + a), duv[a], l);
        dsn = -l / W(k, O_CV);
      } else {
#pragma unroll
        for (int a = 0; a < NU; ++a) { duv[a] = 0; u[a] = 0; }
      }
      double dp[NP];
#pragma unroll
      for (int a = 0; a < NP; ++a) dp[a] = dxv[POSE2X[a]];
      // cost / boxes
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]), e = x[i] - W(k, O_XREF + i);
        fsum += Wx * e * e; gphi += 2 * Wx * e * dxv[i];
        if (k >= 1) {
          double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
          if (is_fin(lo)) {
            double d = x[i] - lo, z = W(k, O_ZXL + i), dz = mu / d - z - z / d * dxv[i];
            gphi -= mu * dxv[i] / d; lp.mul(d);
            if (dxv[i] < 0) ap = fmin(ap, -tau * d / dxv[i]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
          if (is_fin(hi)) {
            double d = hi - x[i], z = W(k, O_ZXU + i), dz = mu / d - z + z / d * dxv[i];
            gphi += mu * dxv[i] / d; lp.mul(d);
            if (dxv[i] > 0) ap = fmin(ap, tau * d / dxv[i]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
        }
      }
      double S1 = os * cfg.S;
      fsum += S1 * s * s; gphi += 2 * S1 * s * dsv;
      if (k < N) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
          double e = u[j] - W(k, O_UREF + j), dl = u[j] - W(k, O_ULAST + j);
          fsum += Rj * e * e + Wj * dl * dl; gphi += (2 * Rj * e + 2 * Wj * dl) * duv[j];
          double lo = W(k, O_ULO + j), hi = W(k, O_UHI + j);
          if (is_fin(lo)) {
            double d = u[j] - lo, z = W(k, O_ZUL + j), dz = mu / d - z - z / d * duv[j];
            gphi -= mu * duv[j] / d; lp.mul(d);
            if (duv[j] < 0) ap = fmin(ap, -tau * d / duv[j]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
          if (is_fin(hi)) {
            double d = hi - u[j], z = W(k, O_ZUU + j), dz = mu / d - z + z / d * duv[j];
            gphi += mu * duv[j] / d; lp.mul(d);
            if (duv[j] > 0) ap = fmin(ap, tau * d / duv[j]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) theta += fabs(W(k, O_DFC + i));
      }
      FK f; f.cp = W(k, O_FK + 0); f.sp = W(k, O_FK + 1);
#pragma unroll
      for (int q = 0; q < 3; ++q) { f.vr[q] = W(k, O_FK + 2 + q); f.vh[q] = W(k, O_FK + 5 + q); }
      // rows: dt_i = -res_i - (grad h_i . dx - ds)
      auto row_step = [&](int r, double h, double gd) {
        double t = W(k, o_t(r)), z = W(k, o_z(r));
        double res = h - s + t;
        double dtv = -res - (gd - dsv);
        W(k, o_dt(r)) = dtv;
        double dz = (mu - z * (t + dtv)) / t;
        theta += fabs(res); gphi -= mu * dtv / t; lp.mul(t);
        if (dtv < 0) ap = fmin(ap, -tau * t / dtv);
        if (dz < 0) ad = fmin(ad, -tau * z / dz);
      };
      for (int i = 0; i < nobs; ++i) {
        double ddx = x[0] - circ(k, i, 0), ddy = x[1] - circ(k, i, 1);
        double d2 = ddx * ddx + ddy * ddy, inv = rsqrt(d2), d = d2 * inv;
        row_step(i, (circ(k, i, 2) + cfg.base_radius) - d, -(ddx * dp[0] + ddy * dp[1]) * inv);
      }
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        double d2 = p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2], inv = rsqrt(d2), d = d2 * inv;
        double n[3] = {p.P[0] * inv, p.P[1] * inv, p.P[2] * inv}, g[NP];
        point_grad(f, p, n, g);
        double gd = 0;
#pragma unroll
        for (int a = 0; a < NP; ++a) gd = fma(g[a], dp[a], gd);
        row_step(nobs + m, cfg.self_collision_radius - d, -gd);
      }
      if (npl > 0) {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; double h = plane_row(p, jb);
          double n[3] = {G(G_PL + 6 * jb + 3), G(G_PL + 6 * jb + 4), G(G_PL + 6 * jb + 5)}, g[NP];
          point_grad(f, p, n, g);
          double gd = 0;
#pragma unroll
          for (int a = 0; a < NP; ++a) gd = fma(g[a], dp[a], gd);
          row_step(nobs + 4 + i, h, gd);
        }
      }
      // next stage:  dx+ = A dx + B du + d ;  lam+_{k+1} = Pxx dx+ + a ds+ + pxx
      if (k < N) {
        ACoef c = acoef(k);
        double nx_[NX];
        nx_[0] = dxv[0] + dt * dxv[3]; nx_[1] = dxv[1] + dt * dxv[4]; nx_[2] = dxv[2] + dt * dxv[5];
        nx_[3] = dxv[3] + c.a32 * dxv[2] + c.a34 * dxv[4] + c.a35 * dxv[5] + dt * c.cp * duv[0];
        nx_[4] = dxv[4] + c.a42 * dxv[2] + c.a43 * dxv[3] + c.a45 * dxv[5] + dt * c.sp * duv[0];
        nx_[5] = dxv[5] + dt * duv[1];
        nx_[6] = dxv[6] + dt * duv[2]; nx_[7] = dxv[7] + dt * duv[3]; nx_[8] = dxv[8] + dt * duv[4];
#pragma unroll
        for (int i = 0; i < NX; ++i) dxv[i] = nx_[i] + W(k, O_DFC + i);
        dsv = dsn;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = W(k + 1, O_PV + i);
#pragma unroll
          for (int j = 0; j < NX; ++j) v = fma(W(k + 1, O_P + sidx(i, j)), dxv[j], v);
          if (i < 3) v = fma(W(k + 1, O_A + i), dsv, v);
          if (i >= 6) v = fma(W(k + 1, O_A + (i - 3)), dsv, v);
          W(k + 1, O_LAMN + i) = v;
        }
      }
    }
    ls_ap = ap; ls_ad = ad; ls_gphi = gphi; ls_theta = theta; ls_f = fsum; ls_logsum = lp.value();
  }

  struct Trial { double theta, f, logsum; bool ok; };

  // values-only evaluation at  w + alpha d  for the filter line search
  __device__ void trial(double alpha, Trial& tr) {
    double theta = 0, fsum = 0; bool ok = true;
    LogProd lp; lp.init();
    double xk[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xk[i] = fma(alpha, W(0, O_DX + i), W(0, O_X + i));
    for (int k = 0; k <= N; ++k) {
      double x[NX], u[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xk[i];
      double s = fma(alpha, W(k, O_DS), W(k, O_S));
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]), e = x[i] - W(k, O_XREF + i);
        fsum += Wx * e * e;
        if (k >= 1) {
          double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
          if (is_fin(lo)) { double d = x[i] - lo; if (d <= 0) ok = false; else lp.mul(d); }
          if (is_fin(hi)) { double d = hi - x[i]; if (d <= 0) ok = false; else lp.mul(d); }
        }
      }
      fsum += os * cfg.S * s * s;
      if (k < N) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          u[j] = fma(alpha, W(k, O_DU + j), W(k, O_U + j));
          double e = u[j] - W(k, O_UREF + j), dl = u[j] - W(k, O_ULAST + j);
          fsum += os * (cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl);
          double lo = W(k, O_ULO + j), hi = W(k, O_UHI + j);
          if (is_fin(lo)) { double d = u[j] - lo; if (d <= 0) ok = false; else lp.mul(d); }
          if (is_fin(hi)) { double d = hi - u[j]; if (d <= 0) ok = false; else lp.mul(d); }
        }
        double xn[NX]; dyn_f(x, u, dt, f.cp, f.sp, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          xk[i] = fma(alpha, W(k + 1, O_DX + i), W(k + 1, O_X + i));
          theta += fabs(xn[i] - xk[i]);
        }
      }
      auto row_val = [&](int r, double h) {
        double tt = fma(alpha, W(k, o_dt(r)), W(k, o_t(r)));
        tt = fmax(tt, s - h);  // slack reset
        theta += fabs(h - s + tt);
        if (tt <= 0) ok = false; else lp.mul(tt);
      };
      for (int i = 0; i < nobs; ++i) {
        double ddx = x[0] - circ(k, i, 0), ddy = x[1] - circ(k, i, 1);
        row_val(i, (circ(k, i, 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy));
      }
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        row_val(nobs + m, cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]));
      }
      if (npl > 0) {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; row_val(nobs + 4 + i, plane_row(p, jb));
        }
      }
    }
    tr.theta = theta; tr.f = fsum; tr.logsum = lp.value();
    tr.ok = ok && (fsum == fsum) && (theta == theta);
  }

  // results: sol.value(U/X/s/cost) :317,:329-330 (the iterate in the workspace is committed)
  __device__ void finish(int status) {
    double fsum = 0;
    for (int k = 0; k <= N; ++k) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = W(k, O_X + i), e = v - W(k, O_XREF + i);
        fsum += (k < N ? cfg.Qd[i] : cfg.Pd[i]) * e * e;
        if (P.X) P.X[((long long)b * (N + 1) + k) * NX + i] = v;
      }
      if (k < N)
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double v = W(k, O_U + j), e = v - W(k, O_UREF + j), dl = v - W(k, O_ULAST + j);
          fsum += cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl;
          P.U[((long long)b * N + k) * NU + j] = v;
        }
      double s = W(k, O_S);
      fsum += cfg.S * s * s;
      if (P.s) P.s[(long long)b * (N + 1) + k] = s;
    }
    if (P.cost) P.cost[b] = fsum;
    if (P.kkt) P.kkt[b] = E0;
    if (P.iters) P.iters[b] = it;
    P.status[b] = status;
  }
};

// The per-lane driver: identical on the GPU (32 lanes in lock-step, warp votes keep them on the
// same phase) and under tests/emu (one lane at a time, votes are the identity).
__device__ inline void lane_main(const LParams& P, double* lane_base) {
  Lane S(P);
  const MmpcConfig& cfg = P.cfg;
  const double kap_eps = 10, kap_mu = 0.2, th_mu = 1.5, tau_min = 0.99;
  const double tol = cfg.tol;
  bool active = false, drained = false;
  for (;;) {
    if (!active && !drained) {
      unsigned b = lane_next_instance(P.counter);
      if (b < (unsigned)P.B) { S.bind(lane_base, (int)b); S.init(); active = true; }
      else drained = true;
    }
    sync_warp();
    if (!warp_any(active)) break;
    // ---- evaluation, convergence test, barrier update ----
    if (active) {
      KktParts kp;
      S.eval(kp);
      S.E0 = kkt_error(kp, 0.0);
      if (!(S.E0 == S.E0)) { S.finish(MMPC_STATUS_NAN); active = false; }
      else if (S.E0 <= tol) { S.finish(MMPC_STATUS_CONVERGED); active = false; }
      else if (S.it >= cfg.max_iter) { S.finish(MMPC_STATUS_MAX_ITER); active = false; }
      else {
        bool mu_changed = false;
        while (kkt_error(kp, S.mu) <= kap_eps * S.mu && S.mu > tol / 10) {
          S.mu = fmax(tol / 10, fmin(kap_mu * S.mu, pow(S.mu, th_mu))); mu_changed = true;
        }
        if (mu_changed) S.nfilt = 0;
      }
    }
    sync_warp();
    // ---- Riccati factorisation with inertia correction ----
    {
      bool need = active; double reg = 0; int tries = 0;
      while (warp_any(need)) {
        if (need) {
          int fail = S.riccati(reg);
          if (!fail) { need = false; if (reg > 0) S.reg_last = reg; }
          else {
            if (reg == 0) reg = (S.reg_last == 0) ? 1e-4 : fmax(1e-20, S.reg_last / 3);
            else reg *= (S.reg_last == 0 ? 100 : 8);
            if (++tries > 40 || reg > 1e20) { S.finish(MMPC_STATUS_FACTOR); active = false; need = false; }
          }
        }
        sync_warp();
      }
    }
    // ---- Newton step, fraction to the boundary ----
    double tau = fmax(tau_min, 1 - S.mu);
    if (active) S.forward(tau);
    sync_warp();
    // ---- filter line search (Waechter & Biegler 2006, Alg. A without second-order correction) ----
    {
      double theta_k = S.ls_theta, phi0 = S.ls_f - S.mu * S.ls_logsum, gphi = S.ls_gphi;
      if (active && S.theta_max < 0) { S.theta_max = 1e4 * fmax(1.0, theta_k); S.theta_min = 1e-4 * fmax(1.0, theta_k); }
      double alpha = S.ls_ap; bool need = active, ftype = false; int ls = 0;
      while (warp_any(need)) {
        if (need) {
          Lane::Trial tr; S.trial(alpha, tr);
          double th1 = tr.theta, ph1 = tr.f - S.mu * tr.logsum;
          bool ok = tr.ok && th1 < S.theta_max;
          for (int q = 0; ok && q < S.nfilt; ++q)
            if (th1 >= S.G(G_FILT + 2 * q) && ph1 >= S.G(G_FILT + 2 * q + 1)) ok = false;
          if (ok) {
            bool sw = (gphi < 0) && (alpha * pow(-gphi, 2.3) > pow(theta_k, 1.1));
            if (theta_k <= S.theta_min && sw) {
              ok = ph1 <= phi0 + 1e-8 * alpha * gphi + 10 * 2.220446049250313e-16 * fabs(phi0); ftype = ok;
            } else {
              ok = (th1 <= (1 - 1e-5) * theta_k) || (ph1 <= phi0 - 1e-8 * theta_k); ftype = false;
            }
          }
          if (ok) need = false;
          else {
            alpha *= 0.5;
            if (++ls >= 50) { S.finish(MMPC_STATUS_LINESEARCH); active = false; need = false; }
          }
        }
        sync_warp();
      }
      if (active) {
        if (!ftype) {
          if (S.nfilt == 16) {
            for (int q = 0; q < 30; ++q) S.G(G_FILT + q) = S.G(G_FILT + q + 2);
            S.nfilt--;
          }
          S.G(G_FILT + 2 * S.nfilt) = (1 - 1e-5) * theta_k; S.G(G_FILT + 2 * S.nfilt + 1) = phi0 - 1e-8 * theta_k; S.nfilt++;
        }
        S.have_step = true; S.alpha_p = alpha; S.alpha_d = S.ls_ad; S.mu_prev = S.mu; S.it++;
      }
    }
  }
}

#ifndef MMPC_EMULATE_LANE
__global__ void __launch_bounds__(256, 1) lane_kernel(const __grid_constant__ LParams P) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double* base = P.ws + (long long)warp * P.warp_stride + (threadIdx.x & 31);
  lane_main(P, base);
}
#endif

}  // namespace mmpc
