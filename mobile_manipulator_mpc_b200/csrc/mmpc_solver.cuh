// mmpc_solver.cuh -- warp-per-instance primal-dual interior-point solve of the whole-body MPC
// NLP (controllers/mpc_wholebody_qref.py:142-285), replacing opti.solve() (:315) i.e. the
// CasADi/IPOPT/MUMPS stack, for thousands of independent instances.
//
// One warp owns one instance for the whole solve (persistent grid, instances handed out by an
// atomic counter).  Iterate and stage QPs live in shared memory as [field][stage]; per-row
// slacks/multipliers and the Riccati factors live in a per-warp-slot global workspace that stays
// L2-resident.  Phases of one interior-point iteration:
//   eval_full   lane = stage: dynamics, FK, every inequality row with gradient and Hessian,
//               barrier condensation, closed-form elimination of the stage slack s_k, KKT error
//   riccati     lanes cooperate on one stage at a time: backward recursion (nx=9, nu=5, A/B
//               sparsity hard-coded), inertia check on the 5x5 pivots, delta_w restarts
//   forward     sequential roll-out of the Newton step, new costates
//   step_info   lane = stage: slack steps of every row, fraction-to-boundary, merit ingredients
//   trial       lane = stage: values-only evaluation for the filter line search (warp reductions)
// Algorithm = oracle/mmpc_oracle.c (IPOPT-style: monotone mu, tau = max(0.99, 1-mu), bound push,
// gradient-based objective scaling, filter line search with slack reset), "clean" NLP variant.
#pragma once
#include <stdint.h>
#include "../../include/mmpc.h"
#include "mmpc_model.cuh"
#include "mmpc_warp.cuh"
#include "mmpc_ipm.cuh"

namespace mmpc {

struct KParams {
  MmpcConfig cfg;
  int B;
  const double *x_init, *x_ref, *u_ref, *u_last, *u_guess, *circles, *planes;
  const int32_t* n_pl_inst;
  const uint8_t* flags;
  double *U, *X, *s, *cost, *kkt;
  int32_t *iters, *status;
  double* ws;           // per-slot workspace
  long long ws_stride;  // doubles per slot
  unsigned* counter;
  int SP, KP, R;  // smem stage pitch (N+1), global lane pitch, rows per stage
};

// sparse layout of the condensed stage Hessian (32 slots): pose 6x6 packed (21), velocity
// diagonal (3), (dx,dpsi), (dy,dpsi), control diagonal (5), (psi,u0)
constexpr int H_VD = 21, H_X35 = 24, H_X45 = 25, H_UU = 26, H_PU = 31, NH = 32;
constexpr int SCRATCH = 400;

__host__ __device__ inline int smem_doubles(int N) { return (24 + NH + 28 + 9 + 6 + 4 + 8) * (N + 1) + SCRATCH; }
__host__ __device__ inline long long ws_doubles(int N, int KP, int R) {
  return (long long)(3 * R + 28) * KP + (long long)N * 50 + (long long)(N + 1) * 54;
}

// (i,j) state pair -> slot in the sparse Hessian, -1 if structurally zero
__device__ __forceinline__ int qidx(int i, int j) {
  if (i > j) { int t = i; i = j; j = t; }
  const int x2p[9] = {0, 1, 2, -1, -1, -1, 3, 4, 5};
  int pi = x2p[i], pj = x2p[j];
  if (pi >= 0 && pj >= 0) return pidx(pi, pj);
  if (i == j) return H_VD + (i - 3);
  if (i == 3 && j == 5) return H_X35;
  if (i == 4 && j == 5) return H_X45;
  return -1;
}
// upper-triangle enumeration of a symmetric 9x9: e -> (i,j), i<=j
__device__ __forceinline__ void tri9(int e, int& i, int& j) {
  int r = 0, base = 0;
  while (e >= base + (9 - r)) { base += 9 - r; ++r; }
  i = r; j = r + (e - base);
}

struct Solver {
  const KParams& P;
  const MmpcConfig& cfg;
  int lane, N, SP, KP, R, nobs, npl;
  double dt, os;  // objective scale (IPOPT nlp_scaling_max_gradient = 100)
  // shared memory
  double *sx, *su, *ss, *slam, *sH, *sdx, *sdu, *sds, *slamn, *sgA, *sgB, *sdfc, *sa, *sc, *sb0, *sb1, *sfk, *scr;
  // global workspace of this warp slot
  double *gt, *gz, *gdt, *gzxl, *gzxu, *gzul, *gzuu, *gK, *gP;
  // instance inputs
  const double *xref, *uref, *ulast, *circ, *planes;
  int circ_kstride;
  // lazy dual update state
  bool have_step;
  double alpha_p, alpha_d, mu_prev;

  __device__ Solver(const KParams& p) : P(p), cfg(p.cfg) {}

  __device__ __forceinline__ const double* circle_at(int k, int i) const { return circ + (long long)k * circ_kstride + 3 * i; }
  __device__ __forceinline__ void ubox(int k, int j, double& lo, double& hi) const {
    double ul = ldg(ulast + k * NU + j);
    lo = fmax(cfg.ulim[0][j], ul + cfg.dulim[0][j]);  // mpc_wholebody_qref.py:203 and :205 merged
    hi = fmin(cfg.ulim[1][j], ul + cfg.dulim[1][j]);
  }
  __device__ __forceinline__ double plane_off(const double* pl) const {
    return pl[3] * (pl[0] - cfg.obstacle_expand_dist * pl[3]) + pl[4] * (pl[1] - cfg.obstacle_expand_dist * pl[4]) +
           pl[5] * (pl[2] - cfg.obstacle_expand_dist * pl[5]);
  }
  // -max_j c[i][j] for body point i (:76-87); returns the arg-max plane
  __device__ __forceinline__ double plane_row(const Point& p, int& jbest) const {
    double cb = 0; jbest = 0;
    for (int j = 0; j < npl; ++j) {
      const double* pl = planes + 6 * j;
      double c = plane_off(pl) - (pl[3] * p.P[0] + pl[4] * p.P[1] + pl[5] * p.P[2]);
      bool take = (j == 0) || (npl == 2 ? !(cb > c) : (c > cb));  // if_else(c0 > c1, c0, c1) :85 ; mmax :87
      if (take) { cb = c; jbest = j; }
    }
    return -cb;
  }

  // ------------------------------------------------------------------------------------------
  // init: reference initial guess (:302-304) + IPOPT bound push; s lifted so every row starts
  // strictly feasible; objective scaling.
  __device__ void init(int b) {
    double gmax = 0;
    for (int k = lane; k <= N; k += 32) {
      double x[NX], u[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = fmax(fmin(ldg(P.x_init + (long long)b * NX + i), cfg.xlim[1][i]), cfg.xlim[0][i]);  // :290-291
        if (k >= 1) v = push_in(v, cfg.xlim[0][i], cfg.xlim[1][i]);
        x[i] = v; sx[i * SP + k] = v; slam[i * SP + k] = 0;
        gzxl[i * KP + k] = 1; gzxu[i * KP + k] = 1;
        double Wx = (k < N ? cfg.Qd[i] : cfg.Pd[i]);
        if (k >= 1) gmax = fmax(gmax, fabs(2 * Wx * (v - ldg(xref + k * NX + i))));
      }
      if (k < N) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double lo, hi; ubox(k, j, lo, hi);
          double v = P.u_guess ? ldg(P.u_guess + ((long long)b * N + k) * NU + j) : ldg(ulast + k * NU + j);
          v = push_in(v, lo, hi);
          u[j] = v; su[j * SP + k] = v; gzul[j * KP + k] = 1; gzuu[j * KP + k] = 1;
          gmax = fmax(gmax, fabs(2 * cfg.Rd[j] * (v - ldg(uref + k * NU + j)) + 2 * cfg.Wd[j] * (v - ldg(ulast + k * NU + j))));
        }
      }
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
      double hmax = -1e300;
      for (int i = 0; i < nobs; ++i) {
        const double* c = circle_at(k, i);
        double ddx = x[0] - ldg(c), ddy = x[1] - ldg(c + 1);
        double h = (ldg(c + 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy);
        gt[i * KP + k] = h; hmax = fmax(hmax, h);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        double h = cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]);
        gt[(nobs + m) * KP + k] = h; hmax = fmax(hmax, h);
      }
      if (npl > 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; double h = plane_row(p, jb);
          gt[(nobs + 4 + i) * KP + k] = h; hmax = fmax(hmax, h);
        }
      }
      double s = fmax(0.0, hmax + 1e-2);
      ss[k] = s;
      gmax = fmax(gmax, fabs(2 * cfg.S * s));
      for (int r = 0; r < R; ++r) { gt[r * KP + k] = s - gt[r * KP + k]; gz[r * KP + k] = 1.0; }
    }
    gmax = warp_max(gmax);
    os = (gmax > 100.0) ? fmax(100.0 / gmax, 1e-8) : 1.0;
    have_step = false; alpha_p = alpha_d = 0; mu_prev = 0;
    sync_warp();
  }

  // lazy multiplier update of a bound  v - lo >= 0  (sign = +1) or  hi - v >= 0  (sign = -1)
  __device__ __forceinline__ double box_z(double* gzp, double v, double dv, double bound, double sgn) const {
    double z = *gzp;
    if (have_step) {
      double d_old = sgn * ((v - alpha_p * dv) - bound), d_new = sgn * (v - bound);
      double dz = mu_prev / d_old - z - sgn * (z / d_old) * dv;
      z += alpha_d * dz;
      z = fmax(fmin(z, 1e10 * mu_prev / d_new), mu_prev / (1e10 * d_new));
      *gzp = z;
    }
    return z;
  }

  struct RowAcc {
    double H[21], a[NP], bA[NP], bB[NP], st[NP];
    double csum, be0, be1, sumz, zrows, chi, clo, prim;
    int nz;
  };

  // bookkeeping of one slack row: lazy (t,z) update with slack reset, returns sigma etc.
  __device__ __forceinline__ void row_state(int r, int k, double h, double s, double& t, double& z, double& it, double& res,
                                            RowAcc& A) const {
    double t_old = gt[r * KP + k];
    z = gz[r * KP + k];
    t = t_old;
    if (have_step) {
      double dtv = gdt[r * KP + k];
      double dz = (mu_prev - z * (t_old + dtv)) / t_old;
      z += alpha_d * dz;
      t = fmax(t_old + alpha_p * dtv, s - h);  // slack reset (Nocedal & Wright 19.30)
      z = fmax(fmin(z, 1e10 * mu_prev / t), mu_prev / (1e10 * t));
      gt[r * KP + k] = t; gz[r * KP + k] = z;
    }
    it = 1.0 / t;
    res = h - s + t;
    A.prim = fmax(A.prim, fabs(res));
    double zt = z * t;
    A.chi = fmax(A.chi, zt); A.clo = fmin(A.clo, zt);
    A.sumz += z; A.zrows += z; A.nz++;
    double sig = z * it;
    A.csum += sig; A.be0 += sig * res; A.be1 += it;
  }

  // ------------------------------------------------------------------------------------------
  __device__ void eval_full(KktParts& kp) {
    RowAcc A;
    double e_stat = 0, e_prim = 0, c_hi = -1e300, c_lo = 1e300, sum_lam = 0, sum_z = 0;
    int n_z = 0, n_eq = 0;
    for (int k = lane; k <= N; k += 32) {
      double x[NX], u[NU], dxo[NX], duo[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = sx[i * SP + k]; dxo[i] = have_step ? sdx[i * SP + k] : 0.0; }
#pragma unroll
      for (int j = 0; j < NU; ++j) { u[j] = (k < N) ? su[j * SP + k] : 0.0; duo[j] = (have_step && k < N) ? sdu[j * SP + k] : 0.0; }
      double s = ss[k];
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
      sfk[0 * SP + k] = f.cp; sfk[1 * SP + k] = f.sp;
#pragma unroll
      for (int q = 0; q < 3; ++q) { sfk[(2 + q) * SP + k] = f.vr[q]; sfk[(5 + q) * SP + k] = f.vh[q]; }
      double Hd[14], gA[14], gB[14], st[14];  // diagonal Hessian adds, gradient pieces, stationarity
      A.chi = -1e300; A.clo = 1e300; A.prim = 0; A.sumz = 0; A.zrows = 0; A.nz = 0; A.csum = A.be0 = A.be1 = 0;
      // cost and boxes -- :192-205, :240-245
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]);
        double gr = 2 * Wx * (x[i] - ldg(xref + k * NX + i));
        Hd[i] = 2 * Wx; gA[i] = gr; gB[i] = 0; st[i] = gr;
        if (k >= 1) {
          double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
          if (is_fin(lo)) {
            double z = box_z(gzxl + i * KP + k, x[i], dxo[i], lo, 1.0), d = x[i] - lo, id = 1.0 / d;
            Hd[i] += z * id; gB[i] -= id; st[i] -= z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
          if (is_fin(hi)) {
            double z = box_z(gzxu + i * KP + k, x[i], dxo[i], hi, -1.0), d = hi - x[i], id = 1.0 / d;
            Hd[i] += z * id; gB[i] += id; st[i] += z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        Hd[9 + j] = 0; gA[9 + j] = 0; gB[9 + j] = 0; st[9 + j] = 0;
        if (k < N) {
          double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
          double gr = 2 * Rj * (u[j] - ldg(uref + k * NU + j)) + 2 * Wj * (u[j] - ldg(ulast + k * NU + j));
          Hd[9 + j] = 2 * Rj + 2 * Wj; gA[9 + j] = gr; st[9 + j] = gr;
          double lo, hi; ubox(k, j, lo, hi);
          if (is_fin(lo)) {
            double z = box_z(gzul + j * KP + k, u[j], duo[j], lo, 1.0), d = u[j] - lo, id = 1.0 / d;
            Hd[9 + j] += z * id; gB[9 + j] -= id; st[9 + j] -= z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
          if (is_fin(hi)) {
            double z = box_z(gzuu + j * KP + k, u[j], duo[j], hi, -1.0), d = hi - u[j], id = 1.0 / d;
            Hd[9 + j] += z * id; gB[9 + j] += id; st[9 + j] += z;
            A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
          }
        }
      }
      // dynamics :180 -- defect, curvature weighted by the costate, multiplier terms
      double hpp = 0, hpu = 0, h35 = 0, h45 = 0;
      if (k < N) {
        double lam[NX], xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) lam[i] = slam[i * SP + k + 1];
        dyn_f(x, u, dt, f.cp, f.sp, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double d = xn[i] - sx[i * SP + k + 1];
          sdfc[i * SP + k] = d; A.prim = fmax(A.prim, fabs(d)); sum_lam += fabs(lam[i]);
        }
        n_eq += NX;
        hpp = -dt * u[0] * (lam[3] * f.cp + lam[4] * f.sp);
        hpu = dt * (-lam[3] * f.sp + lam[4] * f.cp);
        h45 = -dt * lam[3];  // (dy,dpsi)
        h35 = dt * lam[4];   // (dx,dpsi)
        // A^T lam, B^T lam
        st[0] += lam[0]; st[1] += lam[1];
        st[2] += lam[2] + dt * u[0] * (-f.sp * lam[3] + f.cp * lam[4]);
        st[3] += dt * lam[0] + lam[3] + dt * x[5] * lam[4];
        st[4] += dt * lam[1] - dt * x[5] * lam[3] + lam[4];
        st[5] += dt * lam[2] - dt * x[4] * lam[3] + dt * x[3] * lam[4] + lam[5];
        st[6] += lam[6]; st[7] += lam[7]; st[8] += lam[8];
        st[9] += dt * (f.cp * lam[3] + f.sp * lam[4]);
        st[10] += dt * lam[5];
        st[11] += dt * lam[6]; st[12] += dt * lam[7]; st[13] += dt * lam[8];
      }
      if (k >= 1) {
#pragma unroll
        for (int i = 0; i < NX; ++i) st[i] -= slam[i * SP + k];
      }
      // inequality rows with slack
#pragma unroll
      for (int e = 0; e < 21; ++e) A.H[e] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) A.a[a] = A.bA[a] = A.bB[a] = A.st[a] = 0;
      for (int i = 0; i < nobs; ++i) {  // obsAvoid :49-54
        const double* c = circle_at(k, i);
        double ddx = x[0] - ldg(c), ddy = x[1] - ldg(c + 1);
        double d2 = ddx * ddx + ddy * ddy, inv = rsqrt(d2), d = d2 * inv;
        double h = (ldg(c + 2) + cfg.base_radius) - d;
        double t, z, it, res; row_state(i, k, h, s, t, z, it, res, A);
        double sig = z * it, nx = ddx * inv, ny = ddy * inv, zd = z * inv;
        A.H[pidx(0, 0)] += sig * nx * nx - zd * (1 - nx * nx);
        A.H[pidx(0, 1)] += (sig + zd) * nx * ny;
        A.H[pidx(1, 1)] += sig * ny * ny - zd * (1 - ny * ny);
        double ca = sig, cb = sig * res;
        A.a[0] -= ca * nx; A.a[1] -= ca * ny; A.bA[0] -= cb * nx; A.bA[1] -= cb * ny;
        A.bB[0] -= it * nx; A.bB[1] -= it * ny; A.st[0] -= z * nx; A.st[1] -= z * ny;
      }
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {  // self collision :219-222
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        double d2 = p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2], inv = rsqrt(d2), d = d2 * inv;
        double h = cfg.self_collision_radius - d;
        double t, z, it, res; row_state(nobs + m, k, h, s, t, z, it, res, A);
        double sig = z * it, zd = z * inv;
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
        for (int a = 0; a < NP; ++a) { A.a[a] -= sig * g[a]; A.bA[a] -= cb * g[a]; A.bB[a] -= it * g[a]; A.st[a] -= z * g[a]; }
      }
      if (npl > 0) {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {  // obsAvoidConvex :57-89 (proper row)
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; double h = plane_row(p, jb);
          double t, z, it, res; row_state(nobs + 4 + i, k, h, s, t, z, it, res, A);
          double sig = z * it;
          const double* pl = planes + 6 * jb;
          double n[3] = {pl[3], pl[4], pl[5]}, g[NP];
          point_grad(f, p, n, g);  // grad h = +g
#pragma unroll
          for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += sig * g[a] * g[c];
          point_hess_acc(f, p, n, z, A.H);
          double cb = sig * res;
#pragma unroll
          for (int a = 0; a < NP; ++a) { A.a[a] += sig * g[a]; A.bA[a] += cb * g[a]; A.bB[a] += it * g[a]; A.st[a] += z * g[a]; }
        }
      }
      // eliminate s_k: c = 2S + sum sigma, beta = mu*be1 + (be0 - 2 S s)
      double S2 = 2 * os * cfg.S;
      double c = S2 + A.csum, ic = 1.0 / c;
      double b0 = A.be0 - S2 * s;
      sc[k] = c; sb0[k] = b0; sb1[k] = A.be1;
#pragma unroll
      for (int a = 0; a < NP; ++a) {
        sa[a * SP + k] = A.a[a];
        gA[POSE2X[a]] += A.bA[a] - A.a[a] * b0 * ic;
        gB[POSE2X[a]] += A.bB[a] - A.a[a] * A.be1 * ic;
        st[POSE2X[a]] += A.st[a];
      }
      // store the condensed stage QP
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int cc = a; cc < NP; ++cc) {
          double v = A.H[pidx(a, cc)] - A.a[a] * A.a[cc] * ic;
          if (a == cc) v += Hd[POSE2X[a]];
          if (a == 2 && cc == 2) v += hpp;
          sH[pidx(a, cc) * SP + k] = v;
        }
      sH[(H_VD + 0) * SP + k] = Hd[3]; sH[(H_VD + 1) * SP + k] = Hd[4]; sH[(H_VD + 2) * SP + k] = Hd[5];
      sH[H_X35 * SP + k] = h35; sH[H_X45 * SP + k] = h45; sH[H_PU * SP + k] = hpu;
#pragma unroll
      for (int j = 0; j < NU; ++j) sH[(H_UU + j) * SP + k] = Hd[9 + j];
#pragma unroll
      for (int i = 0; i < 14; ++i) { sgA[i * SP + k] = gA[i]; sgB[i * SP + k] = gB[i]; }
      // stationarity residual of this stage
      double es = 0;
      if (k >= 1)
#pragma unroll
        for (int i = 0; i < NX; ++i) es = fmax(es, fabs(st[i]));
      if (k < N)
#pragma unroll
        for (int j = 0; j < NU; ++j) es = fmax(es, fabs(st[9 + j]));
      es = fmax(es, fabs(S2 * s - A.zrows));
      e_stat = fmax(e_stat, es); e_prim = fmax(e_prim, A.prim);
      c_hi = fmax(c_hi, A.chi); c_lo = fmin(c_lo, A.clo); sum_z += A.sumz; n_z += A.nz;
    }
    kp.e_stat = warp_max(e_stat); kp.e_prim = warp_max(e_prim);
    kp.c_hi = warp_max(c_hi); kp.c_lo = warp_min(c_lo);
    kp.sum_lam = warp_sum(sum_lam); kp.sum_z = warp_sum(sum_z);
    kp.n_z = warp_sum(n_z); kp.n_eq = warp_sum(n_eq);
    sync_warp();
  }
  // ------------------------------------------------------------------------------------------
  // Riccati backward recursion over the condensed stage QPs.  Scratch layout (doubles):
  //   Pm 0..80 | pv 81..89 | Wm 90..170 | pd 171..179 | Mxx 180..260 | Mux 261..305 | Muu 306..330
  //   | mx 331..339 | mu 340..344 | Kc 345..394
  __device__ __forceinline__ double gval(int i, int k, double mu) const { return sgA[i * SP + k] + mu * sgB[i * SP + k]; }
  __device__ __forceinline__ double hval(int i, int j, int k) const {
    int q = qidx(i, j);
    return q >= 0 ? sH[q * SP + k] : 0.0;
  }

  struct ACoef { double dt, a32, a42, a34, a43, a35, a45, cp, sp; };
  __device__ __forceinline__ ACoef acoef(int k) const {
    ACoef c; c.dt = dt; c.cp = sfk[0 * SP + k]; c.sp = sfk[1 * SP + k];
    double u0 = su[0 * SP + k], x3 = sx[3 * SP + k], x4 = sx[4 * SP + k], x5 = sx[5 * SP + k];
    c.a32 = -dt * u0 * c.sp; c.a42 = dt * u0 * c.cp; c.a34 = -dt * x5; c.a43 = dt * x5; c.a35 = -dt * x4; c.a45 = dt * x3;
    return c;
  }
  // (X A)[i][j], X row-major 9x9
  __device__ __forceinline__ static double colA(const double* X, int i, int j, const ACoef& c) {
    const double* r = X + 9 * i;
    switch (j) {
      case 2: return r[2] + c.a32 * r[3] + c.a42 * r[4];
      case 3: return r[3] + c.dt * r[0] + c.a43 * r[4];
      case 4: return r[4] + c.dt * r[1] + c.a34 * r[3];
      case 5: return r[5] + c.dt * r[2] + c.a35 * r[3] + c.a45 * r[4];
      default: return r[j];
    }
  }
  // (A^T X)[i][j], X with row stride ld
  __device__ __forceinline__ static double rowA(const double* X, int ld, int i, int j, const ACoef& c) {
    const double* q = X + j;
    switch (i) {
      case 2: return q[2 * ld] + c.a32 * q[3 * ld] + c.a42 * q[4 * ld];
      case 3: return q[3 * ld] + c.dt * q[0] + c.a43 * q[4 * ld];
      case 4: return q[4 * ld] + c.dt * q[1 * ld] + c.a34 * q[3 * ld];
      case 5: return q[5 * ld] + c.dt * q[2 * ld] + c.a35 * q[3 * ld] + c.a45 * q[4 * ld];
      default: return q[i * ld];
    }
  }
  // (B^T X)[a][j]
  __device__ __forceinline__ static double rowB(const double* X, int ld, int a, int j, const ACoef& c) {
    const double* q = X + j;
    if (a == 0) return c.dt * (c.cp * q[3 * ld] + c.sp * q[4 * ld]);
    return c.dt * q[(4 + a) * ld];
  }

  __device__ int riccati_backward(double reg, double mu) {
    double *Pm = scr, *pv = scr + 81, *Wm = scr + 90, *pd = scr + 171, *Mxx = scr + 180, *Mux = scr + 261,
           *Muu = scr + 306, *mx = scr + 331, *mu_ = scr + 340, *Kc = scr + 345;
    for (int e = lane; e < 90; e += 32) {
      if (e < 81) { int i = e / 9, j = e % 9; Pm[e] = hval(i, j, N) + (i == j ? reg : 0.0); }
      else pv[e - 81] = gval(e - 81, N, mu);
    }
    sync_warp();
    for (int w = lane; w < 54; w += 32) {
      if (w < 45) { int i, j; tri9(w, i, j); gP[(long long)N * 54 + w] = Pm[i * 9 + j]; }
      else gP[(long long)N * 54 + w] = pv[w - 45];
    }
    for (int k = N - 1; k >= 0; --k) {
      ACoef c = acoef(k);
      // step 1: W = P A, pd = p + P d
      for (int e = lane; e < 90; e += 32) {
        if (e < 81) Wm[e] = colA(Pm, e / 9, e % 9, c);
        else {
          int i = e - 81; double v = pv[i];
#pragma unroll
          for (int q = 0; q < 9; ++q) v = fma(Pm[i * 9 + q], sdfc[q * SP + k], v);
          pd[i] = v;
        }
      }
      sync_warp();
      // step 2: M = H + [A B]^T P [A B], m = g + [A B]^T pd
      for (int w = lane; w < 119; w += 32) {
        if (w < 45) {
          int i, j; tri9(w, i, j);
          double v = hval(i, j, k) + rowA(Wm, 9, i, j, c) + (i == j ? reg : 0.0);
          Mxx[i * 9 + j] = v; Mxx[j * 9 + i] = v;
        } else if (w < 90) {
          int e = w - 45, a = e / 9, j = e % 9;
          Mux[e] = ((a == 0 && j == 2) ? sH[H_PU * SP + k] : 0.0) + rowB(Wm, 9, a, j, c);
        } else if (w < 105) {
          int e = w - 90, a = 0, base = 0;
          while (e >= base + (5 - a)) { base += 5 - a; ++a; }
          int b2 = a + (e - base);
          // (Bd^T P Bd)[a][b]
          int jb = (b2 == 0) ? 3 : 4 + b2;
          double pr = (a == 0) ? c.cp * Pm[3 * 9 + jb] + c.sp * Pm[4 * 9 + jb] : Pm[(4 + a) * 9 + jb];
          if (b2 == 0) {
            double pr4 = (a == 0) ? c.cp * Pm[3 * 9 + 4] + c.sp * Pm[4 * 9 + 4] : Pm[(4 + a) * 9 + 4];
            pr = c.cp * pr + c.sp * pr4;
          }
          double v = c.dt * c.dt * pr + (a == b2 ? sH[(H_UU + a) * SP + k] + reg : 0.0);
          Muu[a * 5 + b2] = v; Muu[b2 * 5 + a] = v;
        } else if (w < 114) {
          int i = w - 105;
          mx[i] = gval(i, k, mu) + rowA(pd, 1, i, 0, c);
        } else {
          int a = w - 114;
          mu_[a] = gval(9 + a, k, mu) + rowB(pd, 1, a, 0, c);
        }
      }
      sync_warp();
      // step 3: LDL^T of the 5x5 pivot block (redundantly in every lane), gains by lanes 0..9
      double L[5][5], D[5];
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double d = Muu[j * 5 + j];
#pragma unroll
        for (int q = 0; q < j; ++q) d -= L[j][q] * L[j][q] * D[q];
        bad = bad || !(d > 1e-13);
        D[j] = d;
        double id = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 5; ++i) {
          double v = Muu[i * 5 + j];
#pragma unroll
          for (int q = 0; q < j; ++q) v -= L[i][q] * L[j][q] * D[q];
          L[i][j] = v * id;
        }
      }
      if (warp_any(bad)) return 1;
      if (lane < 10) {
        double y[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) y[a] = (lane < 9) ? Mux[a * 9 + lane] : mu_[a];
#pragma unroll
        for (int i = 1; i < 5; ++i)
#pragma unroll
          for (int q = 0; q < i; ++q) y[i] -= L[i][q] * y[q];
#pragma unroll
        for (int i = 0; i < 5; ++i) y[i] /= D[i];
#pragma unroll
        for (int i = 3; i >= 0; --i)
#pragma unroll
          for (int q = i + 1; q < 5; ++q) y[i] -= L[q][i] * y[q];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          int o = (lane < 9) ? a * 9 + lane : 45 + a;
          Kc[o] = -y[a]; gK[(long long)k * 50 + o] = -y[a];
        }
      }
      sync_warp();
      // step 4: P_k = Mxx + Mxu K, p_k = mx + Mxu kff
      for (int w = lane; w < 54; w += 32) {
        if (w < 45) {
          int i, j; tri9(w, i, j);
          double v = Mxx[i * 9 + j];
#pragma unroll
          for (int a = 0; a < 5; ++a) v = fma(Mux[a * 9 + i], Kc[a * 9 + j], v);
          Pm[i * 9 + j] = v; Pm[j * 9 + i] = v; gP[(long long)k * 54 + w] = v;
        } else {
          int i = w - 45; double v = mx[i];
#pragma unroll
          for (int a = 0; a < 5; ++a) v = fma(Mux[a * 9 + i], Kc[45 + a], v);
          pv[i] = v; gP[(long long)k * 54 + w] = v;
        }
      }
      sync_warp();
    }
    return 0;
  }

  // forward roll-out of the Newton step and the new costates lam+ = P dx + p
  __device__ void forward() {
    if (lane < 9) { sdx[lane * SP + 0] = 0.0; slamn[lane * SP + 0] = 0.0; }
    sync_warp();
    for (int k = 0; k < N; ++k) {
      if (lane < 5) {
        const double* Kk = gK + (long long)k * 50;
        double v = Kk[45 + lane];
#pragma unroll
        for (int j = 0; j < 9; ++j) v = fma(Kk[lane * 9 + j], sdx[j * SP + k], v);
        sdu[lane * SP + k] = v;
      }
      sync_warp();
      if (lane < 9) {
        ACoef c = acoef(k);
        double d[9], du[5];
#pragma unroll
        for (int j = 0; j < 9; ++j) d[j] = sdx[j * SP + k];
#pragma unroll
        for (int j = 0; j < 5; ++j) du[j] = sdu[j * SP + k];
        double v;
        switch (lane) {
          case 0: v = d[0] + dt * d[3]; break;
          case 1: v = d[1] + dt * d[4]; break;
          case 2: v = d[2] + dt * d[5]; break;
          case 3: v = d[3] + c.a32 * d[2] + c.a34 * d[4] + c.a35 * d[5] + dt * c.cp * du[0]; break;
          case 4: v = d[4] + c.a42 * d[2] + c.a43 * d[3] + c.a45 * d[5] + dt * c.sp * du[0]; break;
          case 5: v = d[5] + dt * du[1]; break;
          default: v = d[lane] + dt * du[lane - 4]; break;
        }
        sdx[lane * SP + k + 1] = v + sdfc[lane * SP + k];
      }
      sync_warp();
      if (lane < 9) {
        const double* Pk = gP + (long long)(k + 1) * 54;
        double v = Pk[45 + lane];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          int i0 = lane < j ? lane : j, j0 = lane < j ? j : lane;
          v = fma(Pk[i0 * 9 - i0 * (i0 - 1) / 2 + (j0 - i0)], sdx[j * SP + k + 1], v);
        }
        slamn[lane * SP + k + 1] = v;
      }
    }
    sync_warp();
  }

  struct StepInfo { double ap, ad, gphi, theta, f, logsum; };

  // slack / multiplier steps of every row and bound, fraction-to-boundary, merit ingredients
  __device__ void step_info(double mu, double tau, StepInfo& si) {
    double ap = 1.0, ad = 1.0, gphi = 0, theta = 0, fsum = 0;
    LogProd lp; lp.init();
    for (int k = lane; k <= N; k += 32) {
      double x[NX], dxv[NX], u[NU], duv[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = sx[i * SP + k]; dxv[i] = sdx[i * SP + k]; }
#pragma unroll
      for (int j = 0; j < NU; ++j) { u[j] = (k < N) ? su[j * SP + k] : 0.0; duv[j] = (k < N) ? sdu[j * SP + k] : 0.0; }
      double s = ss[k];
      double dp[NP];
#pragma unroll
      for (int a = 0; a < NP; ++a) dp[a] = dxv[POSE2X[a]];
      double dsv = mu * sb1[k] + sb0[k];
#pragma unroll
      for (int a = 0; a < NP; ++a) dsv = fma(sa[a * SP + k], dp[a], dsv);
      dsv /= sc[k];
      sds[k] = dsv;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]), e = x[i] - ldg(xref + k * NX + i);
        fsum += Wx * e * e; gphi += 2 * Wx * e * dxv[i];
        if (k >= 1) {
          double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
          if (is_fin(lo)) {
            double d = x[i] - lo, z = gzxl[i * KP + k], dz = mu / d - z - z / d * dxv[i];
            gphi -= mu * dxv[i] / d; lp.mul(d);
            if (dxv[i] < 0) ap = fmin(ap, -tau * d / dxv[i]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
          if (is_fin(hi)) {
            double d = hi - x[i], z = gzxu[i * KP + k], dz = mu / d - z + z / d * dxv[i];
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
          double e = u[j] - ldg(uref + k * NU + j), dl = u[j] - ldg(ulast + k * NU + j);
          fsum += Rj * e * e + Wj * dl * dl; gphi += (2 * Rj * e + 2 * Wj * dl) * duv[j];
          double lo, hi; ubox(k, j, lo, hi);
          if (is_fin(lo)) {
            double d = u[j] - lo, z = gzul[j * KP + k], dz = mu / d - z - z / d * duv[j];
            gphi -= mu * duv[j] / d; lp.mul(d);
            if (duv[j] < 0) ap = fmin(ap, -tau * d / duv[j]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
          if (is_fin(hi)) {
            double d = hi - u[j], z = gzuu[j * KP + k], dz = mu / d - z + z / d * duv[j];
            gphi += mu * duv[j] / d; lp.mul(d);
            if (duv[j] > 0) ap = fmin(ap, tau * d / duv[j]);
            if (dz < 0) ad = fmin(ad, -tau * z / dz);
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) theta += fabs(sdfc[i * SP + k]);
      }
      FK f; f.cp = sfk[0 * SP + k]; f.sp = sfk[1 * SP + k];
#pragma unroll
      for (int q = 0; q < 3; ++q) { f.vr[q] = sfk[(2 + q) * SP + k]; f.vh[q] = sfk[(5 + q) * SP + k]; }
      // rows: dt_i = -res_i - (grad h_i . dw - ds)
      auto row_step = [&](int r, double h, double gd) {
        double t = gt[r * KP + k], z = gz[r * KP + k];
        double res = h - s + t;
        double dtv = -res - (gd - dsv);
        gdt[r * KP + k] = dtv;
        double dz = (mu - z * (t + dtv)) / t;
        theta += fabs(res); gphi -= mu * dtv / t; lp.mul(t);
        if (dtv < 0) ap = fmin(ap, -tau * t / dtv);
        if (dz < 0) ad = fmin(ad, -tau * z / dz);
      };
      for (int i = 0; i < nobs; ++i) {
        const double* c = circle_at(k, i);
        double ddx = x[0] - ldg(c), ddy = x[1] - ldg(c + 1);
        double d2 = ddx * ddx + ddy * ddy, inv = rsqrt(d2), d = d2 * inv;
        row_step(i, (ldg(c + 2) + cfg.base_radius) - d, -(ddx * dp[0] + ddy * dp[1]) * inv);
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
          const double* pl = planes + 6 * jb;
          double n[3] = {pl[3], pl[4], pl[5]}, g[NP];
          point_grad(f, p, n, g);
          double gd = 0;
#pragma unroll
          for (int a = 0; a < NP; ++a) gd = fma(g[a], dp[a], gd);
          row_step(nobs + 4 + i, h, gd);
        }
      }
    }
    si.ap = warp_min(ap); si.ad = warp_min(ad);
    si.gphi = warp_sum(gphi); si.theta = warp_sum(theta);
    si.f = warp_sum(fsum); si.logsum = warp_sum(lp.value());
    sync_warp();
  }

  struct Trial { double theta, f, logsum; bool ok; };

  // values-only evaluation at  w + alpha d  for the filter line search
  __device__ void trial(double alpha, Trial& tr) {
    double theta = 0, fsum = 0; bool ok = true;
    LogProd lp; lp.init();
    for (int k = lane; k <= N; k += 32) {
      double x[NX], u[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = fma(alpha, sdx[i * SP + k], sx[i * SP + k]);
#pragma unroll
      for (int j = 0; j < NU; ++j) u[j] = (k < N) ? fma(alpha, sdu[j * SP + k], su[j * SP + k]) : 0.0;
      double s = fma(alpha, sds[k], ss[k]);
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]), e = x[i] - ldg(xref + k * NX + i);
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
          double e = u[j] - ldg(uref + k * NU + j), dl = u[j] - ldg(ulast + k * NU + j);
          fsum += os * (cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl);
          double lo, hi; ubox(k, j, lo, hi);
          if (is_fin(lo)) { double d = u[j] - lo; if (d <= 0) ok = false; else lp.mul(d); }
          if (is_fin(hi)) { double d = hi - u[j]; if (d <= 0) ok = false; else lp.mul(d); }
        }
        double xn[NX]; dyn_f(x, u, dt, f.cp, f.sp, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) theta += fabs(xn[i] - fma(alpha, sdx[i * SP + k + 1], sx[i * SP + k + 1]));
      }
      auto row_val = [&](int r, double h) {
        double tt = fma(alpha, gdt[r * KP + k], gt[r * KP + k]);
        tt = fmax(tt, s - h);  // slack reset
        theta += fabs(h - s + tt);
        if (tt <= 0) ok = false; else lp.mul(tt);
      };
      for (int i = 0; i < nobs; ++i) {
        const double* c = circle_at(k, i);
        double ddx = x[0] - ldg(c), ddy = x[1] - ldg(c + 1);
        row_val(i, (ldg(c + 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy));
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
    tr.theta = warp_sum(theta); tr.f = warp_sum(fsum); tr.logsum = warp_sum(lp.value());
    tr.ok = !warp_any(!ok) && (tr.f == tr.f) && (tr.theta == tr.theta);
  }

  __device__ void commit(double alpha) {
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        sx[i * SP + k] = fma(alpha, sdx[i * SP + k], sx[i * SP + k]);
        if (k >= 1) slam[i * SP + k] += alpha * (slamn[i * SP + k] - slam[i * SP + k]);
      }
      if (k < N)
#pragma unroll
        for (int j = 0; j < NU; ++j) su[j * SP + k] = fma(alpha, sdu[j * SP + k], su[j * SP + k]);
      ss[k] = fma(alpha, sds[k], ss[k]);
    }
    sync_warp();
  }

  // ------------------------------------------------------------------------------------------
  __device__ void run(int b) {
    init(b);
    double mu = cfg.mu_init, tol = cfg.tol, reg_last = 0;
    const double kap_eps = 10, kap_mu = 0.2, th_mu = 1.5, tau_min = 0.99;
    double fth = 0, fph = 0;  // lane q holds filter entry q
    int nfilt = 0;
    double theta_max = -1, theta_min = -1, E0 = 1e300;
    int status = MMPC_STATUS_MAX_ITER, it = 0;
    KktParts kp;
    for (it = 0; it <= cfg.max_iter; ++it) {
      eval_full(kp);
      E0 = kkt_error(kp, 0.0);
      if (!(E0 == E0)) { status = MMPC_STATUS_NAN; break; }
      if (E0 <= tol) { status = MMPC_STATUS_CONVERGED; break; }
      if (it == cfg.max_iter) break;
      bool mu_changed = false;
      while (kkt_error(kp, mu) <= kap_eps * mu && mu > tol / 10) {
        mu = fmax(tol / 10, fmin(kap_mu * mu, pow(mu, th_mu))); mu_changed = true;
      }
      if (mu_changed) nfilt = 0;
      double tau = fmax(tau_min, 1 - mu);
      double reg = 0; int tries = 0, fail;
      while ((fail = riccati_backward(reg, mu)) != 0) {
        if (reg == 0) reg = (reg_last == 0) ? 1e-4 : fmax(1e-20, reg_last / 3);
        else reg *= (reg_last == 0 ? 100 : 8);
        if (++tries > 40 || reg > 1e20) break;
      }
      if (fail) { status = MMPC_STATUS_FACTOR; break; }
      if (reg > 0) reg_last = reg;
      forward();
      StepInfo si; step_info(mu, tau, si);
      double theta_k = si.theta, phi0 = si.f - mu * si.logsum, gphi = si.gphi;
      if (theta_max < 0) { theta_max = 1e4 * fmax(1.0, theta_k); theta_min = 1e-4 * fmax(1.0, theta_k); }
      double alpha = si.ap; bool accepted = false, ftype = false;
      for (int ls = 0; ls < 50; ++ls) {
        Trial tr; trial(alpha, tr);
        double th1 = tr.theta, ph1 = tr.f - mu * tr.logsum;
        bool ok = tr.ok && th1 < theta_max;
        bool dominated = (lane < nfilt) && (th1 >= fth) && (ph1 >= fph);
        if (warp_any(dominated)) ok = false;
        if (ok) {
          bool sw = (gphi < 0) && (alpha * pow(-gphi, 2.3) > pow(theta_k, 1.1));
          if (theta_k <= theta_min && sw) {
            ok = ph1 <= phi0 + 1e-8 * alpha * gphi + 10 * 2.220446049250313e-16 * fabs(phi0); ftype = ok;
          } else {
            ok = (th1 <= (1 - 1e-5) * theta_k) || (ph1 <= phi0 - 1e-8 * theta_k); ftype = false;
          }
        }
        if (ok) { accepted = true; break; }
        alpha *= 0.5;
      }
      if (!accepted) { status = MMPC_STATUS_LINESEARCH; break; }
      if (!ftype) {
        if (lane == (nfilt & 31)) { fth = (1 - 1e-5) * theta_k; fph = phi0 - 1e-8 * theta_k; }
        if (nfilt < 32) nfilt++;
      }
      commit(alpha);
      have_step = true; alpha_p = alpha; alpha_d = si.ad; mu_prev = mu;
    }
    // results: sol.value(U/X/s/cost) :317,:329-330
    double fsum = 0;
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = sx[i * SP + k], e = v - ldg(xref + k * NX + i);
        fsum += (k < N ? cfg.Qd[i] : cfg.Pd[i]) * e * e;
        if (P.X) P.X[((long long)b * (N + 1) + k) * NX + i] = v;
      }
      if (k < N)
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double v = su[j * SP + k], e = v - ldg(uref + k * NU + j), dl = v - ldg(ulast + k * NU + j);
          fsum += cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl;
          P.U[((long long)b * N + k) * NU + j] = v;
        }
      double s = ss[k];
      fsum += cfg.S * s * s;
      if (P.s) P.s[(long long)b * (N + 1) + k] = s;
    }
    fsum = warp_sum(fsum);
    if (lane == 0) {
      if (P.cost) P.cost[b] = fsum;
      if (P.kkt) P.kkt[b] = E0;
      if (P.iters) P.iters[b] = it;
      P.status[b] = status;
    }
    sync_warp();
  }

  __device__ void bind(double* smem, int slot, int b) {
    N = cfg.N; SP = P.SP; KP = P.KP; R = P.R; nobs = cfg.n_obs; dt = cfg.dt;
    npl = P.n_pl_inst ? ldg(P.n_pl_inst + b) : cfg.n_pl;
    double* q = smem;
    sx = q; q += 9 * SP; su = q; q += 5 * SP; ss = q; q += SP; slam = q; q += 9 * SP;
    sH = q; sdx = q; sdu = q + 9 * SP; sds = q + 14 * SP; slamn = q + 15 * SP; q += NH * SP;
    sgA = q; q += 14 * SP; sgB = q; q += 14 * SP; sdfc = q; q += 9 * SP; sa = q; q += 6 * SP;
    sc = q; q += SP; sb0 = q; q += SP; sb1 = q; q += SP; q += SP; sfk = q; q += 8 * SP; scr = q;
    double* g = P.ws + (long long)slot * P.ws_stride;
    gt = g; g += (long long)R * KP; gz = g; g += (long long)R * KP; gdt = g; g += (long long)R * KP;
    gzxl = g; g += 9 * KP; gzxu = g; g += 9 * KP; gzul = g; g += 5 * KP; gzuu = g; g += 5 * KP;
    gK = g; g += (long long)N * 50; gP = g;
    xref = P.x_ref + (long long)b * (N + 1) * NX; uref = P.u_ref + (long long)b * N * NU; ulast = P.u_last + (long long)b * N * NU;
    long long cper = (long long)nobs * 3 * (cfg.obs_per_stage ? (N + 1) : 1);
    circ = P.circles ? P.circles + (long long)b * cper : nullptr;
    circ_kstride = cfg.obs_per_stage ? nobs * 3 : 0;
    planes = P.planes ? P.planes + (long long)b * cfg.n_pl * 6 : nullptr;
  }
};

// persistent kernel: one warp per block, one instance per warp at a time
#ifndef MMPC_EMULATE
extern __shared__ double mmpc_smem[];
__global__ void __launch_bounds__(32) solve_kernel(const __grid_constant__ KParams P) {
  Solver S(P);
  S.lane = lane_id();
  for (;;) {
    unsigned b = next_instance(P.counter);
    if (b >= (unsigned)P.B) break;
    S.bind(mmpc_smem, blockIdx.x, (int)b);
    S.run((int)b);
  }
}
#endif

}  // namespace mmpc
