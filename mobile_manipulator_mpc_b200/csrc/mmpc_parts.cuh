// mmpc_parts.cuh -- the two stage-parallel phases of a round (step, trial+evaluation), with every
// (instance, stage) item split into PARTS that different warps of a block execute:
//
//     part 0            base-x   states: candidate x, lam, dynamics defect and costates, state cost and boxes, FK cache
//     part 1            base-u   controls: candidate u, control cost and boxes
//     parts 2..         circles  groups of up to 8 ground-circle rows                 (obsAvoid :49-54)
//     next 2 parts      self     2 + 2 self-collision rows                            (:216-222)
//     next 2 parts      planes   3 + 3 body points against the plane set (if any)     (obsAvoidConvex :57-89)
//
// A block owns a tile of 32 items (lane = item, warp = part), so every warp runs one role without
// divergence and the fields of 32 neighbouring instances still load as coalesced lines.  Compared
// with one fat thread per item (247 registers, ~270 loads, 10 k instructions in one dependency chain)
// a part has ~100 registers, ~30 loads that are all issued up front, and ~1 k instructions: 8x more
// warps in flight, each 5-8x shorter.  The parts' contributions to the stage QP / KKT / merit partials
// are combined in shared memory in a fixed part order (deterministic), the packed stage-QP record is
// assembled in shared memory and written to HBM as contiguous 640-byte records.
//
// The arithmetic of every row is the same as in Inst::trial_eval / Inst::step (mmpc_staged.cuh); only
// the order in which the contributions are summed differs.
#pragma once
#include "mmpc_staged.cuh"

namespace mmpc {

constexpr int CIRC_PER_PART = 8;
// accumulated fields of one item (shared memory, stride ACC_STRIDE doubles per item)
constexpr int F_H = 0, F_A = 21, F_GA = 27, F_GB = 33, F_ST = 39, F_CSUM = 45, F_BE0 = 46, F_BE1 = 47, F_ZROWS = 48,
              F_THETA = 49, F_FSUM = 50, F_LOG = 51, F_OK = 52, F_PRIM = 53, F_CHI = 54, F_CLO = 55, F_SUMZ = 56, F_NZ = 57,
              F_ES = 58, F_SUMLAM = 59, F_NEQ = 60, F_S = 61, F_AP = 62, F_AD = 63, F_GPHI = 64, NF = 65, ACC_STRIDE = 67;
constexpr int QREC_STRIDE = QS + 1;

struct PartPlan {
  int n_parts, circ_parts, self0, plane0;  // first part index of the self / plane groups (plane0 = -1: none)
};
__host__ __device__ inline PartPlan part_plan(const MmpcConfig& c) {
  PartPlan p;
  p.circ_parts = (c.n_obs + CIRC_PER_PART - 1) / CIRC_PER_PART;
  p.self0 = 2 + p.circ_parts;
  p.plane0 = c.n_pl > 0 ? p.self0 + 2 : -1;
  p.n_parts = p.self0 + 2 + (c.n_pl > 0 ? 2 : 0);
  return p;
}

// how a part's value enters the accumulated field: the first part stores, later parts combine
__device__ __forceinline__ void acc_sum(double* a, int f, double v, bool first) { a[f] = first ? v : a[f] + v; }
__device__ __forceinline__ void acc_max(double* a, int f, double v, bool first) { a[f] = first ? v : fmax(a[f], v); }
__device__ __forceinline__ void acc_min(double* a, int f, double v, bool first) { a[f] = first ? v : fmin(a[f], v); }

struct Parts {
  Inst S;
  int k, it, jt;
  double os, mu, alpha, ad, tau;
  bool trial;  // true: trial + evaluation of the candidate; false: step of the current iterate

  __device__ __forceinline__ Parts(const SParams& p, int b, int k_, bool trial_) : S(p, b), k(k_), trial(trial_) {
    S.load_npl();
    it = S.J(J_CUR) * S.ITSZ; jt = (1 - S.J(J_CUR)) * S.ITSZ;
    os = S.D(D_OS); mu = S.D(D_MU);
    alpha = trial ? S.D(D_ALPHA) : 0.0; ad = trial ? S.D(D_AD) : 0.0;
    tau = fmax(0.99, 1 - mu);
  }

  // scalars every part produces
  struct Scal {
    double theta, fsum, ok, prim, chi, clo, sumz, nz, es, gphi;
    MinRatio rp, rd;  // fraction to the boundary: primal, dual
    LogProd lp;
    __device__ __forceinline__ void init() {
      theta = 0; fsum = 0; ok = 1.0; prim = 0; chi = -1e300; clo = 1e300; sumz = 0; nz = 0; es = 0; gphi = 0;
      rp.init(); rd.init(); lp.init();
    }
  };
  __device__ __forceinline__ void commit_scal(const Scal& c, double* a, bool first) const {
    acc_sum(a, F_THETA, c.theta, first); acc_sum(a, F_FSUM, c.fsum, first); acc_sum(a, F_LOG, c.lp.value(), first);
    acc_min(a, F_OK, c.ok, first); acc_max(a, F_PRIM, c.prim, first); acc_max(a, F_CHI, c.chi, first);
    acc_min(a, F_CLO, c.clo, first); acc_sum(a, F_SUMZ, c.sumz, first); acc_sum(a, F_NZ, c.nz, first);
    acc_max(a, F_ES, c.es, first); if (!trial) { acc_min(a, F_AP, c.rp.value(tau), first); acc_min(a, F_AD, c.rd.value(tau), first); }
    acc_sum(a, F_GPHI, c.gphi, first);
  }

  // candidate pose (x y psi q1 q2 q3), slack and their steps, as the row parts need them
  struct Pose { double x[NP], dp[NP], s, ds; };
  __device__ __forceinline__ void load_pose(Pose& p, int npose) const {
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      if (a < npose) {
        double xo = S.W(k, it + I_X + POSE2X[a]), d = S.W2(k, S_DX + POSE2X[a]);
        p.dp[a] = d; p.x[a] = trial ? fma(alpha, d, xo) : xo;
      } else { p.x[a] = 0; p.dp[a] = 0; }
    }
    double so = S.W(k, it + I_S); p.ds = S.W2(k, S_DS);
    p.s = trial ? fma(alpha, p.ds, so) : so;
  }

  // bookkeeping of one slack row  h - s + t = 0:
  //   step : dt = -res - (grad h . dx - ds), dz, fraction to the boundary, merit ingredients of the iterate
  //   trial: candidate (t, z) with slack reset and multiplier safeguard, merit and KKT ingredients
  struct RowIO { double t, z, dtv; };
  __device__ __forceinline__ void row_load(int r, RowIO& io) const {
    io.t = S.W(k, it + I_T + r); io.z = S.W(k, it + I_T + S.R + r);
    io.dtv = trial ? S.W2(k, S_DT + r) : 0.0;
  }
  __device__ __forceinline__ void row_step(int r, const RowIO& io, double h, double gd_, const Pose& p, Scal& c) const {
    double res = h - p.s + io.t;
    double dtv = -res - (gd_ - p.ds);
    S.W2(k, S_DT + r) = dtv;
    double itv = rcp(io.t), dz = (mu - io.z * (io.t + dtv)) * itv;
    c.theta += fabs(res); c.gphi -= mu * dtv * itv; c.lp.mul(io.t);
    if (dtv < 0) c.rp.add(io.t, -dtv);
    if (dz < 0) c.rd.add(io.z, -dz);
  }
  struct RowOut { double z, it_, res; };
  __device__ __forceinline__ RowOut row_trial(int r, const RowIO& io, double h, const Pose& p, Scal& c,
                                              double& csum, double& be0, double& be1, double& zrows) const {
    RowOut o;
    double tt = fmax(fma(alpha, io.dtv, io.t), p.s - h);  // slack reset (Nocedal & Wright 19.30)
    double dz = (mu - io.z * (io.t + io.dtv)) * rcp(io.t);
    o.it_ = rcp(tt);
    double z = zclamp(io.z + ad * dz, mu, o.it_);
    S.W(k, jt + I_T + r) = tt; S.W(k, jt + I_T + S.R + r) = z;
    o.res = h - p.s + tt; o.z = z;
    c.theta += fabs(o.res);
    if (tt <= 0) c.ok = 0.0; else c.lp.mul(tt);
    c.prim = fmax(c.prim, fabs(o.res));
    double zt = z * tt;
    c.chi = fmax(c.chi, zt); c.clo = fmin(c.clo, zt); c.sumz += z; c.nz += 1.0; zrows += z;
    double sig = z * o.it_;
    csum += sig; be0 += sig * o.res; be1 += o.it_;
    return o;
  }

  // ---- circles ------------------------------------------------------------------------------------------
  struct CircAcc { double h00, h01, h11, a0, a1, gA0, gA1, gB0, gB1, st0, st1, csum, be0, be1, zrows; Scal c; };
  __device__ void part_circles(int i0, int i1, CircAcc& A) const {
    A.h00 = A.h01 = A.h11 = A.a0 = A.a1 = A.gA0 = A.gA1 = A.gB0 = A.gB1 = A.st0 = A.st1 = A.csum = A.be0 = A.be1 = A.zrows = 0;
    A.c.init();
    Pose p; load_pose(p, 2);
    RowIO io[CIRC_PER_PART]; double cx[CIRC_PER_PART], cy[CIRC_PER_PART], cr[CIRC_PER_PART];
#pragma unroll
    for (int q = 0; q < CIRC_PER_PART; ++q) {
      int i = i0 + q;
      if (i < i1) { row_load(i, io[q]); cx[q] = S.circ(k, i, 0); cy[q] = S.circ(k, i, 1); cr[q] = S.circ(k, i, 2); }
      else { io[q].t = 1; io[q].z = 0; io[q].dtv = 0; cx[q] = cy[q] = cr[q] = 0; }
    }
#pragma unroll
    for (int q = 0; q < CIRC_PER_PART; ++q) {
      int i = i0 + q;
      if (i < i1) {
        double ddx = p.x[0] - cx[q], ddy = p.x[1] - cy[q];
        double d2 = ddx * ddx + ddy * ddy, inv = rsq(d2), d = d2 * inv;
        double h = (cr[q] + S.cfg.base_radius) - d;
        if (!trial) { row_step(i, io[q], h, -(ddx * p.dp[0] + ddy * p.dp[1]) * inv, p, A.c); continue; }
        RowOut o = row_trial(i, io[q], h, p, A.c, A.csum, A.be0, A.be1, A.zrows);
        double sig = o.z * o.it_, nx = ddx * inv, ny = ddy * inv, zd = o.z * inv;
        A.h00 += sig * nx * nx - zd * (1 - nx * nx);
        A.h01 += (sig + zd) * nx * ny;
        A.h11 += sig * ny * ny - zd * (1 - ny * ny);
        double cb = sig * o.res;
        A.a0 += sig * nx; A.a1 += sig * ny; A.gA0 -= cb * nx; A.gA1 -= cb * ny;
        A.gB0 -= o.it_ * nx; A.gB1 -= o.it_ * ny; A.st0 -= o.z * nx; A.st1 -= o.z * ny;
      }
    }
  }
  __device__ __forceinline__ void commit_circles(const CircAcc& A, double* a, bool first) const {
    acc_sum(a, F_H + pidx(0, 0), A.h00, first); acc_sum(a, F_H + pidx(0, 1), A.h01, first); acc_sum(a, F_H + pidx(1, 1), A.h11, first);
    acc_sum(a, F_A + 0, A.a0, first); acc_sum(a, F_A + 1, A.a1, first); acc_sum(a, F_GA + 0, A.gA0, first); acc_sum(a, F_GA + 1, A.gA1, first);
    acc_sum(a, F_GB + 0, A.gB0, first); acc_sum(a, F_GB + 1, A.gB1, first); acc_sum(a, F_ST + 0, A.st0, first); acc_sum(a, F_ST + 1, A.st1, first);
    acc_sum(a, F_CSUM, A.csum, first); acc_sum(a, F_BE0, A.be0, first); acc_sum(a, F_BE1, A.be1, first); acc_sum(a, F_ZROWS, A.zrows, first);
    commit_scal(A.c, a, first);
  }

  // ---- self-collision and plane rows ----------------------------------------------------------------------
  struct RowAcc2 { double H[21], a[NP], gA[NP], gB[NP], st[NP], csum, be0, be1, zrows; Scal c; };
  __device__ __forceinline__ static void clear(RowAcc2& A) {
#pragma unroll
    for (int e = 0; e < 21; ++e) A.H[e] = 0;
#pragma unroll
    for (int q = 0; q < NP; ++q) A.a[q] = A.gA[q] = A.gB[q] = A.st[q] = 0;
    A.csum = A.be0 = A.be1 = A.zrows = 0; A.c.init();
  }
  __device__ __forceinline__ FK pose_fk(const Pose& p) const {
    FK f;
    if (trial) fk_eval(p.x[2], p.x[3], p.x[4], p.x[5], f);
    else {  // the FK cache of the current iterate
      f.cp = S.W2(k, S_FK + 0); f.sp = S.W2(k, S_FK + 1);
#pragma unroll
      for (int q = 0; q < 3; ++q) { f.vr[q] = S.W2(k, S_FK + 2 + q); f.vh[q] = S.W2(k, S_FK + 5 + q); }
    }
    return f;
  }
  __device__ void part_self(int m0, int m1, RowAcc2& A) const {
    clear(A);
    Pose p; load_pose(p, NP);
    RowIO io[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) row_load(S.nobs + m0 + q, io[q]);
    FK f = pose_fk(p);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int m = m0 + q;
      if (m >= m1) continue;
      const RowIO& r = io[q];
      Point pt; point_eval(p.x[0], p.x[1], f, SELFD[m], pt);
      double d2 = pt.P[0] * pt.P[0] + pt.P[1] * pt.P[1] + pt.P[2] * pt.P[2], inv = rsq(d2);
      double h = S.cfg.self_collision_radius - d2 * inv;
      double n[3] = {pt.P[0] * inv, pt.P[1] * inv, pt.P[2] * inv}, g[NP];
      point_grad(f, pt, n, g);  // grad h = -g
      if (!trial) {
        double gd_ = 0;
#pragma unroll
        for (int a = 0; a < NP; ++a) gd_ = fma(g[a], p.dp[a], gd_);
        row_step(S.nobs + m, r, h, -gd_, p, A.c);
        continue;
      }
      RowOut o = row_trial(S.nobs + m, r, h, p, A.c, A.csum, A.be0, A.be1, A.zrows);
      double sig = o.z * o.it_, zd = o.z * inv, cgg = sig + zd;
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += cgg * g[a] * g[c];
      point_jtj_acc(f, pt, -zd, A.H);
      point_hess_acc(f, pt, n, -o.z, A.H);
      double cb = sig * o.res;
#pragma unroll
      for (int a = 0; a < NP; ++a) { A.a[a] += sig * g[a]; A.gA[a] -= cb * g[a]; A.gB[a] -= o.it_ * g[a]; A.st[a] -= o.z * g[a]; }
    }
  }
  __device__ void part_planes(int i0, int i1, RowAcc2& A) const {
    clear(A);
    Pose p; load_pose(p, NP);
    RowIO io[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) row_load(S.nobs + 4 + i0 + q, io[q]);
    FK f = pose_fk(p);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = i0 + q;
      if (i >= i1) continue;
      const RowIO& r = io[q];
      Point pt; point_eval(p.x[0], p.x[1], f, BODY[i], pt);
      int jb; double h = S.plane_row(pt, jb);
      double n[3] = {S.D(D_PL + 6 * jb + 3), S.D(D_PL + 6 * jb + 4), S.D(D_PL + 6 * jb + 5)}, g[NP];
      point_grad(f, pt, n, g);  // the row is  -max c <= s, c = off - n.P  =>  grad h = +g
      if (!trial) {
        double gd_ = 0;
#pragma unroll
        for (int a = 0; a < NP; ++a) gd_ = fma(g[a], p.dp[a], gd_);
        row_step(S.nobs + 4 + i, r, h, gd_, p, A.c);
        continue;
      }
      RowOut o = row_trial(S.nobs + 4 + i, r, h, p, A.c, A.csum, A.be0, A.be1, A.zrows);
      double sig = o.z * o.it_;
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += sig * g[a] * g[c];
      point_hess_acc(f, pt, n, o.z, A.H);
      double cb = sig * o.res;
#pragma unroll
      for (int a = 0; a < NP; ++a) { A.a[a] -= sig * g[a]; A.gA[a] += cb * g[a]; A.gB[a] += o.it_ * g[a]; A.st[a] += o.z * g[a]; }
    }
  }
  __device__ __forceinline__ void commit_rows(const RowAcc2& A, double* a, bool first) const {
#pragma unroll
    for (int e = 0; e < 21; ++e) acc_sum(a, F_H + e, A.H[e], first);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      acc_sum(a, F_A + q, A.a[q], first); acc_sum(a, F_GA + q, A.gA[q], first);
      acc_sum(a, F_GB + q, A.gB[q], first); acc_sum(a, F_ST + q, A.st[q], first);
    }
    acc_sum(a, F_CSUM, A.csum, first); acc_sum(a, F_BE0, A.be0, first); acc_sum(a, F_BE1, A.be1, first); acc_sum(a, F_ZROWS, A.zrows, first);
    commit_scal(A.c, a, first);
  }

  // candidate multiplier of a bound (trial) / its step and the fraction to the boundary (step)
  __device__ __forceinline__ double box(double zold, double d_old, double d, double dv_in, int slot, Scal& c, double& id) const {
    // dv_in: step of the variable towards the interior: +dv for a lower bound, -dv for an upper bound
    const double ido = rcp(d_old);
    const double dz = mu * ido - zold - zold * ido * dv_in;
    if (!trial) {
      id = ido;
      c.gphi -= mu * dv_in * ido; c.lp.mul(d_old);
      if (dv_in < 0) c.rp.add(d_old, -dv_in);
      if (dz < 0) c.rd.add(zold, -dz);
      return zold;
    }
    id = rcp(d);
    double z = zclamp(zold + ad * dz, mu, id);
    S.W(k, slot) = z;
    if (d <= 0) c.ok = 0.0; else c.lp.mul(d);
    c.chi = fmax(c.chi, z * d); c.clo = fmin(c.clo, z * d); c.sumz += z; c.nz += 1.0;
    return z;
  }

  // ---- base-x: states ----------------------------------------------------------------------------------------
  struct BaseX { double Hd[NP], gA[NP], gB[NP], st[NP], sum_lam, n_eq, s; Scal c; };
  __device__ void part_base_x(BaseX& B, double* qrec) const {
    const MmpcConfig& cfg = S.cfg; const int N = S.N; const double dt = S.dt;
    B.c.init(); B.sum_lam = 0; B.n_eq = 0;
    double x[NX], xo[NX], dxo[NX], lam[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      xo[i] = S.W(k, it + I_X + i); dxo[i] = S.W2(k, S_DX + i);
      x[i] = trial ? fma(alpha, dxo[i], xo[i]) : xo[i];
      lam[i] = 0;
      if (trial) {
        S.W(k, jt + I_X + i) = x[i];
        if (k >= 1) {
          double l = S.W(k, it + I_LAM + i);
          lam[i] = l + alpha * (S.W2(k, S_LAMN + i) - l);
          S.W(k, jt + I_LAM + i) = lam[i];
        }
      }
    }
    const double so = S.W(k, it + I_S), dsv = S.W2(k, S_DS);
    const double s = trial ? fma(alpha, dsv, so) : so;
    B.s = s;
    if (trial) S.W(k, jt + I_S) = s;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double uo = (k < N) ? S.W(k, it + I_U + j) : 0.0, du = (k < N) ? S.W2(k, S_DU + j) : 0.0;
      u[j] = trial ? fma(alpha, du, uo) : uo;
    }
    double hpp = 0, stx[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) stx[i] = 0;
    if (trial) {
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
      S.W2(k, S_FK + 0) = f.cp; S.W2(k, S_FK + 1) = f.sp;
#pragma unroll
      for (int q = 0; q < 3; ++q) { S.W2(k, S_FK + 2 + q) = f.vr[q]; S.W2(k, S_FK + 5 + q) = f.vh[q]; }
      if (k < N) {  // dynamics :180 at the candidate -- defect and costate terms A^T lam_{k+1}
        double xn[NX], lam1[NX];
        dyn_f(x, u, dt, f.cp, f.sp, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double l1 = S.W(k + 1, it + I_LAM + i);
          lam1[i] = l1 + alpha * (S.W2(k + 1, S_LAMN + i) - l1);
          double x1 = fma(alpha, S.W2(k + 1, S_DX + i), S.W(k + 1, it + I_X + i));
          double d = xn[i] - x1;
          S.W2(k, S_DFC + i) = d; B.c.prim = fmax(B.c.prim, fabs(d)); B.sum_lam += fabs(lam1[i]);
          B.c.theta += fabs(d);
        }
        B.n_eq = NX;
        hpp = -dt * u[0] * (lam1[3] * f.cp + lam1[4] * f.sp);
        qrec[Q_HPU] = dt * (-lam1[3] * f.sp + lam1[4] * f.cp);
        qrec[Q_H45] = -dt * lam1[3];  // (dy,dpsi)
        qrec[Q_H35] = dt * lam1[4];   // (dx,dpsi)
        stx[0] = lam1[0]; stx[1] = lam1[1];
        stx[2] = lam1[2] + dt * u[0] * (-f.sp * lam1[3] + f.cp * lam1[4]);
        stx[3] = dt * lam1[0] + lam1[3] + dt * x[5] * lam1[4];
        stx[4] = dt * lam1[1] - dt * x[5] * lam1[3] + lam1[4];
        stx[5] = dt * lam1[2] - dt * x[4] * lam1[3] + dt * x[3] * lam1[4] + lam1[5];
        stx[6] = lam1[6]; stx[7] = lam1[7]; stx[8] = lam1[8];
      } else {
        qrec[Q_HPU] = 0; qrec[Q_H45] = 0; qrec[Q_H35] = 0;
      }
    } else if (k < N) {
#pragma unroll
      for (int i = 0; i < NX; ++i) B.c.theta += fabs(S.W2(k, S_DFC + i));
    }
    const bool teq = S.term_eq(k);
    double nu[2] = {0, 0};
    if (teq && trial) {  // candidate multipliers of the terminal equality
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double no = S.W(0, it + I_LAM + i);
        nu[i] = no + alpha * (S.W2(0, S_LAMN + i) - no);
        S.W(0, jt + I_LAM + i) = nu[i];
      }
    }
    // cost and boxes -- :192-205, :240-245
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double Wx = os * (k < N ? cfg.Qd[i] : cfg.Pd[i]);
      double e = x[i] - S.W2(k, IN_XREF + i);
      B.c.fsum += Wx * e * e;
      double gr = 2 * Wx * e;
      if (!trial) B.c.gphi += gr * dxo[i];
      double Hd = 2 * Wx, gA = gr, gB = 0, st = gr + stx[i] - lam[i];
      if (k >= 1) {
        double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
        if (is_fin(lo)) {
          double d = x[i] - lo, id, z = box(S.W(k, it + I_ZXL + i), xo[i] - lo, d, dxo[i], jt + I_ZXL + i, B.c, id);
          Hd += z * id; gB -= id; st -= z;
        }
        if (is_fin(hi)) {
          double d = hi - x[i], id, z = box(S.W(k, it + I_ZXU + i), hi - xo[i], d, -dxo[i], jt + I_ZXU + i, B.c, id);
          Hd += z * id; gB += id; st += z;
        }
      }
      if (i < 2 && teq) {  // terminal equality row  x_N[i] - xref_N[i] = 0  with multiplier nu[i]
        Hd += 1.0 / MMPC_DELTA_C; gA += nu[i] + e / MMPC_DELTA_C; st += nu[i];
        B.c.theta += fabs(e);
        if (trial) { B.c.prim = fmax(B.c.prim, fabs(e)); B.sum_lam += fabs(nu[i]); B.n_eq += 1; }
      }
      if (i < 3 || i >= 6) {
        const int a = (i < 3) ? i : i - 3;
        B.Hd[a] = Hd + (i == 2 ? hpp : 0.0); B.gA[a] = gA; B.gB[a] = gB; B.st[a] = st;
      } else if (trial) {
        qrec[Q_HVD + (i - 3)] = Hd; qrec[Q_GA + i] = gA; qrec[Q_GB + i] = gB;
        if (k >= 1) B.c.es = fmax(B.c.es, fabs(st));
      }
    }
    B.c.fsum += os * cfg.S * s * s;
    if (!trial) B.c.gphi += 2 * os * cfg.S * s * dsv;
  }
  __device__ __forceinline__ void commit_base_x(const BaseX& B, double* a, bool first) const {
#pragma unroll
    for (int e = 0; e < 21; ++e) acc_sum(a, F_H + e, 0.0, first);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      a[F_H + pidx(q, q)] = first ? B.Hd[q] : a[F_H + pidx(q, q)] + B.Hd[q];
      acc_sum(a, F_A + q, 0.0, first); acc_sum(a, F_GA + q, B.gA[q], first);
      acc_sum(a, F_GB + q, B.gB[q], first); acc_sum(a, F_ST + q, B.st[q], first);
    }
    acc_sum(a, F_CSUM, 0.0, first); acc_sum(a, F_BE0, 0.0, first); acc_sum(a, F_BE1, 0.0, first); acc_sum(a, F_ZROWS, 0.0, first);
    acc_sum(a, F_SUMLAM, B.sum_lam, first); acc_sum(a, F_NEQ, B.n_eq, first); acc_sum(a, F_S, B.s, first);
    commit_scal(B.c, a, first);
  }

  // ---- base-u: controls -------------------------------------------------------------------------------------
  __device__ void part_base_u(Scal& c, double* qrec) const {
    const MmpcConfig& cfg = S.cfg; const int N = S.N; const double dt = S.dt;
    c.init();
    if (k >= N) {
      if (trial) {
#pragma unroll
        for (int j = 0; j < NU; ++j) { qrec[Q_HUU + j] = 0; qrec[Q_GA + SGY_U + j] = 0; qrec[Q_GB + SGY_U + j] = 0; }
      }
      return;
    }
    double stu[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) stu[j] = 0;
    if (trial) {  // B^T lam_{k+1} at the candidate
      double psi = fma(alpha, S.W2(k, S_DX + 2), S.W(k, it + I_X + 2)), sp, cp;
      sincos(psi, &sp, &cp);
      double lam1[NX];
#pragma unroll
      for (int i = 3; i < NX; ++i) {
        double l1 = S.W(k + 1, it + I_LAM + i);
        lam1[i] = l1 + alpha * (S.W2(k + 1, S_LAMN + i) - l1);
      }
      stu[0] = dt * (cp * lam1[3] + sp * lam1[4]);
      stu[1] = dt * lam1[5];
      stu[2] = dt * lam1[6]; stu[3] = dt * lam1[7]; stu[4] = dt * lam1[8];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double uo = S.W(k, it + I_U + j), du = S.W2(k, S_DU + j);
      double u = trial ? fma(alpha, du, uo) : uo;
      if (trial) S.W(k, jt + I_U + j) = u;
      double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
      double e = u - S.W2(k, IN_UREF + j), dl = u - S.W2(k, IN_ULAST + j);
      c.fsum += os * (cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl);
      double gr = 2 * Rj * e + 2 * Wj * dl;
      if (!trial) c.gphi += gr * du;
      double Hd = 2 * Rj + 2 * Wj, gA = gr, gB = 0, st = gr + stu[j];
      double lo = S.W2(k, IN_ULO + j), hi = S.W2(k, IN_UHI + j);
      if (is_fin(lo)) {
        double d = u - lo, id, z = box(S.W(k, it + I_ZUL + j), uo - lo, d, du, jt + I_ZUL + j, c, id);
        Hd += z * id; gB -= id; st -= z;
      }
      if (is_fin(hi)) {
        double d = hi - u, id, z = box(S.W(k, it + I_ZUU + j), hi - uo, d, -du, jt + I_ZUU + j, c, id);
        Hd += z * id; gB += id; st += z;
      }
      if (trial) {
        c.es = fmax(c.es, fabs(st));
        qrec[Q_HUU + j] = Hd; qrec[Q_GA + SGY_U + j] = gA; qrec[Q_GB + SGY_U + j] = gB;
      }
    }
  }

  // ---- after all parts are combined: the slack column, the stationarity norm, the partial slots ---------
  __device__ void finalize(const double* a, double* qrec) const {
    if (!trial) {
      S.W2(k, S_PART + 0) = a[F_AP]; S.W2(k, S_PART + 1) = a[F_AD]; S.W2(k, S_PART + 2) = a[F_GPHI];
      S.W2(k, S_PART + 3) = a[F_THETA]; S.W2(k, S_PART + 4) = a[F_FSUM]; S.W2(k, S_PART + 5) = a[F_LOG];
      return;
    }
    const double s = a[F_S];
    double es = a[F_ES];
    double S2 = 2 * os * S.cfg.S;
    qrec[Q_C] = S2 + a[F_CSUM];
    qrec[Q_GA + SGY_S] = S2 * s - a[F_BE0];
    qrec[Q_GB + SGY_S] = -a[F_BE1];
    qrec[Q_HVV] = 0; qrec[Q_GA + SGY_V] = 0; qrec[Q_GB + SGY_V] = 0;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      qrec[Q_A + q] = a[F_A + q]; qrec[Q_BV + q] = 0;
      qrec[Q_GA + POSE2X[q]] = a[F_GA + q]; qrec[Q_GB + POSE2X[q]] = a[F_GB + q];
      if (k >= 1) es = fmax(es, fabs(a[F_ST + q]));
    }
#pragma unroll
    for (int e = 0; e < 21; ++e) qrec[Q_HP + e] = a[F_H + e];
    es = fmax(es, fabs(S2 * s - a[F_ZROWS]));
    S.W2(k, S_PART + 0) = es; S.W2(k, S_PART + 1) = a[F_PRIM]; S.W2(k, S_PART + 2) = a[F_CHI]; S.W2(k, S_PART + 3) = a[F_CLO];
    S.W2(k, S_PART + 4) = a[F_SUMLAM]; S.W2(k, S_PART + 5) = a[F_SUMZ]; S.W2(k, S_PART + 6) = a[F_NZ]; S.W2(k, S_PART + 7) = a[F_NEQ];
    const double theta = a[F_THETA], fsum = a[F_FSUM];
    bool fin = (a[F_OK] > 0.5) && (fsum == fsum) && (theta == theta);
    S.W2(k, S_PART + PT_MERIT + 0) = theta; S.W2(k, S_PART + PT_MERIT + 1) = fsum;
    S.W2(k, S_PART + PT_MERIT + 2) = fin ? a[F_LOG] : 0.0; S.W2(k, S_PART + PT_MERIT + 3) = fin ? 1.0 : 0.0;
  }
};

// One part of one item: compute, then (when it is this part's turn) combine into the item's accumulators.
// The GPU kernel calls run() from the part's warp and commit() inside the barrier-separated phases; the CPU
// emulation calls both back to back, part after part.
struct PartRunner {
  Parts::BaseX bx; Parts::Scal bu; Parts::CircAcc ca; Parts::RowAcc2 ra;
  int kind;  // 0 base-x, 1 base-u, 2 circles, 3 self / planes
  __device__ void run(const Parts& T, const PartPlan& pl, int part, double* qrec) {
    const int nobs = T.S.nobs;
    if (part == 0) { kind = 0; T.part_base_x(bx, qrec); }
    else if (part == 1) { kind = 1; T.part_base_u(bu, qrec); }
    else if (part < pl.self0) { kind = 2; int i0 = (part - 2) * CIRC_PER_PART; T.part_circles(i0, min(nobs, i0 + CIRC_PER_PART), ca); }
    else if (part < pl.self0 + 2) { kind = 3; int m0 = (part - pl.self0) * 2; T.part_self(m0, m0 + 2, ra); }
    else { kind = 3; int i0 = (part - pl.plane0) * 3; T.part_planes(i0, i0 + 3, ra); }
  }
  __device__ void commit(const Parts& T, double* acc, bool first) const {
    if (kind == 0) T.commit_base_x(bx, acc, first);
    else if (kind == 1) T.commit_scal(bu, acc, first);
    else if (kind == 2) T.commit_circles(ca, acc, first);
    else T.commit_rows(ra, acc, first);
  }
};

// CPU-emulation / reference composition of one item: all parts in order, then finalize and write-out
__device__ inline void body_parts_item(const SParams& P, int b, int k, bool trial) {
  double acc[NF], qrec[QS];
  for (int f = 0; f < QS; ++f) qrec[f] = 0;
  Parts T(P, b, k, trial);
  PartPlan pl = part_plan(P.cfg);
  for (int part = 0; part < pl.n_parts; ++part) {
    PartRunner R;
    R.run(T, pl, part, qrec);
    R.commit(T, acc, part == 0);
  }
  T.finalize(acc, qrec);
  if (trial)
    for (int f = 0; f < QS - 2; ++f) T.S.Qw(k, f) = qrec[f];
}

#if !defined(MMPC_EMULATE) && !defined(MMPC_EMULATE_LANE)
// block-wide barrier that warps reach from different code paths (one path per part kind)
__device__ __forceinline__ void block_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// grid-stride over tiles of 32 items; blockDim = 32 * n_parts.  Each warp keeps only its own kind of
// partial in registers; the partials are combined in part order, one barrier-separated phase per part.
template <bool TRIAL>
#ifndef MMPC_PARTS_MINB
#define MMPC_PARTS_MINB 1
#endif
__global__ void __launch_bounds__(256, MMPC_PARTS_MINB) staged_parts_kernel(const __grid_constant__ SParams P) {
  __shared__ double acc[32 * ACC_STRIDE];
  __shared__ double qrec[32 * QREC_STRIDE];
  const PartPlan pl = part_plan(P.cfg);
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5, nthr = blockDim.x;
  const int n = TRIAL ? P.cnt[P.tsel] : P.cnt[0];
  const int* list = TRIAL ? list_T(P) : list_E(P);
  const long long tot = (long long)n * (P.cfg.N + 1);
  double* accp = acc + lane * ACC_STRIDE;
  double* qr = qrec + lane * QREC_STRIDE;
#define MMPC_PHASES(COMMIT)                                  \
  for (int p = 0; p < pl.n_parts; ++p) {                     \
    if (p == part && valid) { const bool first = (p == 0); COMMIT; } \
    block_sync(nthr);                                        \
  }
  for (long long tile = blockIdx.x; tile * 32 < tot; tile += gridDim.x) {
    const long long t = tile * 32 + lane;
    bool valid = t < tot;
    int b = 0, k = 0;
    if (valid) {
      b = list[(int)(t % n)]; k = (int)(t / n);
      if (!TRIAL) {
        const int st = inst_state(P, b);
        if (st == ST_FINISH && part == 0) { Inst F(P, b); F.finish_stage(k); }  // left the solve in this round
        if (st != ST_ACTIVE) valid = false;
      }
    }
    Parts T(P, b, k, TRIAL);
    if (part == 0) {
      Parts::BaseX bx;
      if (valid) T.part_base_x(bx, qr);
      MMPC_PHASES(T.commit_base_x(bx, accp, first))
    } else if (part == 1) {
      Parts::Scal bu;
      if (valid) T.part_base_u(bu, qr);
      MMPC_PHASES(T.commit_scal(bu, accp, first))
    } else if (part < pl.self0) {
      Parts::CircAcc ca;
      const int i0 = (part - 2) * CIRC_PER_PART;
      if (valid) T.part_circles(i0, min(T.S.nobs, i0 + CIRC_PER_PART), ca);
      MMPC_PHASES(T.commit_circles(ca, accp, first))
    } else {
      Parts::RowAcc2 ra;
      if (valid) {
        if (part < pl.self0 + 2) { const int m0 = (part - pl.self0) * 2; T.part_self(m0, m0 + 2, ra); }
        else { const int i0 = (part - pl.plane0) * 3; T.part_planes(i0, i0 + 3, ra); }
      }
      MMPC_PHASES(T.commit_rows(ra, accp, first))
    }
#undef MMPC_PHASES
    if (part == 0 && valid) T.finalize(accp, qr);
    block_sync(nthr);
    if (TRIAL) {
      // the 32 stage-QP records of the tile, each 78 contiguous doubles, written by whole warps
      for (int i = part; i < 32; i += pl.n_parts) {
        const long long ti = tile * 32 + i;
        if (ti < tot) {
          const int bi = list[(int)(ti % n)], ki = (int)(ti / n);
          double* dst = P.qp + ((long long)ki * P.LS + bi) * QS;
          for (int f = lane; f < QS - 2; f += 32) dst[f] = qrec[i * QREC_STRIDE + f];
        }
      }
      block_sync(nthr);
    }
  }
}
#endif

}  // namespace mmpc
