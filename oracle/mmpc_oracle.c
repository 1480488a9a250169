/*
 * mmpc_oracle.c -- CPU oracle: plain-C restatement of the whole-body MPC NLP and an
 * IPOPT-style primal-dual interior-point solve of it.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * leg may load this; the product (mobile_manipulator_mpc_b200/) never does.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in the casadi==3.6.4 wheel (IPOPT + MUMPS,
 * requirements.txt:9) which cannot be installed in this image, and the reference has no tests
 * or golden vectors.  What is restated here, with citations relative to /root/reference:
 *   NLP        controllers/mpc_wholebody_qref.py:142-285 (reset) and :287-331 (solve)
 *   dynamics   robot_models/base.py:19-26, robot_models/manipulator_3DoF.py:189-191
 *   FK         robot_models/manipulator_3DoF.py:18-77, robot_models/mobile_manipulator.py:28-55
 *   solver     IPOPT's published algorithm (Waechter & Biegler 2006): slack reformulation of
 *              inequalities, monotone barrier update (mu_init 0.1, kappa_mu 0.2, theta_mu 1.5,
 *              kappa_eps 10), fraction-to-boundary tau = max(0.99, 1-mu), bound_push 0.01,
 *              inertia correction by delta_w, scaled error E_mu with s_max = 100, tol 1e-8.
 *              The globalisation is an l1-merit Armijo backtracking (IPOPT uses a filter), so
 *              iterates differ from IPOPT's; the KKT point reached from the same start is what
 *              is compared.
 * The oracle is pinned by oracle/model.py (expression-level restatement, itself pinned by the
 * sympy DH derivation utils/dh_to_kinematics.py), by finite differences of every derivative
 * here, and by independent SciPy solves of the restated NLP (tests/golden/).
 *
 * The KKT system is solved by a dense Riccati recursion over the augmented stage vector
 *   y_k = (x_k[9], s_k, u_k[5], v_k)   with  s_{k+1} = v_k,
 * which makes the reference's cross-stage rows ordinary
 * stage rows: a stale-column plane row on (x_{k-1}, s_k) is a mixed row on (x_{k-1}, v_{k-1}),
 * and the terminal self-collision row on (x_N, s_{N-1}) is a row on (x_{N-1}, u_{N-1}, s_{N-1})
 * because the pose part of f_kinematics is linear in (x, u).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include "../include/mmpc.h"
#include <stdio.h>
int mmpc_oracle_verbose = 0;

#define NX 9
#define NU 5
#define NXA 10 /* x, s          */
#define NUA 6  /* u, v          */
#define NY 16  /* x, s, u, v    */
#define IS 9   /* index of s in y */
#define IU 10  /* index of u0 in y */
#define IV 15  /* index of v in y */
#define NP 6   /* pose variables x y psi q1 q2 q3 */
#define DELTA_C 1e-6 /* IPOPT's delta_c: regularisation of the terminal-equality block of the KKT matrix */

static const int POSE2X[NP] = {0, 1, 2, 6, 7, 8};

/* manipulator_3DoF.py:18-22, mobile_manipulator.py:14-15 */
static const double A2 = 0.316, A3 = 0.0825, A5 = 0.384, A6 = 0.088, A7 = 0.107;
static const double BX = -0.007, BZ = 0.606 + 0.333;

/* body points pos_i = P(kappa; w) (mpc_wholebody_qref.py:216-217) and self-collision
 * differences chk_m - endpoint (:219-221) in the (kappa, w1, w2, w3) parametrisation:
 *   P = (kappa*x + R cos(psi), kappa*y + R sin(psi), Z),
 *   R = sum_s w_s v_s.r + kappa*BX,  Z = sum_s w_s v_s.h + kappa*BZ                        */
static const double BODY[6][4] = {{0.5, 0.5, 0, 0}, {1, 1, 0, 0}, {1, 1, 0.5, 0},
                                  {1, 1, 1, 0},     {1, 1, 1, 0.5}, {1, 1, 1, 1}};
static const double SELFD[4][4] = {{-1, -1, -1, -1}, {-0.5, -0.5, -1, -1}, {0, 0, -1, -1}, {0, 0, -0.5, -1}};

typedef struct {
  double cp, sp;           /* cos/sin psi */
  double vr[3], vh[3];     /* arm segments in the arm plane (r, h) */
} FK;

static void fk_eval(const double* x, FK* f) {
  double t1 = x[6], t2 = x[6] - x[7], t3 = x[6] - x[7] - x[8];
  f->cp = cos(x[2]); f->sp = sin(x[2]);
  double s1 = sin(t1), c1 = cos(t1), s2 = sin(t2), c2 = cos(t2), s3 = sin(t3), c3 = cos(t3);
  f->vr[0] = A2 * s1 + A3 * c1;  f->vh[0] = A2 * c1 - A3 * s1;
  f->vr[1] = -A3 * c2 + A5 * s2; f->vh[1] = A3 * s2 + A5 * c2;
  f->vr[2] = A6 * c3 - A7 * s3;  f->vh[2] = -A6 * s3 - A7 * c3;
}

typedef struct {
  double P[3];
  double J[3][NP];
  double R, Rth[3], Zth[3];
} Point;

static void point_eval(const double* x, const FK* f, const double kw[4], Point* p) {
  double kap = kw[0];
  double R = kap * BX, Z = kap * BZ;
  for (int s = 0; s < 3; ++s) {
    R += kw[1 + s] * f->vr[s]; Z += kw[1 + s] * f->vh[s];
    p->Rth[s] = kw[1 + s] * f->vh[s];   /* dR/dtheta_s */
    p->Zth[s] = -kw[1 + s] * f->vr[s];  /* dZ/dtheta_s */
  }
  p->R = R;
  p->P[0] = kap * x[0] + R * f->cp; p->P[1] = kap * x[1] + R * f->sp; p->P[2] = Z;
  double Rq[3] = {p->Rth[0] + p->Rth[1] + p->Rth[2], -p->Rth[1] - p->Rth[2], -p->Rth[2]};
  double Zq[3] = {p->Zth[0] + p->Zth[1] + p->Zth[2], -p->Zth[1] - p->Zth[2], -p->Zth[2]};
  memset(p->J, 0, sizeof p->J);
  p->J[0][0] = kap; p->J[1][1] = kap;
  p->J[0][2] = -R * f->sp; p->J[1][2] = R * f->cp;
  for (int j = 0; j < 3; ++j) {
    p->J[0][3 + j] = Rq[j] * f->cp; p->J[1][3 + j] = Rq[j] * f->sp; p->J[2][3 + j] = Zq[j];
  }
}

/* G = sum_c n_c * Hessian(P_c) (6x6 symmetric, full storage) */
static void point_hess(const FK* f, const Point* p, const double n[3], double G[NP][NP]) {
  double npar = n[0] * f->cp + n[1] * f->sp, nperp = -n[0] * f->sp + n[1] * f->cp;
  memset(G, 0, sizeof(double) * NP * NP);
  double Rq[3] = {p->Rth[0] + p->Rth[1] + p->Rth[2], -p->Rth[1] - p->Rth[2], -p->Rth[2]};
  G[2][2] = -p->R * npar;
  for (int j = 0; j < 3; ++j) G[2][3 + j] = G[3 + j][2] = Rq[j] * nperp;
  double gam[3];
  for (int s = 0; s < 3; ++s) gam[s] = p->Zth[s] * npar - p->Rth[s] * n[2];
  double g23 = gam[1] + gam[2], g123 = gam[0] + g23;
  G[3][3] = g123; G[3][4] = G[4][3] = -g23; G[3][5] = G[5][3] = -gam[2];
  G[4][4] = g23;  G[4][5] = G[5][4] = gam[2]; G[5][5] = gam[2];
}

/* ---- MMPC_MODEL_POSEREF (controllers/mpc_wholebody.py:79-86, :104-107): the tracking cost  e^T diag(W) e  on the end-point pose
 * e = forward_tranformation(x)[0] - X_ref[k] = (P_e - r_xyz, psi - r_psi); value, gradient and exact Hessian wrt the 6 pose
 * variables (x, y, psi, q1, q2, q3).  Wp: 4 weights; r: 4 reference values. ---- */
static double pose_cost(const double* x, const FK* f, const double* Wp, const double* r, double g[NP], double H[NP][NP]) {
  Point p; point_eval(x, f, BODY[5], &p);
  double e[3], n[3], val = 0;
  for (int c = 0; c < 3; ++c) { e[c] = p.P[c] - r[c]; n[c] = 2 * Wp[c] * e[c]; val += Wp[c] * e[c] * e[c]; }
  const double ep = x[2] - r[3];
  val += Wp[3] * ep * ep;
  if (g) {
    for (int a = 0; a < NP; ++a) { g[a] = 0; for (int c = 0; c < 3; ++c) g[a] += n[c] * p.J[c][a]; }
    g[2] += 2 * Wp[3] * ep;
  }
  if (H) {
    point_hess(f, &p, n, H);
    for (int a = 0; a < NP; ++a)
      for (int b = 0; b < NP; ++b)
        for (int c = 0; c < 3; ++c) H[a][b] += 2 * Wp[c] * p.J[c][a] * p.J[c][b];
    H[2][2] += 2 * Wp[3];
  }
  return val;
}

/* ---- row evaluation: value, gradient and Hessian wrt the 6 pose variables ---- */
static void row_circle(const double* x, const double* c, double base_r, double* h, double g[NP], double H[NP][NP]) {
  double dx = x[0] - c[0], dy = x[1] - c[1];
  double d = sqrt(dx * dx + dy * dy);
  *h = (c[2] + base_r) - d; /* mpc_wholebody_qref.py:53 */
  if (!g) return;
  double nx = dx / d, ny = dy / d;
  memset(g, 0, sizeof(double) * NP); memset(H, 0, sizeof(double) * NP * NP);
  g[0] = -nx; g[1] = -ny;
  H[0][0] = -(1 - nx * nx) / d; H[0][1] = H[1][0] = nx * ny / d; H[1][1] = -(1 - ny * ny) / d;
}

static void row_selfcoll(const double* x, const FK* f, int m, double rad, double* h, double g[NP], double H[NP][NP]) {
  Point p; point_eval(x, f, SELFD[m], &p);
  double d = sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]);
  *h = rad - d; /* :221-222 */
  if (!g) return;
  double n[3] = {p.P[0] / d, p.P[1] / d, p.P[2] / d};
  double G[NP][NP]; point_hess(f, &p, n, G);
  double Jn[NP];
  for (int a = 0; a < NP; ++a) Jn[a] = n[0] * p.J[0][a] + n[1] * p.J[1][a] + n[2] * p.J[2][a];
  for (int a = 0; a < NP; ++a) {
    g[a] = -Jn[a];
    for (int b = 0; b < NP; ++b) {
      double JJ = p.J[0][a] * p.J[0][b] + p.J[1][a] * p.J[1][b] + p.J[2][a] * p.J[2][b];
      H[a][b] = -((JJ - Jn[a] * Jn[b]) / d + G[a][b]);
    }
  }
}

/* plane margin c[i][j] = n_j . ((p_j - expand n_j) - pos_i)   (:78-80); value only */
static double plane_margin(const double* x, const FK* f, int i, const double* pl, double expand) {
  Point p; point_eval(x, f, BODY[i], &p);
  double c = 0;
  for (int a = 0; a < 3; ++a) c += pl[3 + a] * ((pl[a] - expand * pl[3 + a]) - p.P[a]);
  return c;
}
/* derivatives of h = -c[i][j] */
static void plane_derivs(const double* x, const FK* f, int i, const double* pl, double g[NP], double H[NP][NP]) {
  Point p; point_eval(x, f, BODY[i], &p);
  const double* n = pl + 3;
  for (int a = 0; a < NP; ++a) g[a] = n[0] * p.J[0][a] + n[1] * p.J[1][a] + n[2] * p.J[2][a];
  point_hess(f, &p, n, H);
}

/* ---- dynamics ---- */
static void dyn_f(const double* x, const double* u, double dt, double* xn) {
  double c = cos(x[2]), s = sin(x[2]);
  xn[0] = x[0] + dt * x[3]; xn[1] = x[1] + dt * x[4]; xn[2] = x[2] + dt * x[5];
  xn[3] = x[3] + dt * (u[0] * c - x[4] * x[5]);
  xn[4] = x[4] + dt * (u[0] * s + x[3] * x[5]);
  xn[5] = x[5] + dt * u[1];
  xn[6] = x[6] + u[2] * dt; xn[7] = x[7] + u[3] * dt; xn[8] = x[8] + u[4] * dt;
}
static void dyn_AB(const double* x, const double* u, double dt, double A[NX][NX], double B[NX][NU]) {
  double c = cos(x[2]), s = sin(x[2]);
  memset(A, 0, sizeof(double) * NX * NX); memset(B, 0, sizeof(double) * NX * NU);
  for (int i = 0; i < NX; ++i) A[i][i] = 1;
  A[0][3] = dt; A[1][4] = dt; A[2][5] = dt;
  A[3][2] = -dt * u[0] * s; A[3][4] = -dt * x[5]; A[3][5] = -dt * x[4];
  A[4][2] = dt * u[0] * c;  A[4][3] = dt * x[5];  A[4][5] = dt * x[3];
  B[3][0] = dt * c; B[4][0] = dt * s; B[5][1] = dt;
  B[6][2] = dt; B[7][3] = dt; B[8][4] = dt;
}

/* ---- problem / iterate ---- */
enum { ROW_CIRCLE = 0, ROW_SELF = 1, ROW_SELF_NEXT = 2, ROW_PLANE = 3, ROW_PLANE_STALE = 4 };
typedef struct { int kind, i, j; } RowDesc;

typedef struct {
  const MmpcConfig* cfg;
  int N, nobs, npl, mode, rmax;
  double dt;
  const double *xref, *uref, *ulast, *circles, *planes;
  double obj_scale;
  int circ_stride; /* doubles between stages of the circle table (0 = static) */
  unsigned flags;
  double x0[NX];
  double nu[2], nun[2]; /* multipliers of the terminal xy equality (flags bit 0; interface_wholebody_qref.py:167) and their Newton target */
  double ulo[64][NU], uhi[64][NU]; /* merged u box (ulim and dulim about u_last), filled per stage */
  /* iterate */
  double *x, *u, *s, *lam;       /* x[(N+1)*9], u[N*5], s[N+1], lam[(N+1)*9] */
  double *t, *z;                 /* [(N+1)*rmax] */
  double *zxl, *zxu, *zul, *zuu; /* box multipliers */
  int* nrow; RowDesc* rows;      /* rows of slack m: rows[m*rmax + r] */
  /* direction */
  double *dx, *du, *ds, *lamn, *dtt, *dz, *dzxl, *dzxu, *dzul, *dzuu;
  /* stage QP */
  double *H, *g, *K, *kff, *P, *p, *dfc; /* H[N+1][16][16], g[N+1][16], K[N][6][10], kff[N][6], P[N+1][10][10], p[N+1][10], dfc[N][9] */
  double *res; /* per row residual h - s + t */
  double *gy;  /* per row gradient in y space of its stage: [(N+1)*rmax][16] */
  int* rstage; /* per row: stage whose y-vector the row lives in */
  FK* fk;
} Work;

#define XK(w, k) ((w)->x + (k) * NX)
#define UK(w, k) ((w)->u + (k) * NU)

static const double* circle_at(const Work* w, int k, int i) {
  return w->circles + (size_t)k * w->circ_stride + 3 * i;
}

static int is_fin(double v) { return v > -1e300 && v < 1e300; }

/* angleDiff (controllers/mpc_base.py:59-84 = mpc_wholebody_qref.py:92-117): a - b folded to the nearest representative */
static double angle_diff(double a, double b) {
  const double PI = 3.14159265358979323846;
  a = fmod(a + PI, 2 * PI) - PI; b = fmod(b + PI, 2 * PI) - PI;
  double d = a - b;
  if (a * b >= 0) return d;
  if (a > b) return d <= PI ? d : d - 2 * PI;
  return d > -PI ? d : d + 2 * PI;
}
/* state error of the cost: the yaw error of MPCBase goes through angleDiff (mpc_base.py:129-133, :146-150) */
static double xerr(const MmpcConfig* c, int i, double x, double xr) {
  return (i == 2 && c->model == MMPC_MODEL_BASE) ? angle_diff(x, xr) : x - xr;
}

static void build_rows(Work* w) {
  int N = w->N;
  for (int m = 0; m <= N; ++m) {
    RowDesc* r = w->rows + (size_t)m * w->rmax; int n = 0;
    for (int i = 0; i < w->nobs; ++i) r[n++] = (RowDesc){ROW_CIRCLE, i, 0};
    int term_on_prev = (w->mode == MMPC_MODE_REFERENCE); /* quirk 3, :263-265 */
    if (w->cfg->terminal_rows_on_sN) term_on_prev = 0; /* rows on s_N: the variant the GPU's reference mode solves */
    const int nself = (w->cfg->model == MMPC_MODEL_WHOLEBODY) ? 4 : 0; /* MPCBase has no arm (controllers/mpc_base.py); the pose-reference controller no self-collision rows (controllers/mpc_wholebody.py:100 "TODO") */
    if (m < N || !term_on_prev)
      for (int i = 0; i < nself; ++i) r[n++] = (RowDesc){ROW_SELF, i, 0};
    if (m == N - 1 && term_on_prev)
      for (int i = 0; i < nself; ++i) r[n++] = (RowDesc){ROW_SELF_NEXT, i, 0};
    if (w->npl > 0) {
      for (int i = 0; i < 6; ++i) {
        if (w->mode == MMPC_MODE_REFERENCE && m >= 1) /* quirk 1 (:89 inside the j loop); k=0 rows vacuous (quirk 2) */
          for (int j = 0; j < w->npl - 1; ++j) r[n++] = (RowDesc){ROW_PLANE_STALE, i, j};
        r[n++] = (RowDesc){ROW_PLANE, i, w->npl - 1};
      }
    }
    w->nrow[m] = n;
  }
}

/* argmax with casadi semantics: n==2 -> if_else(c0 > c1, c0, c1) (:85); otherwise first max */
static int argmax_c(const double* c, int n) {
  if (n == 2) return c[0] > c[1] ? 0 : 1;
  int b = 0; for (int j = 1; j < n; ++j) if (c[j] > c[b]) b = j; return b;
}

/* Evaluate one row of slack m at the trial/current iterate (xs = states array).
 * Returns h; if g != NULL also gradient/Hessian in pose space of stage *kst, and the kind of
 * embedding (*emb): 0 = state row (x_kst, s_kst), 1 = mixed row (x_kst, v_kst), 2 = next-state row
 * (pose of x_{kst+1} = linear map of (x_kst, u_kst), slack s_kst). */
static double row_eval(const Work* w, const double* xs, const FK* fks, int m, const RowDesc* rd,
                       int* kst, int* emb, double g[NP], double H[NP][NP]) {
  const MmpcConfig* c = w->cfg; double h = 0;
  switch (rd->kind) {
    case ROW_CIRCLE:
      *kst = m; *emb = 0;
      row_circle(xs + m * NX, circle_at(w, m, rd->i), c->base_radius, &h, g, H);
      break;
    case ROW_SELF:
      *kst = m; *emb = 0;
      row_selfcoll(xs + m * NX, fks + m, rd->i, c->self_collision_radius, &h, g, H);
      break;
    case ROW_SELF_NEXT:
      *kst = m; *emb = 2;
      row_selfcoll(xs + (m + 1) * NX, fks + m + 1, rd->i, c->self_collision_radius, &h, g, H);
      break;
    case ROW_PLANE:
    case ROW_PLANE_STALE: {
      double cv[MMPC_MAX_PLANES]; int st[MMPC_MAX_PLANES];
      for (int j = 0; j < w->npl; ++j) {
        st[j] = (rd->kind == ROW_PLANE_STALE && j > rd->j) ? m - 1 : m;
        cv[j] = plane_margin(xs + st[j] * NX, fks + st[j], rd->i, w->planes + 6 * j, c->obstacle_expand_dist);
      }
      int b = argmax_c(cv, w->npl);
      h = -cv[b]; *kst = st[b]; *emb = (st[b] == m) ? 0 : 1;
      if (g) plane_derivs(xs + st[b] * NX, fks + st[b], rd->i, w->planes + 6 * b, g, H);
      break;
    }
  }
  return h;
}

/* cost of the current iterate: mpc_wholebody_qref.py:192-201,227,240-242,270 */
static double cost_eval(const Work* w, const double* x, const double* u, const double* s) {
  const MmpcConfig* c = w->cfg; int N = w->N; double J = 0;
  for (int k = 0; k <= N; ++k) {
    const double* Wx = (k < N) ? c->Qd : c->Pd;
    if (c->model == MMPC_MODEL_POSEREF) { FK f; fk_eval(x + k * NX, &f); J += pose_cost(x + k * NX, &f, Wx, w->xref + k * NX, NULL, NULL); }
    else for (int i = 0; i < NX; ++i) { double e = xerr(c, i, x[k * NX + i], w->xref[k * NX + i]); J += Wx[i] * e * e; }
    if (k < N)
      for (int j = 0; j < NU; ++j) {
        double e = u[k * NU + j] - w->uref[k * NU + j], dl = u[k * NU + j] - w->ulast[k * NU + j];
        J += c->Rd[j] * e * e + c->Wd[j] * dl * dl;
      }
    J += c->S * s[k] * s[k];
  }
  return J;
}

/* ---- merit ingredients at a trial point (values only) ---- */
typedef struct { double f, logsum, theta; int ok; } Merit;

static Merit merit_eval(const Work* w, const double* x, const double* u, const double* s, double* t, FK* fks, int reset) {
  Merit mt = {0, 0, 0, 1}; int N = w->N; const MmpcConfig* c = w->cfg;
  for (int k = 0; k <= N; ++k) fk_eval(x + k * NX, fks + k);
  mt.f = cost_eval(w, x, u, s);
  for (int k = 0; k < N; ++k) {
    double xn[NX]; dyn_f(x + k * NX, u + k * NU, w->dt, xn);
    for (int i = 0; i < NX; ++i) mt.theta += fabs(xn[i] - x[(k + 1) * NX + i]);
    for (int j = 0; j < NU; ++j) {
      double lo = w->ulo[k][j], hi = w->uhi[k][j], v = u[k * NU + j];
      if (is_fin(lo)) { if (v - lo <= 0) mt.ok = 0; else mt.logsum += log(v - lo); }
      if (is_fin(hi)) { if (hi - v <= 0) mt.ok = 0; else mt.logsum += log(hi - v); }
    }
  }
  for (int k = 1; k <= N; ++k)
    for (int i = 0; i < NX; ++i) {
      double lo = c->xlim[0][i], hi = c->xlim[1][i], v = x[k * NX + i];
      if (is_fin(lo)) { if (v - lo <= 0) mt.ok = 0; else mt.logsum += log(v - lo); }
      if (is_fin(hi)) { if (hi - v <= 0) mt.ok = 0; else mt.logsum += log(hi - v); }
    }
  if (w->flags & 1)
    for (int i = 0; i < 2; ++i) mt.theta += fabs(x[N * NX + i] - w->xref[N * NX + i]);
  for (int m = 0; m <= N; ++m)
    for (int r = 0; r < w->nrow[m]; ++r) {
      int kst, emb; double tt = t[m * w->rmax + r];
      double h = row_eval(w, x, fks, m, w->rows + (size_t)m * w->rmax + r, &kst, &emb, NULL, NULL);
      if (reset && s[m] - h > tt) { tt = s[m] - h; t[m * w->rmax + r] = tt; } /* slack reset (Nocedal & Wright 19.30) */
      mt.theta += fabs(h - s[m] + tt);
      if (tt <= 0) mt.ok = 0; else mt.logsum += log(tt);
    }
  if (!(mt.f == mt.f) || !(mt.theta == mt.theta)) mt.ok = 0;
  return mt;
}

/* dense symmetric solve helpers for the Riccati step: eliminate the last NUA variables of a
 * NYxNY system (Cholesky of the uu block).  Returns 0 on success, 1 if a pivot is not positive. */
static int chol6(double M[NUA][NUA], double L[NUA][NUA]) {
  memset(L, 0, sizeof(double) * NUA * NUA);
  for (int j = 0; j < NUA; ++j) {
    double d = M[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 1e-13)) return 1;
    L[j][j] = sqrt(d);
    for (int i = j + 1; i < NUA; ++i) {
      double v = M[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / L[j][j];
    }
  }
  return 0;
}
static void chol6_solve(double L[NUA][NUA], double* b) { /* in place */
  for (int i = 0; i < NUA; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[i][k] * b[k]; b[i] = v / L[i][i]; }
  for (int i = NUA - 1; i >= 0; --i) { double v = b[i]; for (int k = i + 1; k < NUA; ++k) v -= L[k][i] * b[k]; b[i] = v / L[i][i]; }
}

typedef struct {
  double e_stat, e_prim, e_comp_hi, e_comp_lo; /* raw inf-norms; comp as max/min of z*t */
  double sum_lam, sum_z; int n_z, n_all;
} KktParts;

static double kkt_error(const KktParts* k, double mu) {
  const double smax = 100.0;
  double sd = fmax(smax, (k->sum_lam + k->sum_z) / fmax(1, k->n_all + k->n_z)) / smax;
  double sc = fmax(smax, k->sum_z / fmax(1, k->n_z)) / smax;
  double ec = fmax(fabs(k->e_comp_hi - mu), fabs(k->e_comp_lo - mu));
  if (k->n_z == 0) ec = 0;
  return fmax(fmax(k->e_stat / sd, k->e_prim), ec / sc);
}

/* Full evaluation at the current iterate: builds the condensed stage QPs (H, g), the per-row
 * gradients in y space, the dynamics defects, and the KKT error ingredients. */
static void evaluate(Work* w, double mu, KktParts* kp) {
  const MmpcConfig* c = w->cfg; int N = w->N; double dt = w->dt;
  memset(w->H, 0, sizeof(double) * (N + 1) * NY * NY);
  memset(w->g, 0, sizeof(double) * (N + 1) * NY);
  double* stat = (double*)calloc((size_t)(N + 1) * NY, sizeof(double)); /* stationarity accumulators */
  memset(kp, 0, sizeof *kp); kp->e_comp_lo = 1e300; kp->e_comp_hi = -1e300;
  for (int k = 0; k <= N; ++k) fk_eval(XK(w, k), w->fk + k);

  /* cost and boxes */
  for (int k = 0; k <= N; ++k) {
    double* H = w->H + (size_t)k * NY * NY; double* g = w->g + k * NY; double* st = stat + k * NY;
    const double* Wx = (k < N) ? c->Qd : c->Pd;
    const int pose = c->model == MMPC_MODEL_POSEREF;
    if (pose) {
      double gp[NP], Hp[NP][NP];
      pose_cost(XK(w, k), w->fk + k, Wx, w->xref + k * NX, gp, Hp);
      for (int a = 0; a < NP; ++a) {
        g[POSE2X[a]] += gp[a]; st[POSE2X[a]] += gp[a];
        for (int b = 0; b < NP; ++b) H[POSE2X[a] * NY + POSE2X[b]] += Hp[a][b];
      }
    }
    for (int i = 0; i < NX; ++i) {
      double gr = pose ? 0.0 : 2 * Wx[i] * xerr(w->cfg, i, XK(w, k)[i], w->xref[k * NX + i]);
      H[i * NY + i] += pose ? 0.0 : 2 * Wx[i]; g[i] += gr; st[i] += gr;
      if (k >= 1) {
        double lo = c->xlim[0][i], hi = c->xlim[1][i], v = XK(w, k)[i];
        if (is_fin(lo)) { double d = v - lo, zz = w->zxl[k * NX + i]; H[i * NY + i] += zz / d; g[i] -= mu / d; st[i] -= zz;
          kp->e_comp_hi = fmax(kp->e_comp_hi, zz * d); kp->e_comp_lo = fmin(kp->e_comp_lo, zz * d); kp->sum_z += zz; kp->n_z++; }
        if (is_fin(hi)) { double d = hi - v, zz = w->zxu[k * NX + i]; H[i * NY + i] += zz / d; g[i] += mu / d; st[i] += zz;
          kp->e_comp_hi = fmax(kp->e_comp_hi, zz * d); kp->e_comp_lo = fmin(kp->e_comp_lo, zz * d); kp->sum_z += zz; kp->n_z++; }
      }
    }
    H[IS * NY + IS] += 2 * c->S; g[IS] += 2 * c->S * w->s[k]; st[IS] += 2 * c->S * w->s[k];
    if (k < N)
      for (int j = 0; j < NU; ++j) {
        double v = UK(w, k)[j];
        double gr = 2 * c->Rd[j] * (v - w->uref[k * NU + j]) + 2 * c->Wd[j] * (v - w->ulast[k * NU + j]);
        int a = IU + j;
        H[a * NY + a] += 2 * c->Rd[j] + 2 * c->Wd[j]; g[a] += gr; st[a] += gr;
        double lo = w->ulo[k][j], hi = w->uhi[k][j];
        if (is_fin(lo)) { double d = v - lo, zz = w->zul[k * NU + j]; H[a * NY + a] += zz / d; g[a] -= mu / d; st[a] -= zz;
          kp->e_comp_hi = fmax(kp->e_comp_hi, zz * d); kp->e_comp_lo = fmin(kp->e_comp_lo, zz * d); kp->sum_z += zz; kp->n_z++; }
        if (is_fin(hi)) { double d = hi - v, zz = w->zuu[k * NU + j]; H[a * NY + a] += zz / d; g[a] += mu / d; st[a] += zz;
          kp->e_comp_hi = fmax(kp->e_comp_hi, zz * d); kp->e_comp_lo = fmin(kp->e_comp_lo, zz * d); kp->sum_z += zz; kp->n_z++; }
      }
  }
  /* terminal equality (x_N, y_N) = (xref_N[0], xref_N[1]):  d nu = (dx + c) / delta_c */
  if (w->flags & 1)
    for (int i = 0; i < 2; ++i) {
      double cq = XK(w, N)[i] - w->xref[N * NX + i];
      w->H[(size_t)N * NY * NY + i * NY + i] += 1.0 / DELTA_C;
      w->g[N * NY + i] += w->nu[i] + cq / DELTA_C;
      stat[N * NY + i] += w->nu[i];
      kp->e_prim = fmax(kp->e_prim, fabs(cq)); kp->sum_lam += fabs(w->nu[i]); kp->n_all += 1;
    }
  /* dynamics: defects, second-order terms, multiplier terms of the stationarity residual */
  for (int k = 0; k < N; ++k) {
    double A[NX][NX], B[NX][NU], xn[NX];
    const double *x = XK(w, k), *u = UK(w, k), *lam = w->lam + (k + 1) * NX;
    dyn_f(x, u, dt, xn); dyn_AB(x, u, dt, A, B);
    double* H = w->H + (size_t)k * NY * NY; double* st = stat + k * NY;
    for (int i = 0; i < NX; ++i) {
      double d = xn[i] - XK(w, k + 1)[i];
      w->dfc[k * NX + i] = d; kp->e_prim = fmax(kp->e_prim, fabs(d));
      kp->sum_lam += fabs(lam[i]);
    }
    kp->n_all += NX;
    double cp = w->fk[k].cp, sp = w->fk[k].sp;
    double hpp = -dt * u[0] * (lam[3] * cp + lam[4] * sp), hpu = dt * (-lam[3] * sp + lam[4] * cp);
    H[2 * NY + 2] += hpp;
    H[2 * NY + IU] += hpu; H[IU * NY + 2] += hpu;
    H[4 * NY + 5] += -dt * lam[3]; H[5 * NY + 4] += -dt * lam[3];
    H[3 * NY + 5] += dt * lam[4];  H[5 * NY + 3] += dt * lam[4];
    for (int i = 0; i < NX; ++i) {
      double v = 0; for (int r = 0; r < NX; ++r) v += A[r][i] * lam[r];
      st[i] += v;
    }
    for (int j = 0; j < NU; ++j) {
      double v = 0; for (int r = 0; r < NX; ++r) v += B[r][j] * lam[r];
      st[IU + j] += v;
    }
    for (int i = 0; i < NX; ++i) stat[(k + 1) * NY + i] -= lam[i];
  }
  /* inequality rows with slack t */
  for (int m = 0; m <= N; ++m)
    for (int r = 0; r < w->nrow[m]; ++r) {
      size_t ri = (size_t)m * w->rmax + r;
      int kst, emb; double gp[NP], Hp[NP][NP];
      double h = row_eval(w, w->x, w->fk, m, w->rows + ri, &kst, &emb, gp, Hp);
      double t = w->t[ri], z = w->z[ri];
      double res = h - w->s[m] + t, sig = z / t;
      kp->e_prim = fmax(kp->e_prim, fabs(res));
      kp->e_comp_hi = fmax(kp->e_comp_hi, z * t); kp->e_comp_lo = fmin(kp->e_comp_lo, z * t);
      kp->sum_z += z; kp->n_z++;
      double* gy = w->gy + ri * NY; memset(gy, 0, sizeof(double) * NY);
      double Hy[NY][NY]; memset(Hy, 0, sizeof Hy);
      w->rstage[ri] = kst;
      if (emb == 2) {
        /* pose(x_{k+1}) = pose(x_k) + dt*(dx, dy, dpsi, u2, u3, u4): linear map L (6 x NY) */
        int c1[NP] = {0, 1, 2, 6, 7, 8}, c2[NP] = {3, 4, 5, IU + 2, IU + 3, IU + 4};
        for (int a = 0; a < NP; ++a) { gy[c1[a]] += gp[a]; gy[c2[a]] += dt * gp[a]; }
        for (int a = 0; a < NP; ++a)
          for (int b = 0; b < NP; ++b) {
            Hy[c1[a]][c1[b]] += Hp[a][b]; Hy[c1[a]][c2[b]] += dt * Hp[a][b];
            Hy[c2[a]][c1[b]] += dt * Hp[a][b]; Hy[c2[a]][c2[b]] += dt * dt * Hp[a][b];
          }
        gy[IS] = -1;
      } else {
        for (int a = 0; a < NP; ++a) {
          gy[POSE2X[a]] = gp[a];
          for (int b = 0; b < NP; ++b) Hy[POSE2X[a]][POSE2X[b]] = Hp[a][b];
        }
        gy[emb == 0 ? IS : IV] = -1;
      }
      double* H = w->H + (size_t)kst * NY * NY; double* g = w->g + kst * NY; double* st = stat + kst * NY;
      double coef = mu / t + sig * res;
      w->res[ri] = res;
      for (int a = 0; a < NY; ++a) {
        g[a] += coef * gy[a]; st[a] += z * gy[a];
        for (int b = 0; b < NY; ++b) H[a * NY + b] += sig * gy[a] * gy[b] + z * Hy[a][b];
      }
    }
  /* fold v_k stationarity into s_{k+1}; gather inf-norm */
  for (int k = 0; k <= N; ++k) {
    double* st = stat + k * NY;
    if (k < N) { stat[(k + 1) * NY + IS] += st[IV]; st[IV] = 0; }
  }
  for (int k = 0; k <= N; ++k) {
    double* st = stat + k * NY;
    for (int a = 0; a < IV; ++a) {
      if (k == 0 && a < NX) continue;      /* x_0 is fixed */
      if (k == N && a >= IU) continue;     /* no control at the terminal stage */
      kp->e_stat = fmax(kp->e_stat, fabs(st[a]));
    }
  }
  free(stat);
}

/* Riccati factorisation + solve of the stage QPs.  reg = delta_w.  Returns 0 or 1 (bad inertia). */
static int riccati(Work* w, double reg) {
  int N = w->N;
  double* PN = w->P + (size_t)N * NXA * NXA; double* pN = w->p + N * NXA;
  const double* HN = w->H + (size_t)N * NY * NY;
  for (int a = 0; a < NXA; ++a) {
    for (int b = 0; b < NXA; ++b) PN[a * NXA + b] = HN[a * NY + b];
    if (a != IS) PN[a * NXA + a] += reg; /* s_k is never regularised: its pivot 2S + sum(sigma) is always positive */
    pN[a] = w->g[N * NY + a];
  }
  for (int k = N - 1; k >= 0; --k) {
    double A[NX][NX], B[NX][NU];
    dyn_AB(XK(w, k), UK(w, k), w->dt, A, B);
    /* G = [Atilde Btilde] (NXA x NY) in y order */
    double G[NXA][NY]; memset(G, 0, sizeof G);
    for (int r = 0; r < NX; ++r) { for (int i = 0; i < NX; ++i) G[r][i] = A[r][i]; for (int j = 0; j < NU; ++j) G[r][IU + j] = B[r][j]; }
    G[IS][IV] = 1;
    const double* Pn = w->P + (size_t)(k + 1) * NXA * NXA; const double* pn = w->p + (k + 1) * NXA;
    double PG[NXA][NY], M[NY][NY], mv[NY], pd[NXA];
    for (int r = 0; r < NXA; ++r) {
      double v = pn[r];
      for (int q = 0; q < NX; ++q) v += Pn[r * NXA + q] * w->dfc[k * NX + q];
      pd[r] = v;
      for (int a = 0; a < NY; ++a) { double s = 0; for (int q = 0; q < NXA; ++q) s += Pn[r * NXA + q] * G[q][a]; PG[r][a] = s; }
    }
    const double* H = w->H + (size_t)k * NY * NY;
    for (int a = 0; a < NY; ++a) {
      for (int b = 0; b < NY; ++b) { double s = H[a * NY + b]; for (int r = 0; r < NXA; ++r) s += G[r][a] * PG[r][b]; M[a][b] = s; }
      if (a != IV && a != IS) M[a][a] += reg;
      double s = w->g[k * NY + a]; for (int r = 0; r < NXA; ++r) s += G[r][a] * pd[r]; mv[a] = s;
    }
    double Muu[NUA][NUA], L[NUA][NUA];
    for (int a = 0; a < NUA; ++a) for (int b = 0; b < NUA; ++b) Muu[a][b] = M[IU + a][IU + b];
    if (chol6(Muu, L)) return 1;
    double* K = w->K + (size_t)k * NUA * NXA; double* kff = w->kff + k * NUA;
    double col[NUA];
    for (int i = 0; i < NXA; ++i) {
      for (int a = 0; a < NUA; ++a) col[a] = -M[IU + a][i];
      chol6_solve(L, col);
      for (int a = 0; a < NUA; ++a) K[a * NXA + i] = col[a];
    }
    for (int a = 0; a < NUA; ++a) col[a] = -mv[IU + a];
    chol6_solve(L, col);
    for (int a = 0; a < NUA; ++a) kff[a] = col[a];
    double* Pk = w->P + (size_t)k * NXA * NXA; double* pk = w->p + k * NXA;
    for (int i = 0; i < NXA; ++i) {
      for (int j = 0; j < NXA; ++j) { double s = M[i][j]; for (int a = 0; a < NUA; ++a) s += M[i][IU + a] * K[a * NXA + j]; Pk[i * NXA + j] = s; }
      double s = mv[i]; for (int a = 0; a < NUA; ++a) s += M[i][IU + a] * kff[a]; pk[i] = s;
    }
    for (int i = 0; i < NXA; ++i) for (int j = i + 1; j < NXA; ++j) { double s = 0.5 * (Pk[i * NXA + j] + Pk[j * NXA + i]); Pk[i * NXA + j] = Pk[j * NXA + i] = s; }
  }
  /* stage 0: dx_0 = 0, ds_0 free */
  double* P0 = w->P; double* p0 = w->p;
  if (!(P0[IS * NXA + IS] > 1e-13)) return 1;
  double dxa[NXA]; memset(dxa, 0, sizeof dxa);
  dxa[IS] = -p0[IS] / P0[IS * NXA + IS];
  for (int i = 0; i < NX; ++i) w->dx[i] = 0;
  for (int k = 0; k < N; ++k) {
    w->ds[k] = dxa[IS];
    const double* K = w->K + (size_t)k * NUA * NXA; const double* kff = w->kff + k * NUA;
    double dua[NUA];
    for (int a = 0; a < NUA; ++a) { double s = kff[a]; for (int i = 0; i < NXA; ++i) s += K[a * NXA + i] * dxa[i]; dua[a] = s; }
    for (int j = 0; j < NU; ++j) w->du[k * NU + j] = dua[j];
    double A[NX][NX], B[NX][NU]; dyn_AB(XK(w, k), UK(w, k), w->dt, A, B);
    double nxt[NXA];
    for (int r = 0; r < NX; ++r) {
      double s = w->dfc[k * NX + r];
      for (int i = 0; i < NX; ++i) s += A[r][i] * dxa[i];
      for (int j = 0; j < NU; ++j) s += B[r][j] * dua[j];
      nxt[r] = s;
    }
    nxt[IS] = dua[NUA - 1];
    memcpy(dxa, nxt, sizeof nxt);
    for (int i = 0; i < NX; ++i) w->dx[(k + 1) * NX + i] = dxa[i];
    const double* Pn = w->P + (size_t)(k + 1) * NXA * NXA; const double* pn = w->p + (k + 1) * NXA;
    for (int i = 0; i < NX; ++i) { double s = pn[i]; for (int q = 0; q < NXA; ++q) s += Pn[i * NXA + q] * dxa[q]; w->lamn[(k + 1) * NX + i] = s; }
  }
  w->ds[N] = dxa[IS];
  if (w->flags & 1)
    for (int i = 0; i < 2; ++i)
      w->nun[i] = w->nu[i] + (dxa[i] + XK(w, N)[i] - w->xref[N * NX + i]) / DELTA_C;
  return 0;
}

static void push_in(double* v, double lo, double hi) {
  const double k1 = 1e-2, k2 = 1e-2;
  if (is_fin(lo) && is_fin(hi)) {
    double pl = fmin(k1 * fmax(1, fabs(lo)), k2 * (hi - lo)), pu = fmin(k1 * fmax(1, fabs(hi)), k2 * (hi - lo));
    if (*v < lo + pl) *v = lo + pl; if (*v > hi - pu) *v = hi - pu;
  } else if (is_fin(lo)) { double pl = k1 * fmax(1, fabs(lo)); if (*v < lo + pl) *v = lo + pl; }
  else if (is_fin(hi)) { double pu = k1 * fmax(1, fabs(hi)); if (*v > hi - pu) *v = hi - pu; }
}

#define ALLOC(p, n) p = calloc((size_t)(n), sizeof *(p))

static int solve_one(const MmpcConfig* cfg, int npl, const double* x_init, const double* x_ref, const double* u_ref,
                     const double* u_last, const double* u_guess, const double* x_guess, const double* circles, const double* planes,
                     unsigned flags, double* Uo, double* Xo, double* so, double* cost, double* kkt, int32_t* iters) {
  Work W, *w = &W; memset(w, 0, sizeof W);
  int N = cfg->N; const double dt = cfg->dt;
  MmpcConfig sc = *cfg; w->cfg = &sc; w->obj_scale = 1.0; w->N = N; w->nobs = cfg->n_obs; w->npl = npl; w->mode = cfg->mode; w->dt = dt;
  w->xref = x_ref; w->uref = u_ref; w->ulast = u_last; w->circles = circles; w->planes = planes; w->flags = flags;
  w->circ_stride = cfg->obs_per_stage ? 3 * cfg->n_obs : 0;
  w->rmax = cfg->n_obs + 8 + 6 * MMPC_MAX_PLANES;
  if (N > 64) return MMPC_STATUS_NAN;
  size_t nr = (size_t)(N + 1) * w->rmax;
  ALLOC(w->x, (N + 1) * NX); ALLOC(w->u, N * NU + 1); ALLOC(w->s, N + 1); ALLOC(w->lam, (N + 1) * NX);
  ALLOC(w->t, nr); ALLOC(w->z, nr); ALLOC(w->zxl, (N + 1) * NX); ALLOC(w->zxu, (N + 1) * NX); ALLOC(w->zul, N * NU + 1); ALLOC(w->zuu, N * NU + 1);
  ALLOC(w->nrow, N + 1); ALLOC(w->rows, nr);
  ALLOC(w->dx, (N + 1) * NX); ALLOC(w->du, N * NU + 1); ALLOC(w->ds, N + 1); ALLOC(w->lamn, (N + 1) * NX);
  ALLOC(w->dtt, nr); ALLOC(w->dz, nr); ALLOC(w->dzxl, (N + 1) * NX); ALLOC(w->dzxu, (N + 1) * NX); ALLOC(w->dzul, N * NU + 1); ALLOC(w->dzuu, N * NU + 1);
  ALLOC(w->H, (size_t)(N + 1) * NY * NY); ALLOC(w->g, (N + 1) * NY); ALLOC(w->K, (size_t)N * NUA * NXA + 1); ALLOC(w->kff, N * NUA + 1);
  ALLOC(w->P, (size_t)(N + 1) * NXA * NXA); ALLOC(w->p, (N + 1) * NXA); ALLOC(w->dfc, N * NX + 1);
  ALLOC(w->gy, nr * NY); ALLOC(w->res, nr); ALLOC(w->rstage, nr); ALLOC(w->fk, N + 1);
  double *xt, *ut, *st, *tt; FK* fkt;
  ALLOC(xt, (N + 1) * NX); ALLOC(ut, N * NU + 1); ALLOC(st, N + 1); ALLOC(tt, nr); ALLOC(fkt, N + 1);
  build_rows(w);

  /* solve(): clip x_init to xlim (:290-291) */
  for (int i = 0; i < NX; ++i) w->x0[i] = fmax(fmin(x_init[i], cfg->xlim[1][i]), cfg->xlim[0][i]);
  /* initial guess (:302-304): X <- tile(x_init), U <- u_latest, s <- 0; then IPOPT's bound push */
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < NU; ++j) {
      double ul = u_last[k * NU + j];
      w->ulo[k][j] = fmax(cfg->ulim[0][j], ul + cfg->dulim[0][j]);   /* :203 and :205 merged */
      w->uhi[k][j] = fmin(cfg->ulim[1][j], ul + cfg->dulim[1][j]);
      double v = (u_guess ? u_guess : u_last)[k * NU + j];
      push_in(&v, w->ulo[k][j], w->uhi[k][j]); UK(w, k)[j] = v;
      w->zul[k * NU + j] = 1; w->zuu[k * NU + j] = 1;
    }
  for (int k = 0; k <= N; ++k)
    for (int i = 0; i < NX; ++i) {
      double v = w->x0[i];
      if (k >= 1 && x_guess) v = x_guess[k * NX + i]; /* MPCBase: X <- previous solution (mpc_base.py:196-201) */
      if (k >= 1) push_in(&v, cfg->xlim[0][i], cfg->xlim[1][i]);
      XK(w, k)[i] = v; w->zxl[k * NX + i] = 1; w->zxu[k * NX + i] = 1;
    }
  for (int k = 0; k <= N; ++k) fk_eval(XK(w, k), w->fk + k);
  for (int m = 0; m <= N; ++m) {
    /* s <- 0 (:304) unless a row is violated at the start: then s_m is lifted so that every row of
     * slack m starts strictly feasible (s is a free variable, so this is only a choice of start) */
    double hmax = -1e300;
    for (int r = 0; r < w->nrow[m]; ++r) {
      int kst, emb; size_t ri = (size_t)m * w->rmax + r;
      double h = row_eval(w, w->x, w->fk, m, w->rows + ri, &kst, &emb, NULL, NULL);
      w->t[ri] = h; hmax = fmax(hmax, h);
    }
    w->s[m] = fmax(0.0, hmax + 1e-2);
    for (int r = 0; r < w->nrow[m]; ++r) {
      size_t ri = (size_t)m * w->rmax + r;
      w->t[ri] = w->s[m] - w->t[ri]; w->z[ri] = 1;
    }
  }

  /* IPOPT gradient-based objective scaling: nlp_scaling_max_gradient = 100 at the starting point */
  {
    double gmax = 0;
    for (int k = 1; k <= N; ++k) {
      const double* Wx = (k < N) ? cfg->Qd : cfg->Pd;
      if (cfg->model == MMPC_MODEL_POSEREF) {
        double gp[NP]; pose_cost(XK(w, k), w->fk + k, Wx, x_ref + k * NX, gp, NULL);
        for (int a = 0; a < NP; ++a) gmax = fmax(gmax, fabs(gp[a]));
      } else
      for (int i = 0; i < NX; ++i) gmax = fmax(gmax, fabs(2 * Wx[i] * xerr(cfg, i, XK(w, k)[i], x_ref[k * NX + i])));
    }
    for (int k = 0; k < N; ++k)
      for (int j = 0; j < NU; ++j)
        gmax = fmax(gmax, fabs(2 * cfg->Rd[j] * (UK(w, k)[j] - u_ref[k * NU + j]) + 2 * cfg->Wd[j] * (UK(w, k)[j] - u_last[k * NU + j])));
    for (int k = 0; k <= N; ++k) gmax = fmax(gmax, fabs(2 * cfg->S * w->s[k]));
    double os = (gmax > 100.0) ? fmax(100.0 / gmax, 1e-8) : 1.0;
    w->obj_scale = os;
    for (int i = 0; i < NX; ++i) { sc.Qd[i] *= os; sc.Pd[i] *= os; }
    for (int j = 0; j < NU; ++j) { sc.Rd[j] *= os; sc.Wd[j] *= os; }
    sc.S *= os;
  }
  enum { MAXFILT = 16 };
  double filt_th[MAXFILT], filt_ph[MAXFILT], theta_max = -1, theta_min = -1; int nfilt = 0, n_freset = 0, cnt_frej = 0, last_rej_filter = 0;
  double mu = cfg->mu_init, tol = cfg->tol, nu = 1.0, reg_last = 0;
  const double kap_eps = 10, kap_mu = 0.2, th_mu = 1.5, tau_min = 0.99;
  int it = 0, status = MMPC_STATUS_MAX_ITER; double E0 = 1e300;
  KktParts kp;
  for (it = 0; it <= cfg->max_iter; ++it) {
    evaluate(w, mu, &kp);
    E0 = kkt_error(&kp, 0.0);
    if (!(E0 == E0)) { status = MMPC_STATUS_NAN; break; }
    if (E0 <= tol) { status = MMPC_STATUS_CONVERGED; break; }
    if (it == cfg->max_iter) break;
    int mu_changed = 0;
    while (kkt_error(&kp, mu) <= kap_eps * mu && mu > tol / 10) {
      mu = fmax(tol / 10, fmin(kap_mu * mu, pow(mu, th_mu))); mu_changed = 1;
    }
    if (mu_changed) { evaluate(w, mu, &kp); nfilt = 0; cnt_frej = 0; last_rej_filter = 0; }
    double tau = fmax(tau_min, 1 - mu);
    /* factorise with inertia correction */
    double reg = 0; int tries = 0, fail;
    while ((fail = riccati(w, reg)) != 0) {
      if (reg == 0) reg = (reg_last == 0) ? 1e-4 : fmax(1e-20, reg_last / 3);
      else reg *= (reg_last == 0 ? 100 : 8);
      if (++tries > 40 || reg > 1e20) break;
    }
    if (fail) { status = MMPC_STATUS_FACTOR; break; }
    if (reg > 0) reg_last = reg;
    /* recover row / box steps and step-to-boundary */
    double ap = 1, ad = 1;
    double gphi = 0; /* directional derivative of the barrier objective */
    for (int k = 0; k <= N; ++k) {
      const MmpcConfig* c = w->cfg;
      const double* Wx = (k < N) ? c->Qd : c->Pd;
      if (c->model == MMPC_MODEL_POSEREF) {
        double gp[NP]; pose_cost(XK(w, k), w->fk + k, Wx, x_ref + k * NX, gp, NULL);
        for (int a = 0; a < NP; ++a) gphi += gp[a] * w->dx[k * NX + POSE2X[a]];
      }
      for (int i = 0; i < NX; ++i) {
        double dxi = w->dx[k * NX + i];
        if (c->model != MMPC_MODEL_POSEREF) gphi += 2 * Wx[i] * xerr(c, i, XK(w, k)[i], x_ref[k * NX + i]) * dxi;
        if (k >= 1) {
          double lo = c->xlim[0][i], hi = c->xlim[1][i], v = XK(w, k)[i];
          if (is_fin(lo)) { double d = v - lo, zz = w->zxl[k * NX + i]; double dz = mu / d - zz - zz / d * dxi; w->dzxl[k * NX + i] = dz;
            gphi -= mu * dxi / d; if (dxi < 0) ap = fmin(ap, -tau * d / dxi); if (dz < 0) ad = fmin(ad, -tau * zz / dz); }
          if (is_fin(hi)) { double d = hi - v, zz = w->zxu[k * NX + i]; double dz = mu / d - zz + zz / d * dxi; w->dzxu[k * NX + i] = dz;
            gphi += mu * dxi / d; if (dxi > 0) ap = fmin(ap, tau * d / dxi); if (dz < 0) ad = fmin(ad, -tau * zz / dz); }
        }
      }
      gphi += 2 * c->S * w->s[k] * w->ds[k];
      if (k < N)
        for (int j = 0; j < NU; ++j) {
          double v = UK(w, k)[j], duj = w->du[k * NU + j];
          gphi += (2 * c->Rd[j] * (v - u_ref[k * NU + j]) + 2 * c->Wd[j] * (v - u_last[k * NU + j])) * duj;
          double lo = w->ulo[k][j], hi = w->uhi[k][j];
          if (is_fin(lo)) { double d = v - lo, zz = w->zul[k * NU + j]; double dz = mu / d - zz - zz / d * duj; w->dzul[k * NU + j] = dz;
            gphi -= mu * duj / d; if (duj < 0) ap = fmin(ap, -tau * d / duj); if (dz < 0) ad = fmin(ad, -tau * zz / dz); }
          if (is_fin(hi)) { double d = hi - v, zz = w->zuu[k * NU + j]; double dz = mu / d - zz + zz / d * duj; w->dzuu[k * NU + j] = dz;
            gphi += mu * duj / d; if (duj > 0) ap = fmin(ap, tau * d / duj); if (dz < 0) ad = fmin(ad, -tau * zz / dz); }
        }
    }
    double theta0 = 0;
    for (int k = 0; k < N; ++k) for (int i = 0; i < NX; ++i) theta0 += fabs(w->dfc[k * NX + i]);
    for (int m = 0; m <= N; ++m)
      for (int r = 0; r < w->nrow[m]; ++r) {
        size_t ri = (size_t)m * w->rmax + r; int kst = w->rstage[ri]; const double* gy = w->gy + ri * NY;
        double dy[NY];
        for (int i = 0; i < NX; ++i) dy[i] = w->dx[kst * NX + i];
        dy[IS] = w->ds[kst];
        for (int j = 0; j < NU; ++j) dy[IU + j] = (kst < N) ? w->du[kst * NU + j] : 0;
        dy[IV] = (kst < N) ? w->ds[kst + 1] : 0;
        double gd = 0; for (int a = 0; a < NY; ++a) gd += gy[a] * dy[a];
        double t = w->t[ri], z = w->z[ri], res = w->res[ri];
        double dtr = -res - gd, dzr = mu / t - z + (z / t) * (res + gd);
        w->dtt[ri] = dtr; w->dz[ri] = dzr; theta0 += fabs(res);
        gphi -= mu * dtr / t;
        if (dtr < 0) ap = fmin(ap, -tau * t / dtr);
        if (dzr < 0) ad = fmin(ad, -tau * z / dzr);
      }
    /* filter line search (Waechter & Biegler 2006, Alg. A without second-order correction) */
    Merit m0 = merit_eval(w, w->x, w->u, w->s, w->t, fkt, 0);
    double theta_k = m0.theta, phi0 = m0.f - mu * m0.logsum, D = gphi;
    if (theta_max < 0) { theta_max = 1e4 * fmax(1, theta_k); theta_min = 1e-4 * fmax(1, theta_k); }
    /* IPOPT's filter reset heuristic (filter_reset_trigger 5, max_filter_resets 5): if in 5 successive iterations the last
     * rejected trial point was rejected by the filter although it made sufficient progress, the filter is cleared */
    if (n_freset < 5) {
      if (last_rej_filter) { if (++cnt_frej >= 5) { nfilt = 0; n_freset++; cnt_frej = 0; } }
      else cnt_frej = 0;
    }
    last_rej_filter = 0;
    /* IPOPT's minimal step size (IpFilterLSAcceptor::CalculateAlphaMin): below it the line search is given up */
    double alpha_min = 1e-5;
    if (gphi < 0) {
      alpha_min = fmin(1e-5, 1e-8 * theta_k / (-gphi));
      if (theta_k <= theta_min) alpha_min = fmin(alpha_min, pow(theta_k, 1.1) / pow(-gphi, 2.3));
    }
    alpha_min *= 0.05;
    double alpha = ap; int accepted = 0, ftype = 0;
    for (int ls = 0; ls < 50; ++ls) {
      for (int i = 0; i < (N + 1) * NX; ++i) xt[i] = w->x[i] + alpha * w->dx[i];
      for (int i = 0; i < N * NU; ++i) ut[i] = w->u[i] + alpha * w->du[i];
      for (int k = 0; k <= N; ++k) st[k] = w->s[k] + alpha * w->ds[k];
      for (size_t i = 0; i < nr; ++i) tt[i] = w->t[i] + alpha * w->dtt[i];
      Merit m1 = merit_eval(w, xt, ut, st, tt, fkt, 1);
      double th1 = m1.theta, ph1 = m1.f - mu * m1.logsum;
      int ok = m1.ok && th1 < theta_max;
      int dom = 0;
      for (int q = 0; q < nfilt; ++q) if (th1 >= filt_th[q] && ph1 >= filt_ph[q]) dom = 1;
      if (ok) {
        int sw = (gphi < 0) && (alpha * pow(-gphi, 2.3) > pow(theta_k, 1.1));
        if (theta_k <= theta_min && sw) {
          ok = ph1 <= phi0 + 1e-8 * alpha * gphi + 10 * DBL_EPSILON * fabs(phi0); ftype = ok;
        } else {
          ok = (th1 <= (1 - 1e-5) * theta_k) || (ph1 <= phi0 - 1e-8 * theta_k); ftype = 0;
        }
      }
      if (ok && !dom) { accepted = 1; break; }
      last_rej_filter = ok && dom;
      alpha *= 0.5;
      if (alpha <= alpha_min) ls = 49;   /* give up (or, below, clear the filter and search once more) */
      if (ls == 49 && n_freset < 5 && nfilt > 0) { nfilt = 0; n_freset++; cnt_frej = 0; alpha = ap; ls = -1; }  /* in place of IPOPT's restoration phase: clear the filter, search again */
    }
    if (accepted && !ftype) {
      if (nfilt == MAXFILT) { memmove(filt_th, filt_th + 1, sizeof(double) * (MAXFILT - 1)); memmove(filt_ph, filt_ph + 1, sizeof(double) * (MAXFILT - 1)); nfilt--; }
      filt_th[nfilt] = (1 - 1e-5) * theta_k; filt_ph[nfilt] = phi0 - 1e-8 * theta_k; nfilt++;
    }
    if (mmpc_oracle_verbose)
      printf("it %3d mu %.2e E0 %.3e Emu %.3e stat %.2e prim %.2e reg %.1e ap %.3f ad %.3f alpha %.4f nu %.2e theta %.2e D %.2e f %.6f\n",
             it, mu, E0, kkt_error(&kp, mu), kp.e_stat, kp.e_prim, reg, ap, ad, alpha, (double)nfilt, m0.theta, D, m0.f / w->obj_scale);
    if (!accepted) { status = MMPC_STATUS_LINESEARCH; break; }
    /* update */
    memcpy(w->x, xt, sizeof(double) * (N + 1) * NX); memcpy(w->u, ut, sizeof(double) * N * NU);
    memcpy(w->s, st, sizeof(double) * (N + 1)); memcpy(w->t, tt, sizeof(double) * nr);
    for (int i = NX; i < (N + 1) * NX; ++i) w->lam[i] += alpha * (w->lamn[i] - w->lam[i]);
    for (int i = 0; i < 2; ++i) w->nu[i] += alpha * (w->nun[i] - w->nu[i]);
    const double ks = 1e10;
#define ZUPD(zv, dzv, dist) do { double zn = (zv) + ad * (dzv); double dd = (dist); zn = fmax(fmin(zn, ks * mu / dd), mu / (ks * dd)); (zv) = zn; } while (0)
    for (size_t i = 0; i < nr; ++i) if (w->t[i] > 0) ZUPD(w->z[i], w->dz[i], w->t[i]);
    for (int k = 1; k <= N; ++k)
      for (int i = 0; i < NX; ++i) {
        double lo = cfg->xlim[0][i], hi = cfg->xlim[1][i], v = XK(w, k)[i];
        if (is_fin(lo)) ZUPD(w->zxl[k * NX + i], w->dzxl[k * NX + i], v - lo);
        if (is_fin(hi)) ZUPD(w->zxu[k * NX + i], w->dzxu[k * NX + i], hi - v);
      }
    for (int k = 0; k < N; ++k)
      for (int j = 0; j < NU; ++j) {
        double lo = w->ulo[k][j], hi = w->uhi[k][j], v = UK(w, k)[j];
        if (is_fin(lo)) ZUPD(w->zul[k * NU + j], w->dzul[k * NU + j], v - lo);
        if (is_fin(hi)) ZUPD(w->zuu[k * NU + j], w->dzuu[k * NU + j], hi - v);
      }
  }
  memcpy(Uo, w->u, sizeof(double) * N * NU);
  if (Xo) memcpy(Xo, w->x, sizeof(double) * (N + 1) * NX);
  if (so) memcpy(so, w->s, sizeof(double) * (N + 1));
  if (cost) *cost = cost_eval(w, w->x, w->u, w->s) / w->obj_scale;
  if (kkt) *kkt = E0;
  if (iters) *iters = it;
  free(w->x); free(w->u); free(w->s); free(w->lam); free(w->t); free(w->z); free(w->zxl); free(w->zxu); free(w->zul); free(w->zuu);
  free(w->nrow); free(w->rows); free(w->dx); free(w->du); free(w->ds); free(w->lamn); free(w->dtt); free(w->dz);
  free(w->dzxl); free(w->dzxu); free(w->dzul); free(w->dzuu); free(w->H); free(w->g); free(w->K); free(w->kff);
  free(w->P); free(w->p); free(w->dfc); free(w->gy); free(w->res); free(w->rstage); free(w->fk);
  free(xt); free(ut); free(st); free(tt); free(fkt);
  return status;
}

/* ---- exported (host pointers; same batch layout as include/mmpc.h) ---- */
int mmpc_oracle_solve(const MmpcConfig* cfg, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out, int32_t nthreads) {
  int N = cfg->N;
  size_t cs = (size_t)cfg->n_obs * 3 * (cfg->obs_per_stage ? (N + 1) : 1);
  (void)nthreads; /* threading is done by the caller (oracle/solver.py splits the batch over a thread pool) */
  for (int b = 0; b < B; ++b) {
    int npl = in->n_pl_inst ? in->n_pl_inst[b] : cfg->n_pl;
    double c = 0, e = 0; int32_t it = 0;
    int st = solve_one(cfg, npl, in->x_init + (size_t)b * NX, in->x_ref + (size_t)b * (N + 1) * NX,
                       in->u_ref + (size_t)b * N * NU, in->u_last + (size_t)b * N * NU,
                       in->u_guess ? in->u_guess + (size_t)b * N * NU : NULL,
                       in->x_guess ? in->x_guess + (size_t)b * (N + 1) * NX : NULL,
                       in->circles ? in->circles + b * cs : NULL,
                       in->planes ? in->planes + (size_t)b * cfg->n_pl * 6 : NULL,
                       in->flags ? in->flags[b] : 0,
                       out->U + (size_t)b * N * NU, out->X ? out->X + (size_t)b * (N + 1) * NX : NULL,
                       out->s ? out->s + (size_t)b * (N + 1) : NULL, &c, &e, &it);
    if (out->cost) out->cost[b] = c;
    if (out->kkt) out->kkt[b] = e;
    if (out->iters) out->iters[b] = it;
    out->status[b] = st;
  }
  return 0;
}

/* single-row derivative probe for the finite-difference / sympy tests:
 * kind 0 circle (par = ox, oy, r), 1 self-collision (idx), 2 plane (idx = body point, par = point, normal) */
void mmpc_oracle_row(int kind, int idx, const double* x, const double* par, double base_r, double rad,
                     double expand, double* h, double* g, double* H) {
  FK f; fk_eval(x, &f); double Hm[NP][NP], gm[NP];
  if (kind == 0) row_circle(x, par, base_r, h, gm, Hm);
  else if (kind == 1) row_selfcoll(x, &f, idx, rad, h, gm, Hm);
  else { *h = -plane_margin(x, &f, idx, par, expand); plane_derivs(x, &f, idx, par, gm, Hm); }
  memcpy(g, gm, sizeof gm); memcpy(H, Hm, sizeof Hm);
}

void mmpc_oracle_dynamics(const double* x, const double* u, double dt, double* xn, double* A, double* B) {
  double Am[NX][NX], Bm[NX][NU];
  dyn_f(x, u, dt, xn); dyn_AB(x, u, dt, Am, Bm);
  memcpy(A, Am, sizeof Am); memcpy(B, Bm, sizeof Bm);
}
