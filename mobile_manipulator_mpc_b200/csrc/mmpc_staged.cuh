// mmpc_staged.cuh -- batch-synchronous ("staged") interior-point solve of the whole-body MPC NLP
// (controllers/mpc_wholebody_qref.py:142-285), replacing opti.solve() (:315), i.e. the
// CasADi/IPOPT/MUMPS stack, for tens of thousands of independent instances at once.
//
// Mapping (B200).  One interior-point iteration of the whole batch is a short sequence of kernels
// over compacted lists of still-active instances; all per-instance state lives in HBM in a
// tile-major, lane-interleaved layout  ws[((tile*(N+1) + stage)*STG + field)*32 + lane]  (see Inst)
// so that every access of a warp of neighbouring instances is one coalesced 256-byte line:
//
//   eval        thread per (instance, stage)   dynamics, FK, every inequality row with gradient and
//                                              Hessian, barrier condensation -> stage QP, KKT partials
//   riccati     16 lanes per instance          KKT reduction + convergence test + barrier update,
//               (mmpc_team.cuh; one thread     column-parallel register-resident Riccati recursion (inertia
//               per instance here as A/B)      correction by delta_w), roll-out of the Newton step
//   step        thread per (instance, stage)   slack / multiplier steps of every row and bound,
//                                              fraction to the boundary, merit ingredients
//   ctrl_step   thread per instance            reduction, line-search start
//   trial       thread per (instance, stage)   candidate iterate  w + alpha d  (primal and dual) into
//                                              the other half of a ping-pong buffer, merit values, and
//                                              (fused, the default) the evaluation of the next iteration
// The step and trial kernels stream their row-like inputs through per-thread cp.async rings in shared
// memory (Inst::RING_*); the trial kernel also parks the bound multipliers and stage inputs there.
//   ctrl_trial  thread per instance            filter acceptance test; accepted instances flip their
//                                              buffer and join the next round's eval list, rejected
//                                              ones halve alpha and join the next round's trial list
//
// The stage-parallel kernels expose (N+1) x more threads than instances and carry no sequential
// dependency; the sequential part (the Riccati sweep) is isolated in one kernel.  Instances leave
// the lists when they finish, so the tail of slow instances costs launch latency, not idle lanes.
// Algorithm = oracle/mmpc_oracle.c (IPOPT-style: monotone mu, tau = max(0.99, 1-mu), bound push,
// gradient-based objective scaling, filter line search with slack reset).
//
// The phase bodies are plain __device__ functions of (params, instance, stage) so that tests/emu can
// run the same source on the CPU (g++), one work item at a time.
#pragma once
#include <stdint.h>
#include "../../include/mmpc.h"
#include "mmpc_model.cuh"
#include "mmpc_warp.cuh"
#include "mmpc_ipm.cuh"

namespace mmpc {

// log2 of the lane stride of the per-instance workspace: 32 instances interleaved in HBM (tile-major layout, see Inst), one
// instance alone in shared memory in the resident build
#ifdef MMPC_RESIDENT
constexpr int LSH = 0;
#else
constexpr int LSH = 5;
#endif

// The caller's arrays of one mmpc_solve call (include/mmpc.h: MmpcBatchIn / MmpcBatchOut).  They live in a small block of
// DEVICE memory that a one-thread kernel rewrites in front of every solve, so the kernel parameters -- and with them the
// CUDA graph of the solve (mmpc_api.cu) -- do not change when the caller passes other buffers.  Only init() reads the
// inputs and only finish() / finish_stage() / ctrl_step() write the outputs.
struct SIO {
  const double *x_init, *x_ref, *u_ref, *u_last, *u_guess, *circles, *planes;
  const int32_t* n_pl_inst;
  const uint8_t* flags;
  const double* x_guess;
  double *U, *X, *s, *cost, *kkt;
  int32_t *iters, *status;
  int B;           // instances of this call
};

struct SParams {
  MmpcConfig cfg;
  int B;           // capacity the launch configurations were chosen for (>= io->B)
  const SIO* io;   // device memory (host memory under tests/emu)
  double* ws;      // per-stage records, field-major / instance-minor
  double* qp;      // stage QP records      qp[(k*LS + b)*QS + f]
  double* rk;      // Riccati records       rk[(k*LS + b)*RS + f]
  double* gd;      // per-instance doubles  gd[field * LS + b]
  int* gi;         // per-instance ints     gi[field * LS + b]
  int* lists;      // 3 lists of LS ints: E (next phase solve), T0 / T1 (next phase trial, ping-pong over rounds)
  int* cnt;        // their lengths
  int tsel;        // which T list the trial kernels of this round read (1 or 2)
  long long LS;    // B_max rounded up to a multiple of 32 (whole tiles)
  int ND;          // per-instance doubles (staged_inst_doubles)
  int R, STG, ITSZ;
  int parts;       // 1: step and trial run as warp-specialised parts (mmpc_parts.cuh); needs fused
  int fused;       // 1: the trial kernel also evaluates the next iteration's derivatives (no eval kernel per round)
  int team;        // 1: 16-lane column-parallel Riccati (mmpc_team.cuh), 0: one thread per instance
};

// What the resident TAIL kernel (mmpc_resident.cu) needs of the staged solver's state in HBM: once the active set of a staged
// solve has thinned out to what the resident kernel holds in flight, the remaining instances are loaded into shared memory
// and finished there (mmpc_api.cu, graph_build).
struct ResTail {
  const double *ws, *qp, *gd;   // tile-major workspace, stage-QP records, per-instance doubles (SParams::ws / qp / gd)
  const int* gi;                // per-instance ints
  const int* list;              // the trial list of the last round: every instance still active (and the ones that just finished)
  const int* cnt;               // list lengths; cnt[2] is this list's
  unsigned* queue;              // work counter (zeroed in front of every solve)
  long long LS;
  int STG, ITSZ, ND;            // the staged layout's stage stride, iterate size, per-instance doubles
};

// ---- iterate buffer (two copies, ping-pong): offsets inside one copy --------------------------
constexpr int I_X = 0, I_U = 9, I_S = 14, I_LAM = 15, I_ZXL = 24, I_ZXU = 33, I_ZUL = 42, I_ZUU = 47, I_T = 52;
// ---- after the two iterate copies (field-major, like the iterate) -----------------------------------
constexpr int S_DX = 0, S_DU = 9, S_DS = 14, S_LAMN = 15, S_FK = 24, S_DFC = 32, S_PART = 41;  // 12 partial slots
constexpr int IN_XREF = 53, IN_UREF = 62, IN_ULAST = 67, IN_ULO = 72, IN_UHI = 77;
constexpr int S_DT = 82;  // dt[R], then (moving obstacles) circles[3*nobs]
constexpr int S_FIXED = 82;
constexpr int PT_MERIT = 8;  // partial slots 0..7: KKT parts (eval) or step parts; 8..11: merit of the trial point
// ---- stage QP record, contiguous per (stage, instance):  qp[(k*LS + b)*QS + f] ------------------------
// pose Hessian (21 packed) | velocity diagonal 3 | (dx,dpsi) (dy,dpsi) | control diagonal 5 | (psi,u0) |
// a = H[pose][s] 6 | c = H[s][s] | bv = H[pose][v] 6 | hvv | gA 16 | gB 16   (y = x9 s u5 v)
constexpr int Q_HP = 0, Q_HVD = 21, Q_H35 = 24, Q_H45 = 25, Q_HUU = 26, Q_HPU = 31, Q_A = 32, Q_C = 38,
              Q_BV = 39, Q_HVV = 45, Q_GA = 46, Q_GB = 62, QS = 80;
// ---- Riccati record, contiguous per (stage, instance):  rk[(k*LS + b)*RS + f] -------------------------
// K 45 | kff 5 | w 14 | cv | l0 | Pxx 45 (packed) | pxx 9
constexpr int R_K = 0, R_KFF = 45, R_W = 50, R_CV = 64, R_L0 = 65, R_P = 66, R_PV = 111, RS = 120;
constexpr int SGY_S = 9, SGY_U = 10, SGY_V = 15;
// per-instance doubles
constexpr int D_MU = 0, D_REGLAST = 1, D_THMAX = 2, D_THMIN = 3, D_E0 = 4, D_OS = 5, D_ALPHA = 6, D_AD = 7,
              D_GPHI = 8, D_THETA = 9, D_PHI0 = 10, D_FILT = 11, D_AP0 = 43, D_PL = 44, D_CIRC = 44 + 6 * MMPC_MAX_PLANES;
// per-instance ints.  J_FRST: IPOPT's filter reset heuristic (filter_reset_trigger 5, max_filter_resets 5), packed:
// bits 0-7 successive iterations whose last rejection was the filter's, bits 8-15 resets so far, bit 16 the last rejection of
// the running line search was the filter's
constexpr int J_STATE = 0, J_IT = 1, J_NFILT = 2, J_LS = 3, J_CUR = 4, J_NPL = 5, J_STATUS = 6, J_FLAGS = 7, J_FRST = 8,
              J_REGF = 9 /* 1: the factorisation of the last iteration needed delta_w > 0 */, J_NFIELDS = 10;
constexpr int FILTER_RESET_TRIGGER = 5, MAX_FILTER_RESETS = 5;
constexpr int ST_ACTIVE = 0, ST_DONE = 1, ST_TRIAL = 2, ST_FINISH = 3;  // FINISH: results are written by the step kernels of this round
//  // ACTIVE: next phase is eval; TRIAL: next phase is a trial
// partial slots
constexpr int NPART = 8;
// Terminal equality  (x_N, y_N) = (xref_N[0], xref_N[1])  (flags bit 0; interface_wholebody_qref.py:167):
// multipliers nu[2] with the regularised Newton step  d nu = (dx_N + c) / delta_c  (IPOPT's delta_c), i.e.
// the terminal Hessian gets 1/delta_c on x and y.  nu lives in the otherwise unused lam slots of stage 0
// (x_0 is fixed, it has no costate), its Newton target in stage 0's lam+ slots.
#define MMPC_DELTA_C 1e-6

// rows per stage: circles | 4 self-collision | 6 proper plane rows | (MMPC_MODE_REFERENCE) 6 x (n_pl - 1) rows with
// stale plane columns, row (i, j) at index n_obs + 10 + i * (n_pl - 1) + j
__host__ __device__ inline int staged_stale_rows(const MmpcConfig& c) {
  return (c.mode == MMPC_MODE_REFERENCE && c.n_pl > 1) ? 6 * (c.n_pl - 1) : 0;
}
// self-collision rows of a stage: 4; none in the base-only model (controllers/mpc_base.py has no arm) and in the pose-reference
// model (controllers/mpc_wholebody.py:100 "onstacles 3D: TODO")
__host__ __device__ inline int staged_self_rows(const MmpcConfig& c) { return c.model == MMPC_MODEL_WHOLEBODY ? 4 : 0; }
__host__ __device__ inline int staged_rows(const MmpcConfig& c) { return c.n_obs + staged_self_rows(c) + (c.n_pl > 0 ? 6 : 0) + staged_stale_rows(c); }
__host__ __device__ inline int staged_itsz(const MmpcConfig& c) { return I_T + 2 * staged_rows(c); }
// MMPC_MODE_REFERENCE: the plane margins c[i][j] of the stage's six body points (6 x n_pl, i-major), written once per
// evaluated point by the pose kernel and read by the three threads (stages k-1, k, k+1) whose stale-column rows need them
__host__ __device__ inline int staged_marg_doubles(const MmpcConfig& c) { return staged_stale_rows(c) > 0 ? 6 * c.n_pl : 0; }
__host__ __device__ inline int staged_stage_doubles(const MmpcConfig& c) {
  return 2 * staged_itsz(c) + S_FIXED + staged_rows(c) + (c.obs_per_stage ? 3 * c.n_obs : 0) + staged_marg_doubles(c);
}
__host__ __device__ inline int staged_inst_doubles(const MmpcConfig& c) { return D_CIRC + (c.obs_per_stage ? 0 : 3 * c.n_obs); }

__host__ __device__ constexpr int ssidx(int i, int j) { return i <= j ? i * 9 - i * (i - 1) / 2 + (j - i) : j * 9 - j * (j - 1) / 2 + (i - j); }
struct SACoef { double dt, a32, a42, a34, a43, a35, a45, cp, sp; };

// step kernel: what is parked in shared memory next to the row ring (A/B): 0 nothing, 1 the 29 stage inputs, 2 also the FK
// cache and the defect (17)
#ifndef MMPC_STEP_PARK
#define MMPC_STEP_PARK 1
#endif
constexpr int STEP_PARK = MMPC_STEP_PARK;
constexpr int STAGED_STEP_PARKED = (STEP_PARK >= 1 ? S_DT - IN_XREF : 0) + (STEP_PARK >= 2 ? S_PART - S_FK : 0);

struct Inst {
  const SParams& P;
  const MmpcConfig& cfg;
  double* w;   // ws + b
  double* gd;  // this instance's lane in its tile of P.gd
  int* gi;     // this instance's lane in its tile of P.gi
  long long LS;
  int N, R, STG, ITSZ, B2, nobs, npl, b;   // b: index into the workspace
  int bio;                                  // index into the caller's arrays (P.io); == b except in the resident build
  double dt;
  // Row ring (step and trial kernels): the inputs of the "row-like" items of a stage -- circle rows with their circle,
  // bound multipliers, self-collision and plane rows -- are streamed through a per-thread ring in shared memory
  // with cp.async, RING_D items ahead of their use.  The items are consumed strictly in order; one commit group
  // per item, so `wait_group RING_D - 1` is "the oldest item has landed".  (Register prefetching does not work
  // here: ptxas puts the prefetch loads on the scoreboard their consumer waits on.)
  static constexpr int RING_D = 4, RING_W = 6;
  static constexpr int RING_DT = 4;  // depth of the trial kernel's circle-row ring
  double* sm = nullptr;  // this thread's lane of the ring: item q, value c at sm[((q % RING_D) * RING_W + c) * bs]
  int bs = 1;            // threads per block
  __device__ __forceinline__ double* ring_slot(int q) const { return sm + ((q & (RING_D - 1)) * RING_W) * bs; }
  __device__ __forceinline__ const double* circ_ptr(int k, int i, int c) const {
    return cfg.obs_per_stage ? &W2(k, S_DT + R + 3 * i + c) : &D(D_CIRC + 3 * i + c);
  }

  __device__ __forceinline__ Inst(const SParams& p, int b_) : P(p), cfg(p.cfg) {
    b = b_; LS = p.LS;
    // tile-major layout: the 32 instances of a tile keep all their fields together, lane-interleaved
    //   ws[((tile*(N+1) + k)*STG + f)*32 + lane]   gd[(tile*ND + f)*32 + lane]   gi[(tile*J_NFIELDS + f)*32 + lane]
    // so a warp of neighbouring instances reads one 256-byte line per field, consecutive fields are
    // consecutive lines (DRAM page locality), and the lane stride is a compile-time constant.
#ifdef MMPC_RESIDENT
    // resident build (mmpc_resident.cu): the workspace is ONE instance in shared memory, fields contiguous per stage; b_ only
    // names the caller's arrays (bio)
    b = 0; bio = b_;
    w = p.ws; gd = p.gd; gi = p.gi; rkp = p.rk; rks = RS;
#else
    bio = b_;
    const long long tile = b_ >> 5; const int ln = b_ & 31;
    w = p.ws + ((tile * (cfg.N + 1) * p.STG) << 5) + ln;
    gd = p.gd + ((tile * p.ND) << 5) + ln;
    gi = p.gi + ((tile * J_NFIELDS) << 5) + ln;
#endif
    N = cfg.N; R = p.R; STG = p.STG; ITSZ = p.ITSZ; B2 = 2 * p.ITSZ; nobs = cfg.n_obs; dt = cfg.dt;
    npl = 0;
    MG = S_DT + R + (cfg.obs_per_stage ? 3 * nobs : 0);
    nself = staged_self_rows(cfg);
  }
  int nself;  // self-collision rows per stage (0 in the base-only model)
  // state error of the cost: component 2 (yaw) of the base-only model goes through angleDiff (controllers/mpc_base.py:129-133)
  __device__ __forceinline__ double xerr(int i, double x, double xr) const {
    return (i == 2 && cfg.model == MMPC_MODEL_BASE) ? angle_diff(x, xr) : x - xr;
  }
  // MMPC_MODEL_POSEREF (controllers/mpc_wholebody.py:79-86, :104-107; compiled into the pose build of the resident kernel only,
  // mmpc_resident_pose.cu): the tracking cost is  e^T diag(W) e  on the end-point pose  e = forward_tranformation(x)[0] - X_ref[k]
  // = (P_e - r_xyz, psi - r_psi), W = Qd[0..3] (Pd at k = N); the states carry no tracking weight of their own.
#ifdef MMPC_POSEREF
  __device__ __forceinline__ bool pose_model() const { return cfg.model == MMPC_MODEL_POSEREF; }
#else
  __device__ __forceinline__ constexpr bool pose_model() const { return false; }
#endif
  __device__ __forceinline__ double xweight(int k, int i) const { return pose_model() ? 0.0 : (k < N ? cfg.Qd[i] : cfg.Pd[i]); }
#ifdef MMPC_POSEREF
  // value of the pose cost of stage k (weights times `scale`), its gradient g wrt the pose (x, y, psi, q1, q2, q3) and -- if H
  // is given -- its exact Hessian added to the packed pose block H[21]:  2 J^T W J + sum_c 2 W_c e_c Hess(P_c)
  __device__ __forceinline__ double pose_cost(int k, double scale, double px, double py, double psi, const FK& f, const double (&r)[4],
                                              double (&g)[NP], double* H) const {
    const double* Wv = k < N ? cfg.Qd : cfg.Pd;
    Point p; point_eval(px, py, f, BODY[5], p);   // the end point
    double n[3], val = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const double e = p.P[c] - r[c], w = scale * Wv[c]; n[c] = 2 * w * e; val += w * e * e; }
    const double ep = psi - r[3], w3 = scale * Wv[3];
    val += w3 * ep * ep;
    point_grad(f, p, n, g);
    g[2] += 2 * w3 * ep;
    if (H) {
      point_hess_acc(f, p, n, 1.0, H);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double u[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
        double gc[NP]; point_grad(f, p, u, gc);
        const double w2 = 2 * scale * Wv[c];
#pragma unroll
        for (int a = 0; a < NP; ++a)
#pragma unroll
          for (int b = a; b < NP; ++b) H[pidx(a, b)] += w2 * gc[a] * gc[b];
      }
      H[pidx(2, 2)] += 2 * w3;
    }
    return val;
  }
#endif
#ifdef MMPC_POSEREF
  // unscaled pose cost of stage k at the iterate `it` (results: sol.value of the objective)
  __device__ __forceinline__ double pose_cost_value(int k, int it) const {
    const double px = W(k, it + I_X + 0), py = W(k, it + I_X + 1), psi = W(k, it + I_X + 2);
    FK f; fk_eval(psi, W(k, it + I_X + 6), W(k, it + I_X + 7), W(k, it + I_X + 8), f);
    const double r[4] = {W2(k, IN_XREF + 0), W2(k, IN_XREF + 1), W2(k, IN_XREF + 2), W2(k, IN_XREF + 3)};
    double g[NP];
    return pose_cost(k, 1.0, px, py, psi, f, r, g, nullptr);
  }
#endif
  int MG;  // offset (second block) of the plane-margin cache of a stage, see staged_marg_doubles
  __device__ __forceinline__ double& W(int k, int o) const { return w[(k * STG + o) << LSH]; }
  __device__ __forceinline__ double& W2(int k, int o) const { return w[(k * STG + B2 + o) << LSH]; }
  // base pointer of a block of fields of stage k (field f of the block is p[f << LSH]): with it the field offsets
  // of the hot kernels are compile-time immediates instead of an index computation per access
  __device__ __forceinline__ double* stage_ptr(int k, int base) const { return w + ((k * STG + base) << LSH); }
  __device__ __forceinline__ double& Qw(int k, int o) const { return P.qp[((long long)k * LS + b) * QS + o]; }
#ifdef MMPC_RESIDENT
  double* rkp;   // the Riccati records this thread works on and their stage stride: the resident kernel keeps two sets
  int rks;       // (speculative delta_w, mmpc_team.cuh), the second one inside the dead copy of the iterate
  __device__ __forceinline__ double& Rw(int k, int o) const { return rkp[k * rks + o]; }
#else
  __device__ __forceinline__ double& Rw(int k, int o) const { return P.rk[((long long)k * LS + b) * RS + o]; }
#endif
  __device__ __forceinline__ double& D(int o) const { return gd[o << LSH]; }
  __device__ __forceinline__ int& J(int o) const { return gi[o << LSH]; }
  template <bool NC = true>
  __device__ __forceinline__ double circ(int k, int i, int c) const {  // read-only after init: non-coherent load
    const double* p = cfg.obs_per_stage ? &W2(k, S_DT + R + 3 * i + c) : &D(D_CIRC + 3 * i + c);
    return NC ? ldg(p) : *p;
  }
  __device__ __forceinline__ void load_npl() { npl = J(J_NPL); }
  // MMPC_MODE_REFERENCE to the letter (controllers/mpc_wholebody_qref.py:263-265): the four terminal
  // self-collision rows  h_m(x_N) <= s[N-1]  (the reference's leaked loop variable), not s[N].  Thread (instance, N) still
  // evaluates them (their pose terms belong to x_N), but with the slack of stage N-1, and their slack-column sums
  //     abar = H[pose(x_N)][s_{N-1}],  cbar = sum sigma,  gbar = -(sum sigma res) - mu sum 1/t
  // go to the v-slots of stage N's record (Q_BV, Q_HVV, Q_GA/GB[SGY_V]: stage N has no v), from where the Riccati sweep
  // folds them into the slack column of stage N-1 (mmpc_team.cuh).  The stationarity residual of s[N-1] is the sum of two
  // threads' terms; they are exchanged through the unused defect slots of stage N (S_DFC + 0 / + 1).
  __device__ __forceinline__ bool q3() const { return cfg.mode == MMPC_MODE_REFERENCE && cfg.terminal_rows_on_sN == 0; }
  // plane data (point, normal) of this instance: written by init only, so the phase kernels read it non-coherently
  // (init itself, which has just written it, passes NC = false)
  template <bool NC = true>
  __device__ __forceinline__ double PL(int o) const { return NC ? ldg(&gd[(D_PL + o) << LSH]) : gd[(D_PL + o) << LSH]; }
  __device__ __forceinline__ bool term_eq(int k) const { return k == N && (J(J_FLAGS) & 1); }
  // L2 prefetch of every field of stage k this thread is going to read (current iterate, step, references,
  // rows): the loads further down then find their lines in L2 instead of paying a full HBM round trip
  // one after the other.  A warp touches one 256-byte line per field, so this is one request per field.
  __device__ __forceinline__ void prefetch_stage(int k, int it, bool with_step) const {
#ifdef MMPC_PREFETCH_L2
    const int n_it = I_T + 2 * R;
    for (int o = 0; o < n_it; ++o) prefetch_l2(&W(k, it + o));
    const int o0 = with_step ? S_DX : S_FK, o1 = S_DT + (with_step ? R : 0) + (cfg.obs_per_stage ? 3 * nobs : 0);
    for (int o = o0; o < S_DFC + 9; ++o) prefetch_l2(&W2(k, o));
    for (int o = IN_XREF; o < o1; ++o) prefetch_l2(&W2(k, o));
    if (with_step && k < N) {
      for (int i = 0; i < NX; ++i) { prefetch_l2(&W(k + 1, it + I_X + i)); prefetch_l2(&W(k + 1, it + I_LAM + i)); }
      for (int i = 0; i < NX; ++i) { prefetch_l2(&W2(k + 1, S_DX + i)); prefetch_l2(&W2(k + 1, S_LAMN + i)); }
    }
#endif
  }

  // -max_j c[i][j] for body point i (:76-87); returns the arg-max plane
  template <bool NC = true>
  __device__ __forceinline__ double plane_row(const Point& p, int& jbest) const {
    double cb = 0; jbest = 0;
    for (int j = 0; j < npl; ++j) {
      double n0 = PL<NC>(6 * j + 3), n1 = PL<NC>(6 * j + 4), n2 = PL<NC>(6 * j + 5);
      double e = cfg.obstacle_expand_dist;
      double off = n0 * (PL<NC>(6 * j + 0) - e * n0) + n1 * (PL<NC>(6 * j + 1) - e * n1) + n2 * (PL<NC>(6 * j + 2) - e * n2);
      double c = off - (n0 * p.P[0] + n1 * p.P[1] + n2 * p.P[2]);
      bool take = (j == 0) || (npl == 2 ? !(cb > c) : (c > cb));  // if_else(c0 > c1, c0, c1) :85 ; mmax :87
      if (take) { cb = c; jbest = j; }
    }
    return -cb;
  }

  // ------------------------------------------------------------------------------------------
  // init (thread per instance): copy the instance into the field-major layout, reference initial
  // guess (:302-304) + IPOPT bound push; s lifted so every row starts strictly feasible;
  // objective scaling.
  __device__ void init() {
    const double* xref = P.io->x_ref + (long long)bio * (N + 1) * NX;
    const double* uref = P.io->u_ref + (long long)bio * N * NU;
    const double* ulast = P.io->u_last + (long long)bio * N * NU;
    npl = P.io->n_pl_inst ? ldg(P.io->n_pl_inst + bio) : cfg.n_pl;
    npl = npl < 0 ? 0 : (npl > cfg.n_pl ? cfg.n_pl : npl);  // a caller's value outside [0, n_pl] must not index past the plane tables
    J(J_NPL) = npl;
    for (int j = 0; j < cfg.n_pl; ++j)
      for (int c = 0; c < 6; ++c) D(D_PL + 6 * j + c) = ldg(P.io->planes + ((long long)bio * cfg.n_pl + j) * 6 + c);
    if (!cfg.obs_per_stage)
      for (int i = 0; i < 3 * nobs; ++i) D(D_CIRC + i) = ldg(P.io->circles + (long long)bio * 3 * nobs + i);
    double gmax = 0;
    const bool refmode = cfg.mode == MMPC_MODE_REFERENCE && npl > 1;
    double cprev[6][MMPC_MAX_PLANES], ccur[6][MMPC_MAX_PLANES];
    for (int k = 0; k <= N; ++k) {
      double x[NX];
      if (cfg.obs_per_stage)
        for (int i = 0; i < 3 * nobs; ++i) W2(k, S_DT + R + i) = ldg(P.io->circles + ((long long)bio * (N + 1) + k) * 3 * nobs + i);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = fmax(fmin(ldg(P.io->x_init + (long long)bio * NX + i), cfg.xlim[1][i]), cfg.xlim[0][i]);  // :290-291
        if (k >= 1 && P.io->x_guess) v = ldg(P.io->x_guess + ((long long)bio * (N + 1) + k) * NX + i);       // mpc_base.py:196-201
        if (k >= 1) v = push_in(v, cfg.xlim[0][i], cfg.xlim[1][i]);
        double xr = ldg(xref + k * NX + i);
        x[i] = v; W(k, I_X + i) = v; W(k, I_LAM + i) = 0; W2(k, IN_XREF + i) = xr;
        W(k, I_ZXL + i) = 1; W(k, I_ZXU + i) = 1;
        double Wx = xweight(k, i);
        if (k >= 1) gmax = fmax(gmax, fabs(2 * Wx * xerr(i, v, xr)));
      }
      if (k < N) {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double ul = ldg(ulast + k * NU + j), ur = ldg(uref + k * NU + j);
          double lo = fmax(cfg.ulim[0][j], ul + cfg.dulim[0][j]);  // mpc_wholebody_qref.py:203 and :205 merged
          double hi = fmin(cfg.ulim[1][j], ul + cfg.dulim[1][j]);
          double v = P.io->u_guess ? ldg(P.io->u_guess + ((long long)bio * N + k) * NU + j) : ul;
          v = push_in(v, lo, hi);
          W(k, I_U + j) = v; W2(k, IN_UREF + j) = ur; W2(k, IN_ULAST + j) = ul; W2(k, IN_ULO + j) = lo; W2(k, IN_UHI + j) = hi;
          W(k, I_ZUL + j) = 1; W(k, I_ZUU + j) = 1;
          gmax = fmax(gmax, fabs(2 * cfg.Rd[j] * (v - ur) + 2 * cfg.Wd[j] * (v - ul)));
        }
      }
      FK f; fk_eval(x[2], x[6], x[7], x[8], f);
      double hmax = -1e300;
      for (int i = 0; i < nobs; ++i) {
        double ddx = x[0] - circ<false>(k, i, 0), ddy = x[1] - circ<false>(k, i, 1);  // written above by this thread
        double h = (circ<false>(k, i, 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy);
        W(k, I_T + i) = h; hmax = fmax(hmax, h);
      }
#pragma unroll 1
      for (int m = 0; m < nself; ++m) {
        Point p; point_eval(x[0], x[1], f, SELFD[m], p);
        double h = cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]);
        W(k, I_T + nobs + m) = h;
        if (!(k == N && q3())) hmax = fmax(hmax, h);  // the terminal rows of the literal reference NLP are bounded by s[N-1]: they do not lift s[N]
      }
      if (npl > 0) {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
          Point p; point_eval(x[0], x[1], f, BODY[i], p);
          int jb; double h = plane_row<false>(p, jb);
          W(k, I_T + nobs + nself + i) = h; hmax = fmax(hmax, h);
        }
      }
      if (staged_stale_rows(cfg) > 0) {  // rows with stale plane columns: slack k >= 1, j < npl - 1
        const int nst = cfg.n_pl - 1, r0 = nobs + nself + 6;
        for (int r = r0; r < R; ++r) W(k, I_T + r) = 0;
        if (refmode) {
          double pp[NP] = {x[0], x[1], x[2], x[6], x[7], x[8]}; FK ff;
          margins(pp, ff, ccur);
          if (k >= 1)
            for (int i = 0; i < 6; ++i)
              for (int j = 0; j < npl - 1; ++j) {
                int jb; double h = stale_max(ccur, cprev, i, j, jb);
                W(k, I_T + r0 + i * nst + j) = h; hmax = fmax(hmax, h);
              }
          for (int i = 0; i < 6; ++i) for (int j = 0; j < npl; ++j) cprev[i][j] = ccur[i][j];
        }
      }
      double s = fmax(0.0, hmax + 1e-2);
      W(k, I_S) = s;
      gmax = fmax(gmax, fabs(2 * cfg.S * s));
      for (int r = 0; r < R; ++r) {
        // (x_N = x_{N-1} in the starting point, so s[N-1] already clears the terminal rows by the same 1e-2)
        const double sr = (k == N && q3() && r >= nobs && r < nobs + nself) ? W(k - 1, I_S) : s;
        W(k, I_T + r) = sr - W(k, I_T + r); W(k, I_T + R + r) = 1.0;
      }
    }
    D(D_OS) = (gmax > 100.0) ? fmax(100.0 / gmax, 1e-8) : 1.0;
    D(D_MU) = cfg.mu_init; D(D_REGLAST) = 0; D(D_THMAX) = -1; D(D_THMIN) = -1; D(D_E0) = 1e300;
    J(J_STATE) = ST_ACTIVE; J(J_IT) = 0; J(J_NFILT) = 0; J(J_LS) = 0; J(J_CUR) = 0; J(J_FRST) = 0; J(J_REGF) = 0;
    J(J_FLAGS) = P.io->flags ? (int)P.io->flags[bio] : 0;
  }

  struct RowAcc {
    double H[21], a[NP], gA[NP], gB[NP], st[NP];
    double csum, be0, be1, zrows, chi, clo, prim, sumz;
    int nz;
  };

  // ------------------------------------------------------------------------------------------
  // MMPC_MODE_REFERENCE: the rows with stale plane columns (controllers/mpc_wholebody_qref.py:76-89 with the
  // subject_to inside the j loop).  Row (m, i, j), m >= 1, i < 6, j < npl - 1:
  //     -max( c_m[i][0..j], c_{m-1}[i][j+1..] ) <= s_m
  // Thread (instance, k) OWNS the rows of slack s_k (their candidate (t, z), the s_k terms, merit and KKT
  // bookkeeping and, when the arg-max column belongs to stage k, their pose terms) and ADDS the pose
  // terms of the rows of slack s_{k+1} whose arg-max column is a stale one, i.e. belongs to x_k: those
  // couple x_k with v_k = s_{k+1} (Q_BV) in the augmented Riccati form.  Kept out of line: it only
  // runs in reference mode and needs forward kinematics at three stages.
  //   phase 0: evaluation of the current iterate   1: trial + evaluation of the candidate   2: step
  // (Quirk 3, the terminal self-collision rows bounded by s_{N-1}, is handled where those rows are evaluated: see q3().)
  struct StaleIO {
    double theta; LogProd lp; bool ok;       // merit ingredients (phase 1, 2); lp: the barrier's log of the product of the slacks
    MinRatio rp, rd; double gphi;            // phase 2
    __device__ __forceinline__ void reset() { theta = 0; lp.init(); ok = true; gphi = 0; }
  };
  __device__ __forceinline__ void pose_at(int kk, int it, bool cand, double alpha, double (&p)[NP]) const {
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      double v = W(kk, it + I_X + POSE2X[a]);
      p[a] = cand ? fma(alpha, W2(kk, S_DX + POSE2X[a]), v) : v;
    }
  }
  __device__ __forceinline__ void xy_at(int kk, int it, bool cand, double alpha, double (&p)[2]) const {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      double v = W(kk, it + I_X + a);
      p[a] = cand ? fma(alpha, W2(kk, S_DX + a), v) : v;
    }
  }
  __device__ __forceinline__ void load_fk(int kk, FK& f) const {
    f.cp = W2(kk, S_FK + 0); f.sp = W2(kk, S_FK + 1);
#pragma unroll
    for (int q = 0; q < 3; ++q) { f.vr[q] = W2(kk, S_FK + 2 + q); f.vh[q] = W2(kk, S_FK + 5 + q); }
  }
  __device__ __forceinline__ void load_margins(int kk, double (&c)[6][MMPC_MAX_PLANES]) const {
#pragma unroll 1
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < npl; ++j) c[i][j] = W2(kk, MG + i * cfg.n_pl + j);
  }
  // pose kernel (thread per instance and stage, MMPC_MODE_REFERENCE only): forward kinematics and plane margins of stage k
  // at the candidate  w + alpha d  (cand) or at the current iterate, into the stage's FK cache and margin cache.  Runs in
  // front of the evaluation of the starting point and in front of every trial.
  __device__ void pose_pass(int k, bool cand) {
    load_npl();
    const int it = J(J_CUR) * ITSZ;
    const double alpha = cand ? D(D_ALPHA) : 0.0;
    double p[NP]; pose_at(k, it, cand, alpha, p);
    FK f;
    double c[6][MMPC_MAX_PLANES];
    if (npl >= 2) {
      margins(p, f, c);
#pragma unroll 1
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < npl; ++j) W2(k, MG + i * cfg.n_pl + j) = c[i][j];
    } else fk_eval(p[2], p[3], p[4], p[5], f);
    W2(k, S_FK + 0) = f.cp; W2(k, S_FK + 1) = f.sp;
#pragma unroll
    for (int q = 0; q < 3; ++q) { W2(k, S_FK + 2 + q) = f.vr[q]; W2(k, S_FK + 5 + q) = f.vh[q]; }
  }
  // plane margins c[i][j] of the six body points at one pose (:78-80)
  __device__ __forceinline__ void margins(const double (&p)[NP], FK& f, double (&c)[6][MMPC_MAX_PLANES]) const {
    fk_eval(p[2], p[3], p[4], p[5], f);
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
      Point pt; point_eval(p[0], p[1], f, BODY[i], pt);
      for (int j = 0; j < npl; ++j) {
        double n0 = D(D_PL + 6 * j + 3), n1 = D(D_PL + 6 * j + 4), n2 = D(D_PL + 6 * j + 5), e = cfg.obstacle_expand_dist;
        double off = n0 * (D(D_PL + 6 * j + 0) - e * n0) + n1 * (D(D_PL + 6 * j + 1) - e * n1) + n2 * (D(D_PL + 6 * j + 2) - e * n2);
        c[i][j] = off - (n0 * pt.P[0] + n1 * pt.P[1] + n2 * pt.P[2]);
      }
    }
  }
  // arg-max column of row (i, j) built from the margins of stage m (columns <= j) and m-1 (columns > j)
  __device__ __forceinline__ double stale_max(const double (&cm)[6][MMPC_MAX_PLANES], const double (&cm1)[6][MMPC_MAX_PLANES],
                                              int i, int j, int& best) const {
    double cb = 0; best = 0;
    for (int jj = 0; jj < npl; ++jj) {
      double c = jj <= j ? cm[i][jj] : cm1[i][jj];
      bool take = (jj == 0) || (npl == 2 ? !(cb > c) : (c > cb));  // if_else(c0 > c1, c0, c1) :85 ; mmax :87
      if (take) { cb = c; best = jj; }
    }
    return -cb;
  }
  // the same for ONE body point whose margins sit in registers (fully unrolled over the plane slots)
  __device__ __forceinline__ double stale_max1(const double (&cm)[MMPC_MAX_PLANES], const double (&cm1)[MMPC_MAX_PLANES], int j, int& best) const {
    double cb = 0; best = 0;
#pragma unroll
    for (int jj = 0; jj < MMPC_MAX_PLANES; ++jj) {
      const double c = jj <= j ? cm[jj] : cm1[jj];
      const bool take = jj < npl && ((jj == 0) || (npl == 2 ? !(cb > c) : (c > cb)));
      if (take) { cb = c; best = jj; }
    }
    return -cb;
  }
  // Where stale_rows() reads the margin caches of the stages k-1, k, k+1 (st = 0, 1, 2) from: the workspace, or a copy the
  // step / trial kernel has prefetched into its shared-memory slots (margins_prefetch).
  const double* mg_b[3] = {nullptr, nullptr, nullptr};
  int mg_s = 1;
  __device__ __forceinline__ void margins_in_place(int k) {
#pragma unroll
    for (int st = 0; st < 3; ++st) { const int kk = k - 1 + st; mg_b[st] = (kk >= 0 && kk <= N) ? stage_ptr(kk, B2 + MG) : nullptr; }
    mg_s = 1 << LSH;
  }
  // cp.async of the margin caches of NST stages from k-1 on into per-thread slots at dst (6 n_pl slots per stage, stride
  // bs); one commit group.  The caller waits for it (async_wait) before stale_rows().
  template <int NST>
  __device__ __forceinline__ void margins_prefetch(int k, double* dst) {
    const int nm = 6 * cfg.n_pl;
    margins_in_place(k);
#pragma unroll
    for (int st = 0; st < NST; ++st) {   // (unrolled: mg_b stays in registers)
      const double* src = mg_b[st];
      double* d = dst + st * nm * bs;
      if (src) {
#pragma unroll 1
        for (int f = 0; f < nm; ++f) async_copy8(d + f * bs, src + (f << LSH));
        mg_b[st] = d;
      }
    }
    async_commit();
    mg_s = bs;
  }
  __device__ __forceinline__ void load_margins1(int st, int i, double (&c)[MMPC_MAX_PLANES]) const {
    const double* p = mg_b[st] + (i * cfg.n_pl) * mg_s;
#pragma unroll
    for (int j = 0; j < MMPC_MAX_PLANES; ++j) c[j] = j < npl ? p[j * mg_s] : 0.0;
  }
  __device__ __forceinline__ static void pick_fk(bool first, const FK& a, const FK& b, FK& o) {
    o.cp = first ? a.cp : b.cp; o.sp = first ? a.sp : b.sp;
#pragma unroll
    for (int q = 0; q < 3; ++q) { o.vr[q] = first ? a.vr[q] : b.vr[q]; o.vh[q] = first ? a.vh[q] : b.vh[q]; }
  }
  // PHASE 0: evaluation of the current iterate   1: trial + evaluation of the candidate   2: step.
  // Inlined into its three call sites with the phase as a compile-time constant: the accumulators stay in registers
  // (out of line they lived in local memory, and every H += was a load and a store).
  // A, bv: the caller's own accumulators of stage k (phases 0 and 1)
  template <int PHASE>
  __device__ __forceinline__ void stale_rows(int k, RowAcc& A, double (&bv)[NP], StaleIO& io) const {
    if (npl < 2) return;
    const int it = J(J_CUR) * ITSZ, jt = (1 - J(J_CUR)) * ITSZ;
    constexpr bool cand = PHASE == 1;
    const double mu = D(D_MU), alpha = cand ? D(D_ALPHA) : 0.0, ad = cand ? D(D_AD) : 0.0;
    const int nst = cfg.n_pl - 1, r0 = nobs + nself + 6;
    // Margins and forward kinematics of the stages k-1, k, k+1 at the point this phase works on (the candidate in phase 1,
    // the current iterate otherwise): computed ONCE per stage by pose_pass() in the kernel launched just before, read here,
    // one body point at a time.  The owner of a row and the neighbour that adds its pose terms see the same stored
    // numbers, so they always agree on the arg-max column.
    double pk[2], pm[2] = {0, 0};
    FK fk, fm;
    xy_at(k, it, cand, alpha, pk); load_fk(k, fk);
    if (k >= 1) { xy_at(k - 1, it, cand, alpha, pm); load_fk(k - 1, fm); } else fm = fk;
    const double s_k = cand ? fma(alpha, W2(k, S_DS), W(k, it + I_S)) : W(k, it + I_S);
    // ---- rows of slack s_k (owned) ----
    if (k >= 1) {
      double dpk[NP], dpm[NP], dsk = 0;
      if (PHASE == 2) {
#pragma unroll
        for (int a = 0; a < NP; ++a) { dpk[a] = W2(k, S_DX + POSE2X[a]); dpm[a] = W2(k - 1, S_DX + POSE2X[a]); }
        dsk = W2(k, S_DS);
      }
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        double ck[MMPC_MAX_PLANES], cm[MMPC_MAX_PLANES];
        load_margins1(1, i, ck); load_margins1(0, i, cm);
#pragma unroll 1
        for (int j = 0; j < npl - 1; ++j) {
          const int r = r0 + i * nst + j;
          int jb; const double h = stale_max1(ck, cm, j, jb);
          const bool here = jb <= j;  // the arg-max column belongs to stage k
          FK ff; pick_fk(here, fk, fm, ff);
          Point pt; point_eval(here ? pk[0] : pm[0], here ? pk[1] : pm[1], ff, BODY[i], pt);
          const double n[3] = {PL(6 * jb + 3), PL(6 * jb + 4), PL(6 * jb + 5)};
          double g[NP]; point_grad(ff, pt, n, g);  // grad h = +g (at the arg-max stage)
          double t = W(k, it + I_T + r), z = W(k, it + I_T + R + r);
          if (PHASE == 2) {
            double gd_ = 0;
#pragma unroll
            for (int a = 0; a < NP; ++a) gd_ = fma(g[a], here ? dpk[a] : dpm[a], gd_);
            double res = h - s_k + t, dtv = -res - (gd_ - dsk);
            W2(k, S_DT + r) = dtv;
            double itv = rcp(t), dz = (mu - z * (t + dtv)) * itv;
            io.theta += fabs(res); io.gphi -= mu * dtv * itv; io.lp.mul(t);
            if (dtv < 0) io.rp.add(t, -dtv);
            if (dz < 0) io.rd.add(z, -dz);
            continue;
          }
          double it_;
          if (cand) {
            double dtv = W2(k, S_DT + r);
            double tt = fmax(fma(alpha, dtv, t), s_k - h);
            double dz = (mu - z * (t + dtv)) * rcp(t);
            it_ = rcp(tt);
            z = zclamp(z + ad * dz, mu, it_);
            W(k, jt + I_T + r) = tt; W(k, jt + I_T + R + r) = z;
            t = tt;
            if (tt <= 0) io.ok = false; else io.lp.mul(tt);
          } else it_ = 1.0 / t;
          const double res = h - s_k + t;
          io.theta += fabs(res);
          A.prim = fmax(A.prim, fabs(res));
          const double zt = z * t;
          A.chi = fmax(A.chi, zt); A.clo = fmin(A.clo, zt); A.sumz += z; A.zrows += z; A.nz++;
          const double sig = z * it_;
          A.csum += sig; A.be0 += sig * res; A.be1 += it_;
          if (here) {
#pragma unroll
            for (int a = 0; a < NP; ++a)
#pragma unroll
              for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += sig * g[a] * g[c];
            point_hess_acc(ff, pt, n, z, A.H);
            const double cb = sig * res;
#pragma unroll
            for (int a = 0; a < NP; ++a) { A.a[a] -= sig * g[a]; A.gA[a] += cb * g[a]; A.gB[a] += it_ * g[a]; A.st[a] += z * g[a]; }
          }
        }
      }
    }
    // ---- rows of slack s_{k+1} whose arg-max column belongs to x_k: pose terms and H[pose][v] of stage k ----
    if (k < N && PHASE != 2) {
      const double s_n = cand ? fma(alpha, W2(k + 1, S_DS), W(k + 1, it + I_S)) : W(k + 1, it + I_S);
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        double ck[MMPC_MAX_PLANES], cn[MMPC_MAX_PLANES];
        load_margins1(1, i, ck); load_margins1(2, i, cn);
#pragma unroll 1
        for (int j = 0; j < npl - 1; ++j) {
          const int r = r0 + i * nst + j;
          int jb; const double h = stale_max1(cn, ck, j, jb);
          if (jb <= j) continue;  // belongs to stage k+1: its owner handles it
          Point pt; point_eval(pk[0], pk[1], fk, BODY[i], pt);
          const double n[3] = {PL(6 * jb + 3), PL(6 * jb + 4), PL(6 * jb + 5)};
          double g[NP]; point_grad(fk, pt, n, g);
          double t = W(k + 1, it + I_T + r), z = W(k + 1, it + I_T + R + r), it_;
          if (cand) {  // the owner's candidate, recomputed with the same arithmetic
            double dtv = W2(k + 1, S_DT + r);
            double tt = fmax(fma(alpha, dtv, t), s_n - h);
            double dz = (mu - z * (t + dtv)) * rcp(t);
            it_ = rcp(tt);
            z = zclamp(z + ad * dz, mu, it_);
            t = tt;
          } else it_ = 1.0 / t;
          const double res = h - s_n + t, sig = z * it_, cb = sig * res;
#pragma unroll
          for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int c = a; c < NP; ++c) A.H[pidx(a, c)] += sig * g[a] * g[c];
          point_hess_acc(fk, pt, n, z, A.H);
#pragma unroll
          for (int a = 0; a < NP; ++a) { bv[a] -= sig * g[a]; A.gA[a] += cb * g[a]; A.gB[a] += it_ * g[a]; A.st[a] += z * g[a]; }
        }
      }
    }
  }

  // bookkeeping of one slack row  h - s + t = 0  with multiplier z
  __device__ __forceinline__ void row_state(int it, int r, int k, double h, double s, double& z, double& it_, double& res, RowAcc& A) const {
    double t = W(k, it + I_T + r);
    z = W(k, it + I_T + R + r);
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
  // eval (thread per instance and stage): full evaluation of stage k at the current iterate.
  template <bool REF>
  __device__ void eval(int k) {
    load_npl();
    const int it = J(J_CUR) * ITSZ;
    const double os = D(D_OS);
    RowAcc A;
    double x[NX], u[NU], lam[NX], lam1[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { x[i] = W(k, it + I_X + i); lam[i] = (k >= 1) ? W(k, it + I_LAM + i) : 0.0; }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = (k < N) ? W(k, it + I_U + j) : 0.0;
    double s = W(k, it + I_S);
    FK f;
    if (REF) load_fk(k, f);  // the pose kernel has just evaluated it
    else {
      fk_eval(x[2], x[6], x[7], x[8], f);
      W2(k, S_FK + 0) = f.cp; W2(k, S_FK + 1) = f.sp;
#pragma unroll
      for (int q = 0; q < 3; ++q) { W2(k, S_FK + 2 + q) = f.vr[q]; W2(k, S_FK + 5 + q) = f.vh[q]; }
    }
    A.chi = -1e300; A.clo = 1e300; A.prim = 0; A.sumz = 0; A.zrows = 0; A.nz = 0; A.csum = A.be0 = A.be1 = 0;
#pragma unroll
    for (int e = 0; e < 21; ++e) A.H[e] = 0;
#pragma unroll
    for (int a = 0; a < NP; ++a) A.a[a] = A.gA[a] = A.gB[a] = A.st[a] = 0;
    double es = 0;  // stationarity inf-norm of this stage
    double hpp = 0, sum_lam = 0;
    int n_eq = 0;
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
        lam1[i] = W(k + 1, it + I_LAM + i);
        double d = xn[i] - W(k + 1, it + I_X + i);
        W2(k, S_DFC + i) = d; A.prim = fmax(A.prim, fabs(d)); sum_lam += fabs(lam1[i]);
      }
      n_eq = NX;
      hpp = -dt * u[0] * (lam1[3] * f.cp + lam1[4] * f.sp);
      Qw(k, Q_HPU) = dt * (-lam1[3] * f.sp + lam1[4] * f.cp);
      Qw(k, Q_H45) = -dt * lam1[3];  // (dy,dpsi)
      Qw(k, Q_H35) = dt * lam1[4];   // (dx,dpsi)
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
      Qw(k, Q_HPU) = 0; Qw(k, Q_H45) = 0; Qw(k, Q_H35) = 0;
    }
    const bool teq = term_eq(k);
    double nu[2] = {0, 0};
    if (teq) { nu[0] = W(0, it + I_LAM + 0); nu[1] = W(0, it + I_LAM + 1); }
    // cost and boxes -- :192-205, :240-245.  Pose components seed the row accumulators, the
    // others are final here.
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double Wx = os * xweight(k, i);
      double gr = 2 * Wx * xerr(i, x[i], W2(k, IN_XREF + i));
      double Hd = 2 * Wx, gA = gr, gB = 0, st = gr + stx[i] - lam[i];
      if (k >= 1) {
        double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
        if (is_fin(lo)) {
          double z = W(k, it + I_ZXL + i), d = x[i] - lo, id = 1.0 / d;
          Hd += z * id; gB -= id; st -= z;
          A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
        }
        if (is_fin(hi)) {
          double z = W(k, it + I_ZXU + i), d = hi - x[i], id = 1.0 / d;
          Hd += z * id; gB += id; st += z;
          A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
        }
      }
      if (i < 2 && teq) {  // terminal equality row  x_N[i] - xref_N[i] = 0  with multiplier nu[i]
        double cq = x[i] - W2(k, IN_XREF + i);
        Hd += 1.0 / MMPC_DELTA_C; gA += nu[i] + cq / MMPC_DELTA_C; st += nu[i];
        A.prim = fmax(A.prim, fabs(cq)); sum_lam += fabs(nu[i]); n_eq += 1;
      }
      if (i < 3 || i >= 6) {
        const int a = (i < 3) ? i : i - 3;
        A.H[pidx(a, a)] = Hd + (i == 2 ? hpp : 0.0); A.gA[a] = gA; A.gB[a] = gB; A.st[a] = st;
      } else {
        Qw(k, Q_HVD + (i - 3)) = Hd; Qw(k, Q_GA + i) = gA; Qw(k, Q_GB + i) = gB;
        if (k >= 1) es = fmax(es, fabs(st));
      }
    }
#ifdef MMPC_POSEREF
    if (pose_model()) {  // the end-point pose cost: gradient and exact Hessian into the pose block
      const double r[4] = {W2(k, IN_XREF + 0), W2(k, IN_XREF + 1), W2(k, IN_XREF + 2), W2(k, IN_XREF + 3)};
      double g[NP];
      pose_cost(k, os, x[0], x[1], x[2], f, r, g, A.H);
#pragma unroll
      for (int a = 0; a < NP; ++a) { A.gA[a] += g[a]; A.st[a] += g[a]; }
    }
#endif
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double Hd = 0, gA = 0, gB = 0;
      if (k < N) {
        double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
        double gr = 2 * Rj * (u[j] - W2(k, IN_UREF + j)) + 2 * Wj * (u[j] - W2(k, IN_ULAST + j));
        Hd = 2 * Rj + 2 * Wj; gA = gr;
        double st = gr + stu[j];
        double lo = W2(k, IN_ULO + j), hi = W2(k, IN_UHI + j);
        if (is_fin(lo)) {
          double z = W(k, it + I_ZUL + j), d = u[j] - lo, id = 1.0 / d;
          Hd += z * id; gB -= id; st -= z;
          A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
        }
        if (is_fin(hi)) {
          double z = W(k, it + I_ZUU + j), d = hi - u[j], id = 1.0 / d;
          Hd += z * id; gB += id; st += z;
          A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
        }
        es = fmax(es, fabs(st));
      }
      Qw(k, Q_HUU + j) = Hd; Qw(k, Q_GA + SGY_U + j) = gA; Qw(k, Q_GB + SGY_U + j) = gB;
    }
    // inequality rows with slack:  h(x_k) - s_k + t = 0
    for (int i = 0; i < nobs; ++i) {  // obsAvoid :49-54
      double ddx = x[0] - circ(k, i, 0), ddy = x[1] - circ(k, i, 1);
      double d2 = ddx * ddx + ddy * ddy, inv = rsq(d2), d = d2 * inv;
      double h = (circ(k, i, 2) + cfg.base_radius) - d;
      double z, it_, res; row_state(it, i, k, h, s, z, it_, res, A);
      double sig = z * it_, nx = ddx * inv, ny = ddy * inv, zd = z * inv;
      A.H[pidx(0, 0)] += sig * nx * nx - zd * (1 - nx * nx);
      A.H[pidx(0, 1)] += (sig + zd) * nx * ny;
      A.H[pidx(1, 1)] += sig * ny * ny - zd * (1 - ny * ny);
      double cb = sig * res;
      A.a[0] += sig * nx; A.a[1] += sig * ny; A.gA[0] -= cb * nx; A.gA[1] -= cb * ny;
      A.gB[0] -= it_ * nx; A.gB[1] -= it_ * ny; A.st[0] -= z * nx; A.st[1] -= z * ny;
    }
    const bool q3n = REF && k == N && q3();  // terminal self-collision rows on s[N-1]
    double s_self = s, sv_c = 0, sv_b0 = 0, sv_b1 = 0, sv_z = 0, sv_a[NP] = {0, 0, 0, 0, 0, 0};
    if (q3n) {
      s_self = W(N - 1, it + I_S);
      sv_c = A.csum; sv_b0 = A.be0; sv_b1 = A.be1; sv_z = A.zrows; A.csum = A.be0 = A.be1 = A.zrows = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) { sv_a[a] = A.a[a]; A.a[a] = 0; }
    }
#pragma unroll 1
    for (int m = 0; m < nself; ++m) {  // self collision :219-222
      Point p; point_eval(x[0], x[1], f, SELFD[m], p);
      double d2 = p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2], inv = rsq(d2);
      double h = cfg.self_collision_radius - d2 * inv;
      double z, it_, res; row_state(it, nobs + m, k, h, s_self, z, it_, res, A);
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
    double q3_c = 0, q3_b0 = 0, q3_b1 = 0, q3_z = 0, q3_a[NP] = {0, 0, 0, 0, 0, 0};
    if (q3n) {  // the sums of the four rows are the slack column they add to stage N-1; stage N's own column goes on without them
      q3_c = A.csum; q3_b0 = A.be0; q3_b1 = A.be1; q3_z = A.zrows; A.csum = sv_c; A.be0 = sv_b0; A.be1 = sv_b1; A.zrows = sv_z;
#pragma unroll
      for (int a = 0; a < NP; ++a) { q3_a[a] = A.a[a]; A.a[a] = sv_a[a]; }
    }
    if (npl > 0) {
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {  // obsAvoidConvex :57-89 (proper row)
        Point p; point_eval(x[0], x[1], f, BODY[i], p);
        int jb; double h = plane_row(p, jb);
        double z, it_, res; row_state(it, nobs + nself + i, k, h, s, z, it_, res, A);
        double sig = z * it_;
        double n[3] = {PL(6 * jb + 3), PL(6 * jb + 4), PL(6 * jb + 5)}, g[NP];
        point_grad(f, p, n, g);  // the row is  -max c <= s, c = off - n.P  =>  grad h = +g
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
    double bv[NP] = {0, 0, 0, 0, 0, 0};
    if (REF) {  // compiled out of the clean-mode kernels
      StaleIO io; io.reset(); io.rp.init(); io.rd.init();
      margins_in_place(k);
      stale_rows<0>(k, A, bv, io);
    }
    // slack column of the stage Hessian: H[s][s] = 2S + sum sigma, H[pose][s] = -sum sigma grad h
    double S2 = 2 * os * cfg.S;
    Qw(k, Q_C) = S2 + A.csum;
    Qw(k, Q_GA + SGY_S) = S2 * s - A.be0;
    Qw(k, Q_GB + SGY_S) = -A.be1;
    Qw(k, Q_HVV) = q3_c; Qw(k, Q_GA + SGY_V) = -q3_b0; Qw(k, Q_GB + SGY_V) = -q3_b1;
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      Qw(k, Q_A + a) = A.a[a]; Qw(k, Q_BV + a) = q3n ? q3_a[a] : bv[a];
      Qw(k, Q_GA + POSE2X[a]) = A.gA[a]; Qw(k, Q_GB + POSE2X[a]) = A.gB[a];
      if (k >= 1) es = fmax(es, fabs(A.st[a]));
    }
#pragma unroll
    for (int e = 0; e < 21; ++e) Qw(k, Q_HP + e) = A.H[e];
    if (REF && q3() && k >= N - 1) {  // d/ds[N-1] of the Lagrangian: this thread's share, summed in the solve kernel
      if (k == N) { W2(N, S_DFC + 0) = q3_z; es = fmax(es, fabs(S2 * s - A.zrows)); }
      else W2(N, S_DFC + 1) = S2 * s - A.zrows;
    } else es = fmax(es, fabs(S2 * s - A.zrows));
    W2(k, S_PART + 0) = es; W2(k, S_PART + 1) = A.prim; W2(k, S_PART + 2) = A.chi; W2(k, S_PART + 3) = A.clo;
    W2(k, S_PART + 4) = sum_lam; W2(k, S_PART + 5) = A.sumz; W2(k, S_PART + 6) = (double)A.nz; W2(k, S_PART + 7) = (double)n_eq;
  }

  __device__ __forceinline__ SACoef acoef(int k, int it) const {
    SACoef c; c.dt = dt; c.cp = W2(k, S_FK + 0); c.sp = W2(k, S_FK + 1);
    double u0 = W(k, it + I_U + 0), x3 = W(k, it + I_X + 3), x4 = W(k, it + I_X + 4), x5 = W(k, it + I_X + 5);
    c.a32 = -dt * u0 * c.sp; c.a42 = dt * u0 * c.cp; c.a34 = -dt * x5; c.a43 = dt * x5; c.a35 = -dt * x4; c.a45 = dt * x3;
    return c;
  }
  // v <- A^T v  (in place on a 9-vector)
  __device__ __forceinline__ static void at_mul(double* v, const SACoef& c) {
    double v2 = v[2] + c.a32 * v[3] + c.a42 * v[4];
    double v3 = v[3] + c.dt * v[0] + c.a43 * v[4];
    double v4 = v[4] + c.dt * v[1] + c.a34 * v[3];
    double v5 = v[5] + c.dt * v[2] + c.a35 * v[3] + c.a45 * v[4];
    v[2] = v2; v[3] = v3; v[4] = v4; v[5] = v5;
  }
  // o <- B^T v
  __device__ __forceinline__ static void bt_mul(const double* v, double* o, const SACoef& c) {
    o[0] = c.dt * (c.cp * v[3] + c.sp * v[4]); o[1] = c.dt * v[5];
    o[2] = c.dt * v[6]; o[3] = c.dt * v[7]; o[4] = c.dt * v[8];
  }

  // ------------------------------------------------------------------------------------------
  // Riccati backward recursion in registers.  Stage k eliminates v_k = s_{k+1} (scalar pivot cv)
  // and u_k (5x5 LDL^T); reg = delta_w on the x and u diagonals.  Returns 0, or 1 on a
  // non-positive pivot (wrong inertia).
  __device__ int riccati(double reg, double mu, int it) {
    double Pm[45], pv[NX], an[NP], cn, gsn;  // cost-to-go of stage k+1: Pxx, pxx, slack column a, c, g_s
    {
#pragma unroll
      for (int e = 0; e < 45; ++e) Pm[e] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int c = a; c < NP; ++c) Pm[ssidx(POSE2X[a], POSE2X[c])] = Qw(N, Q_HP + pidx(a, c));
#pragma unroll
      for (int q = 0; q < 3; ++q) Pm[ssidx(3 + q, 3 + q)] = Qw(N, Q_HVD + q);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        Pm[ssidx(i, i)] += reg;
        pv[i] = Qw(N, Q_GA + i) + mu * Qw(N, Q_GB + i);
      }
#pragma unroll
      for (int a = 0; a < NP; ++a) an[a] = Qw(N, Q_A + a);
      cn = Qw(N, Q_C); gsn = Qw(N, Q_GA + SGY_S) + mu * Qw(N, Q_GB + SGY_S);
#pragma unroll
      for (int e = 0; e < 45; ++e) Rw(N, R_P + e) = Pm[e];
#pragma unroll
      for (int i = 0; i < NX; ++i) Rw(N, R_PV + i) = pv[i];
    }
    int bad = 0;
    for (int k = N - 1; k >= 0; --k) {
      SACoef c = acoef(k, it);
      double d[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) d[i] = W2(k, S_DFC + i);
      // pd = p + P d ;  l0 = a.d + g_s(k+1) + g_v(k)
      double pd[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = pv[i];
#pragma unroll
        for (int q = 0; q < NX; ++q) v = fma(Pm[ssidx(i, q)], d[q], v);
        pd[i] = v;
      }
      double l0 = gsn + Qw(k, Q_GA + SGY_V) + mu * Qw(k, Q_GB + SGY_V);
#pragma unroll
      for (int a = 0; a < NP; ++a) l0 = fma(an[a], d[POSE2X[a]], l0);
      // C = (P A)[:, 2..5]
      double C[NX][4];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double p0 = Pm[ssidx(i, 0)], p1 = Pm[ssidx(i, 1)], p2 = Pm[ssidx(i, 2)], p3 = Pm[ssidx(i, 3)], p4 = Pm[ssidx(i, 4)], p5 = Pm[ssidx(i, 5)];
        C[i][0] = p2 + c.a32 * p3 + c.a42 * p4;
        C[i][1] = p3 + c.dt * p0 + c.a43 * p4;
        C[i][2] = p4 + c.dt * p1 + c.a34 * p3;
        C[i][3] = p5 + c.dt * p2 + c.a35 * p3 + c.a45 * p4;
      }
#define PA_(l, j) (((j) >= 2 && (j) <= 5) ? C[l][(j) - 2] : Pm[ssidx(l, j)])
      // Mux = B^T (P A) + H_ux ; Muu = B^T P B + H_uu + reg ; m_u
      double Mux[NU][NX], Muu[15], mvu[NU];
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        Mux[0][j] = c.dt * (c.cp * PA_(3, j) + c.sp * PA_(4, j));
        Mux[1][j] = c.dt * PA_(5, j);
        Mux[2][j] = c.dt * PA_(6, j); Mux[3][j] = c.dt * PA_(7, j); Mux[4][j] = c.dt * PA_(8, j);
      }
      Mux[0][2] += Qw(k, Q_HPU);
      {
        const double dd = c.dt * c.dt;
        double q33 = Pm[ssidx(3, 3)], q34 = Pm[ssidx(3, 4)], q44 = Pm[ssidx(4, 4)];
        Muu[0] = dd * (c.cp * (c.cp * q33 + c.sp * q34) + c.sp * (c.cp * q34 + c.sp * q44));
        Muu[1] = dd * (c.cp * Pm[ssidx(3, 5)] + c.sp * Pm[ssidx(4, 5)]);
        Muu[2] = dd * (c.cp * Pm[ssidx(3, 6)] + c.sp * Pm[ssidx(4, 6)]);
        Muu[3] = dd * (c.cp * Pm[ssidx(3, 7)] + c.sp * Pm[ssidx(4, 7)]);
        Muu[4] = dd * (c.cp * Pm[ssidx(3, 8)] + c.sp * Pm[ssidx(4, 8)]);
        // rows 1..4 <-> states 5..8
#pragma unroll
        for (int a = 1; a < NU; ++a)
#pragma unroll
          for (int b2 = a; b2 < NU; ++b2) Muu[a * 5 - a * (a - 1) / 2 + (b2 - a)] = dd * Pm[ssidx(4 + a, 4 + b2)];
#pragma unroll
        for (int a = 0; a < NU; ++a) Muu[a * 5 - a * (a - 1) / 2] += Qw(k, Q_HUU + a) + reg;
      }
      bt_mul(pd, mvu, c);
#pragma unroll
      for (int a = 0; a < NU; ++a) mvu[a] += Qw(k, Q_GA + SGY_U + a) + mu * Qw(k, Q_GB + SGY_U + a);
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
          Mxx[ssidx(i, j)] = v;
        }
#undef PA_
#pragma unroll
      for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int c2 = a; c2 < NP; ++c2) Mxx[ssidx(POSE2X[a], POSE2X[c2])] += Qw(k, Q_HP + pidx(a, c2));
#pragma unroll
      for (int q = 0; q < 3; ++q) Mxx[ssidx(3 + q, 3 + q)] += Qw(k, Q_HVD + q);
      Mxx[ssidx(3, 5)] += Qw(k, Q_H35); Mxx[ssidx(4, 5)] += Qw(k, Q_H45);
#pragma unroll
      for (int i = 0; i < NX; ++i) { Mxx[ssidx(i, i)] += reg; mvx[i] = pd[i]; }
      at_mul(mvx, c);
#pragma unroll
      for (int i = 0; i < NX; ++i) mvx[i] += Qw(k, Q_GA + i) + mu * Qw(k, Q_GB + i);
      // eliminate v_k = s_{k+1}:  w = [A^T a(k+1) + bv(k) ; B^T a(k+1)],  cv = hvv(k) + c(k+1)
      double wx[NX], wu[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) wx[i] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) wx[POSE2X[a]] = an[a];
      bt_mul(wx, wu, c);
      at_mul(wx, c);
#pragma unroll
      for (int a = 0; a < NP; ++a) wx[POSE2X[a]] += Qw(k, Q_BV + a);
      double cv = Qw(k, Q_HVV) + cn, icv = 1.0 / cv;
      bad |= !(cv > 1e-13);
      {
        double l0c = l0 * icv;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double wi = wx[i] * icv;
#pragma unroll
          for (int j = i; j < NX; ++j) Mxx[ssidx(i, j)] = fma(-wi, wx[j], Mxx[ssidx(i, j)]);
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
      double L[NU][NU], Dg[NU], iD[NU];
#define MUU_(i, j) Muu[(j) * 5 - (j) * ((j) - 1) / 2 + ((i) - (j))]  // i >= j
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double dj = MUU_(j, j);
#pragma unroll
        for (int q = 0; q < j; ++q) dj -= L[j][q] * L[j][q] * Dg[q];
        bad |= !(dj > 1e-13);
        Dg[j] = dj; iD[j] = 1.0 / dj;
#pragma unroll
        for (int i = j + 1; i < NU; ++i) {
          double v = MUU_(i, j);
#pragma unroll
          for (int q = 0; q < j; ++q) v -= L[i][q] * L[j][q] * Dg[q];
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
          if (j < NX) { Kc[a][j < NX ? j : 0] = -y[a]; Rw(k, R_K + a * NX + (j < NX ? j : 0)) = -y[a]; }
          else { kff[a] = -y[a]; Rw(k, R_KFF + a) = -y[a]; }
        }
      }
      // P_k = Mxx + Mux^T K ; p_k = m_x + Mux^T kff
#pragma unroll
      for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = i; j < NX; ++j) {
          double v = Mxx[ssidx(i, j)];
#pragma unroll
          for (int a = 0; a < NU; ++a) v = fma(Mux[a][i], Kc[a][j], v);
          Pm[ssidx(i, j)] = v; Rw(k, R_P + ssidx(i, j)) = v;
        }
        double v = mvx[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) v = fma(Mux[a][i], kff[a], v);
        pv[i] = v; Rw(k, R_PV + i) = v;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) Rw(k, R_W + i) = wx[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) Rw(k, R_W + NX + a) = wu[a];
      Rw(k, R_CV) = cv; Rw(k, R_L0) = l0;
#pragma unroll
      for (int a = 0; a < NP; ++a) an[a] = Qw(k, Q_A + a);
      cn = Qw(k, Q_C); gsn = Qw(k, Q_GA + SGY_S) + mu * Qw(k, Q_GB + SGY_S);
    }
    if (!(cn > 1e-13)) return 1;
    return 0;
  }

  // roll-out of the Newton step (thread per instance): dx, du, ds and the new costates
  // lam+ = P [dx; ds] + p
  __device__ void rollout(double mu, int it) {
    double dxv[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { dxv[i] = 0; W2(0, S_DX + i) = 0; }
    double dsv = -(Qw(0, Q_GA + SGY_S) + mu * Qw(0, Q_GB + SGY_S)) / Qw(0, Q_C);
    W2(0, S_DS) = dsv;
    for (int k = 0; k < N; ++k) {
      double duv[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        double v = Rw(k, R_KFF + a);
#pragma unroll
        for (int j = 0; j < NX; ++j) v = fma(Rw(k, R_K + a * NX + j), dxv[j], v);
        duv[a] = v; W2(k, S_DU + a) = v;
      }
      double l = Rw(k, R_L0);
#pragma unroll
      for (int i = 0; i < NX; ++i) l = fma(Rw(k, R_W + i), dxv[i], l);
#pragma unroll
      for (int a = 0; a < NU; ++a) l = fma(Rw(k, R_W + NX + a), duv[a], l);
      dsv = -l / Rw(k, R_CV);
      SACoef c = acoef(k, it);
      double nx_[NX];
      nx_[0] = dxv[0] + dt * dxv[3]; nx_[1] = dxv[1] + dt * dxv[4]; nx_[2] = dxv[2] + dt * dxv[5];
      nx_[3] = dxv[3] + c.a32 * dxv[2] + c.a34 * dxv[4] + c.a35 * dxv[5] + dt * c.cp * duv[0];
      nx_[4] = dxv[4] + c.a42 * dxv[2] + c.a43 * dxv[3] + c.a45 * dxv[5] + dt * c.sp * duv[0];
      nx_[5] = dxv[5] + dt * duv[1];
      nx_[6] = dxv[6] + dt * duv[2]; nx_[7] = dxv[7] + dt * duv[3]; nx_[8] = dxv[8] + dt * duv[4];
#pragma unroll
      for (int i = 0; i < NX; ++i) { dxv[i] = nx_[i] + W2(k, S_DFC + i); W2(k + 1, S_DX + i) = dxv[i]; }
      W2(k + 1, S_DS) = dsv;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = Rw(k + 1, R_PV + i);
#pragma unroll
        for (int j = 0; j < NX; ++j) v = fma(Rw(k + 1, R_P + ssidx(i, j)), dxv[j], v);
        if (i < 3) v = fma(Qw(k + 1, Q_A + i), dsv, v);
        if (i >= 6) v = fma(Qw(k + 1, Q_A + (i - 3)), dsv, v);
        W2(k + 1, S_LAMN + i) = v;
      }
    }
    if (term_eq(N)) {  // Newton target of the terminal-equality multipliers
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double cq = W(N, it + I_X + i) - W2(N, IN_XREF + i);
        W2(0, S_LAMN + i) = W(0, it + I_LAM + i) + (dxv[i] + cq) / MMPC_DELTA_C;
      }
    }
  }

  // results: sol.value(U/X/s/cost) :317,:329-330
  __device__ void finish(int status) {
    const int it = J(J_CUR) * ITSZ;
    double fsum = 0;
    for (int k = 0; k <= N; ++k) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = W(k, it + I_X + i), e = xerr(i, v, W2(k, IN_XREF + i));
        fsum += xweight(k, i) * e * e;
        if (P.io->X) P.io->X[((long long)bio * (N + 1) + k) * NX + i] = v;
      }
#ifdef MMPC_POSEREF
      if (pose_model()) fsum += pose_cost_value(k, it);
#endif
      if (k < N)
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double v = W(k, it + I_U + j), e = v - W2(k, IN_UREF + j), dl = v - W2(k, IN_ULAST + j);
          fsum += cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl;
          P.io->U[((long long)bio * N + k) * NU + j] = v;
        }
      double s = W(k, it + I_S);
      fsum += cfg.S * s * s;
      if (P.io->s) P.io->s[(long long)bio * (N + 1) + k] = s;
    }
    if (P.io->cost) P.io->cost[bio] = fsum;
    if (P.io->kkt) P.io->kkt[bio] = D(D_E0);
    if (P.io->iters) P.io->iters[bio] = J(J_IT);
    P.io->status[bio] = status;
    J(J_STATE) = ST_DONE;
  }

  // ------------------------------------------------------------------------------------------
  // solve (thread per instance): KKT reduction, convergence test, barrier update, Riccati
  // factorisation with inertia correction, roll-out.
  __device__ void solve() {
    const double kap_eps = 10, kap_mu = 0.2, th_mu = 1.5;
    const double tol = cfg.tol;
    const int it = J(J_CUR) * ITSZ;
    KktParts kp;
    kp.e_stat = 0; kp.e_prim = 0; kp.c_hi = -1e300; kp.c_lo = 1e300; kp.sum_lam = 0; kp.sum_z = 0;
    double nz = 0, neq = 0;
    for (int k = N; k >= 0; --k) {
      kp.e_stat = fmax(kp.e_stat, W2(k, S_PART + 0)); kp.e_prim = fmax(kp.e_prim, W2(k, S_PART + 1));
      kp.c_hi = fmax(kp.c_hi, W2(k, S_PART + 2)); kp.c_lo = fmin(kp.c_lo, W2(k, S_PART + 3));
      kp.sum_lam += W2(k, S_PART + 4); kp.sum_z += W2(k, S_PART + 5); nz += W2(k, S_PART + 6); neq += W2(k, S_PART + 7);
    }
    kp.n_z = (int)nz; kp.n_eq = (int)neq;
    double E0 = kkt_error(kp, 0.0);
    D(D_E0) = E0;
    if (!(E0 == E0)) { finish(MMPC_STATUS_NAN); return; }
    if (E0 <= tol) { finish(MMPC_STATUS_CONVERGED); return; }
    if (J(J_IT) >= cfg.max_iter) { finish(MMPC_STATUS_MAX_ITER); return; }
    double mu = D(D_MU);
    bool mu_changed = false;
    while (kkt_error(kp, mu) <= kap_eps * mu && mu > tol / 10) {
      mu = fmax(tol / 10, fmin(kap_mu * mu, pow(mu, th_mu))); mu_changed = true;
    }
    if (mu_changed) { J(J_NFILT) = 0; J(J_FRST) = J(J_FRST) & 0xff00; D(D_MU) = mu; }
    double reg = 0, reg_last = D(D_REGLAST);
    int tries = 0;
    for (;;) {
      int fail = riccati(reg, mu, it);
      if (!fail) { if (reg > 0) D(D_REGLAST) = reg; J(J_REGF) = reg > 0; break; }
      if (reg == 0) reg = (reg_last == 0) ? 1e-4 : fmax(1e-20, reg_last / 3);
      else reg *= (reg_last == 0 ? 100 : 8);
      if (++tries > 40 || reg > 1e20) { finish(MMPC_STATUS_FACTOR); return; }
    }
    rollout(mu, it);
  }

  // ------------------------------------------------------------------------------------------
  // step (thread per instance and stage): slack / multiplier steps of every row and bound,
  // fraction to the boundary, merit ingredients of the current point.
  template <bool REF>
  __device__ void step(int k) {
    load_npl();
    const int it = J(J_CUR) * ITSZ;
    prefetch_stage(k, it, false);
    const double* ci = stage_ptr(k, it); double* c2 = stage_ptr(k, B2);
    // ring items in consumption order: x bounds (zl, zu) x9, u bounds x5, circle rows (t, z, cx, cy, r), other rows (t, z)
    const int q_circ = NX + NU, q_rows = q_circ + nobs, q_end = q_rows + nself + (npl > 0 ? 6 : 0);
    auto ring_issue = [&](int q) {  // one code path (selects, no per-kind branches): the body is inlined at 20 sites
      double* d = ring_slot(q);
      if (q < q_end) {
        const bool isbox = q < q_circ;
        const int r = q - q_circ;
        const int oa = isbox ? (q < NX ? I_ZXL + q : I_ZXU + q) : I_T + r;          // I_ZUL + (q - NX) = I_ZXU + q
        const int ob = isbox ? (q < NX ? I_ZXU + q : I_ZUU - NX + q) : I_T + R + r;
        async_copy8(d, &ci[oa << LSH]); async_copy8(d + bs, &ci[ob << LSH]);
        if (!isbox && q < q_rows) { async_copy8(d + 2 * bs, circ_ptr(k, r, 0)); async_copy8(d + 3 * bs, circ_ptr(k, r, 1)); async_copy8(d + 4 * bs, circ_ptr(k, r, 2)); }
      }
      async_commit();
    };
    int rq = 0;  // next ring item
    auto ring_pop = [&]() -> const double* { async_wait<RING_D - 1>(); return ring_slot(rq); };
    auto ring_next = [&]() { ring_issue(rq + RING_D); ++rq; };
    // The 29 reference / bound inputs of the stage (IN_XREF .. S_DT) and its FK cache + defect (S_FK .. S_DFC + 9) are parked
    // in shared memory by one group of copies in front of the ring (so the first ring_pop also waits for it): read one at a
    // time where they are used, each of those loads exposed a full HBM round trip (34 % of the kernel's stall samples)
#if defined(MMPC_RESIDENT)
    const double* pk_in = &c2[IN_XREF]; const double* pk_fk = &c2[S_FK]; constexpr int ps = 1, pf = 1;   // shared memory already
#else
    const int ps = STEP_PARK >= 1 ? bs : (1 << LSH), pf = STEP_PARK >= 2 ? bs : (1 << LSH);
    const double* pk_in = &c2[IN_XREF << LSH]; const double* pk_fk = &c2[S_FK << LSH];
    if (STEP_PARK >= 1) {
      double* d = sm + (RING_D * RING_W) * bs;
#pragma unroll 1
      for (int f = 0; f < S_DT - IN_XREF; ++f) async_copy8(d + f * bs, &c2[(IN_XREF + f) << LSH]);
      pk_in = d;
      if (STEP_PARK >= 2) {
        d += (S_DT - IN_XREF) * bs;
#pragma unroll 1
        for (int f = 0; f < S_PART - S_FK; ++f) async_copy8(d + f * bs, &c2[(S_FK + f) << LSH]);
        pk_fk = d;
      }
      async_commit();
    }
#endif
#pragma unroll
    for (int q = 0; q < RING_D; ++q) ring_issue(q);
    const double os = D(D_OS), mu = D(D_MU);
    const double tau = fmax(0.99, 1 - mu);
    MinRatio rp, rd; rp.init(); rd.init();  // fraction to the boundary: primal, dual
    double gphi = 0, theta = 0, fsum = 0;
    LogProd lp; lp.init();
    double x[NX], dxv[NX], u[NU], duv[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) { x[i] = ldg(&ci[(I_X + i) << LSH]); dxv[i] = ldg(&c2[(S_DX + i) << LSH]); }
    double s = ldg(&ci[(I_S) << LSH]), dsv = ldg(&c2[(S_DS) << LSH]);
#pragma unroll
    for (int a = 0; a < NU; ++a) { u[a] = (k < N) ? ldg(&ci[(I_U + a) << LSH]) : 0.0; duv[a] = (k < N) ? ldg(&c2[(S_DU + a) << LSH]) : 0.0; }
    double dp[NP];
#pragma unroll
    for (int a = 0; a < NP; ++a) dp[a] = dxv[POSE2X[a]];
    // cost / boxes
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const double* rb = ring_pop(); const double zl_c = rb[0], zu_c = rb[bs]; ring_next();
      double Wx = os * xweight(k, i), e = xerr(i, x[i], pk_in[i * ps]);
      fsum += Wx * e * e; gphi += 2 * Wx * e * dxv[i];
      if (k >= 1) {
        double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
        if (is_fin(lo)) {
          double d = x[i] - lo, id = rcp(d), z = zl_c, dz = mu * id - z - z * id * dxv[i];
          gphi -= mu * dxv[i] * id; lp.mul(d);
          if (dxv[i] < 0) rp.add(d, -dxv[i]);
          if (dz < 0) rd.add(z, -dz);
        }
        if (is_fin(hi)) {
          double d = hi - x[i], id = rcp(d), z = zu_c, dz = mu * id - z + z * id * dxv[i];
          gphi += mu * dxv[i] * id; lp.mul(d);
          if (dxv[i] > 0) rp.add(d, dxv[i]);
          if (dz < 0) rd.add(z, -dz);
        }
      }
    }
    double S1 = os * cfg.S;
    fsum += S1 * s * s; gphi += 2 * S1 * s * dsv;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      const double* rb = ring_pop(); const double zl_c = rb[0], zu_c = rb[bs]; ring_next();
      if (k < N) {
        double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
        double e = u[j] - pk_in[(IN_UREF - IN_XREF + j) * ps], dl = u[j] - pk_in[(IN_ULAST - IN_XREF + j) * ps];
        fsum += Rj * e * e + Wj * dl * dl; gphi += (2 * Rj * e + 2 * Wj * dl) * duv[j];
        double lo = pk_in[(IN_ULO - IN_XREF + j) * ps], hi = pk_in[(IN_UHI - IN_XREF + j) * ps];
        if (is_fin(lo)) {
          double d = u[j] - lo, id = rcp(d), z = zl_c, dz = mu * id - z - z * id * duv[j];
          gphi -= mu * duv[j] * id; lp.mul(d);
          if (duv[j] < 0) rp.add(d, -duv[j]);
          if (dz < 0) rd.add(z, -dz);
        }
        if (is_fin(hi)) {
          double d = hi - u[j], id = rcp(d), z = zu_c, dz = mu * id - z + z * id * duv[j];
          gphi += mu * duv[j] * id; lp.mul(d);
          if (duv[j] > 0) rp.add(d, duv[j]);
          if (dz < 0) rd.add(z, -dz);
        }
      }
    }
    if (k < N) {
#pragma unroll
      for (int i = 0; i < NX; ++i) theta += fabs(pk_fk[(S_DFC - S_FK + i) * pf]);
    }
    if (term_eq(k)) theta += fabs(x[0] - pk_in[0]) + fabs(x[1] - pk_in[ps]);
    FK f; f.cp = pk_fk[0]; f.sp = pk_fk[pf];
#pragma unroll
    for (int q = 0; q < 3; ++q) { f.vr[q] = pk_fk[(2 + q) * pf]; f.vh[q] = pk_fk[(5 + q) * pf]; }
#ifdef MMPC_POSEREF
    if (pose_model()) {  // the end-point pose cost: value and directional derivative along the step
      const double r[4] = {pk_in[0], pk_in[ps], pk_in[2 * ps], pk_in[3 * ps]};
      double g[NP];
      fsum += pose_cost(k, os, x[0], x[1], x[2], f, r, g, nullptr);
#pragma unroll
      for (int a = 0; a < NP; ++a) gphi = fma(g[a], dp[a], gphi);
    }
#endif
#if !defined(MMPC_RESIDENT) && !defined(MMPC_NO_MARGIN_PREFETCH)
    // (reference NLP) the margin caches of stages k-1 and k for the stale-column rows at the end: prefetched into the parked
    // slots, which are free from here on, while the rows below are worked through
    const bool mg_pref = REF && npl >= 2 && k >= 1 && 2 * 6 * cfg.n_pl <= STAGED_STEP_PARKED;
    if (mg_pref) margins_prefetch<2>(k, sm + (RING_D * RING_W) * bs);
#endif
    // rows: dt_i = -res_i - (grad h_i . dx - ds)
    double s_cur = s, ds_cur = dsv;  // slack (and its step) the rows are bounded by
    auto row_step = [&](int r, double h, double gd_, double t, double z) {
      double res = h - s_cur + t;
      double dtv = -res - (gd_ - ds_cur);
      c2[(S_DT + r) << LSH] = dtv;
      double itv = rcp(t), dz = (mu - z * (t + dtv)) * itv;
      theta += fabs(res); gphi -= mu * dtv * itv; lp.mul(t);
      if (dtv < 0) rp.add(t, -dtv);
      if (dz < 0) rd.add(z, -dz);
    };
    for (int i = 0; i < nobs; ++i) {
      const double* rb = ring_pop();
      const double rt = rb[0], rz = rb[bs], ddx = x[0] - rb[2 * bs], ddy = x[1] - rb[3 * bs], rad = rb[4 * bs];
      ring_next();
      double d2 = ddx * ddx + ddy * ddy, inv = rsq(d2), d = d2 * inv;
      row_step(i, (rad + cfg.base_radius) - d, -(ddx * dp[0] + ddy * dp[1]) * inv, rt, rz);
    }
    if (REF && k == N && q3()) { s_cur = W(N - 1, it + I_S); ds_cur = W2(N - 1, S_DS); }  // terminal rows on s[N-1]
#pragma unroll 1
    for (int m = 0; m < nself; ++m) {
      Point p; point_eval(x[0], x[1], f, SELFD[m], p);
      double d2 = p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2], inv = rsq(d2), d = d2 * inv;
      double n[3] = {p.P[0] * inv, p.P[1] * inv, p.P[2] * inv}, g[NP];
      point_grad(f, p, n, g);
      double gd_ = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) gd_ = fma(g[a], dp[a], gd_);
      const double* rb = ring_pop(); const double rt = rb[0], rz = rb[bs]; ring_next();
      row_step(nobs + m, cfg.self_collision_radius - d, -gd_, rt, rz);
    }
    s_cur = s; ds_cur = dsv;
    if (npl > 0) {
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        Point p; point_eval(x[0], x[1], f, BODY[i], p);
        int jb; double h = plane_row(p, jb);
        double n[3] = {PL(6 * jb + 3), PL(6 * jb + 4), PL(6 * jb + 5)}, g[NP];
        point_grad(f, p, n, g);
        double gd_ = 0;
#pragma unroll
        for (int a = 0; a < NP; ++a) gd_ = fma(g[a], dp[a], gd_);
        const double* rb = ring_pop(); const double rt = rb[0], rz = rb[bs]; ring_next();
        row_step(nobs + nself + i, h, gd_, rt, rz);
      }
    }
    double log_extra = 0;
    if (REF) {  // compiled out of the clean-mode kernels
      StaleIO io; io.reset(); io.rp = rp; io.rd = rd;
      RowAcc Adummy; double bvdummy[NP];   // phase 2 touches neither
#if defined(MMPC_RESIDENT) || defined(MMPC_NO_MARGIN_PREFETCH)
      margins_in_place(k);
#else
      if (mg_pref) async_wait<0>(); else margins_in_place(k);
#endif
      stale_rows<2>(k, Adummy, bvdummy, io);
      theta += io.theta; gphi += io.gphi; log_extra = io.lp.value(); rp = io.rp; rd = io.rd;
    }
    c2[(S_PART + 0) << LSH] = rp.value(tau); c2[(S_PART + 1) << LSH] = rd.value(tau); c2[(S_PART + 2) << LSH] = gphi; c2[(S_PART + 3) << LSH] = theta;
    c2[(S_PART + 4) << LSH] = fsum; c2[(S_PART + 5) << LSH] = lp.value() + log_extra;
    async_wait<0>();  // nothing of this item's ring may still be in flight when the thread primes the next one
  }

  // Stage k of the results of an instance that left the solve in this round (state ST_FINISH):
  // sol.value(U/X/s) :329-330 and the stage's share of the unscaled cost :317
  __device__ void finish_stage(int k) const {
    const int it = J(J_CUR) * ITSZ;
    double fsum = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v = W(k, it + I_X + i), e = xerr(i, v, W2(k, IN_XREF + i));
      fsum += xweight(k, i) * e * e;
      if (P.io->X) P.io->X[((long long)bio * (N + 1) + k) * NX + i] = v;
    }
#ifdef MMPC_POSEREF
    if (pose_model()) fsum += pose_cost_value(k, it);
#endif
    if (k < N)
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double v = W(k, it + I_U + j), e = v - W2(k, IN_UREF + j), dl = v - W2(k, IN_ULAST + j);
        fsum += cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl;
        P.io->U[((long long)bio * N + k) * NU + j] = v;
      }
    double s = W(k, it + I_S);
    fsum += cfg.S * s * s;
    if (P.io->s) P.io->s[(long long)bio * (N + 1) + k] = s;
    W2(k, S_PART + 0) = fsum;
  }

  // ctrl_step (thread per instance): reduce the step partials, start the line search.
  // Returns true if the instance goes on to a trial.
  // NL lanes share the reduction over the stages (NL = 32: one warp per instance on the GPU, butterfly sums in a
  // fixed order; NL = 1: one lane does everything, the CPU emulation); lane 0 writes the instance state.
  template <int NL>
  __device__ bool ctrl_step(int lane) {
    // (the state is the same in all lanes; the branches on it are taken on VOTES so that the compiler knows them warp-uniform
    // and does not bracket every shuffle below with WARPSYNC.COLLECTIVE / ENDCOLLECTIVE)
    const int st = J(J_STATE);
    if (warp_all(st == ST_FINISH)) {  // close an instance whose stages were written by finish_stage
      double fsum = 0;
      for (int k = lane; k <= N; k += NL) fsum += W2(k, S_PART + 0);
      fsum = lanes_sum<NL>(fsum);
      if (lane == 0) {
        if (P.io->cost) P.io->cost[bio] = fsum;
        if (P.io->kkt) P.io->kkt[bio] = D(D_E0);
        if (P.io->iters) P.io->iters[bio] = J(J_IT);
        P.io->status[bio] = J(J_STATUS);
        J(J_STATE) = ST_DONE;
      }
      return false;
    }
    if (!warp_all(st == ST_ACTIVE)) return false;
    double ap = 1.0, ad = 1.0, gphi = 0, theta = 0, fsum = 0, logsum = 0;
    for (int k = lane; k <= N; k += NL) {
      ap = fmin(ap, W2(k, S_PART + 0)); ad = fmin(ad, W2(k, S_PART + 1));
      gphi += W2(k, S_PART + 2); theta += W2(k, S_PART + 3); fsum += W2(k, S_PART + 4); logsum += W2(k, S_PART + 5);
    }
    ap = lanes_min<NL>(ap); ad = lanes_min<NL>(ad);
    gphi = lanes_sum<NL>(gphi); theta = lanes_sum<NL>(theta); fsum = lanes_sum<NL>(fsum); logsum = lanes_sum<NL>(logsum);
    if (lane == 0) {
      if (D(D_THMAX) < 0) { D(D_THMAX) = 1e4 * fmax(1.0, theta); D(D_THMIN) = 1e-4 * fmax(1.0, theta); }
      D(D_ALPHA) = ap; D(D_AP0) = ap; D(D_AD) = ad; D(D_GPHI) = gphi; D(D_THETA) = theta; D(D_PHI0) = fsum - D(D_MU) * logsum;
      J(J_LS) = 0;
      {  // filter reset heuristic, at the start of the iteration's line search (IpFilterLSAcceptor: InitThisLineSearch)
        const int fr = J(J_FRST);
        int cnt = fr & 0xff, nres = (fr >> 8) & 0xff;
        if (nres < MAX_FILTER_RESETS) {
          if (fr >> 16 & 1) { if (++cnt >= FILTER_RESET_TRIGGER) { J(J_NFILT) = 0; nres++; cnt = 0; } }
          else cnt = 0;
        }
        J(J_FRST) = cnt | (nres << 8);
      }
      J(J_STATE) = ST_TRIAL;
    }
    return true;
  }

  // ------------------------------------------------------------------------------------------
  // trial (thread per instance and stage): candidate iterate  w + alpha d  (primal and dual, with
  // slack reset and multiplier safeguard) into the other iterate buffer; merit ingredients.
  __device__ void trial(int k) {
    load_npl();
    const int it = J(J_CUR) * ITSZ, jt = (1 - J(J_CUR)) * ITSZ;
    const double os = D(D_OS), mu = D(D_MU), alpha = D(D_ALPHA), ad = D(D_AD);
    double theta = 0, fsum = 0; bool ok = true;
    LogProd lp; lp.init();
    double x[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double xo = W(k, it + I_X + i), dxi = W2(k, S_DX + i);
      x[i] = fma(alpha, dxi, xo);
      W(k, jt + I_X + i) = x[i];
      if (k >= 1) {
        double l = W(k, it + I_LAM + i);
        W(k, jt + I_LAM + i) = l + alpha * (W2(k, S_LAMN + i) - l);
      }
      double Wx = os * xweight(k, i), e = xerr(i, x[i], W2(k, IN_XREF + i));
      fsum += Wx * e * e;
      if (k >= 1) {
        double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
        if (is_fin(lo)) {
          double d_old = xo - lo, d = x[i] - lo, z = W(k, it + I_ZXL + i);
          double dz = mu / d_old - z - (z / d_old) * dxi;
          z += ad * dz; z = fmax(fmin(z, 1e10 * mu / d), mu / (1e10 * d));
          W(k, jt + I_ZXL + i) = z;
          if (d <= 0) ok = false; else lp.mul(d);
        }
        if (is_fin(hi)) {
          double d_old = hi - xo, d = hi - x[i], z = W(k, it + I_ZXU + i);
          double dz = mu / d_old - z + (z / d_old) * dxi;
          z += ad * dz; z = fmax(fmin(z, 1e10 * mu / d), mu / (1e10 * d));
          W(k, jt + I_ZXU + i) = z;
          if (d <= 0) ok = false; else lp.mul(d);
        }
      }
    }
    double s = fma(alpha, W2(k, S_DS), W(k, it + I_S));
    W(k, jt + I_S) = s;
    fsum += os * cfg.S * s * s;
    FK f; fk_eval(x[2], x[6], x[7], x[8], f);
    if (k < N) {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double uo = W(k, it + I_U + j), duj = W2(k, S_DU + j);
        u[j] = fma(alpha, duj, uo);
        W(k, jt + I_U + j) = u[j];
        double e = u[j] - W2(k, IN_UREF + j), dl = u[j] - W2(k, IN_ULAST + j);
        fsum += os * (cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl);
        double lo = W2(k, IN_ULO + j), hi = W2(k, IN_UHI + j);
        if (is_fin(lo)) {
          double d_old = uo - lo, d = u[j] - lo, z = W(k, it + I_ZUL + j);
          double dz = mu / d_old - z - (z / d_old) * duj;
          z += ad * dz; z = fmax(fmin(z, 1e10 * mu / d), mu / (1e10 * d));
          W(k, jt + I_ZUL + j) = z;
          if (d <= 0) ok = false; else lp.mul(d);
        }
        if (is_fin(hi)) {
          double d_old = hi - uo, d = hi - u[j], z = W(k, it + I_ZUU + j);
          double dz = mu / d_old - z + (z / d_old) * duj;
          z += ad * dz; z = fmax(fmin(z, 1e10 * mu / d), mu / (1e10 * d));
          W(k, jt + I_ZUU + j) = z;
          if (d <= 0) ok = false; else lp.mul(d);
        }
      }
      double xn[NX]; dyn_f(x, u, dt, f.cp, f.sp, xn);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double x1 = fma(alpha, W2(k + 1, S_DX + i), W(k + 1, it + I_X + i));
        theta += fabs(xn[i] - x1);
      }
    }
    auto row_val = [&](int r, double h) {
      double t = W(k, it + I_T + r), z = W(k, it + I_T + R + r), dtv = W2(k, S_DT + r);
      double tt = fma(alpha, dtv, t);
      tt = fmax(tt, s - h);  // slack reset (Nocedal & Wright 19.30)
      double dz = (mu - z * (t + dtv)) / t;
      z += ad * dz; z = fmax(fmin(z, 1e10 * mu / tt), mu / (1e10 * tt));
      W(k, jt + I_T + r) = tt; W(k, jt + I_T + R + r) = z;
      theta += fabs(h - s + tt);
      if (tt <= 0) ok = false; else lp.mul(tt);
    };
    for (int i = 0; i < nobs; ++i) {
      double ddx = x[0] - circ(k, i, 0), ddy = x[1] - circ(k, i, 1);
      row_val(i, (circ(k, i, 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy));
    }
#pragma unroll 1
    for (int m = 0; m < nself; ++m) {
      Point p; point_eval(x[0], x[1], f, SELFD[m], p);
      row_val(nobs + m, cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]));
    }
    if (npl > 0) {
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        Point p; point_eval(x[0], x[1], f, BODY[i], p);
        int jb; row_val(nobs + nself + i, plane_row(p, jb));
      }
    }
    bool fin = ok && (fsum == fsum) && (theta == theta);
    W2(k, S_PART + PT_MERIT + 0) = theta; W2(k, S_PART + PT_MERIT + 1) = fsum; W2(k, S_PART + PT_MERIT + 2) = fin ? lp.value() : 0.0;
    W2(k, S_PART + PT_MERIT + 3) = fin ? 1.0 : 0.0;
  }

  // ------------------------------------------------------------------------------------------
  // trial_eval (thread per instance and stage): the trial kernel fused with the evaluation of the
  // NEXT iteration.  Builds the candidate iterate  w + alpha d  (primal and dual, slack reset, multiplier
  // safeguard) in the other iterate buffer, its merit ingredients, and -- because an accepted candidate
  // is exactly where the next iteration linearises -- the stage QP, defect, FK cache and KKT partials at
  // the candidate, all from registers.  A rejected candidate only wastes the derivative arithmetic: its
  // records are overwritten by the next trial before anything reads them.
  template <bool REF>
  __device__ void trial_eval(int k) {
    load_npl();
    const int it = J(J_CUR) * ITSZ, jt = (1 - J(J_CUR)) * ITSZ;
    const double os = D(D_OS), mu = D(D_MU), alpha = D(D_ALPHA), ad = D(D_AD);
    const double* ci = stage_ptr(k, it); double* cj = stage_ptr(k, jt); double* c2 = stage_ptr(k, B2);
    const int k1 = k < N ? k + 1 : k;
    const double* ni = stage_ptr(k1, it); const double* n2 = stage_ptr(k1, B2);
    prefetch_stage(k, it, true);
    // The rows (t, z, dt; circle rows also their circle) stream through a small cp.async ring, RING_DT rows ahead: primed
    // here, consumed after the cost / bound section in three rolled loops.  Not the bound multipliers: neither ringed (14 unrolled
    // issue sites: instruction-cache misses) nor parked in shared memory up front (L1, where the spills live, shrinks) paid.
    // the 28 bound multipliers (fields I_ZXL .. I_T) are parked in shared memory by one group of copies and read back as plain
    // LDS at the sites that use them: each of those sites exposed a full HBM round trip (the compiler does not hoist them)
#ifdef MMPC_RESIDENT
    // (resident build: the workspace is shared memory already, the parked values are read where they lie)
    const double* zb = &ci[I_ZXL]; const double* rb_in = &c2[IN_XREF]; constexpr int ps = 1;
#else
    double* zb = sm + (RING_DT * RING_W) * bs;
#pragma unroll 1
    for (int f = 0; f < I_T - I_ZXL; ++f) async_copy8(zb + f * bs, &ci[(I_ZXL + f) << LSH]);
    // ... and so are the 29 reference / bound inputs of the stage (fields IN_XREF .. S_DT of the second block)
    double* rb_in = zb + (I_T - I_ZXL) * bs;
#pragma unroll 1
    for (int f = 0; f < S_DT - IN_XREF; ++f) async_copy8(rb_in + f * bs, &c2[(IN_XREF + f) << LSH]);
    async_commit();
    const int ps = bs;
#endif
    const int q_end = nobs + nself + (npl > 0 ? 6 : 0);  // ring items = the rows in evaluation order: circles, self-collision, planes
    auto ring_issue = [&](int q) {
      if (q < q_end) {
        double* d = sm + ((q & (RING_DT - 1)) * RING_W) * bs;
        async_copy8(d, &ci[(I_T + q) << LSH]); async_copy8(d + bs, &ci[(I_T + R + q) << LSH]); async_copy8(d + 2 * bs, &c2[(S_DT + q) << LSH]);
        if (q < nobs) { async_copy8(d + 3 * bs, circ_ptr(k, q, 0)); async_copy8(d + 4 * bs, circ_ptr(k, q, 1)); async_copy8(d + 5 * bs, circ_ptr(k, q, 2)); }
      }
      async_commit();
    };
#pragma unroll 1
    for (int q = 0; q < RING_DT; ++q) ring_issue(q);
    double theta = 0, fsum = 0; bool ok = true;
    LogProd lp; lp.init();
    RowAcc A;
    A.chi = -1e300; A.clo = 1e300; A.prim = 0; A.sumz = 0; A.zrows = 0; A.nz = 0; A.csum = A.be0 = A.be1 = 0;
#pragma unroll
    for (int e = 0; e < 21; ++e) A.H[e] = 0;
#pragma unroll
    for (int a = 0; a < NP; ++a) A.a[a] = A.gA[a] = A.gB[a] = A.st[a] = 0;
    double x[NX], u[NU], lam[NX], lam1[NX], xo[NX], dxo[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      xo[i] = ldg(&ci[(I_X + i) << LSH]); dxo[i] = ldg(&c2[(S_DX + i) << LSH]);
      x[i] = fma(alpha, dxo[i], xo[i]);
      cj[(I_X + i) << LSH] = x[i];
      lam[i] = 0;
      if (k >= 1) {
        double l = ldg(&ci[(I_LAM + i) << LSH]);
        lam[i] = l + alpha * (ldg(&c2[(S_LAMN + i) << LSH]) - l);
        cj[(I_LAM + i) << LSH] = lam[i];
      }
    }
    const double s = fma(alpha, ldg(&c2[(S_DS) << LSH]), ldg(&ci[(I_S) << LSH]));
    cj[(I_S) << LSH] = s;
    double uo[NU], duo[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      uo[j] = (k < N) ? ldg(&ci[(I_U + j) << LSH]) : 0.0; duo[j] = (k < N) ? ldg(&c2[(S_DU + j) << LSH]) : 0.0;
      u[j] = fma(alpha, duo[j], uo[j]);
      if (k < N) cj[(I_U + j) << LSH] = u[j];
    }
    FK f;
    if (REF) load_fk(k, f);  // the pose kernel has just evaluated the candidate's forward kinematics
    else {
      fk_eval(x[2], x[6], x[7], x[8], f);
      c2[(S_FK + 0) << LSH] = f.cp; c2[(S_FK + 1) << LSH] = f.sp;
#pragma unroll
      for (int q = 0; q < 3; ++q) { c2[(S_FK + 2 + q) << LSH] = f.vr[q]; c2[(S_FK + 5 + q) << LSH] = f.vh[q]; }
    }
    double es = 0, hpp = 0, sum_lam = 0;
    int n_eq = 0;
    // dynamics :180 at the candidate -- defect and costate terms (A^T lam_{k+1}, B^T lam_{k+1})
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
        double l1 = ldg(&ni[(I_LAM + i) << LSH]);
        lam1[i] = l1 + alpha * (ldg(&n2[(S_LAMN + i) << LSH]) - l1);
        double x1 = fma(alpha, ldg(&n2[(S_DX + i) << LSH]), ldg(&ni[(I_X + i) << LSH]));
        double d = xn[i] - x1;
        c2[(S_DFC + i) << LSH] = d; A.prim = fmax(A.prim, fabs(d)); sum_lam += fabs(lam1[i]);
        theta += fabs(d);
      }
      n_eq = NX;
      hpp = -dt * u[0] * (lam1[3] * f.cp + lam1[4] * f.sp);
      Qw(k, Q_HPU) = dt * (-lam1[3] * f.sp + lam1[4] * f.cp);
      Qw(k, Q_H45) = -dt * lam1[3];  // (dy,dpsi)
      Qw(k, Q_H35) = dt * lam1[4];   // (dx,dpsi)
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
      Qw(k, Q_HPU) = 0; Qw(k, Q_H45) = 0; Qw(k, Q_H35) = 0;
    }
    // candidate multiplier of a bound at distance d_old -> d (sgn = +1 lower, -1 upper); merit + KKT bookkeeping
    auto box = [&](double zold, double d_old, double d, double sgn_dv, int slot, double& id) -> double {
      double ido = rcp(d_old); id = rcp(d);
      double dz = mu * ido - zold - (zold * ido) * sgn_dv;
      double z = zclamp(zold + ad * dz, mu, id);
      W(k, slot) = z;
      if (d <= 0) ok = false; else lp.mul(d);
      A.chi = fmax(A.chi, z * d); A.clo = fmin(A.clo, z * d); A.sumz += z; A.nz++;
      return z;
    };
    const bool teq = term_eq(k);
    double nu[2] = {0, 0};
    if (teq) {  // candidate multipliers of the terminal equality
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double no = W(0, it + I_LAM + i);
        nu[i] = no + alpha * (W2(0, S_LAMN + i) - no);
        W(0, jt + I_LAM + i) = nu[i];
      }
    }
    // cost and boxes -- :192-205, :240-245
    async_wait<RING_DT>();  // the bound multipliers have landed (the row ring's RING_DT groups were committed after them)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double Wx = os * xweight(k, i);
      double e = xerr(i, x[i], rb_in[(IN_XREF - IN_XREF + i) * ps]);
      fsum += Wx * e * e;
      double gr = 2 * Wx * e;
      double Hd = 2 * Wx, gA = gr, gB = 0, st = gr + stx[i] - lam[i];
      if (k >= 1) {
        double lo = cfg.xlim[0][i], hi = cfg.xlim[1][i];
        if (is_fin(lo)) {
          double d = x[i] - lo, id, z = box(zb[i * ps], xo[i] - lo, d, dxo[i], jt + I_ZXL + i, id);
          Hd += z * id; gB -= id; st -= z;
        }
        if (is_fin(hi)) {
          double d = hi - x[i], id, z = box(zb[(I_ZXU - I_ZXL + i) * ps], hi - xo[i], d, -dxo[i], jt + I_ZXU + i, id);
          Hd += z * id; gB += id; st += z;
        }
      }
      if (i < 2 && teq) {  // terminal equality row  x_N[i] - xref_N[i] = 0  with multiplier nu[i]
        Hd += 1.0 / MMPC_DELTA_C; gA += nu[i] + e / MMPC_DELTA_C; st += nu[i];
        A.prim = fmax(A.prim, fabs(e)); theta += fabs(e); sum_lam += fabs(nu[i]); n_eq += 1;
      }
      if (i < 3 || i >= 6) {
        const int a = (i < 3) ? i : i - 3;
        A.H[pidx(a, a)] = Hd + (i == 2 ? hpp : 0.0); A.gA[a] = gA; A.gB[a] = gB; A.st[a] = st;
      } else {
        Qw(k, Q_HVD + (i - 3)) = Hd; Qw(k, Q_GA + i) = gA; Qw(k, Q_GB + i) = gB;
        if (k >= 1) es = fmax(es, fabs(st));
      }
    }
#ifdef MMPC_POSEREF
    if (pose_model()) {  // the end-point pose cost at the candidate: value, gradient and exact Hessian into the pose block
      const double r[4] = {rb_in[0], rb_in[ps], rb_in[2 * ps], rb_in[3 * ps]};
      double g[NP];
      fsum += pose_cost(k, os, x[0], x[1], x[2], f, r, g, A.H);
#pragma unroll
      for (int a = 0; a < NP; ++a) { A.gA[a] += g[a]; A.st[a] += g[a]; }
    }
#endif
    fsum += os * cfg.S * s * s;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double Hd = 0, gA = 0, gB = 0;
      if (k < N) {
        double Rj = os * cfg.Rd[j], Wj = os * cfg.Wd[j];
        double e = u[j] - rb_in[(IN_UREF - IN_XREF + j) * ps], dl = u[j] - rb_in[(IN_ULAST - IN_XREF + j) * ps];
        fsum += os * (cfg.Rd[j] * e * e + cfg.Wd[j] * dl * dl);
        double gr = 2 * Rj * e + 2 * Wj * dl;
        Hd = 2 * Rj + 2 * Wj; gA = gr;
        double st = gr + stu[j];
        double lo = rb_in[(IN_ULO - IN_XREF + j) * ps], hi = rb_in[(IN_UHI - IN_XREF + j) * ps];
        if (is_fin(lo)) {
          double d = u[j] - lo, id, z = box(zb[(I_ZUL - I_ZXL + j) * ps], uo[j] - lo, d, duo[j], jt + I_ZUL + j, id);
          Hd += z * id; gB -= id; st -= z;
        }
        if (is_fin(hi)) {
          double d = hi - u[j], id, z = box(zb[(I_ZUU - I_ZXL + j) * ps], hi - uo[j], d, -duo[j], jt + I_ZUU + j, id);
          Hd += z * id; gB += id; st += z;
        }
        es = fmax(es, fabs(st));
      }
      Qw(k, Q_HUU + j) = Hd; Qw(k, Q_GA + SGY_U + j) = gA; Qw(k, Q_GB + SGY_U + j) = gB;
    }
#if !defined(MMPC_RESIDENT) && !defined(MMPC_NO_MARGIN_PREFETCH)
    // (reference NLP) the margin caches of stages k-1, k, k+1 for the stale-column rows at the end: prefetched into the slots
    // of the parked multipliers and inputs, which are free from here on, while the rows below are worked through (the
    // margin loads of stale_rows() were 10 % of the kernel's stall samples)
    const bool mg_pref = REF && npl >= 2 && 3 * 6 * cfg.n_pl <= (I_T - I_ZXL) + (S_DT - IN_XREF);
    if (mg_pref) margins_prefetch<3>(k, sm + (RING_DT * RING_W) * bs);
#endif
    // one slack row  h - s + t = 0 : candidate (t, z) with slack reset, merit and KKT bookkeeping
    double s_cur = s;  // slack the rows are bounded by (s[N-1] for the terminal self-collision rows of the literal reference NLP)
    auto row_core = [&](int r, double h, double t, double dtv, double& z, double& it_, double& res) {  // z: in = current multiplier
      double tt = fmax(fma(alpha, dtv, t), s_cur - h);  // slack reset (Nocedal & Wright 19.30)
      double dz = (mu - z * (t + dtv)) * rcp(t);
      it_ = rcp(tt);
      z = zclamp(z + ad * dz, mu, it_);
      cj[(I_T + r) << LSH] = tt; cj[(I_T + R + r) << LSH] = z;
      res = h - s_cur + tt;
      theta += fabs(res);
      if (tt <= 0) ok = false; else lp.mul(tt);
      A.prim = fmax(A.prim, fabs(res));
      double zt = z * tt;
      A.chi = fmax(A.chi, zt); A.clo = fmin(A.clo, zt);
      A.sumz += z; A.zrows += z; A.nz++;
      double sig = z * it_;
      A.csum += sig; A.be0 += sig * res; A.be1 += it_;
    };
#pragma unroll 1
    for (int i = 0; i < nobs; ++i) {  // obsAvoid :49-54
      async_wait<RING_DT - 1>();
      const double* rb = sm + ((i & (RING_DT - 1)) * RING_W) * bs;
      const double rt = rb[0], rdt = rb[2 * bs], ddx = x[0] - rb[3 * bs], ddy = x[1] - rb[4 * bs], rad = rb[5 * bs];
      double z = rb[bs], it_, res;
      ring_issue(i + RING_DT);
      double d2 = ddx * ddx + ddy * ddy, inv = rsq(d2), d = d2 * inv;
      double h = (rad + cfg.base_radius) - d;
      row_core(i, h, rt, rdt, z, it_, res);
      double sig = z * it_, nx = ddx * inv, ny = ddy * inv, zd = z * inv;
      A.H[pidx(0, 0)] += sig * nx * nx - zd * (1 - nx * nx);
      A.H[pidx(0, 1)] += (sig + zd) * nx * ny;
      A.H[pidx(1, 1)] += sig * ny * ny - zd * (1 - ny * ny);
      double cb = sig * res;
      A.a[0] += sig * nx; A.a[1] += sig * ny; A.gA[0] -= cb * nx; A.gA[1] -= cb * ny;
      A.gB[0] -= it_ * nx; A.gB[1] -= it_ * ny; A.st[0] -= z * nx; A.st[1] -= z * ny;
    }
    const bool q3n = REF && k == N && q3();  // terminal self-collision rows on s[N-1]
    double sv_c = 0, sv_b0 = 0, sv_b1 = 0, sv_z = 0, sv_a[NP] = {0, 0, 0, 0, 0, 0};
    if (q3n) {
      s_cur = fma(alpha, W2(N - 1, S_DS), W(N - 1, it + I_S));  // candidate s[N-1], as thread N-1 forms it
      sv_c = A.csum; sv_b0 = A.be0; sv_b1 = A.be1; sv_z = A.zrows; A.csum = A.be0 = A.be1 = A.zrows = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) { sv_a[a] = A.a[a]; A.a[a] = 0; }
    }
#pragma unroll 1
    for (int m = 0; m < nself; ++m) {  // self collision :219-222
      Point p; point_eval(x[0], x[1], f, SELFD[m], p);
      double d2 = p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2], inv = rsq(d2);
      double h = cfg.self_collision_radius - d2 * inv;
      async_wait<RING_DT - 1>();
      const double* rb = sm + (((nobs + m) & (RING_DT - 1)) * RING_W) * bs;
      const double rt = rb[0], rdt = rb[2 * bs];
      double z = rb[bs], it_, res;
      ring_issue(nobs + m + RING_DT);
      row_core(nobs + m, h, rt, rdt, z, it_, res);
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
    double q3_c = 0, q3_b0 = 0, q3_b1 = 0, q3_z = 0, q3_a[NP] = {0, 0, 0, 0, 0, 0};
    if (q3n) {
      s_cur = s;
      q3_c = A.csum; q3_b0 = A.be0; q3_b1 = A.be1; q3_z = A.zrows; A.csum = sv_c; A.be0 = sv_b0; A.be1 = sv_b1; A.zrows = sv_z;
#pragma unroll
      for (int a = 0; a < NP; ++a) { q3_a[a] = A.a[a]; A.a[a] = sv_a[a]; }
    }
    if (npl > 0) {
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {  // obsAvoidConvex :57-89 (proper row)
        Point p; point_eval(x[0], x[1], f, BODY[i], p);
        int jb; double h = plane_row(p, jb);
        async_wait<RING_DT - 1>();
        const double* rb = sm + (((nobs + nself + i) & (RING_DT - 1)) * RING_W) * bs;
        const double rt = rb[0], rdt = rb[2 * bs];
        double z = rb[bs], it_, res;
        ring_issue(nobs + nself + i + RING_DT);
        row_core(nobs + nself + i, h, rt, rdt, z, it_, res);
        double sig = z * it_;
        double n[3] = {PL(6 * jb + 3), PL(6 * jb + 4), PL(6 * jb + 5)}, g[NP];
        point_grad(f, p, n, g);  // the row is  -max c <= s, c = off - n.P  =>  grad h = +g
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
    double bv[NP] = {0, 0, 0, 0, 0, 0}, log_extra = 0;
    if (REF) {  // compiled out of the clean-mode kernels
      StaleIO io; io.reset(); io.rp.init(); io.rd.init();
#if defined(MMPC_RESIDENT) || defined(MMPC_NO_MARGIN_PREFETCH)
      margins_in_place(k);
#else
      if (mg_pref) async_wait<0>(); else margins_in_place(k);
#endif
      stale_rows<1>(k, A, bv, io);
      theta += io.theta; log_extra = io.lp.value(); ok = ok && io.ok;
    }
    // slack column of the stage Hessian: H[s][s] = 2S + sum sigma, H[pose][s] = -sum sigma grad h
    double S2 = 2 * os * cfg.S;
    Qw(k, Q_C) = S2 + A.csum;
    Qw(k, Q_GA + SGY_S) = S2 * s - A.be0;
    Qw(k, Q_GB + SGY_S) = -A.be1;
    Qw(k, Q_HVV) = q3_c; Qw(k, Q_GA + SGY_V) = -q3_b0; Qw(k, Q_GB + SGY_V) = -q3_b1;
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      Qw(k, Q_A + a) = A.a[a]; Qw(k, Q_BV + a) = q3n ? q3_a[a] : bv[a];
      Qw(k, Q_GA + POSE2X[a]) = A.gA[a]; Qw(k, Q_GB + POSE2X[a]) = A.gB[a];
      if (k >= 1) es = fmax(es, fabs(A.st[a]));
    }
#pragma unroll
    for (int e = 0; e < 21; ++e) Qw(k, Q_HP + e) = A.H[e];
    if (REF && q3() && k >= N - 1) {  // d/ds[N-1] of the Lagrangian: this thread's share, summed in the solve kernel
      if (k == N) { W2(N, S_DFC + 0) = q3_z; es = fmax(es, fabs(S2 * s - A.zrows)); }
      else W2(N, S_DFC + 1) = S2 * s - A.zrows;
    } else es = fmax(es, fabs(S2 * s - A.zrows));
    c2[(S_PART + 0) << LSH] = es; c2[(S_PART + 1) << LSH] = A.prim; c2[(S_PART + 2) << LSH] = A.chi; c2[(S_PART + 3) << LSH] = A.clo;
    c2[(S_PART + 4) << LSH] = sum_lam; c2[(S_PART + 5) << LSH] = A.sumz; c2[(S_PART + 6) << LSH] = (double)A.nz; c2[(S_PART + 7) << LSH] = (double)n_eq;
    bool fin = ok && (fsum == fsum) && (theta == theta);
    c2[(S_PART + PT_MERIT + 0) << LSH] = theta; c2[(S_PART + PT_MERIT + 1) << LSH] = fsum; c2[(S_PART + PT_MERIT + 2) << LSH] = fin ? lp.value() + log_extra : 0.0;
    c2[(S_PART + PT_MERIT + 3) << LSH] = fin ? 1.0 : 0.0;
    async_wait<0>();  // nothing of this item's ring may still be in flight when the thread primes the next one
  }

  // ctrl_trial (thread per instance): filter acceptance test (Waechter & Biegler 2006, Alg. A
  // without second-order correction).  Returns 0 = accepted (next: eval), 1 = rejected (next: another
  // trial with alpha/2), 2 = finished (line search failed).
  template <int NL>
  __device__ int ctrl_trial(int lane) {
    double th1 = 0, f1 = 0, logsum = 0, okv = 1.0;
    for (int k = lane; k <= N; k += NL) {
      th1 += W2(k, S_PART + PT_MERIT + 0); f1 += W2(k, S_PART + PT_MERIT + 1); logsum += W2(k, S_PART + PT_MERIT + 2);
      okv = fmin(okv, W2(k, S_PART + PT_MERIT + 3));
    }
    th1 = lanes_sum<NL>(th1); f1 = lanes_sum<NL>(f1); logsum = lanes_sum<NL>(logsum); okv = lanes_min<NL>(okv);
    bool ok = okv > 0.5;
    const double mu = D(D_MU), theta_k = D(D_THETA), phi0 = D(D_PHI0), gphi = D(D_GPHI), alpha = D(D_ALPHA);
    double ph1 = f1 - mu * logsum;
    int nfilt = J(J_NFILT);
    const int cur = J(J_CUR), itn = J(J_IT), lsn = J(J_LS);
    bool ftype = false;
    ok = ok && (th1 == th1) && (ph1 == ph1) && th1 < D(D_THMAX);
    // filter: an entry dominates the trial point if it is at least as good in both measures; entry q is checked by lane q mod NL
    double dom = 0.0;
    for (int q = lane; q < nfilt; q += NL)
      if (th1 >= D(D_FILT + 2 * q) && ph1 >= D(D_FILT + 2 * q + 1)) dom = 1.0;
    dom = lanes_max<NL>(dom);
    if (ok) {
      bool sw = (gphi < 0) && (alpha * pow(-gphi, 2.3) > pow(theta_k, 1.1));
      if (theta_k <= D(D_THMIN) && sw) {
        ok = ph1 <= phi0 + 1e-8 * alpha * gphi + 10 * 2.220446049250313e-16 * fabs(phi0); ftype = ok;
      } else {
        ok = (th1 <= (1 - 1e-5) * theta_k) || (ph1 <= phi0 - 1e-8 * theta_k); ftype = false;
      }
    }
    const int fr = J(J_FRST);
    const double ap0 = D(D_AP0);
    lanes_sync<NL>();  // every lane has read the instance state before lane 0 rewrites it
    if (!ok || dom > 0.5) {
      const int lastf = (ok && dom > 0.5) ? 1 : 0;  // sufficient progress, but the filter said no
      const int nres = (fr >> 8) & 0xff;
      // IPOPT's minimal step size (IpFilterLSAcceptor::CalculateAlphaMin): below it the line search is given up
      double alpha_min = 1e-5;
      if (gphi < 0) {
        alpha_min = fmin(1e-5, 1e-8 * theta_k / (-gphi));
        if (theta_k <= D(D_THMIN)) alpha_min = fmin(alpha_min, pow(theta_k, 1.1) / pow(-gphi, 2.3));
      }
      alpha_min *= 0.05;
      if (lsn + 1 >= 50 || alpha * 0.5 <= alpha_min) {
        // IPOPT would enter its restoration phase here; in its place: clear the filter (while resets are left) and search again
        if (nres < MAX_FILTER_RESETS && nfilt > 0) {
          if (lane == 0) { J(J_NFILT) = 0; J(J_FRST) = ((nres + 1) << 8) | (lastf << 16); D(D_ALPHA) = ap0; J(J_LS) = 0; }
          return 1;
        }
        if (lane == 0) finish(MMPC_STATUS_LINESEARCH);
        return 2;
      }
      if (lane == 0) { D(D_ALPHA) = alpha * 0.5; J(J_LS) = lsn + 1; J(J_FRST) = (fr & 0xffff) | (lastf << 16); }
      return 1;
    }
    if (lane == 0) {
      if (!ftype) {
        if (nfilt == 16) {
          for (int q = 0; q < 30; ++q) D(D_FILT + q) = D(D_FILT + q + 2);
          nfilt--;
        }
        D(D_FILT + 2 * nfilt) = (1 - 1e-5) * theta_k; D(D_FILT + 2 * nfilt + 1) = phi0 - 1e-8 * theta_k; nfilt++;
        J(J_NFILT) = nfilt;
      }
      J(J_CUR) = 1 - cur;
      J(J_IT) = itn + 1;
      J(J_STATE) = ST_ACTIVE;
    }
    return 0;
  }
};

// state (ST_*) of instance b
__device__ __forceinline__ int& inst_state(const SParams& P, int b) {
  return P.gi[((((long long)(b >> 5)) * J_NFIELDS + J_STATE) << 5) + (b & 31)];
}

// ---- lists ---------------------------------------------------------------------------------------
// Two lists of instance indices, rebuilt in ascending order (so that the gathers of the phase
// kernels stay as coalesced as the surviving instances allow) by an ordered compaction of the
// per-instance state: E = instances whose next phase is eval, T = instances whose next phase is a trial.
__device__ __forceinline__ int* list_E(const SParams& P) { return P.lists; }
__device__ __forceinline__ int* list_T(const SParams& P) { return P.lists + P.tsel * P.LS; }
__device__ __forceinline__ int* list_n(const SParams& P, int which) { return P.lists + which * P.LS; }
// The Riccati kernel's view of the E list (same entries, list 3): first the instances whose last factorisation needed no
// regularisation, then the ones that did.  Two instances share a warp in lock step and a warp repeats its sweep until both
// are factorised, so pairing like with like keeps the convex instances (3 of 4 iterations) out of the retries, and lets a
// warp of two non-convex ones abandon the delta_w = 0 attempt at the first pivot both have lost.
__device__ __forceinline__ int* list_S(const SParams& P) { return P.lists + 3 * P.LS; }

// ---- phase bodies on list items (shared by the kernels and by tests/emu) -----------------------------
__device__ inline void body_init(const SParams& P, int b) { Inst S(P, b); S.init(); }
template <bool REF>
__device__ inline void body_eval(const SParams& P, int j, int k) { Inst S(P, list_E(P)[j]); S.template eval<REF>(k); }
__device__ inline void body_solve(const SParams& P, int j) { Inst S(P, list_E(P)[j]); S.solve(); }
// doubles of shared memory one thread of the step / trial kernels needs for its row ring
constexpr int STAGED_RING_DOUBLES = Inst::RING_D * Inst::RING_W + STAGED_STEP_PARKED;  // row ring + parked stage inputs, FK cache and defect
constexpr int STAGED_TRIAL_RING_DOUBLES = Inst::RING_DT * Inst::RING_W + (I_T - I_ZXL) + (S_DT - IN_XREF);  // row ring + 28 bound multipliers + 29 stage inputs
template <bool REF>
__device__ inline void body_step(const SParams& P, int j, int k, double* sm, int bs) {
  Inst S(P, list_E(P)[j]); S.sm = sm; S.bs = bs;
  const int st = S.J(J_STATE);
  if (st == ST_FINISH) S.finish_stage(k);
  else if (st == ST_ACTIVE) S.template step<REF>(k);
}
template <int NL>
__device__ inline void body_ctrl_step(const SParams& P, int j, int lane) { Inst S(P, list_E(P)[j]); S.template ctrl_step<NL>(lane); }
template <bool REF>
__device__ inline void body_trial(const SParams& P, int j, int k, double* sm, int bs) {
  Inst S(P, list_T(P)[j]); S.sm = sm; S.bs = bs;
  if (P.fused) S.template trial_eval<REF>(k); else S.trial(k);
}
// which: 0 = the E list at the current iterate (front of the stand-alone evaluation), 1 = the trial list at the candidate
__device__ inline void body_pose(const SParams& P, int j, int k, int which) {
  Inst S(P, which ? list_T(P)[j] : list_E(P)[j]);
  S.pose_pass(k, which != 0);
}
template <int NL>
__device__ inline void body_ctrl_trial(const SParams& P, int j, int lane) { Inst S(P, list_T(P)[j]); S.template ctrl_trial<NL>(lane); }

#ifdef MMPC_EMULATE_LANE
// ordered compaction: list `dst` <- the entries of list `src` whose instance is in state `want`
inline void compact_list(const SParams& P, int dst, int src, int want) {
  int n = 0; int* out = list_n(P, dst); const int* in = list_n(P, src);
  for (int i = 0; i < P.cnt[src]; ++i) if (inst_state(P, in[i]) == want) out[n++] = in[i];
  P.cnt[dst] = n;
  if (dst == 0) for (int i = 0; i < n; ++i) list_S(P)[i] = out[i];  // (the order of the Riccati's list does not change any result)
}
#elif !defined(MMPC_RESIDENT)   // (the resident build has its own single kernel, mmpc_resident.cu)
// Ordered compaction  list dst <- { b in list src : state(b) == want }  by one block of 1024 threads: warp w
// scans a contiguous chunk of the source list 32 entries at a time (coalesced), counts with ballots, the warp
// totals are scanned through shared memory, and the second pass writes the survivors in order.  Every
// active instance is in the previous round's trial list, so the source shrinks with the active set.
// regularised(b): the factorisation of instance b's last iteration needed delta_w > 0
__device__ __forceinline__ bool inst_regularised(const SParams& P, int b) {
  return P.gi[((((long long)(b >> 5)) * J_NFIELDS + J_REGF) << 5) + (b & 31)] != 0;
}
__global__ void __launch_bounds__(1024) staged_compact_kernel(const __grid_constant__ SParams P, int dst, int src, int want) {
  __shared__ int wtot[32], wreg[32];
  int* out = list_n(P, dst); const int* in = list_n(P, src);
  const bool split = dst == 0;  // the E list also gets its Riccati ordering (list_S)
  int* outs = list_S(P);
  const int n = P.cnt[src];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const int chunk = ((n + blockDim.x - 1) / blockDim.x) * 32;  // entries per warp, multiple of 32
  const int lo = warp * chunk, hi = min(n, lo + chunk);
  int c = 0, cr = 0;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    int i = i0 + lane;
    int b = i < hi ? in[i] : 0;
    bool f = i < hi && inst_state(P, b) == want;
    c += __popc(__ballot_sync(FULL, f));
    if (split) cr += __popc(__ballot_sync(FULL, f && inst_regularised(P, b)));
  }
  if (lane == 0) { wtot[warp] = c; wreg[warp] = cr; }
  __syncthreads();
  int off = 0, tot = 0, offr = 0, totr = 0;
  for (int w2 = 0; w2 < nw; ++w2) { int v = wtot[w2], vr = wreg[w2]; if (w2 < warp) { off += v; offr += vr; } tot += v; totr += vr; }
  int off0 = off - offr;          // position among the unregularised ones
  offr += tot - totr;             // the regularised ones follow them
  for (int i0 = lo; i0 < hi; i0 += 32) {
    int i = i0 + lane;
    int b = i < hi ? in[i] : 0;
    bool f = i < hi && inst_state(P, b) == want;
    unsigned m = __ballot_sync(FULL, f);
    if (f) out[off + __popc(m & ((1u << lane) - 1))] = b;
    off += __popc(m);
    if (split) {
      const bool r = f && inst_regularised(P, b);
      unsigned mr = __ballot_sync(FULL, r), m0 = m & ~mr;
      if (r) outs[offr + __popc(mr & ((1u << lane) - 1))] = b;
      else if (f) outs[off0 + __popc(m0 & ((1u << lane) - 1))] = b;
      offr += __popc(mr); off0 += __popc(m0);
    }
  }
  if (threadIdx.x == 0) P.cnt[dst] = tot;
}

// ---- kernels: grid-stride loops over the device-side list counts --------------------------------------
__global__ void __launch_bounds__(128) staged_init_kernel(const __grid_constant__ SParams P) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int B = P.io->B;
  if (b < B) { body_init(P, b); list_n(P, 1)[b] = b; }  // every instance starts in the first source list
  if (b == 0) { P.cnt[1] = B; P.cnt[3] = 0; }
}
template <bool REF>
__global__ void __launch_bounds__(128) staged_eval_kernel(const __grid_constant__ SParams P) {
  const int n = P.cnt[0];
  const long long tot = (long long)n * (P.cfg.N + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x)
    body_eval<REF>(P, (int)(t % n), (int)(t / n));
}
__global__ void __launch_bounds__(64) staged_solve_kernel(const __grid_constant__ SParams P) {
  const int n = P.cnt[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) body_solve(P, j);
}
// resident blocks per SM the step / trial kernels are compiled for (A/B on B200: 3 beats 1 and 4)
#ifndef MMPC_STEP_MINB
#define MMPC_STEP_MINB 4
#endif
#ifndef MMPC_TRIAL_MINB
#define MMPC_TRIAL_MINB 2
#endif
template <bool REF>
__global__ void __launch_bounds__(128, REF ? MMPC_STEP_MINB - 1 : MMPC_STEP_MINB) staged_step_kernel(const __grid_constant__ SParams P) {
  extern __shared__ double ring[];  // STAGED_RING_DOUBLES per thread, thread-interleaved
  const int n = P.cnt[0];
  const long long tot = (long long)n * (P.cfg.N + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x)
    body_step<REF>(P, (int)(t % n), (int)(t / n), ring + threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(128) staged_ctrl_step_kernel(const __grid_constant__ SParams P) {
  const int n = P.cnt[0];  // one warp per instance
  // (the loop condition is a vote: a loop that depends on threadIdx is "divergent" to the compiler, which then brackets every
  // shuffle of the body with WARPSYNC.COLLECTIVE / ENDCOLLECTIVE; a vote result is known to be warp-uniform)
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; __all_sync(FULL, j < n); j += (gridDim.x * blockDim.x) >> 5)
    body_ctrl_step<32>(P, j, threadIdx.x & 31);
}
template <bool REF>
__global__ void __launch_bounds__(128, MMPC_TRIAL_MINB) staged_trial_kernel(const __grid_constant__ SParams P) {
  extern __shared__ double ring[];  // STAGED_RING_DOUBLES per thread, thread-interleaved
  const int n = P.cnt[P.tsel];
  const long long tot = (long long)n * (P.cfg.N + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x)
    body_trial<REF>(P, (int)(t % n), (int)(t / n), ring + threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(128) staged_pose_kernel(const __grid_constant__ SParams P, int which) {
  const int n = which ? P.cnt[P.tsel] : P.cnt[0];
  const long long tot = (long long)n * (P.cfg.N + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x)
    body_pose(P, (int)(t % n), (int)(t / n), which);
}
__global__ void __launch_bounds__(128) staged_ctrl_trial_kernel(const __grid_constant__ SParams P) {
  const int n = P.cnt[P.tsel];  // one warp per instance
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; __all_sync(FULL, j < n); j += (gridDim.x * blockDim.x) >> 5)   // (vote: see staged_ctrl_step_kernel)
    body_ctrl_trial<32>(P, j, threadIdx.x & 31);
}
#endif

}  // namespace mmpc
