// mmpc_resident.cu -- the RESIDENT solver: one thread block owns one instance for its whole solve, with every bit of the
// instance's solver state -- iterate (two copies), multipliers, step, stage QP, Riccati factors, per-instance scalars -- in
// SHARED MEMORY (north_star: "state ... staged through shared memory").  HBM sees the instance's
// inputs once and its outputs once: the algorithmic traffic of section 8(d).  Blocks are persistent and pull instances from an
// atomic work queue, so there are no lists, no compaction, no rounds, and no host in the loop: one launch per solve.
//
// It is the LATENCY path (a single controller, closed loops of up to a couple of thousand robots).  A B200 SM holds two
// instances of this NLP in shared memory (112 KB each at N = 20 with 16 circles), so 296 instances are in flight and an
// iteration costs its dependent-instruction latency (about 115 us); batches beyond ~2,000 instances are faster through the
// streaming "staged" kernels, which keep 24 instances per SM in flight.  mmpc_api.cu picks by batch size.
//
// The inputs of an instance (6.2 KB) come in by TMA: cp.async.bulk + mbarrier into a staging area that aliases the Riccati
// records (dead while an instance is set up), see slice_in() / inputs_issue().
//
// The arithmetic is the staged solver's, literally: this file compiles the same phase bodies (mmpc_staged.cuh, mmpc_team.cuh)
// with MMPC_RESIDENT defined, which turns the tile-major HBM layout into a stride-1 layout (LSH = 0), global-memory loads
// and cp.async staging into plain shared-memory accesses, and gives Inst a separate index for the caller's arrays.  On the
// reference NLP the results equal the staged solver's to the bit (tests/test_gpu_parity.py).  Thread roles inside the block:
// warp 0 = the 16-lane Riccati team -- its second half-warp factorises with the NEXT delta_w of the inertia-correction
// sequence at the same time (Team::solve_spec) -- and the per-instance control steps; threads 32 .. 32 + N = one per stage
// (set-up, evaluation, step, trial).  Phases are separated by __syncthreads().
// mmpc_resident_pose.cu compiles this file a second time with the pose-reference controller's cost (MMPC_POSEREF).
#define MMPC_RESIDENT 1
#define mmpc mmpc_res   // its own namespace: the same inline function names are compiled differently in mmpc_api.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mmpc.h"
#include "mmpc_team.cuh"   // includes mmpc_staged.cuh

using namespace mmpc_res;

namespace mmpc_res {

// ---- TMA (1-D bulk copy) + mbarrier: the instance's inputs come into shared memory by cp.async.bulk, issued by one thread ----
__device__ __forceinline__ unsigned sptr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy accesses of shared memory before, async-proxy (bulk copy) writes after
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sptr(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(sptr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(sptr(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// A caller's array holds n doubles per instance, so an instance's slice starts on an 8-byte boundary and a bulk copy wants 16:
// element i of the slice goes to region[off + i] with off = 1 when the slice starts on an odd double -- then the 16-byte
// aligned middle of the slice lands 16-byte aligned in the (16-byte aligned) region and goes by ONE bulk copy; the odd
// double in front and / or behind it goes by a plain copy of the issuing thread (released by its mbarrier arrive).
__device__ __forceinline__ int slice_off(const double* src) { return (int)(((unsigned long long)src >> 3) & 1ull); }
__device__ __forceinline__ unsigned slice_in(double* region, const double* src, int n, unsigned long long* bar) {
  const int head = slice_off(src);
  const int mid = (n - head) & ~1;
  if (head) region[1] = src[0];
  if (n - head - mid) region[head + n - 1] = src[n - 1];
  if (mid > 0) bulk_g2s(region + 2 * head, src + head, (unsigned)mid * 8u, bar);
  return (unsigned)mid * 8u;
}

// shared-memory plan of one block, in doubles
struct ResPlan {
  int ITSZ;        // doubles of one copy of a stage's iterate, at least RS: the dead copy doubles as the second set of Riccati records
  int STGp;        // stage stride, padded to an odd number of doubles: stage threads hit different banks
  int o_ws, o_qp, o_rk, o_gd, o_gi, o_team, o_ring, total;
  int stage_threads, threads;
  // input staging (aliases the Riccati records, which are dead while an instance is set up): regions of n + 2 doubles
  int i_xinit, i_xref, i_uref, i_ulast, i_uguess, i_circ, i_planes, i_xguess, i_total;
};
__host__ __device__ inline ResPlan res_plan(const MmpcConfig& c) {
  ResPlan p;
  const int K1 = c.N + 1;
  p.ITSZ = staged_itsz(c) > RS ? staged_itsz(c) : RS;
  p.STGp = (staged_stage_doubles(c) + 2 * (p.ITSZ - staged_itsz(c))) | 1;
  p.stage_threads = (K1 + 31) / 32 * 32;
  p.threads = 32 + p.stage_threads;
  int i = 0;
  auto region = [&](int n) { int at = i; i += (n + 3) & ~1; return at; };
  p.i_xinit = region(NX); p.i_xref = region(K1 * NX); p.i_uref = region(c.N * NU); p.i_ulast = region(c.N * NU);
  p.i_uguess = region(c.N * NU); p.i_circ = region((c.obs_per_stage ? K1 : 1) * 3 * c.n_obs); p.i_planes = region(6 * c.n_pl);
  p.i_xguess = region(K1 * NX); p.i_total = i;
  int o = 0;
  p.o_ws = o; o += K1 * p.STGp;
  o = (o + 1) & ~1;
  p.o_rk = o; o += (K1 * RS > p.i_total ? K1 * RS : p.i_total);   // 16-byte aligned: bulk copy destination
  p.o_qp = o; o += K1 * QS;
  p.o_gd = o; o += staged_inst_doubles(c);
  p.o_gi = o; o += (J_NFIELDS + 1) / 2 + 1;          // ints, two per double
  o = (o + 1) & ~1;                                   // the team ring is read in 16-byte pieces
  p.o_team = o; o += Team::SMEM_DOUBLES;              // the team's ring: two backward-sweep rings (one per half-warp) or one roll-out ring
  p.o_ring = o; o += Inst::RING_D * Inst::RING_W * p.stage_threads;   // the per-thread row rings of the step / trial bodies
  p.total = o;
  return p;
}

// One thread: bulk copies of instance b's inputs into the staging regions, then its arrive with the byte count.
__device__ inline void inputs_issue(const SParams& P, const ResPlan& pl, double* in, int b, unsigned long long* bar) {
  const MmpcConfig& c = P.cfg; const SIO& io = *P.io;
  const int K1 = c.N + 1;
  unsigned bytes = 0;
  bytes += slice_in(in + pl.i_xinit, io.x_init + (long long)b * NX, NX, bar);
  bytes += slice_in(in + pl.i_xref, io.x_ref + (long long)b * K1 * NX, K1 * NX, bar);
  bytes += slice_in(in + pl.i_uref, io.u_ref + (long long)b * c.N * NU, c.N * NU, bar);
  bytes += slice_in(in + pl.i_ulast, io.u_last + (long long)b * c.N * NU, c.N * NU, bar);
  if (io.u_guess) bytes += slice_in(in + pl.i_uguess, io.u_guess + (long long)b * c.N * NU, c.N * NU, bar);
  const int ncirc = (c.obs_per_stage ? K1 : 1) * 3 * c.n_obs;
  if (ncirc > 0) bytes += slice_in(in + pl.i_circ, io.circles + (long long)b * ncirc, ncirc, bar);
  if (c.n_pl > 0) bytes += slice_in(in + pl.i_planes, io.planes + (long long)b * 6 * c.n_pl, 6 * c.n_pl, bar);
  if (io.x_guess) bytes += slice_in(in + pl.i_xguess, io.x_guess + (long long)b * K1 * NX, K1 * NX, bar);
  mbar_arrive_expect(bar, bytes);
}

// ---- stage-parallel set-up of an instance: Inst::init() (mmpc_staged.cuh, one thread per instance) split into three passes
// of one thread per stage; the values, and the order of the operations that produce each of them, are init()'s.  Reads the
// staged inputs; pass A also leaves the stage's plane margins in the margin cache for pass B's stale-column rows.
struct InPtr { const double *xinit, *xref, *uref, *ulast, *uguess, *circ, *xguess; };
__device__ __forceinline__ InPtr in_ptrs(const SParams& P, const ResPlan& pl, const double* in, int b) {
  const MmpcConfig& c = P.cfg; const SIO& io = *P.io;
  const int K1 = c.N + 1;
  InPtr q;
  q.xinit = in + pl.i_xinit + slice_off(io.x_init + (long long)b * NX);
  q.xref = in + pl.i_xref + slice_off(io.x_ref + (long long)b * K1 * NX);
  q.uref = in + pl.i_uref + slice_off(io.u_ref + (long long)b * c.N * NU);
  q.ulast = in + pl.i_ulast + slice_off(io.u_last + (long long)b * c.N * NU);
  q.uguess = io.u_guess ? in + pl.i_uguess + slice_off(io.u_guess + (long long)b * c.N * NU) : nullptr;
  q.circ = in + pl.i_circ + slice_off(io.circles + (long long)b * (c.obs_per_stage ? K1 : 1) * 3 * c.n_obs);
  q.xguess = io.x_guess ? in + pl.i_xguess + slice_off(io.x_guess + (long long)b * K1 * NX) : nullptr;
  return q;
}
// per-instance tables (planes, static circles): any thread t of nt
__device__ inline void init_tables(Inst& S, const ResPlan& pl, const double* in, const InPtr& q, int t, int nt) {
  const MmpcConfig& cfg = S.cfg;
  const double* planes = in + pl.i_planes + slice_off(S.P.io->planes + (long long)S.bio * 6 * cfg.n_pl);
  for (int i = t; i < 6 * cfg.n_pl; i += nt) S.D(D_PL + i) = planes[i];
  if (!cfg.obs_per_stage)
    for (int i = t; i < 3 * S.nobs; i += nt) S.D(D_CIRC + i) = q.circ[i];
}
__device__ inline void init_pass_a(Inst& S, const InPtr& q, int k) {
  const MmpcConfig& cfg = S.cfg;
  const int N = S.N, R = S.R, nobs = S.nobs, nself = S.nself, npl = S.npl;
  double gmax = 0;
  const bool refmode = cfg.mode == MMPC_MODE_REFERENCE && npl > 1;
  double x[NX];
  if (cfg.obs_per_stage)
    for (int i = 0; i < 3 * nobs; ++i) S.W2(k, S_DT + R + i) = q.circ[k * 3 * nobs + i];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v = fmax(fmin(q.xinit[i], cfg.xlim[1][i]), cfg.xlim[0][i]);
    if (k >= 1 && q.xguess) v = q.xguess[k * NX + i];
    if (k >= 1) v = push_in(v, cfg.xlim[0][i], cfg.xlim[1][i]);
    double xr = q.xref[k * NX + i];
    x[i] = v; S.W(k, I_X + i) = v; S.W(k, I_LAM + i) = 0; S.W2(k, IN_XREF + i) = xr;
    S.W(k, I_ZXL + i) = 1; S.W(k, I_ZXU + i) = 1;
    double Wx = S.xweight(k, i);
    if (k >= 1) gmax = fmax(gmax, fabs(2 * Wx * S.xerr(i, v, xr)));
  }
#ifdef MMPC_POSEREF
  if (S.pose_model() && k >= 1) {   // gradient of the end-point pose cost at the starting point (objective scaling)
    FK fp; fk_eval(x[2], x[6], x[7], x[8], fp);
    const double r[4] = {q.xref[k * NX + 0], q.xref[k * NX + 1], q.xref[k * NX + 2], q.xref[k * NX + 3]};
    double g[NP];
    S.pose_cost(k, 1.0, x[0], x[1], x[2], fp, r, g, nullptr);
    for (int a = 0; a < NP; ++a) gmax = fmax(gmax, fabs(g[a]));
  }
#endif
  if (k < N) {
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double ul = q.ulast[k * NU + j], ur = q.uref[k * NU + j];
      double lo = fmax(cfg.ulim[0][j], ul + cfg.dulim[0][j]);
      double hi = fmin(cfg.ulim[1][j], ul + cfg.dulim[1][j]);
      double v = q.uguess ? q.uguess[k * NU + j] : ul;
      v = push_in(v, lo, hi);
      S.W(k, I_U + j) = v; S.W2(k, IN_UREF + j) = ur; S.W2(k, IN_ULAST + j) = ul; S.W2(k, IN_ULO + j) = lo; S.W2(k, IN_UHI + j) = hi;
      S.W(k, I_ZUL + j) = 1; S.W(k, I_ZUU + j) = 1;
      gmax = fmax(gmax, fabs(2 * cfg.Rd[j] * (v - ur) + 2 * cfg.Wd[j] * (v - ul)));
    }
  }
  FK f; fk_eval(x[2], x[6], x[7], x[8], f);
  double hmax = -1e300;
  for (int i = 0; i < nobs; ++i) {
    double ddx = x[0] - S.circ<false>(k, i, 0), ddy = x[1] - S.circ<false>(k, i, 1);
    double h = (S.circ<false>(k, i, 2) + cfg.base_radius) - sqrt(ddx * ddx + ddy * ddy);
    S.W(k, I_T + i) = h; hmax = fmax(hmax, h);
  }
#pragma unroll 1
  for (int m = 0; m < nself; ++m) {
    Point p; point_eval(x[0], x[1], f, SELFD[m], p);
    double h = cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]);
    S.W(k, I_T + nobs + m) = h;
    if (!(k == N && S.q3())) hmax = fmax(hmax, h);
  }
  if (npl > 0) {
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
      Point p; point_eval(x[0], x[1], f, BODY[i], p);
      int jb; double h = S.plane_row<false>(p, jb);
      S.W(k, I_T + nobs + nself + i) = h; hmax = fmax(hmax, h);
    }
  }
  if (staged_stale_rows(cfg) > 0 && refmode) {
    double pp[NP] = {x[0], x[1], x[2], x[6], x[7], x[8]}; FK ff;
    double c[6][MMPC_MAX_PLANES];
    S.margins(pp, ff, c);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < npl; ++j) S.W2(k, S.MG + i * cfg.n_pl + j) = c[i][j];
  }
  S.W2(k, S_PART + 0) = hmax; S.W2(k, S_PART + 1) = gmax;
}
__device__ inline void init_pass_b(Inst& S, int k) {
  const MmpcConfig& cfg = S.cfg;
  const int R = S.R, nobs = S.nobs, nself = S.nself, npl = S.npl;
  const bool refmode = cfg.mode == MMPC_MODE_REFERENCE && npl > 1;
  double hmax = S.W2(k, S_PART + 0), gmax = S.W2(k, S_PART + 1);
  if (staged_stale_rows(cfg) > 0) {
    const int nst = cfg.n_pl - 1, r0 = nobs + nself + 6;
    for (int r = r0; r < R; ++r) S.W(k, I_T + r) = 0;
    if (refmode && k >= 1) {
      double ccur[6][MMPC_MAX_PLANES], cprev[6][MMPC_MAX_PLANES];
      S.load_margins(k, ccur); S.load_margins(k - 1, cprev);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < npl - 1; ++j) {
          int jb; double h = S.stale_max(ccur, cprev, i, j, jb);
          S.W(k, I_T + r0 + i * nst + j) = h; hmax = fmax(hmax, h);
        }
    }
  }
  double s = fmax(0.0, hmax + 1e-2);
  S.W(k, I_S) = s;
  gmax = fmax(gmax, fabs(2 * cfg.S * s));
  S.W2(k, S_PART + 1) = gmax;
}
__device__ inline void init_pass_c(Inst& S, int k) {
  const int N = S.N, R = S.R, nobs = S.nobs, nself = S.nself;
  const double s = S.W(k, I_S);
  for (int r = 0; r < R; ++r) {
    const double sr = (k == N && S.q3() && r >= nobs && r < nobs + nself) ? S.W(k - 1, I_S) : s;
    S.W(k, I_T + r) = sr - S.W(k, I_T + r); S.W(k, I_T + R + r) = 1.0;
  }
  if (k == 0) {
    const MmpcConfig& cfg = S.cfg;
    double gmax = 0;
    for (int kk = 0; kk <= N; ++kk) gmax = fmax(gmax, S.W2(kk, S_PART + 1));
    S.D(D_OS) = (gmax > 100.0) ? fmax(100.0 / gmax, 1e-8) : 1.0;
    S.D(D_MU) = cfg.mu_init; S.D(D_REGLAST) = 0; S.D(D_THMAX) = -1; S.D(D_THMIN) = -1; S.D(D_E0) = 1e300;
    S.J(J_NPL) = S.npl;
    S.J(J_STATE) = ST_ACTIVE; S.J(J_IT) = 0; S.J(J_NFILT) = 0; S.J(J_LS) = 0; S.J(J_CUR) = 0; S.J(J_FRST) = 0; S.J(J_REGF) = 0;
    S.J(J_FLAGS) = S.P.io->flags ? (int)S.P.io->flags[S.bio] : 0;
  }
}

// The interior-point iterations of the instance in shared memory, until it has finished (its outputs are written by the
// phase bodies).  in_line_search: the instance comes from the staged solver in the middle of a line search (state ST_TRIAL):
// the first thing it needs is a trial, not a factorisation.
template <bool REF, bool Q3>
__device__ __forceinline__ void resident_iterate(Inst& S, const SParams& P, const ResPlan& pl, int b, bool team_warp, int lane, bool stage, int k,
                                                 double* team_ring, bool in_line_search) {
  for (;;) {
    if (!in_line_search) {
      // ---- KKT test, barrier update, Riccati factorisation (two delta_w at a time), roll-out of the Newton step ----
      if (team_warp) { Team T(P, b, lane & 15, team_ring); T.template solve_spec<Q3>(lane >> 4, &S.W(0, (1 - S.J(J_CUR)) * pl.ITSZ), pl.STGp); }
      __syncthreads();
      const int st = S.J(J_STATE);
      // ---- slack / multiplier steps, fraction to the boundary (or: the results of an instance that has just finished) ----
      if (stage) { if (st == ST_FINISH) S.finish_stage(k); else if (st == ST_ACTIVE) S.template step<REF>(k); }
      __syncthreads();
      if (team_warp) S.template ctrl_step<32>(lane);
      __syncthreads();
      if (S.J(J_STATE) != ST_TRIAL) break;      // finished: the outputs are written
    }
    in_line_search = false;
    // ---- filter line search: candidate + evaluation of the next iteration at the candidate ----
    for (;;) {
      if (REF) { if (stage) S.pose_pass(k, true); __syncthreads(); }
      if (stage) S.template trial_eval<REF>(k);
      __syncthreads();
      if (team_warp) S.template ctrl_trial<32>(lane);
      __syncthreads();
      if (S.J(J_STATE) != ST_TRIAL) break;    // accepted (ST_ACTIVE) or given up (ST_DONE)
    }
    if (S.J(J_STATE) != ST_ACTIVE) break;
  }
}

template <bool REF, bool Q3>
__global__ void __launch_bounds__(96, 1) resident_solve_kernel(const __grid_constant__ SParams P0, unsigned* queue) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_b;
  __shared__ __align__(8) unsigned long long s_bar;   // mbarrier: the inputs of the block's next instance have landed
  const ResPlan pl = res_plan(P0.cfg);
  const int tid = threadIdx.x, lane = tid & 31;
  // warp 0 is the team.  The role predicate is a VOTE, not tid >> 5: a branch on threadIdx is "divergent" to the compiler,
  // which then brackets every shuffle under it with WARPSYNC.COLLECTIVE / ENDCOLLECTIVE (439 brackets in this kernel); a vote
  // result is known to be warp-uniform
  const bool team_warp = __all_sync(0xffffffffu, tid < 32);
  const int N = P0.cfg.N;
  const int k = tid - 32;                       // stage of a stage thread
  const bool stage = tid >= 32 && k <= N;
  SParams P = P0;
  P.ws = smem + pl.o_ws; P.qp = smem + pl.o_qp; P.rk = smem + pl.o_rk; P.gd = smem + pl.o_gd; P.gi = (int*)(smem + pl.o_gi);
  P.LS = 1; P.STG = pl.STGp;
  double* const in = smem + pl.o_rk;            // input staging: the Riccati records are dead while an instance is set up
  double* team_ring = smem + pl.o_team;
  double* ring = smem + pl.o_ring + (tid - 32);
  const int B = P.io->B;
  if (tid == 0) mbar_init(&s_bar, 1);
  unsigned parity = 0;
  for (;;) {
    __syncthreads();                            // everybody is done with the previous instance (and with s_b)
    if (tid == 0) {
      const int nb = (int)atomicAdd(queue, 1u);
      s_b = nb;
      if (nb < B) { fence_async_smem(); inputs_issue(P, pl, in, nb, &s_bar); }
    }
    __syncthreads();
    const int b = s_b;
    if (b >= B) break;
    Inst S(P, b); S.sm = ring; S.bs = pl.stage_threads;
    S.npl = P.io->n_pl_inst ? P.io->n_pl_inst[b] : P.cfg.n_pl;
    S.npl = S.npl < 0 ? 0 : (S.npl > P.cfg.n_pl ? P.cfg.n_pl : S.npl);
    const InPtr q = in_ptrs(P, pl, in, b);
    mbar_wait(&s_bar, parity); parity ^= 1;     // the bulk copies have landed (and the issuing thread's odd doubles with them)
    // ---- the starting point (:302-304), bound push, slack lift, objective scaling: Inst::init(), one thread per stage ----
    init_tables(S, pl, in, q, tid, blockDim.x);
    __syncthreads();
    if (stage) init_pass_a(S, q, k);
    __syncthreads();
    if (stage) init_pass_b(S, k);
    __syncthreads();
    if (stage) init_pass_c(S, k);
    __syncthreads();
    if (REF) { if (stage) S.pose_pass(k, false); __syncthreads(); }
    if (stage) S.template eval<REF>(k);
    __syncthreads();
    resident_iterate<REF, Q3>(S, P, pl, b, team_warp, lane, stage, k, team_ring, false);
  }
}

// The TAIL of a staged solve (mmpc_api.cu, graph_build): when the active set has thinned out to what this kernel holds in
// flight, a staged round costs its launch and dependency latency (~300 us for seven kernels that each serve a handful of
// instances) while an iteration here costs ~120 us -- and a batch runs as many rounds as its slowest instance needs
// iterations (config 3: 274 rounds for a mean of 34).  Each block pulls a still-active instance of the last trial list,
// copies its state from the tile-major HBM workspace into shared memory and iterates it to the end.  Same phase bodies, same
// bits as if the staged rounds had gone on.
template <bool REF, bool Q3>
__global__ void __launch_bounds__(96, 1) resident_tail_kernel(const __grid_constant__ SParams P0, const __grid_constant__ ResTail T) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_b;
  const ResPlan pl = res_plan(P0.cfg);
  const int tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
  const bool team_warp = __all_sync(0xffffffffu, tid < 32);   // (a vote: see resident_solve_kernel)
  const int N = P0.cfg.N;
  const int k = tid - 32;
  const bool stage = tid >= 32 && k <= N;
  SParams P = P0;
  P.ws = smem + pl.o_ws; P.qp = smem + pl.o_qp; P.rk = smem + pl.o_rk; P.gd = smem + pl.o_gd; P.gi = (int*)(smem + pl.o_gi);
  P.LS = 1; P.STG = pl.STGp;
  double* team_ring = smem + pl.o_team;
  double* ring = smem + pl.o_ring + (tid - 32);
  const int n = T.cnt[2];
  for (;;) {
    __syncthreads();
    if (tid == 0) s_b = (int)atomicAdd(T.queue, 1u);
    __syncthreads();
    const int j = s_b;
    if (j >= n) break;
    const int b = T.list[j];
    const long long tile = b >> 5; const int ln = b & 31;
    const int* gih = T.gi + ((tile * J_NFIELDS) << 5) + ln;
    const int state = gih[J_STATE << 5];
    if (state != ST_ACTIVE && state != ST_TRIAL) continue;   // finished in the last staged round
    // ---- the instance's state: tile-major HBM (lane stride 32) -> shared memory (stride 1; the two iterate copies padded) ----
    for (int f = tid; f < J_NFIELDS; f += nt) P.gi[f] = gih[f << 5];
    const double* gdh = T.gd + ((tile * T.ND) << 5) + ln;
    for (int f = tid; f < T.ND; f += nt) P.gd[f] = gdh[f << 5];
    const double* wsh = T.ws + ((tile * (N + 1) * T.STG) << 5) + ln;
    for (int i = tid; i < (N + 1) * T.STG; i += nt) {
      const int kk = i / T.STG, f = i - kk * T.STG;
      const int f2 = f < T.ITSZ ? f : f < 2 * T.ITSZ ? pl.ITSZ + (f - T.ITSZ) : 2 * pl.ITSZ + (f - 2 * T.ITSZ);
      P.ws[kk * pl.STGp + f2] = wsh[(long long)i << 5];
    }
    for (int i = tid; i < (N + 1) * QS; i += nt) {
      const int kk = i / QS, f = i - kk * QS;
      P.qp[i] = T.qp[((long long)kk * T.LS + b) * QS + f];
    }
    __syncthreads();
    Inst S(P, b); S.sm = ring; S.bs = pl.stage_threads;
    S.load_npl();
    resident_iterate<REF, Q3>(S, P, pl, b, team_warp, lane, stage, k, team_ring, state == ST_TRIAL);
  }
}

}  // namespace mmpc_res

// ---- host side (called by mmpc_api.cu) ------------------------------------------------------------------------------------
extern "C" int mmpc_resident_smem_bytes(const MmpcConfig* cfg) {
  return (int)(res_plan(*cfg).total * sizeof(double));
}

static SParams resident_params(const MmpcConfig& c, int B, const void* io_dev) {
  const ResPlan pl = res_plan(c);
  SParams P; memset(&P, 0, sizeof P);
  P.cfg = c; P.B = B; P.io = (const SIO*)io_dev;
  P.team = 1; P.fused = 1; P.parts = 0;
  P.R = staged_rows(c); P.ITSZ = pl.ITSZ; P.STG = pl.STGp; P.ND = staged_inst_doubles(c); P.LS = 1;
  return P;
}

// The tail kernel of a staged solve as a graph node (mmpc_api.cu adds it): function, launch shape, and the kernel's first
// parameter (SParams, written to params_out).  blocks_in_flight = what the GPU holds at once: the hand-over threshold.
extern "C" int mmpc_resident_tail_node(const MmpcConfig* cfg, int B, const void* io_dev, int sm_count, const void** fn_out, int* threads_out,
                                       int* smem_out, int* blocks_in_flight, void* params_out) {
  const MmpcConfig& c = *cfg;
  const ResPlan pl = res_plan(c);
  const bool ref = c.mode == MMPC_MODE_REFERENCE, q3 = ref && c.terminal_rows_on_sN == 0;
  const void* fn = ref ? (q3 ? (const void*)resident_tail_kernel<true, true> : (const void*)resident_tail_kernel<true, false>)
                       : (const void*)resident_tail_kernel<false, false>;
  const int smem = (int)(pl.total * sizeof(double));
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, pl.threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  *fn_out = fn; *threads_out = pl.threads; *smem_out = smem; *blocks_in_flight = per_sm * sm_count;
  *(SParams*)params_out = resident_params(c, B, io_dev);
  return 0;
}

// Launches the resident solve of the B instances described by the device block `io_dev` (SIO of mmpc_staged.cuh; the resident
// and the staged build share its layout).  `queue` is a zeroed device counter.  Returns a cudaError_t.
extern "C" int mmpc_resident_launch(const MmpcConfig* cfg, int B, const void* io_dev, unsigned* queue, int max_blocks, void* stream) {
  const MmpcConfig& c = *cfg;
  const ResPlan pl = res_plan(c);
  SParams P = resident_params(c, B, io_dev);
  const bool ref = c.mode == MMPC_MODE_REFERENCE, q3 = ref && c.terminal_rows_on_sN == 0;
  const void* fn = ref ? (q3 ? (const void*)resident_solve_kernel<true, true> : (const void*)resident_solve_kernel<true, false>)
                       : (const void*)resident_solve_kernel<false, false>;
  const int smem = (int)(pl.total * sizeof(double));
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = 1;   // blocks (= instances) an SM holds: 2 at N = 20 with 16 circles (112 KB each)
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, pl.threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = B < max_blocks * per_sm ? B : max_blocks * per_sm;
  void* args[] = {&P, &queue};
  return (int)cudaLaunchKernel(fn, dim3(grid), dim3(pl.threads), args, smem, (cudaStream_t)stream);
}
