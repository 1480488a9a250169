// emu_staged_main.cpp -- TEST INFRASTRUCTURE: runs the phase bodies of csrc/mmpc_staged.cuh on the
// CPU, one work item at a time, in the same round order as the CUDA host loop (mmpc_api.cu), with
// the same C ABI argument layout as mmpc_solve (host pointers).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "emu_lane_runtime.h"
#include "../../mobile_manipulator_mpc_b200/csrc/mmpc_team.cuh"
#include "../../mobile_manipulator_mpc_b200/csrc/mmpc_parts.cuh"
#include "../../mobile_manipulator_mpc_b200/csrc/mmpc_episode.cuh"

namespace mmpc { EmuTeam* g_team = nullptr; }
using namespace mmpc;

// one instance of the team phase: 16 coroutines in lock step
struct TeamArg { const SParams* P; int j; };
static TeamArg g_targ;
static void team_entry() {
  EmuTeam* w = g_team;
  int me = w->cur;
  static double ring[Team::SMEM_DOUBLES];
  const SParams& TP = *g_targ.P;
  if (TP.cfg.mode == MMPC_MODE_REFERENCE && TP.cfg.terminal_rows_on_sN == 0) body_solve_team<true>(TP, g_targ.j, me, ring);
  else body_solve_team<false>(TP, g_targ.j, me, ring);
  w->done++;
  if (w->done < 16) { int nx = (me + 1) & 15; w->cur = nx; swapcontext(&w->ctx[me], &w->ctx[nx]); }
  else swapcontext(&w->ctx[me], &w->main_ctx);
}
static void run_team(const SParams& P, int j, std::vector<char*>& stacks) {
  EmuTeam* w = g_team;
  g_targ.P = &P; g_targ.j = j;
  const size_t STK = 1 << 18;
  for (int i = 0; i < 16; ++i) {
    getcontext(&w->ctx[i]);
    w->ctx[i].uc_stack.ss_sp = stacks[i]; w->ctx[i].uc_stack.ss_size = STK; w->ctx[i].uc_link = &w->main_ctx;
    makecontext(&w->ctx[i], team_entry, 0);
  }
  w->cur = 0; w->done = 0;
  swapcontext(&w->main_ctx, &w->ctx[0]);
}

extern "C" int mmpc_emu_staged_solve(const MmpcConfig* cfg, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out,
                                     int32_t* rounds_out, int32_t team, int32_t fused, int32_t parts) {
  SParams P; memset(&P, 0, sizeof P);
  static double row_ring[STAGED_RING_DOUBLES > STAGED_TRIAL_RING_DOUBLES ? STAGED_RING_DOUBLES : STAGED_TRIAL_RING_DOUBLES];  // the step / trial row ring of the one emulated thread
  P.cfg = *cfg; P.B = B;
  SIO io;
  io.x_init = in->x_init; io.x_ref = in->x_ref; io.u_ref = in->u_ref; io.u_last = in->u_last; io.u_guess = in->u_guess;
  io.circles = in->circles; io.planes = in->planes; io.n_pl_inst = in->n_pl_inst; io.flags = in->flags; io.x_guess = in->x_guess;
  io.U = out->U; io.X = out->X; io.s = out->s; io.cost = out->cost; io.kkt = out->kkt; io.iters = out->iters; io.status = out->status;
  io.B = B;
  P.io = &io;
  P.R = staged_rows(*cfg); P.ITSZ = staged_itsz(*cfg); P.STG = staged_stage_doubles(*cfg); P.LS = (B + 31) / 32 * 32; P.ND = staged_inst_doubles(*cfg);
  std::vector<double> ws((size_t)(cfg->N + 1) * P.STG * P.LS, 0.0), gd((size_t)staged_inst_doubles(*cfg) * P.LS, 0.0);
  std::vector<double> qp((size_t)(cfg->N + 1) * QS * P.LS, 0.0), rk((size_t)(cfg->N + 1) * RS * P.LS, 0.0);
  P.qp = qp.data(); P.rk = rk.data(); P.team = team; P.fused = fused; P.parts = parts;
  EmuTeam* tw = new EmuTeam(); g_team = tw;
  std::vector<char*> stacks(16);
  for (int i = 0; i < 16; ++i) stacks[i] = (char*)malloc(1 << 18);
  std::vector<int> gi((size_t)J_NFIELDS * P.LS, 0), lists((size_t)4 * P.LS, 0);
  int cnt[4] = {0, 0, 0, 0};
  P.ws = ws.data(); P.gd = gd.data(); P.gi = gi.data(); P.lists = lists.data(); P.cnt = cnt;
  for (int b = 0; b < B; ++b) { body_init(P, b); lists[P.LS + b] = b; }
  cnt[1] = B;
  const bool ref = cfg->mode == MMPC_MODE_REFERENCE;
  if (ref || cfg->model != MMPC_MODEL_WHOLEBODY) parts = 0;  // the part kernels implement the clean whole-body NLP only
  int N = cfg->N, r = 0;
  for (;; ++r) {
    const int tcur = 1 + (r & 1), tnext = 1 + ((r + 1) & 1);
    compact_list(P, 0, tcur, ST_ACTIVE);
    int nE = cnt[0];
    if (ref && (!fused || r == 0)) for (int k = 0; k <= N; ++k) for (int j = 0; j < nE; ++j) body_pose(P, j, k, 0);
    if (!fused || r == 0)  // fused: only the starting point needs the stand-alone evaluation
      for (int k = 0; k <= N; ++k) for (int j = 0; j < nE; ++j) { if (ref) body_eval<true>(P, j, k); else body_eval<false>(P, j, k); }
    for (int j = 0; j < nE; ++j) { if (team) run_team(P, j, stacks); else body_solve(P, j); }
    for (int k = 0; k <= N; ++k) for (int j = 0; j < nE; ++j) {
      if (!parts) { if (ref) body_step<true>(P, j, k, row_ring, 1); else body_step<false>(P, j, k, row_ring, 1); }
      else if (inst_state(P, list_E(P)[j]) == ST_ACTIVE) body_parts_item(P, list_E(P)[j], k, false);
      else if (inst_state(P, list_E(P)[j]) == ST_FINISH) { Inst F(P, list_E(P)[j]); F.finish_stage(k); }
    }
    for (int j = 0; j < nE; ++j) body_ctrl_step<1>(P, j, 0);
    compact_list(P, tnext, tcur, ST_TRIAL);
    P.tsel = tnext;
    int nT = cnt[tnext];
    if (ref) for (int k = 0; k <= N; ++k) for (int j = 0; j < nT; ++j) body_pose(P, j, k, 1);
    for (int k = 0; k <= N; ++k) for (int j = 0; j < nT; ++j) {
      if (!parts) { if (ref) body_trial<true>(P, j, k, row_ring, 1); else body_trial<false>(P, j, k, row_ring, 1); }
      else body_parts_item(P, list_T(P)[j], k, true);
    }
    for (int j = 0; j < nT; ++j) body_ctrl_trial<1>(P, j, 0);
    if (nT == 0) break;
    if (r > 200000) return 1;
  }
  for (int i = 0; i < 16; ++i) free(stacks[i]);
  delete tw; g_team = nullptr;
  if (rounds_out) *rounds_out = r + 1;
  return 0;
}

// csrc/mmpc_episode.cuh on the CPU: the bodies of ik_kernel and episode_update_kernel, one episode at a time (host pointers)
extern "C" int mmpc_emu_ik(int32_t B, const double* q_guess, const double* target, double* q_out, int32_t* status) {
  for (int b = 0; b < B; ++b) {
    int rc = ik_solve(q_guess + (size_t)b * 3, target[(size_t)b * 3 + 0], target[(size_t)b * 3 + 2], q_out + (size_t)b * 3);
    if (status) status[b] = rc;
  }
  return 0;
}
extern "C" int mmpc_emu_episode_update(int32_t N, int32_t B, int32_t M, int32_t n_manip, const MmpcEpisodeIO* io) {
  EpisodeArgs A; A.B = B; A.N = N; A.M = M; A.n_manip = n_manip; A.io = *io;
  for (int b = 0; b < B; ++b) episode_update(A, b);
  return 0;
}
