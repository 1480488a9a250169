// emu_main.cpp -- TEST INFRASTRUCTURE: runs Solver::run() of csrc/mmpc_solver.cuh under the lane
// emulator, with the same C ABI argument layout as mmpc_solve (host pointers).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "emu_runtime.h"
#include "../../mobile_manipulator_mpc_b200/csrc/mmpc_solver.cuh"

namespace mmpc { EmuWarp* g_emu = nullptr; }
using namespace mmpc;

struct LaneArg { const KParams* P; double* smem; };
static LaneArg g_arg;

static void lane_entry() {
  EmuWarp* w = g_emu;
  int me = w->cur;
  {
    Solver S(*g_arg.P);
    S.lane = me;
    for (;;) {
      unsigned b = next_instance(g_arg.P->counter);
      if (b >= (unsigned)g_arg.P->B) break;
      S.bind(g_arg.smem, 0, (int)b);
      S.run((int)b);
    }
  }
  w->done++;
  if (w->done < 32) { int nx = (me + 1) & 31; w->cur = nx; swapcontext(&w->ctx[me], &w->ctx[nx]); }
  else swapcontext(&w->ctx[me], &w->main_ctx);
}

extern "C" int mmpc_emu_solve(const MmpcConfig* cfg, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out) {
  KParams P; memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.B = B;
  P.x_init = in->x_init; P.x_ref = in->x_ref; P.u_ref = in->u_ref; P.u_last = in->u_last; P.u_guess = in->u_guess;
  P.circles = in->circles; P.planes = in->planes; P.n_pl_inst = in->n_pl_inst; P.flags = in->flags;
  P.U = out->U; P.X = out->X; P.s = out->s; P.cost = out->cost; P.kkt = out->kkt; P.iters = out->iters; P.status = out->status;
  int N = cfg->N;
  P.SP = N + 1; P.KP = ((N + 1 + 3) / 4) * 4; P.R = cfg->n_obs + 4 + (cfg->n_pl > 0 ? 6 : 0);
  P.ws_stride = ws_doubles(N, P.KP, P.R);
  std::vector<double> ws(P.ws_stride, 0.0), smem(smem_doubles(N), 0.0);
  unsigned counter = 0;
  P.ws = ws.data(); P.counter = &counter;
  EmuWarp* w = new EmuWarp(); memset(w->slot_d, 0, sizeof w->slot_d); memset(w->slot_i, 0, sizeof w->slot_i);
  g_emu = w; g_arg.P = &P; g_arg.smem = smem.data();
  const size_t STK = 1 << 20;
  std::vector<char*> stacks(32);
  for (int i = 0; i < 32; ++i) {
    stacks[i] = (char*)malloc(STK);
    getcontext(&w->ctx[i]);
    w->ctx[i].uc_stack.ss_sp = stacks[i]; w->ctx[i].uc_stack.ss_size = STK; w->ctx[i].uc_link = &w->main_ctx;
    makecontext(&w->ctx[i], lane_entry, 0);
  }
  w->cur = 0; w->done = 0;
  swapcontext(&w->main_ctx, &w->ctx[0]);
  for (int i = 0; i < 32; ++i) free(stacks[i]);
  delete w; g_emu = nullptr;
  return 0;
}
