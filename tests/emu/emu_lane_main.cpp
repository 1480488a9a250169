// emu_lane_main.cpp -- TEST INFRASTRUCTURE: runs lane_main() of csrc/mmpc_lane.cuh on the CPU,
// one lane at a time, with the same C ABI argument layout as mmpc_solve (host pointers).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "emu_lane_runtime.h"
#include "../../mobile_manipulator_mpc_b200/csrc/mmpc_lane.cuh"

namespace mmpc { EmuTeam* g_team = nullptr; }
using namespace mmpc;

extern "C" int mmpc_emu_lane_solve(const MmpcConfig* cfg, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out) {
  LParams P; memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.B = B;
  P.x_init = in->x_init; P.x_ref = in->x_ref; P.u_ref = in->u_ref; P.u_last = in->u_last; P.u_guess = in->u_guess;
  P.circles = in->circles; P.planes = in->planes; P.n_pl_inst = in->n_pl_inst; P.flags = in->flags;
  P.U = out->U; P.X = out->X; P.s = out->s; P.cost = out->cost; P.kkt = out->kkt; P.iters = out->iters; P.status = out->status;
  P.R = lane_rows(*cfg); P.STG = lane_stage_doubles(*cfg); P.OG = (cfg->N + 1) * P.STG;
  std::vector<double> ws((size_t)lane_instance_doubles(*cfg), 0.0);
  unsigned counter = 0;
  P.ws = ws.data(); P.counter = &counter; P.warp_stride = (long long)ws.size();
  lane_main(P, ws.data());
  return 0;
}
