// mmpc_resident.cu -- the RESIDENT solver: one thread block owns one instance for its whole solve, with every bit of the
// instance's solver state -- iterate (two copies), multipliers, step, stage QP, Riccati factors, per-instance scalars -- in
// SHARED MEMORY (north_star: "state ... staged through shared memory").  HBM sees the instance's
// inputs once and its outputs once: the algorithmic traffic of section 8(d).  Blocks are persistent and pull instances from an
// atomic work queue, so there are no lists, no compaction, no rounds, and no host in the loop: one launch per solve.
//
// It is the LATENCY path (small batches: a single controller, closed loops of a few hundred robots).  A B200 SM holds one
// instance of this NLP in shared memory (about 127 KB at N = 20 with 16 circles), so 148 instances are in flight and an
// iteration costs its dependent-instruction latency; batches beyond a few hundred instances are faster through the streaming
// "staged" kernels, which keep 24 instances per SM in flight.  mmpc_api.cu picks by batch size.
//
// The arithmetic is the staged solver's, literally: this file compiles the same phase bodies (mmpc_staged.cuh, mmpc_team.cuh)
// with MMPC_RESIDENT defined, which turns the tile-major HBM layout into a stride-1 layout (LSH = 0), global-memory loads
// and cp.async staging into plain shared-memory accesses, and gives Inst a separate index for the caller's arrays.  Thread
// roles inside the block: warp 0 = the 16-lane Riccati team (its second half-warp mirrors the first) and the per-instance
// control steps; threads 32 .. 32 + N = one per stage (evaluation, step, trial).  Phases are separated by __syncthreads().
#define MMPC_RESIDENT 1
#define mmpc mmpc_res   // its own namespace: the same inline function names are compiled differently in mmpc_api.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mmpc.h"
#include "mmpc_team.cuh"   // includes mmpc_staged.cuh

using namespace mmpc_res;

namespace mmpc_res {

// shared-memory plan of one block, in doubles
struct ResPlan {
  int STGp;        // stage stride, padded to an odd number of doubles: stage threads hit different banks
  int o_ws, o_qp, o_rk, o_gd, o_gi, o_team, o_ring, total;
  int stage_threads, threads;
};
__host__ __device__ inline ResPlan res_plan(const MmpcConfig& c) {
  ResPlan p;
  const int K1 = c.N + 1;
  p.STGp = staged_stage_doubles(c) | 1;
  p.stage_threads = (K1 + 31) / 32 * 32;
  p.threads = 32 + p.stage_threads;
  int o = 0;
  p.o_ws = o; o += K1 * p.STGp;
  p.o_qp = o; o += K1 * QS;
  p.o_rk = o; o += K1 * RS;
  p.o_gd = o; o += staged_inst_doubles(c);
  p.o_gi = o; o += (J_NFIELDS + 1) / 2 + 1;          // ints, two per double
  o = (o + 1) & ~1;                                   // the team ring is read in 16-byte pieces
  p.o_team = o; o += 2 * Team::SMEM_DOUBLES;          // one ring per half-warp
  p.o_ring = o; o += STAGED_TRIAL_RING_DOUBLES * p.stage_threads;   // the per-thread row rings of the step / trial bodies
  p.total = o;
  return p;
}

template <bool REF, bool Q3>
__global__ void __launch_bounds__(96, 1) resident_solve_kernel(const __grid_constant__ SParams P0, unsigned* queue) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_b;
  const ResPlan pl = res_plan(P0.cfg);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P0.cfg.N;
  const int k = tid - 32;                       // stage of a stage thread
  const bool stage = tid >= 32 && k <= N;
  SParams P = P0;
  P.ws = smem + pl.o_ws; P.qp = smem + pl.o_qp; P.rk = smem + pl.o_rk; P.gd = smem + pl.o_gd; P.gi = (int*)(smem + pl.o_gi);
  P.LS = 1; P.STG = pl.STGp;
  double* team_ring = smem + pl.o_team + (lane >> 4) * Team::SMEM_DOUBLES;
  double* ring = smem + pl.o_ring + (tid - 32);
  const int B = P.io->B;
  for (;;) {
    __syncthreads();                            // everybody is done with the previous instance (and with s_b)
    if (tid == 0) s_b = (int)atomicAdd(queue, 1u);
    __syncthreads();
    const int b = s_b;
    if (b >= B) break;
    Inst S(P, b); S.sm = ring; S.bs = pl.stage_threads;
    if (tid == 0) S.init();                     // the starting point (:302-304), bound push, slack lift, objective scaling
    __syncthreads();
    if (REF) { if (stage) S.pose_pass(k, false); __syncthreads(); }
    if (stage) S.template eval<REF>(k);
    __syncthreads();
    for (;;) {
      // ---- KKT test, barrier update, Riccati factorisation, roll-out of the Newton step ----
      if (warp == 0) { Team T(P, b, lane & 15, team_ring); T.template solve<Q3>(); }
      __syncthreads();
      const int st = S.J(J_STATE);
      // ---- slack / multiplier steps, fraction to the boundary (or: the results of an instance that has just finished) ----
      if (stage) { if (st == ST_FINISH) S.finish_stage(k); else if (st == ST_ACTIVE) S.template step<REF>(k); }
      __syncthreads();
      if (warp == 0) S.template ctrl_step<32>(lane);
      __syncthreads();
      if (S.J(J_STATE) != ST_TRIAL) break;      // finished: the outputs are written
      // ---- filter line search: candidate + evaluation of the next iteration at the candidate ----
      for (;;) {
        if (REF) { if (stage) S.pose_pass(k, true); __syncthreads(); }
        if (stage) S.template trial_eval<REF>(k);
        __syncthreads();
        if (warp == 0) S.template ctrl_trial<32>(lane);
        __syncthreads();
        if (S.J(J_STATE) != ST_TRIAL) break;    // accepted (ST_ACTIVE) or given up (ST_DONE)
      }
      if (S.J(J_STATE) != ST_ACTIVE) break;
    }
  }
}

}  // namespace mmpc_res

// ---- host side (called by mmpc_api.cu) ------------------------------------------------------------------------------------
extern "C" int mmpc_resident_smem_bytes(const MmpcConfig* cfg) {
  return (int)(res_plan(*cfg).total * sizeof(double));
}

// Launches the resident solve of the B instances described by the device block `io_dev` (SIO of mmpc_staged.cuh; the resident
// and the staged build share its layout).  `queue` is a zeroed device counter.  Returns a cudaError_t.
extern "C" int mmpc_resident_launch(const MmpcConfig* cfg, int B, const void* io_dev, unsigned* queue, int max_blocks, void* stream) {
  const MmpcConfig& c = *cfg;
  const ResPlan pl = res_plan(c);
  SParams P; memset(&P, 0, sizeof P);
  P.cfg = c; P.B = B; P.io = (const SIO*)io_dev;
  P.team = 1; P.fused = 1; P.parts = 0;
  P.R = staged_rows(c); P.ITSZ = staged_itsz(c); P.STG = pl.STGp; P.ND = staged_inst_doubles(c); P.LS = 1;
  const bool ref = c.mode == MMPC_MODE_REFERENCE, q3 = ref && c.terminal_rows_on_sN == 0;
  const void* fn = ref ? (q3 ? (const void*)resident_solve_kernel<true, true> : (const void*)resident_solve_kernel<true, false>)
                       : (const void*)resident_solve_kernel<false, false>;
  const int smem = (int)(pl.total * sizeof(double));
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = B < max_blocks ? B : max_blocks;
  void* args[] = {&P, &queue};
  return (int)cudaLaunchKernel(fn, dim3(grid), dim3(pl.threads), args, smem, (cudaStream_t)stream);
}
