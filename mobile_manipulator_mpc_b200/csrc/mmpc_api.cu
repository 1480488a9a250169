// mmpc_api.cu -- host side of the C ABI declared in include/mmpc.h (sm_100a only).
// No torch types, no exceptions across the boundary, no CPU fallback: every entry point that
// computes needs a CUDA device and fails with MMPC_ERR_NO_DEVICE / MMPC_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "../../include/mmpc.h"
#include "mmpc_staged.cuh"
#include "mmpc_team.cuh"
#include "mmpc_parts.cuh"
#include "mmpc_episode.cuh"

using namespace mmpc;

// csrc/mmpc_resident.cu (its own translation unit: the same phase bodies compiled for a shared-memory workspace)
extern "C" int mmpc_resident_smem_bytes(const MmpcConfig* cfg);
extern "C" int mmpc_resident_launch(const MmpcConfig* cfg, int B, const void* io_dev, unsigned* queue, int max_blocks, void* stream);
extern "C" int mmpc_resident_tail_node(const MmpcConfig* cfg, int B, const void* io_dev, int sm_count, const void** fn_out, int* threads_out,
                                       int* smem_out, int* blocks_in_flight, void* params_out);
// csrc/mmpc_resident_pose.cu (the same kernel with the end-point pose cost of MMPC_MODEL_POSEREF compiled in)
extern "C" int mmpc_resident_pose_smem_bytes(const MmpcConfig* cfg);
extern "C" int mmpc_resident_pose_launch(const MmpcConfig* cfg, int B, const void* io_dev, unsigned* queue, int max_blocks, void* stream);

struct MmpcHandle {
  MmpcConfig cfg;
  int device, B_max, sm_count;
  long long launches;
  int kernel, sg_team, sg_fused, sg_parts;
  // staged (batch-synchronous) solver: field-major state, per-instance scalars, lists, counters
  struct { double *ws, *qp, *rk, *gd; int *gi, *lists, *cnt; SIO* io; long long LS; int* pin; cudaEvent_t ev[8]; bool ready;
           int rounds; } sg;
  // the solve as one CUDA graph (built for a batch size, rebuilt when B, the weights or the kernel selection change)
  struct { struct { cudaGraph_t graph; cudaGraphExec_t exec; int cap, classes, tail; } e[4]; int next, pending_classes, pending_tail; bool pending; } gr;
  int hostloop;  // MMPC_KERNEL_STAGED_HOSTLOOP: the host sequences the rounds
  int resident;  // MMPC_KERNEL_RESIDENT forced
  int autosel;   // MMPC_KERNEL_AUTO: resident for small batches, staged otherwise
  int smem_optin; // largest dynamic shared memory of a block on this device
  unsigned* queue;  // work queue counter of the resident kernel
  int last_solver;  // MMPC_KERNEL_RESIDENT / MMPC_KERNEL_STAGED: what the last solve ran on
  // per-phase device timing of the staged solver (mmpc_set_profile / mmpc_phase_times)
  int profile;
  std::vector<cudaEvent_t>* prof_ev;
  double phase_ms[MMPC_NPHASE];
  long long phase_launches[MMPC_NPHASE];
  // staging for mmpc_solve_host
  struct { double *x_init, *x_ref, *u_ref, *u_last, *u_guess, *circles, *planes, *x_guess, *U, *X, *s, *cost, *kkt;
           int32_t *n_pl_inst, *iters, *status; uint8_t* flags; } d, h;
  cudaStream_t stream;
  char err[256];
};

static thread_local char g_err[256] = "";

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      snprintf(g_err, sizeof g_err, "%s failed: %s", #call, cudaGetErrorString(e_));               \
      return MMPC_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

extern "C" int mmpc_version(void) { return 200; }

extern "C" const char* mmpc_error_string(int code) {
  switch (code) {
    case MMPC_OK: return "ok";
    case MMPC_ERR_ARG: return "invalid argument";
    case MMPC_ERR_CUDA: return g_err[0] ? g_err : "CUDA error";
    case MMPC_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case MMPC_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

extern "C" void mmpc_default_config(MmpcConfig* c) {
  memset(c, 0, sizeof *c);
  c->N = 20; c->n_obs = 3; c->n_pl = 3; c->mode = MMPC_MODE_REFERENCE; c->obs_per_stage = 0; c->max_iter = 2000; c->model = MMPC_MODEL_WHOLEBODY;
  c->dt = 0.1;
  const double q[9] = {25, 25, 0, 0, 0, 5, 5, 5, 5};
  const double r[5] = {0.1, 0.1, 0, 0, 0}, w[5] = {0, 0, 0.1, 0.1, 0.1};
  for (int i = 0; i < 9; ++i) c->Qd[i] = c->Pd[i] = q[i];
  for (int j = 0; j < 5; ++j) { c->Rd[j] = r[j]; c->Wd[j] = w[j]; }
  c->S = 1e5;
  const double pi = 3.14159265358979323846, inf = INFINITY;
  const double ul[5] = {2, pi, 1, 1, 1};
  for (int j = 0; j < 5; ++j) { c->ulim[0][j] = -ul[j]; c->ulim[1][j] = ul[j]; }
  const double xl[9] = {-100, -100, -inf, -2, -2, -pi, -pi / 2, -pi, 0}, xh[9] = {100, 100, inf, 2, 2, pi, pi / 2, 0, 3 * pi / 2};
  for (int i = 0; i < 9; ++i) { c->xlim[0][i] = xl[i]; c->xlim[1][i] = xh[i]; }
  const double dl[5] = {inf, inf, 0.5, 0.5, 0.5};
  for (int j = 0; j < 5; ++j) { c->dulim[0][j] = -dl[j]; c->dulim[1][j] = dl[j]; }
  c->base_radius = 0.4; c->self_collision_radius = 0.05; c->obstacle_expand_dist = 0.03;
  c->tol = 1e-8; c->mu_init = 0.1; c->acceptable_tol = 1e-8;
}

extern "C" int mmpc_struct_sizes(int32_t* cfg_bytes, int32_t* in_bytes, int32_t* out_bytes) {
  *cfg_bytes = (int32_t)sizeof(MmpcConfig); *in_bytes = (int32_t)sizeof(MmpcBatchIn); *out_bytes = (int32_t)sizeof(MmpcBatchOut);
  return MMPC_OK;
}

static int set_kernel_attributes();
static void graph_destroy(MmpcHandle* h);
static void graph_account(MmpcHandle* h);
static size_t circles_per_instance(const MmpcConfig& c) { return (size_t)c.n_obs * 3 * (c.obs_per_stage ? c.N + 1 : 1); }

extern "C" int mmpc_create(const MmpcConfig* cfg, int32_t B_max, int32_t device, MmpcHandle** out) {
  if (!cfg || !out || B_max < 1) return MMPC_ERR_ARG;
  if (cfg->N < 1 || cfg->N > 63 || cfg->n_obs < 0 || cfg->n_pl < 0 || cfg->n_pl > MMPC_MAX_PLANES) return MMPC_ERR_ARG;
  if (cfg->mode != MMPC_MODE_CLEAN && cfg->mode != MMPC_MODE_REFERENCE) return MMPC_ERR_ARG;
  if (cfg->model != MMPC_MODEL_WHOLEBODY && cfg->model != MMPC_MODEL_BASE && cfg->model != MMPC_MODEL_POSEREF) return MMPC_ERR_ARG;
  // the pose-reference controller (controllers/mpc_wholebody.py) has ground circles only; it runs on the resident kernel
  if (cfg->model == MMPC_MODEL_POSEREF && (cfg->n_pl != 0 || cfg->mode != MMPC_MODE_CLEAN)) return MMPC_ERR_UNSUPPORTED;
  // the base-only controller (controllers/mpc_base.py) has no arm: no plane rows, no bug-for-bug rows
  if (cfg->model == MMPC_MODEL_BASE && (cfg->n_pl != 0 || cfg->mode != MMPC_MODE_CLEAN)) return MMPC_ERR_UNSUPPORTED;
  // the literal reference NLP bounds the terminal self-collision rows by s[N-1]: the starting point relies on x_N = x_{N-1},
  // which needs a stage N-1 >= 1 (stage 0 is the un-pushed initial state)
  if (cfg->mode == MMPC_MODE_REFERENCE && cfg->terminal_rows_on_sN == 0 && cfg->N < 2) return MMPC_ERR_UNSUPPORTED;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MMPC_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return MMPC_ERR_ARG;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    snprintf(g_err, sizeof g_err, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return MMPC_ERR_CUDA;
  }
  MmpcHandle* h = new (std::nothrow) MmpcHandle();
  if (!h) return MMPC_ERR_ARG;
  memset(h, 0, sizeof *h);
  h->cfg = *cfg; h->device = device; h->B_max = B_max; h->sm_count = prop.multiProcessorCount; h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "cudaStreamCreateWithFlags failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return MMPC_ERR_CUDA;
  }
  h->kernel = MMPC_KERNEL_AUTO; h->sg_team = 1; h->sg_fused = 1; h->sg_parts = 1; h->autosel = 1;
  {
    int rc = set_kernel_attributes();  // per device (the trial kernels need 81 KB of dynamic shared memory)
    if (rc != MMPC_OK) { cudaStreamDestroy(h->stream); delete h; return rc; }
  }
  *out = h;
  return MMPC_OK;
}

static void free_staging(MmpcHandle* h) {
  void** dp = (void**)&h->d; void** hp = (void**)&h->h;
  for (size_t i = 0; i < sizeof(h->d) / sizeof(void*); ++i) {
    if (dp[i]) cudaFree(dp[i]);
    if (hp[i]) cudaFreeHost(hp[i]);
    dp[i] = hp[i] = nullptr;
  }
}

extern "C" int mmpc_destroy(MmpcHandle* h) {
  if (!h) return MMPC_ERR_ARG;
  cudaSetDevice(h->device);
  free_staging(h);
  if (h->prof_ev) { for (cudaEvent_t e : *h->prof_ev) cudaEventDestroy(e); delete h->prof_ev; }
  graph_destroy(h);
  for (void* q : {(void*)h->sg.ws, (void*)h->sg.qp, (void*)h->sg.rk, (void*)h->sg.gd, (void*)h->sg.gi, (void*)h->sg.lists, (void*)h->sg.cnt, (void*)h->sg.io, (void*)h->queue})
    if (q) cudaFree(q);
  if (h->sg.pin) cudaFreeHost(h->sg.pin);
  for (int i = 0; i < 8; ++i) if (h->sg.ev[i]) cudaEventDestroy(h->sg.ev[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return MMPC_OK;
}

extern "C" int mmpc_set_weights(MmpcHandle* h, const double* Qd, const double* Pd, const double* Rd, const double* Wd, double S) {
  if (!h) return MMPC_ERR_ARG;
  if (Qd) memcpy(h->cfg.Qd, Qd, sizeof h->cfg.Qd);
  if (Pd) memcpy(h->cfg.Pd, Pd, sizeof h->cfg.Pd);
  if (Rd) memcpy(h->cfg.Rd, Rd, sizeof h->cfg.Rd);
  if (Wd) memcpy(h->cfg.Wd, Wd, sizeof h->cfg.Wd);
  if (S == S && S > 0) h->cfg.S = S;
  cudaSetDevice(h->device);
  graph_destroy(h);  // the weights are kernel parameters of every node
  return MMPC_OK;
}

extern "C" int mmpc_set_kernel(MmpcHandle* h, int32_t kernel) {
  if (!h || (kernel != MMPC_KERNEL_AUTO && (kernel < MMPC_KERNEL_STAGED || kernel > MMPC_KERNEL_RESIDENT))) return MMPC_ERR_ARG;
  if (kernel == MMPC_KERNEL_RESIDENT && mmpc_resident_smem_bytes(&h->cfg) > h->smem_optin) return MMPC_ERR_UNSUPPORTED;
  h->resident = kernel == MMPC_KERNEL_RESIDENT;
  h->autosel = kernel == MMPC_KERNEL_AUTO;
  cudaSetDevice(h->device);
  graph_destroy(h);
  h->hostloop = kernel == MMPC_KERNEL_STAGED_HOSTLOOP;
  h->sg_parts = kernel != MMPC_KERNEL_STAGED_FAT;
  h->sg_team = kernel != MMPC_KERNEL_STAGED_THREAD;
  h->sg_fused = kernel != MMPC_KERNEL_STAGED_UNFUSED;
  h->kernel = MMPC_KERNEL_STAGED;
  return MMPC_OK;
}

// ---- staged solver: launch machinery ------------------------------------------------------------------------------
// One interior-point round of the whole batch = seven kernels over the device-side lists of active instances (eight in
// round 0, which also evaluates the starting point).  The kernels read the list lengths on the device and stride over
// their grids, so a launch configuration only needs an UPPER BOUND of the active instances.  Two drivers issue the same
// rounds through `Issuer`:
//   graph      (default) the whole solve is ONE CUDA graph: round 0, then a chain of conditional WHILE nodes, one per size
//              class of the active set (the bound halves ... down to the thin classes that switch to the spill-free
//              Riccati instantiation and the warp-specialised part kernels).  The loop condition is written by a
//              one-thread kernel at the end of each body from the device-side counts, so the host is out of the loop:
//              one cudaGraphLaunch per solve, no per-round synchronisation, any number of solver contexts per host thread.
//   host loop  (profiling, MMPC_HOSTLOOP=1) the host sequences rounds and learns the list lengths from a 16-byte copy that
//              trails the launches by two rounds; used when every launch is bracketed by timing events.
__global__ void staged_set_io_kernel(SIO io, SIO* dst, unsigned* tail_queue) { *dst = io; *tail_queue = 0u; }   // (tail_queue: the resident tail kernel's work counter)
// cnt[0] E list, cnt[1] / cnt[2] trial lists, cnt[3] rounds executed
__global__ void staged_cond_kernel(cudaGraphConditionalHandle hnd, int* cnt, int which, int above, int add_rounds) {
  cnt[3] += add_rounds;
  cudaGraphSetConditional(hnd, cnt[which] > above ? 1u : 0u);
}

struct Issuer {
  cudaStream_t st;          // stream mode
  cudaGraph_t g;            // graph mode: nodes are chained one after the other
  cudaGraphNode_t last; bool has_last;
  long long issued;
  int rc;
  void launch(const void* fn, dim3 grid, dim3 block, size_t smem, void** args) {
    if (rc != MMPC_OK) return;
    cudaError_t e;
    if (g) {
      cudaKernelNodeParams kp; memset(&kp, 0, sizeof kp);
      kp.func = const_cast<void*>(fn); kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem; kp.kernelParams = args;
      cudaGraphNode_t n;
      e = cudaGraphAddKernelNode(&n, g, has_last ? &last : nullptr, has_last ? 1 : 0, &kp);
      last = n; has_last = true;
    } else {
      e = cudaLaunchKernel(fn, grid, block, args, smem, st);
    }
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "kernel launch / node failed: %s", cudaGetErrorString(e)); rc = MMPC_ERR_CUDA; }
    ++issued;
  }
};

static int ensure_workspace(MmpcHandle* h) {
  if (h->sg.ready) return MMPC_OK;
  const MmpcConfig& cfg = h->cfg;
  const int N = cfg.N, STG = staged_stage_doubles(cfg);
  long long LS = ((long long)h->B_max + 31) / 32 * 32;
  h->sg.LS = LS;
  CK(cudaMalloc(&h->sg.ws, (size_t)(N + 1) * STG * LS * sizeof(double)));
  CK(cudaMalloc(&h->sg.qp, (size_t)(N + 1) * QS * LS * sizeof(double)));
  CK(cudaMalloc(&h->sg.rk, (size_t)(N + 1) * RS * LS * sizeof(double)));
  CK(cudaMalloc(&h->sg.gd, (size_t)staged_inst_doubles(cfg) * LS * sizeof(double)));
  CK(cudaMalloc(&h->sg.gi, (size_t)J_NFIELDS * LS * sizeof(int)));
  CK(cudaMalloc(&h->sg.lists, (size_t)4 * LS * sizeof(int)));   // E, two trial lists, the Riccati's ordering of E
  CK(cudaMalloc(&h->sg.cnt, 4 * sizeof(int)));
  CK(cudaMalloc(&h->sg.io, sizeof(SIO)));
  CK(cudaMalloc(&h->queue, 2 * sizeof(unsigned)));   // [0] the resident kernel's work counter, [1] the resident tail kernel's
  CK(cudaMallocHost(&h->sg.pin, (8 * 4 + 4) * sizeof(int)));
  for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&h->sg.ev[i], cudaEventDisableTiming | cudaEventBlockingSync));  // the host thread sleeps, it does not spin: several contexts per GPU and ranks per box share the cores
  h->sg.ready = true;
  return MMPC_OK;
}

static SParams staged_params(const MmpcHandle* h, int32_t B) {
  const MmpcConfig& cfg = h->cfg;
  SParams P; memset(&P, 0, sizeof P);
  P.cfg = cfg; P.B = B; P.io = h->sg.io;
  P.ws = h->sg.ws; P.qp = h->sg.qp; P.rk = h->sg.rk; P.team = h->sg_team; P.fused = h->sg_fused;
  const PartPlan plan = part_plan(cfg);
  P.parts = (h->sg_parts && h->sg_fused && plan.n_parts <= 10 && cfg.mode == MMPC_MODE_CLEAN && cfg.model == MMPC_MODEL_WHOLEBODY) ? 1 : 0;
  P.gd = h->sg.gd; P.gi = h->sg.gi; P.lists = h->sg.lists; P.cnt = h->sg.cnt; P.LS = h->sg.LS;
  P.R = staged_rows(cfg); P.ITSZ = staged_itsz(cfg); P.STG = staged_stage_doubles(cfg); P.ND = staged_inst_doubles(cfg);
  return P;
}

static const int ring_smem = 128 * STAGED_RING_DOUBLES * (int)sizeof(double);              // row ring of the step kernels: 24 KB per block of 128
static const int trial_ring_smem = 128 * STAGED_TRIAL_RING_DOUBLES * (int)sizeof(double);  // row ring + parked multipliers and inputs of the trial kernels: 81 KB

// cudaFuncSetAttribute is per device: called by mmpc_create after cudaSetDevice (idempotent, so concurrent creates are fine)
static int set_kernel_attributes() {
  CK(cudaFuncSetAttribute(staged_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_smem));
  CK(cudaFuncSetAttribute(staged_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_smem));
  CK(cudaFuncSetAttribute(staged_trial_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, trial_ring_smem));
  CK(cudaFuncSetAttribute(staged_trial_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, trial_ring_smem));
  return MMPC_OK;
}

// thin classes: every instance gets its own half warp at TEAM_WARPS_THIN warps per SM -> the spill-free Riccati; every
// tile of 32 (instance, stage) items gets its own SM -> the warp-specialised part kernels (clean NLP only)
static bool class_thin_team(const MmpcHandle* h, long long ub) { return ub * 16 <= (long long)h->sm_count * TEAM_WARPS_THIN * 32; }
static bool class_thin_parts(const MmpcHandle* h, const SParams& P, long long ub) {
  static const long long parts_tiles = getenv("MMPC_PARTS_TILES") ? atoll(getenv("MMPC_PARTS_TILES")) : -1;  // A/B knob
  return P.parts && (ub * (P.cfg.N + 1) + 31) / 32 <= (parts_tiles >= 0 ? parts_tiles : (long long)h->sm_count);
}

// Issues round r for at most `ub` active instances.  `mark` (host loop with profiling) is called in front of every launch.
template <class Mark>
static void issue_round(MmpcHandle* h, Issuer& I, SParams P, int r, long long ub, bool with_eval, Mark&& mark) {
  const MmpcConfig& cfg = h->cfg;
  const int N = cfg.N;
  static const int cap_mult = getenv("MMPC_GRID_CAP") ? atoi(getenv("MMPC_GRID_CAP")) : 128;  // blocks per SM before grid-striding (A/B: 128 beats 16 by 1.7 %)
  static const int team_cap = getenv("MMPC_TEAM_GRID") ? atoi(getenv("MMPC_TEAM_GRID")) : 0;  // A/B: blocks per SM of the team kernel (0: one block per 2 instances)
  const int cap = h->sm_count * cap_mult;
  const bool ref = cfg.mode == MMPC_MODE_REFERENCE;
  const bool q3 = ref && cfg.terminal_rows_on_sN == 0;  // terminal self-collision rows on s[N-1]
  const PartPlan plan = part_plan(cfg);
  auto clampi = [](long long v, long long lo, long long hi) { return (int)(v < lo ? lo : v > hi ? hi : v); };
  const long long items = ub * (N + 1);
  const int gs = clampi((items + 127) / 128, 1, cap);
  const int g64 = clampi((ub + 63) / 64, 1, 1 << 30);
  int gt = clampi((ub * 16 + MMPC_TEAM_BLOCK - 1) / MMPC_TEAM_BLOCK, 1, 1 << 30);
  if (team_cap > 0 && gt > team_cap * h->sm_count) gt = team_cap * h->sm_count;
  const int gw = clampi((ub * 32 + 127) / 128, 1, cap);  // one warp per instance
  const int gtile = clampi((items + 31) / 32, 1, 8LL * cap);
  const int tcur = 1 + (r & 1), tnext = 1 + ((r + 1) & 1);
  const int cthreads = ub > 16384 ? 1024 : ub > 2048 ? 256 : 64;
  const bool thin_team = class_thin_team(h, ub), thin = class_thin_parts(h, P, ub);
  int dst, src, want;
  void* a1[] = {&P};
  void* a4[] = {&P, &dst, &src, &want};
  mark(MMPC_PHASE_COMPACT);
  dst = 0; src = tcur; want = ST_ACTIVE;
  I.launch((const void*)staged_compact_kernel, dim3(1), dim3(cthreads), 0, a4);
  int which = 0;
  void* a2[] = {&P, &which};
  if (with_eval) {  // fused: the trial kernel has already evaluated every later accepted point
    if (ref) { mark(MMPC_PHASE_POSE); I.launch((const void*)staged_pose_kernel, dim3(gs), dim3(128), 0, a2); }
    mark(MMPC_PHASE_EVAL);
    I.launch(ref ? (const void*)staged_eval_kernel<true> : (const void*)staged_eval_kernel<false>, dim3(gs), dim3(128), 0, a1);
  }
  mark(MMPC_PHASE_SOLVE);
  if (P.team) {
    const void* fn = q3 ? (thin_team ? (const void*)staged_solve_team_kernel<true, TEAM_WARPS_THIN> : (const void*)staged_solve_team_kernel<true, TEAM_WARPS_BULK>)
                        : (thin_team ? (const void*)staged_solve_team_kernel<false, TEAM_WARPS_THIN> : (const void*)staged_solve_team_kernel<false, TEAM_WARPS_BULK>);
    I.launch(fn, dim3(gt), dim3(MMPC_TEAM_BLOCK), 0, a1);
  } else I.launch((const void*)staged_solve_kernel, dim3(g64), dim3(64), 0, a1);
  mark(MMPC_PHASE_STEP);
  if (thin) I.launch((const void*)staged_parts_kernel<false>, dim3(gtile), dim3(32 * plan.n_parts), 0, a1);
  else I.launch(ref ? (const void*)staged_step_kernel<true> : (const void*)staged_step_kernel<false>, dim3(gs), dim3(128), ring_smem, a1);
  mark(MMPC_PHASE_CTRL_STEP);
  I.launch((const void*)staged_ctrl_step_kernel, dim3(gw), dim3(128), 0, a1);
  mark(MMPC_PHASE_COMPACT);
  dst = tnext; src = tcur; want = ST_TRIAL;
  I.launch((const void*)staged_compact_kernel, dim3(1), dim3(cthreads), 0, a4);
  P.tsel = tnext;
  if (ref) { which = 1; mark(MMPC_PHASE_POSE); I.launch((const void*)staged_pose_kernel, dim3(gs), dim3(128), 0, a2); }
  mark(MMPC_PHASE_TRIAL);
  if (thin) I.launch((const void*)staged_parts_kernel<true>, dim3(gtile), dim3(32 * plan.n_parts), 0, a1);
  else I.launch(ref ? (const void*)staged_trial_kernel<true> : (const void*)staged_trial_kernel<false>, dim3(gs), dim3(128), trial_ring_smem, a1);
  mark(MMPC_PHASE_CTRL_TRIAL);
  I.launch((const void*)staged_ctrl_trial_kernel, dim3(gw), dim3(128), 0, a1);
}

static void issue_init(MmpcHandle* h, Issuer& I, SParams P) {
  void* a1[] = {&P};
  I.launch((const void*)staged_init_kernel, dim3((P.B + 127) / 128), dim3(128), 0, a1);
}

// ---- driver 1: the whole solve as one CUDA graph with device-side loops --------------------------------------------------
static void graph_destroy(MmpcHandle* h) {
  graph_account(h);
  for (auto& e : h->gr.e) {
    if (e.exec) cudaGraphExecDestroy(e.exec);
    if (e.graph) cudaGraphDestroy(e.graph);
    e.exec = nullptr; e.graph = nullptr; e.cap = 0;
  }
}

// launch configurations are chosen for a capacity, not for the batch size: B_max itself, else the next power of two (>= 64)
static int graph_capacity(const MmpcHandle* h, int32_t B) {
  long long c = 64;
  while (c < B) c *= 2;
  return (int)(c >= h->B_max ? h->B_max : c);
}

static int graph_build(MmpcHandle* h, int32_t B, int slot) {
  auto& E = h->gr.e[slot];
  if (E.exec) cudaGraphExecDestroy(E.exec);
  if (E.graph) cudaGraphDestroy(E.graph);
  E.exec = nullptr; E.graph = nullptr; E.cap = 0;
  SParams P = staged_params(h, B);
  cudaGraph_t g;
  CK(cudaGraphCreate(&g, 0));
  E.graph = g;
  Issuer I; memset(&I, 0, sizeof I); I.g = g; I.rc = MMPC_OK;
  auto nomark = [](int) {};
  issue_init(h, I, P);
  issue_round(h, I, P, 0, B, true, nomark);
  // size classes of the active set: B, B/2, B/4, ... and the two thin thresholds; a loop per class, entered in turn
  // (the active set only shrinks).  Class c runs while  count > lower bound of c  with launch bounds for `ub[c]`.
  // The tail: once the active set is down to a few times what the resident kernel holds in flight, it is handed to that
  // kernel (one graph node after the loops; csrc/mmpc_resident.cu, resident_tail_kernel) instead of being walked through ~300 us rounds one
  // iteration at a time.  Same results to the bit.  (MMPC_RESIDENT_TAIL=0: A/B, the staged rounds run to the end.)
  static const int tail_on = getenv("MMPC_RESIDENT_TAIL") ? atoi(getenv("MMPC_RESIDENT_TAIL")) : 1;
  const void* tail_fn = nullptr; int tail_threads = 0, tail_smem = 0, tail_blocks = 0, tail_hand = 0;   // tail_hand: active instances at the hand-over
  SParams Ptail;
  if (tail_on && h->sg_fused && h->cfg.model != MMPC_MODEL_POSEREF && mmpc_resident_smem_bytes(&h->cfg) <= h->smem_optin) {
    cudaError_t e = (cudaError_t)mmpc_resident_tail_node(&h->cfg, B, h->sg.io, h->sm_count, &tail_fn, &tail_threads, &tail_smem, &tail_blocks, &Ptail);
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "resident tail kernel: %s", cudaGetErrorString(e)); return MMPC_ERR_CUDA; }
    // hand over at this many times the blocks in flight.  A/B on a B200 (65,536-batch / closed-loop sub-batch of 5,461, ms):
    // 1: 373 / 32.0, 2: 367 / 31.3, 4: 362 / 31.1, 8: 358 / 31.2, 16: 358 / 31.2 -- the queue keeps 296 blocks busy while the
    // many instances that need only a few more iterations drain, and none of them waits for a 300 us round
    static const int tail_mult = getenv("MMPC_TAIL_MULT") ? atoi(getenv("MMPC_TAIL_MULT")) : 8;
    tail_hand = tail_blocks * (tail_mult > 0 ? tail_mult : 1);
    // ... but never more than an eighth of the batch: the bulk belongs to the streaming kernels (B = 4,096 in three contexts:
    // 187 k solves/s with the hand-over at 296, 177 k at 2,368)
    if (tail_hand > B / 8) tail_hand = B / 8 > tail_blocks ? B / 8 : tail_blocks;
  }
  std::vector<long long> ub;
  {
    const long long t_team = (long long)h->sm_count * TEAM_WARPS_THIN * 32 / 16;                    // largest thin-team count
    const long long t_parts = P.parts ? ((long long)h->sm_count * 32) / (h->cfg.N + 1) : 0;          // largest part-kernel count
    const long long t_floor = 64;                                                                    // last class: grids for <= 64 instances
    long long v = B;
    ub.push_back(v);
    while (v > 2 * t_team) { v = (v + 1) / 2; ub.push_back(v); }
    for (long long t : {t_team, t_parts, t_floor}) if (t > 0 && t < ub.back() && t > tail_hand) ub.push_back(t);
  }
  for (size_t c = 0; c < ub.size(); ++c) {
    const int low = c + 1 < ub.size() ? (int)ub[c + 1] : tail_hand;   // run this class while more than `low` instances are active
    cudaGraphConditionalHandle hnd;
    CK(cudaGraphConditionalHandleCreate(&hnd, g, 0, 0));
    // entry condition: the trial list of the last round (parity 0 rounds write list 2) holds every active instance
    int* cnt = h->sg.cnt; int which = 2, above = low, add = 0;
    void* ac[] = {&hnd, &cnt, &which, &above, &add};
    I.launch((const void*)staged_cond_kernel, dim3(1), dim3(1), 0, ac);
    if (I.rc != MMPC_OK) return I.rc;
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = hnd; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    cudaGraphNode_t wn;
    CK(cudaGraphAddNode(&wn, g, &I.last, 1, &cp));
    I.last = wn;
    Issuer Bd; memset(&Bd, 0, sizeof Bd); Bd.g = cp.conditional.phGraph_out[0]; Bd.rc = MMPC_OK;
    issue_round(h, Bd, P, 1, ub[c], false, nomark);   // lists ping-pong over rounds: a body is an odd and an even round
    issue_round(h, Bd, P, 2, ub[c], false, nomark);
    add = 2;
    Bd.launch((const void*)staged_cond_kernel, dim3(1), dim3(1), 0, ac);
    if (Bd.rc != MMPC_OK) return Bd.rc;
  }
  if (tail_fn) {
    const long long LSs = P.LS;
    ResTail T; memset(&T, 0, sizeof T);
    T.ws = h->sg.ws; T.qp = h->sg.qp; T.gd = h->sg.gd; T.gi = h->sg.gi; T.list = h->sg.lists + 2 * LSs; T.cnt = h->sg.cnt; T.queue = h->queue + 1;
    T.LS = LSs; T.STG = P.STG; T.ITSZ = P.ITSZ; T.ND = P.ND;
    void* at[] = {&Ptail, &T};
    I.launch(tail_fn, dim3(tail_blocks), dim3(tail_threads), tail_smem, at);
    if (I.rc != MMPC_OK) return I.rc;
  }
  // the counters of the solve (rounds executed) for mmpc_launch_count / mmpc_phase_times
  {
    cudaGraphNode_t mn;
    CK(cudaGraphAddMemcpyNode1D(&mn, g, &I.last, 1, h->sg.pin + 32, h->sg.cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost));
  }
  CK(cudaGraphInstantiate(&E.exec, g, 0));
  E.cap = B; E.classes = (int)ub.size(); E.tail = tail_fn ? 1 : 0;
  return MMPC_OK;
}

static int launch_staged_graph(MmpcHandle* h, int32_t B, cudaStream_t st) {
  const int cap = graph_capacity(h, B);
  int slot = -1;
  for (int i = 0; i < 4; ++i) if (h->gr.e[i].exec && h->gr.e[i].cap == cap) slot = i;
  if (slot < 0) {
    slot = h->gr.next; h->gr.next = (h->gr.next + 1) & 3;
    int rc = graph_build(h, cap, slot);
    if (rc != MMPC_OK) { graph_destroy(h); return rc; }
  }
  h->sg.pin[32 + 3] = -1;
  CK(cudaGraphLaunch(h->gr.e[slot].exec, st));
  h->gr.pending = true; h->gr.pending_classes = h->gr.e[slot].classes; h->gr.pending_tail = h->gr.e[slot].tail;
  return MMPC_OK;
}

// launches the GPU executed for graph solves: 1 (io) + 1 (init) + 8 (round 0) + 7 per later round (+ the pose kernel in
// reference mode) + one condition kernel per class entry and per loop body; valid once the caller has synchronised the stream of the solve
static void graph_account(MmpcHandle* h) {
  if (!h->gr.pending) return;
  const int rounds = h->sg.pin[32 + 3];
  if (rounds < 0) return;  // the solve has not finished yet
  h->gr.pending = false;
  h->sg.rounds = 1 + rounds;
  const int ref = h->cfg.mode == MMPC_MODE_REFERENCE;   // the pose kernel: once more in round 0, once per round
  h->launches += 2 + 8 + 2 * ref + (7LL + ref) * rounds + h->gr.pending_classes + rounds / 2 + h->gr.pending_tail;
}

// ---- driver 2: the host sequences rounds (profiling; MMPC_HOSTLOOP=1) ----------------------------------------------------
static int launch_staged_hostloop(MmpcHandle* h, int32_t B, cudaStream_t st) {
  SParams P = staged_params(h, B);
  Issuer I; memset(&I, 0, sizeof I); I.st = st; I.rc = MMPC_OK;
  // profiling: one timing event in front of every launch; the time up to the next event is
  // charged to that launch's phase (events are stream-ordered, so this is device time)
  std::vector<int> marks;
  size_t nmark = 0;
  int mrc = MMPC_OK;
  auto mark = [&](int phase) {
    h->phase_launches[phase < 0 ? 0 : phase] += (phase >= 0);
    if (!h->profile || mrc != MMPC_OK) return;
    if (!h->prof_ev) h->prof_ev = new std::vector<cudaEvent_t>();
    if (nmark == h->prof_ev->size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { mrc = MMPC_ERR_CUDA; return; }
      h->prof_ev->push_back(e);
    }
    if (cudaEventRecord((*h->prof_ev)[nmark++], st) != cudaSuccess) { mrc = MMPC_ERR_CUDA; return; }
    marks.push_back(phase);
  };
  for (int i = 0; i < MMPC_NPHASE; ++i) { h->phase_ms[i] = 0; h->phase_launches[i] = 0; }
  mark(MMPC_PHASE_INIT);
  issue_init(h, I, P);
  const int LAG = 2;
  long long ub = B;  // upper bound of the active instances (the lists only shrink)
  int r = 0;
  for (;; ++r) {
    issue_round(h, I, P, r, ub, !P.fused || r == 0, mark);
    if (I.rc != MMPC_OK) return I.rc;
    if (mrc != MMPC_OK) { snprintf(g_err, sizeof g_err, "profiling event failed"); return mrc; }
    // the list lengths of this round trail the launches by LAG rounds
    int slot = r & 7;
    CK(cudaMemcpyAsync(h->sg.pin + 4 * slot, h->sg.cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->sg.ev[slot], st));
    if (r >= LAG) {
      int qs = (r - LAG) & 7;
      CK(cudaEventSynchronize(h->sg.ev[qs]));
      // every instance still active after a round is in that round's trial list, and the
      // active set only shrinks: its length bounds every later list
      long long nT = h->sg.pin[4 * qs + 1 + ((r - LAG + 1) & 1)];
      if (nT == 0) break;
      if (nT < ub) ub = nT;
    }
    if (r > 4000000) break;
  }
  if (h->profile) {
    mark(-1);
    CK(cudaEventSynchronize((*h->prof_ev)[nmark - 1]));
    for (size_t i = 0; i + 1 < nmark; ++i) {
      float ms = 0; CK(cudaEventElapsedTime(&ms, (*h->prof_ev)[i], (*h->prof_ev)[i + 1]));
      if (marks[i] >= 0) h->phase_ms[marks[i]] += ms;
    }
  }
  h->launches += I.issued + 1;
  h->sg.rounds = r + 1;
  return MMPC_OK;
}

static int launch_staged(MmpcHandle* h, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out, cudaStream_t st) {
  int rc = ensure_workspace(h);
  if (rc != MMPC_OK) return rc;
  graph_account(h);
  SIO io; memset(&io, 0, sizeof io);
  io.x_init = in->x_init; io.x_ref = in->x_ref; io.u_ref = in->u_ref; io.u_last = in->u_last; io.u_guess = in->u_guess;
  io.circles = in->circles; io.planes = in->planes; io.n_pl_inst = in->n_pl_inst; io.flags = in->flags; io.x_guess = in->x_guess;
  io.U = out->U; io.X = out->X; io.s = out->s; io.cost = out->cost; io.kkt = out->kkt; io.iters = out->iters; io.status = out->status;
  io.B = B;
  staged_set_io_kernel<<<1, 1, 0, st>>>(io, h->sg.io, h->queue + 1);
  CK(cudaGetLastError());
  // the resident kernel: forced, or AUTO's choice for small batches when the instance fits in shared memory.  Measured on a
  // B200 (reference NLP, resident / staged with its resident tail, ms; scripts/crossover.py): config 3 shapes B = 1 3.5 / -,
  // 148 12.8 / 13.5, 592 16.6 / 23.5, 1,024 28.2 / 26.3, 2,048 43.6 / 29.9; 3 circles, 2 planes 592 17.2 / 17.5, 1,480
  // 26.9 / 22.3; N = 40 with moving obstacles 444 59.7 / 64.5, 1,480 172 / 130.  Since the staged solve hands its tail to the
  // resident kernel the two differ little for a few hundred instances; beyond, the bulk belongs to the streaming kernels.
  // MMPC_AUTO_RESIDENT = largest batch, in instances per SM, AUTO sends to the resident kernel (0: never).
  static const int auto_resident = getenv("MMPC_AUTO_RESIDENT") ? atoi(getenv("MMPC_AUTO_RESIDENT")) : 4;
  const bool pose = h->cfg.model == MMPC_MODEL_POSEREF;   // compiled into the pose build of the resident kernel only
  const bool fits = (pose ? mmpc_resident_pose_smem_bytes(&h->cfg) : mmpc_resident_smem_bytes(&h->cfg)) <= h->smem_optin;
  if (pose && !fits) { snprintf(g_err, sizeof g_err, "MMPC_MODEL_POSEREF runs on the resident kernel; this horizon does not fit in shared memory"); return MMPC_ERR_UNSUPPORTED; }
  if (pose || (!h->profile && h->sg_fused && fits && (h->resident || (h->autosel && (long long)B <= (long long)auto_resident * h->sm_count)))) {
    CK(cudaMemsetAsync(h->queue, 0, sizeof(unsigned), st));
    cudaError_t e = (cudaError_t)(pose ? mmpc_resident_pose_launch : mmpc_resident_launch)(&h->cfg, B, h->sg.io, h->queue, h->sm_count, st);
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "resident kernel launch failed: %s", cudaGetErrorString(e)); return MMPC_ERR_CUDA; }
    h->launches += 2;
    h->sg.rounds = 0;
    h->last_solver = MMPC_KERNEL_RESIDENT;
    return MMPC_OK;
  }
  h->last_solver = MMPC_KERNEL_STAGED;
  static const bool force_hostloop = getenv("MMPC_HOSTLOOP") && atoi(getenv("MMPC_HOSTLOOP")) != 0;
  // (the unfused A/B variant evaluates in every round: host loop only)
  if (h->profile || force_hostloop || h->hostloop || !h->sg_fused) return launch_staged_hostloop(h, B, st);
  return launch_staged_graph(h, B, st);
}

extern "C" int mmpc_solve(MmpcHandle* h, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out, void* stream) {
  if (!h || !in || !out || B < 0 || B > h->B_max) return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;  // the empty batch needs no arrays
  if (!in->x_init || !in->x_ref || !in->u_ref || !in->u_last || !out->U || !out->status) return MMPC_ERR_ARG;
  if ((h->cfg.n_obs > 0 && !in->circles) || (h->cfg.n_pl > 0 && !in->planes)) return MMPC_ERR_ARG;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  // the terminal equality (flags) and the reference NLP's bug-for-bug rows (MMPC_MODE_REFERENCE) are implemented by
  // the fused trial + evaluation kernel only
  if ((in->flags || h->cfg.mode != MMPC_MODE_CLEAN) && !h->sg_fused) return MMPC_ERR_UNSUPPORTED;
  return launch_staged(h, B, in, out, st);
}

template <class T>
static int stage_alloc(T** d, T** hp, size_t n) {
  if (n == 0) n = 1;
  CK(cudaMalloc((void**)d, n * sizeof(T)));
  CK(cudaMallocHost((void**)hp, n * sizeof(T)));
  return MMPC_OK;
}

static int ensure_staging(MmpcHandle* h) {
  if (h->d.x_init) return MMPC_OK;
  size_t B = h->B_max, N = h->cfg.N;
  int rc;
#define SA(name, n) if ((rc = stage_alloc(&h->d.name, &h->h.name, (n))) != MMPC_OK) return rc
  SA(x_init, B * 9); SA(x_ref, B * (N + 1) * 9); SA(u_ref, B * N * 5); SA(u_last, B * N * 5); SA(u_guess, B * N * 5);
  SA(circles, B * circles_per_instance(h->cfg)); SA(planes, B * (size_t)h->cfg.n_pl * 6); SA(x_guess, B * (N + 1) * 9);
  SA(U, B * N * 5); SA(X, B * (N + 1) * 9); SA(s, B * (N + 1)); SA(cost, B); SA(kkt, B);
  SA(n_pl_inst, B); SA(iters, B); SA(status, B); SA(flags, B);
#undef SA
  return MMPC_OK;
}

extern "C" int mmpc_solve_host(MmpcHandle* h, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out) {
  if (!h || !in || !out || B < 0 || B > h->B_max) return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;  // the empty batch needs no arrays
  if (!in->x_init || !in->x_ref || !in->u_ref || !in->u_last || !out->U || !out->status) return MMPC_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc = ensure_staging(h);
  if (rc != MMPC_OK) return rc;
  size_t N = h->cfg.N, b = B;
  cudaStream_t st = h->stream;
  MmpcBatchIn din; memset(&din, 0, sizeof din);
  MmpcBatchOut dout; memset(&dout, 0, sizeof dout);
  // a caller's array that already is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) is copied by DMA
  // straight from / to where it lies; a pageable one goes through this handle's pinned staging buffer
  auto pinned = [](const void* q) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
#define UP(name, n, T)                                                                       \
  if (in->name) {                                                                            \
    const void* src_ = in->name;                                                             \
    if (!pinned(src_)) { memcpy(h->h.name, src_, (n) * sizeof(T)); src_ = h->h.name; }       \
    CK(cudaMemcpyAsync(h->d.name, src_, (n) * sizeof(T), cudaMemcpyHostToDevice, st));       \
    din.name = h->d.name;                                                                    \
  }
  UP(x_init, b * 9, double); UP(x_ref, b * (N + 1) * 9, double); UP(u_ref, b * N * 5, double);
  UP(u_last, b * N * 5, double); UP(u_guess, b * N * 5, double);
  UP(circles, b * circles_per_instance(h->cfg), double); UP(planes, b * (size_t)h->cfg.n_pl * 6, double);
  UP(n_pl_inst, b, int32_t); UP(flags, b, uint8_t); UP(x_guess, b * (N + 1) * 9, double);
#undef UP
  dout.U = h->d.U; dout.status = h->d.status;
  if (out->X) dout.X = h->d.X;
  if (out->s) dout.s = h->d.s;
  if (out->cost) dout.cost = h->d.cost;
  if (out->kkt) dout.kkt = h->d.kkt;
  if (out->iters) dout.iters = h->d.iters;
  rc = mmpc_solve(h, B, &din, &dout, st);
  if (rc != MMPC_OK) return rc;
  bool direct[7]; int di = 0;
#define DOWN(name, n, T)                                                                                      \
  {                                                                                                            \
    direct[di] = out->name && pinned(out->name);                                                               \
    if (out->name) CK(cudaMemcpyAsync(direct[di] ? (void*)out->name : (void*)h->h.name, h->d.name, (n) * sizeof(T), cudaMemcpyDeviceToHost, st)); \
    ++di;                                                                                                      \
  }
  DOWN(U, b * N * 5, double); DOWN(X, b * (N + 1) * 9, double); DOWN(s, b * (N + 1), double);
  DOWN(cost, b, double); DOWN(kkt, b, double); DOWN(iters, b, int32_t); DOWN(status, b, int32_t);
#undef DOWN
  CK(cudaStreamSynchronize(st));
  di = 0;
#define OUTC(name, n, T) { if (out->name && !direct[di]) memcpy(out->name, h->h.name, (n) * sizeof(T)); ++di; }
  OUTC(U, b * N * 5, double); OUTC(X, b * (N + 1) * 9, double); OUTC(s, b * (N + 1), double);
  OUTC(cost, b, double); OUTC(kkt, b, double); OUTC(iters, b, int32_t); OUTC(status, b, int32_t);
#undef OUTC
  graph_account(h);
  return MMPC_OK;
}

// ---- model evaluation / shift / plant step: one thread per instance ----------------------------
__global__ void eval_model_kernel(MmpcConfig cfg, int M, const double* x, const double* u, const double* circles,
                                  const double* planes, double* fo, double* fko, double* rows) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double xs[NX], us[NU] = {0, 0, 0, 0, 0};
  for (int i = 0; i < NX; ++i) xs[i] = x[(size_t)m * NX + i];
  if (u) for (int j = 0; j < NU; ++j) us[j] = u[(size_t)m * NU + j];
  FK f; fk_eval(xs[2], xs[6], xs[7], xs[8], f);
  if (fo) { double xn[NX]; dyn_f(xs, us, cfg.dt, f.cp, f.sp, xn); for (int i = 0; i < NX; ++i) fo[(size_t)m * NX + i] = xn[i]; }
  if (fko) {
    Point e, j2, j3;
    point_eval(xs[0], xs[1], f, BODY[5], e); point_eval(xs[0], xs[1], f, BODY[1], j2); point_eval(xs[0], xs[1], f, BODY[3], j3);
    double* o = fko + (size_t)m * 10;
    o[0] = e.P[0]; o[1] = e.P[1]; o[2] = e.P[2]; o[3] = xs[2];
    o[4] = j2.P[0]; o[5] = j2.P[1]; o[6] = j2.P[2]; o[7] = j3.P[0]; o[8] = j3.P[1]; o[9] = j3.P[2];
  }
  if (rows) {
    int nr = cfg.n_obs + 4 + 6 * cfg.n_pl;
    double* o = rows + (size_t)m * nr;
    for (int i = 0; i < cfg.n_obs; ++i) {
      const double* c = circles + ((size_t)m * cfg.n_obs + i) * 3;
      double dx = xs[0] - c[0], dy = xs[1] - c[1];
      o[i] = (c[2] + cfg.base_radius) - sqrt(dx * dx + dy * dy);
    }
    for (int q = 0; q < 4; ++q) {
      Point p; point_eval(xs[0], xs[1], f, SELFD[q], p);
      o[cfg.n_obs + q] = cfg.self_collision_radius - sqrt(p.P[0] * p.P[0] + p.P[1] * p.P[1] + p.P[2] * p.P[2]);
    }
    for (int i = 0; i < 6; ++i) {
      Point p; point_eval(xs[0], xs[1], f, BODY[i], p);
      for (int j = 0; j < cfg.n_pl; ++j) {
        const double* pl = planes + ((size_t)m * cfg.n_pl + j) * 6;
        double c = 0;
        for (int a = 0; a < 3; ++a) c += pl[3 + a] * ((pl[a] - cfg.obstacle_expand_dist * pl[3 + a]) - p.P[a]);
        o[cfg.n_obs + 4 + i * cfg.n_pl + j] = c;
      }
    }
  }
}

extern "C" int mmpc_eval_model(MmpcHandle* h, int32_t M, const double* x, const double* u, const double* circles,
                               const double* planes, double* f, double* fk, double* rows, void* stream) {
  if (!h || !x || M < 0) return MMPC_ERR_ARG;
  if (rows && ((h->cfg.n_obs > 0 && !circles) || (h->cfg.n_pl > 0 && !planes))) return MMPC_ERR_ARG;
  if (M == 0) return MMPC_OK;
  CK(cudaSetDevice(h->device));
  eval_model_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->cfg, M, x, u, circles, planes, f, fk, rows);
  CK(cudaGetLastError());
  h->launches += 1;
  return MMPC_OK;
}

__global__ void shift_kernel(int B, int N, const double* U, double* ug) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)B * N * NU;
  if (i >= tot) return;
  int j = (int)(i % NU); long long r = i / NU; int k = (int)(r % N); long long b = r / N;
  int ks = k + 1 < N ? k + 1 : N - 1;
  ug[i] = U[(b * N + ks) * NU + j];
}

extern "C" int mmpc_shift(MmpcHandle* h, int32_t B, const double* U, double* u_guess, void* stream) {
  if (!h || !U || !u_guess || B < 0) return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;
  CK(cudaSetDevice(h->device));
  long long tot = (long long)B * h->cfg.N * NU;
  shift_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, h->cfg.N, U, u_guess);
  CK(cudaGetLastError());
  h->launches += 1;
  return MMPC_OK;
}

__global__ void plant_kernel(int B, double dt, const double* x, const double* u0, double* xn) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double xs[NX], us[NU], o[NX];
  for (int i = 0; i < NX; ++i) xs[i] = x[(size_t)b * NX + i];
  for (int j = 0; j < NU; ++j) us[j] = u0[(size_t)b * NU + j];
  double sp, cp; sincos(xs[2], &sp, &cp);
  dyn_f(xs, us, dt, cp, sp, o);
  for (int i = 0; i < NX; ++i) xn[(size_t)b * NX + i] = o[i];
}

extern "C" int mmpc_plant_step(MmpcHandle* h, int32_t B, const double* x, const double* u0, double* x_next, void* stream) {
  if (!h || !x || !u0 || !x_next || B < 0) return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;
  CK(cudaSetDevice(h->device));
  plant_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, h->cfg.dt, x, u0, x_next);
  CK(cudaGetLastError());
  h->launches += 1;
  return MMPC_OK;
}

extern "C" int mmpc_workspace_bytes(const MmpcHandle* h, int64_t* bytes) {
  if (!h || !bytes) return MMPC_ERR_ARG;
  const MmpcConfig& c = h->cfg;
  long long LS = ((long long)h->B_max + 31) / 32 * 32, st = c.N + 1;
  *bytes = (int64_t)(8 * LS * (st * (staged_stage_doubles(c) + QS + RS) + staged_inst_doubles(c)) + 4 * LS * (J_NFIELDS + 2));
  return MMPC_OK;
}

extern "C" int mmpc_set_profile(MmpcHandle* h, int32_t on) {
  if (!h) return MMPC_ERR_ARG;
  h->profile = on != 0;
  return MMPC_OK;
}

extern "C" int mmpc_phase_times(const MmpcHandle* h, double* ms, int64_t* launches, int32_t* rounds) {
  if (!h) return MMPC_ERR_ARG;
  for (int i = 0; i < MMPC_NPHASE; ++i) {
    if (ms) ms[i] = h->phase_ms[i];
    if (launches) launches[i] = h->phase_launches[i];
  }
  graph_account(const_cast<MmpcHandle*>(h));
  if (rounds) *rounds = h->sg.rounds;
  return MMPC_OK;
}

// calcLocalRefTraj (interface_wholebody_qref.py:353-396) on device: nearest row of the instance's global
// reference by Euclidean distance over the state indices in idx_mask (first minimum, like np.argmin), rows
// [i*, i*+N], the last row repeated past the end (:385-389); u_ref window likewise (NULL u_glob -> zeros).
__global__ void window_kernel(int B, int N, int M, int idx_mask, int shared_ref, const double* x, const double* x_glob,
                              const double* u_glob, double* x_ref, double* u_ref, int32_t* i_star) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* g = x_glob + (shared_ref ? 0 : (size_t)b * M * NX);
  double xs[NX];
  for (int i = 0; i < NX; ++i) xs[i] = x[(size_t)b * NX + i];
  double best = 1e300; int ib = 0;
  for (int j = 0; j < M; ++j) {
    double d2 = 0;
    for (int i = 0; i < NX; ++i) if (idx_mask >> i & 1) { double e = g[(size_t)j * NX + i] - xs[i]; d2 += e * e; }
    double d = sqrt(d2);  // np.linalg.norm: compare the same quantity the reference compares
    if (d < best) { best = d; ib = j; }
  }
  if (i_star) i_star[b] = ib;
  for (int k = 0; k <= N; ++k) {
    int r = ib + k < M - 1 ? ib + k : M - 1;
    for (int i = 0; i < NX; ++i) x_ref[((size_t)b * (N + 1) + k) * NX + i] = g[(size_t)r * NX + i];
  }
  if (u_ref)
    for (int k = 0; k < N; ++k) {
      int r = ib + k < M - 2 ? ib + k : M - 2;
      for (int j = 0; j < NU; ++j)
        u_ref[((size_t)b * N + k) * NU + j] = u_glob ? u_glob[(shared_ref ? 0 : (size_t)b * (M - 1) * NU) + (size_t)r * NU + j] : 0.0;
    }
}

extern "C" int mmpc_window(MmpcHandle* h, int32_t B, int32_t M, int32_t idx_mask, int32_t shared_ref, const double* x,
                           const double* x_glob, const double* u_glob, double* x_ref, double* u_ref, int32_t* i_star,
                           void* stream) {
  if (!h || !x || !x_glob || !x_ref || B < 0 || M < 2 || !(idx_mask & 0x1ff)) return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;
  CK(cudaSetDevice(h->device));
  window_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, h->cfg.N, M, idx_mask, shared_ref, x, x_glob, u_glob,
                                                                    x_ref, u_ref, i_star);
  CK(cudaGetLastError());
  h->launches += 1;
  return MMPC_OK;
}

extern "C" int mmpc_ik(MmpcHandle* h, int32_t B, const double* q_guess, const double* target, double* q_out, int32_t* status,
                       void* stream) {
  if (!h || !q_guess || !target || !q_out || B < 0) return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;
  CK(cudaSetDevice(h->device));
  ik_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, q_guess, target, q_out, status);
  CK(cudaGetLastError());
  h->launches += 1;
  return MMPC_OK;
}

extern "C" int mmpc_episode_update(MmpcHandle* h, int32_t B, int32_t M, int32_t n_manip, const MmpcEpisodeIO* io, void* stream) {
  if (!h || !io || B < 0 || n_manip < 1 || M < n_manip + 1) return MMPC_ERR_ARG;
  if (!io->x || !io->pose_target || !io->traj || !io->traj_len || !io->task || !io->flags || !io->wset || !io->active ||
      !io->x_ref || !io->u_ref)
    return MMPC_ERR_ARG;
  if (B == 0) return MMPC_OK;
  CK(cudaSetDevice(h->device));
  EpisodeArgs A; A.B = B; A.N = h->cfg.N; A.M = M; A.n_manip = n_manip; A.io = *io;
  episode_update_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A);
  CK(cudaGetLastError());
  h->launches += 1;
  return MMPC_OK;
}

extern "C" int64_t mmpc_launch_count(const MmpcHandle* h) {
  if (!h) return 0;
  graph_account(const_cast<MmpcHandle*>(h));  // rounds of the last graph solve (the caller has synchronised its stream)
  return h->launches;
}

extern "C" int mmpc_last_solver(const MmpcHandle* h) { return h ? h->last_solver : MMPC_KERNEL_AUTO; }

extern "C" int mmpc_occupancy(const MmpcHandle* h, int32_t* sm_count, int32_t* blocks_per_sm, int32_t* smem_bytes) {
  if (!h) return MMPC_ERR_ARG;
  if (sm_count) *sm_count = h->sm_count;
  if (blocks_per_sm) *blocks_per_sm = TEAM_WARPS_BULK;                                  // resident warps per SM of the Riccati kernel
  if (smem_bytes) *smem_bytes = (int32_t)(Team::SMEM_DOUBLES * 2 * sizeof(double));     // its ring, per warp (two instances)
  return MMPC_OK;
}

// ---- FP64 FMA peak micro-benchmark: the denominator of the roofline (MEASURED_PEAKS.json has no
// FP64 entry).  8 independent DFMA chains per thread. -----------------------------------------
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
    v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
    v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

extern "C" int mmpc_bench_fp64(int32_t device, double* tflops) {
  if (!tflops) return MMPC_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MMPC_ERR_NO_DEVICE; }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* out; CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = (double)blocks * threads * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return MMPC_OK;
}
