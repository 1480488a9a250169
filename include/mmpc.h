/*
 * mmpc.h -- C ABI of the B200-native batched whole-body MPC solver.
 *
 * This is the drop-in boundary for the ONE hot path of HsinyuG/mobile-manipulator-mpc:
 *     MPCWholeBody.solve(x_init, traj_ref, u_ref)      controllers/mpc_wholebody_qref.py:287-331
 * whose NLP is defined in MPCWholeBody.reset()          controllers/mpc_wholebody_qref.py:142-285
 * (all file:line citations in this header are relative to the reference checkout).
 *
 * The reference has no FFI of its own (it is pure Python over the casadi wheel), so the entry
 * points below are what a ctypes binding of that class binds -- see INTEGRATION.md for the
 * stub a maintainer would add.  Plain pointers and sizes only; no torch types; no exceptions
 * across the ABI.  Return value: 0 = MMPC_OK, anything else = API misuse / CUDA failure
 * (mmpc_error_string()).  Numerical failure of an instance is NOT an error code: it is
 * reported per instance in `status`.
 *
 * Batch layout ("instance-major"): every per-instance array is the reference's own NumPy
 * array with one leading batch axis, C-contiguous, float64:
 *     x_init [B][9]        solve() arg 1            (:287)   x y psi dx dy dpsi q1 q2 q3
 *     x_ref  [B][N+1][9]   solve() arg 2 traj_ref   (:306)
 *     u_ref  [B][N][5]     solve() arg 3            (:307)   dV dw dq1 dq2 dq3
 *     u_last [B][N][5]     self.u_latest            (:295-310)  previous U*, zeros on first call
 *     u_guess[B][N][5]     initial guess of U (NULL -> u_last, the reference behaviour :303)
 *     circles[B][n_obs][3] or [B][N+1][n_obs][3]   Obstacles(x, y, radius)  robot_models/obstacles.py:6-10
 *     planes [B][n_pl][6]  obstacle_manipulation_list entries (point xyz, normal xyz)  demo_wholebody_qref.py:21-33
 *     n_pl_inst[B] int32   planes actually used by instance b (NULL -> cfg.n_pl for all)
 *     flags  [B] uint8     bit0: terminal xy equality  interface_wholebody_qref.py:167
 *     x_guess[B][N+1][9]   initial guess of X (NULL -> tile(x_init), the reference behaviour :302)
 * Outputs:
 *     U [B][N][5]  X [B][N+1][9]  s [B][N+1]   sol.value(U/X/s)   (:329-330)
 *     cost[B]  sol.value(cost) (:317)     kkt[B]  final scaled NLP error (IPOPT E_0)
 *     iters[B] int32   status[B] int32 (MMPC_STATUS_*)
 */
#ifndef MMPC_H_
#define MMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMPC_NX 9
#define MMPC_NU 5
#define MMPC_MAX_PLANES 4

/* return codes */
enum {
  MMPC_OK = 0,
  MMPC_ERR_ARG = 1,       /* NULL pointer / bad size / B > B_max            */
  MMPC_ERR_CUDA = 2,      /* CUDA runtime error (see mmpc_error_string)     */
  MMPC_ERR_NO_DEVICE = 3, /* no sm_100-class CUDA device: there is NO CPU fallback */
  MMPC_ERR_UNSUPPORTED = 4
};

/* per-instance status */
enum {
  MMPC_STATUS_CONVERGED = 0,   /* scaled KKT error E_0 <= tol  (IPOPT Solve_Succeeded)               */
  MMPC_STATUS_MAX_ITER = 1,    /* iteration cap reached        (IPOPT Maximum_Iterations_Exceeded)   */
  MMPC_STATUS_LINESEARCH = 2,  /* step size underflow          (IPOPT Restoration_Failed analogue)   */
  MMPC_STATUS_FACTOR = 3,      /* regularisation cap reached in the Riccati factorisation            */
  MMPC_STATUS_NAN = 4,         /* non-finite iterate                                                  */
  MMPC_STATUS_ACCEPTABLE = 5   /* E_0 <= acceptable_tol for acceptable_iter iterations (:283-284)    */
};

/* NLP variants: SURVEY.md section 8(a) rows 7-9 */
enum {
  MMPC_MODE_REFERENCE = 0, /* bug-for-bug: stale plane columns, terminal self-collision on s[N-1] */
  MMPC_MODE_CLEAN = 1      /* stage-separable: -max_j c_k[i,j] <= s_k only, terminal rows on s[N]  */
};

/* Which of the reference's controllers the handle solves.
 *   WHOLEBODY  controllers/mpc_wholebody_qref.py::MPCWholeBody (the hot path; everything above)
 *   BASE       controllers/mpc_base.py::MPCBase (:114-189): the unicycle base alone -- 6 states x y psi dx dy dpsi, 2 controls
 *              dV dw, ground circles against one slack per stage (weight M), boxes on x y dx dy dpsi and on u, the yaw error
 *              of the cost taken through angleDiff (:59-84), and X warm-started from the previous solution (:196-201).  It is
 *              solved as the whole-body NLP with the arm taken out: arrays keep the 9 / 5 layout, the three arm states and
 *              controls are unbounded, carry a unit control weight and no state weight, so they stay at their initial
 *              values and do not interact with the base (no FK row exists: no self-collision rows, n_pl must be 0, mode
 *              MMPC_MODE_CLEAN).  S plays the role of M.  Host class: mobile_manipulator_mpc_b200/controllers/mpc_base.py. */
enum { MMPC_MODEL_WHOLEBODY = 0, MMPC_MODEL_BASE = 1,
       MMPC_MODEL_POSEREF = 2   /* controllers/mpc_wholebody.py:49-128: the whole-body model with the tracking cost on the END-POINT
                                   POSE (x, y, z, psi) = forward_tranformation(x)[0] instead of on the state; circle rows only (no
                                   self-collision rows, n_pl must be 0).  x_ref[B, N+1, 9] carries the pose reference in its first
                                   four columns, Qd[0..3] / Pd[0..3] are the pose weights (the other entries must be 0), x_guess
                                   warm-starts X (:138-141).  Runs on the resident kernel (any batch size) */ };

/* execution strategy of mmpc_solve (same algorithm, same results to the bit):
 *   STAGED   batch-synchronous rounds of phase kernels over compacted lists of active instances (stage-parallel
 *            evaluation / step / trial, 16-lane column-parallel Riccati), the whole solve one CUDA graph whose
 *            conditional WHILE nodes loop over the rounds on the device: the throughput path, AUTO's choice for batches
 *   RESIDENT the same phase bodies inside one persistent thread block per instance with the state in shared memory: the
 *            latency path, AUTO's choice for small batches
 * The others are A/B references of single design decisions.  (Values 1 and 2 named two superseded kernels.) */
enum { MMPC_KERNEL_AUTO = 0, MMPC_KERNEL_STAGED = 3,
       MMPC_KERNEL_STAGED_THREAD = 4,   /* STAGED with the one-thread-per-instance Riccati                         */
       MMPC_KERNEL_STAGED_UNFUSED = 5,  /* STAGED with separate eval and trial kernels (host loop)                 */
       MMPC_KERNEL_STAGED_FAT = 6,      /* STAGED with one thread per (instance, stage) item in the thin rounds too */
       MMPC_KERNEL_STAGED_HOSTLOOP = 7, /* STAGED with the host sequencing the rounds instead of the CUDA graph    */
       MMPC_KERNEL_RESIDENT = 8         /* one thread block per instance, the whole solver state of the instance in shared memory,
                                           persistent blocks on an atomic work queue (csrc/mmpc_resident.cu): the latency path.
                                           AUTO takes it for batches of at most four instances per SM when the state fits
                                           (N = 20 with 16 circles: 112 KB, two blocks per SM); MMPC_ERR_UNSUPPORTED if it
                                           does not fit (N = 63) */ };

typedef struct MmpcConfig {
  int32_t N;             /* horizon; demo_wholebody_qref.py:11 uses 20, class default 10 (:11)   */
  int32_t n_obs;         /* ground circles per instance                                          */
  int32_t n_pl;          /* max planes per instance (<= MMPC_MAX_PLANES), 0 = none (:224)        */
  int32_t mode;          /* MMPC_MODE_*                                                          */
  int32_t obs_per_stage; /* 0: circles[B][n_obs][3]; 1: circles[B][N+1][n_obs][3] (moving)       */
  int32_t max_iter;      /* 'ipopt.max_iter': 2000 (:280)                                        */
  int32_t terminal_rows_on_sN; /* MMPC_MODE_REFERENCE: 0 (default) = the four terminal self-collision rows are bounded by s[N-1],
                                  the reference's leaked loop variable (:263-265); 1 = by s[N] (what a
                                  reader of the reference would expect).  Same optimum unless the barrier path forks. */
  int32_t model;         /* MMPC_MODEL_*: 0 = MPCWholeBody (controllers/mpc_wholebody_qref.py), 1 = MPCBase (controllers/mpc_base.py) */
  double dt;             /* robot.dt, demo_wholebody_qref.py:10                                  */
  double Qd[9], Pd[9], Rd[5], Wd[5], S; /* diagonals of Q,P,R,W and S (:12-16); setWeight :119   */
  double ulim[2][5];     /* (:17)  rows: lower, upper                                            */
  double xlim[2][9];     /* (:18-21) +-inf allowed                                               */
  double dulim[2][5];    /* (:22)                                                                */
  double base_radius;    /* robot_models/base.py:15                                              */
  double self_collision_radius; /* :43 */
  double obstacle_expand_dist;  /* :44 */
  double tol;            /* IPOPT tol default 1e-8                                               */
  double mu_init;        /* IPOPT mu_init default 0.1                                            */
  double acceptable_tol; /* 'ipopt.acceptable_tol': 1e-8 (:283)                                  */
} MmpcConfig;

typedef struct MmpcBatchIn {
  const double* x_init;
  const double* x_ref;
  const double* u_ref;
  const double* u_last;
  const double* u_guess;    /* may be NULL */
  const double* circles;    /* may be NULL iff n_obs == 0 */
  const double* planes;     /* may be NULL iff n_pl == 0 */
  const int32_t* n_pl_inst; /* may be NULL */
  const uint8_t* flags;     /* may be NULL */
  const double* x_guess;    /* may be NULL: X <- tile(x_init), what MPCWholeBody.solve does (:302).  [B][N+1][9]: initial guess of
                               X[1..N] (row 0 is ignored: X[0] = x_init); MPCBase warm-starts X like this (mpc_base.py:196-201) */
} MmpcBatchIn;

typedef struct MmpcBatchOut {
  double* U;
  double* X;      /* may be NULL */
  double* s;      /* may be NULL */
  double* cost;   /* may be NULL */
  double* kkt;    /* may be NULL */
  int32_t* iters; /* may be NULL */
  int32_t* status;
} MmpcBatchOut;

typedef struct MmpcHandle MmpcHandle;

int mmpc_version(void);
const char* mmpc_error_string(int code);

/* Fill cfg with the reference defaults: controllers/mpc_wholebody_qref.py:11-22,43-44,280-285,
 * N = 20 and dt = 0.1 from demo_wholebody_qref.py:10-11. */
void mmpc_default_config(MmpcConfig* cfg);

/* Replaces MPCWholeBody.__init__/reset() (:7-46,142-285): fixes the NLP shape, allocates the
 * per-instance solver workspace for up to B_max instances on CUDA device `device`. */
int mmpc_create(const MmpcConfig* cfg, int32_t B_max, int32_t device, MmpcHandle** out);
int mmpc_destroy(MmpcHandle* h);

/* Replaces MPCWholeBody.setWeight (:119-139); diagonals only (every use in the reference is diagonal). */
int mmpc_set_weights(MmpcHandle* h, const double* Qd, const double* Pd, const double* Rd,
                     const double* Wd, double S);

/* Selects the execution strategy of the following mmpc_solve calls (MMPC_KERNEL_*; default AUTO). */
int mmpc_set_kernel(MmpcHandle* h, int32_t kernel);

/* Replaces MPCWholeBody.solve (:287-331) for B independent instances.  All pointers are DEVICE
 * pointers owned by the caller; asynchronous on `stream` (a cudaStream_t).  x_init is clipped to
 * xlim on device exactly as :290-291 does (written back in place to in->x_init's q entries is the
 * caller's business: the host wrapper does it). */
int mmpc_solve(MmpcHandle* h, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out, void* stream);

/* Same, with HOST pointers: stages through pinned buffers, copies host->device, solves and
 * copies the results back; synchronous.  This is the call the reference-facing Python class uses. */
int mmpc_solve_host(MmpcHandle* h, int32_t B, const MmpcBatchIn* in, const MmpcBatchOut* out);

/* Model evaluation for parity tests (DEVICE pointers), M independent (x,u) pairs:
 *   f   [M][9]   MobileManipulator.f_kinematics          robot_models/mobile_manipulator.py:57-75
 *   fk  [M][10]  forward_tranformation: endpoint x y z psi, joint2 xyz, joint3 xyz  (:17-55)
 *   rows[M][n_obs + 4 + 6*n_pl]  circle rows (:49-54), self-collision rows (:219-222),
 *                plane margins c[i][j], i-major (:76-80), all evaluated at x
 * circles [M][n_obs][3], planes [M][n_pl][6]; any output may be NULL. */
int mmpc_eval_model(MmpcHandle* h, int32_t M, const double* x, const double* u, const double* circles,
                    const double* planes, double* f, double* fk, double* rows, void* stream);

/* Warm-start shift of an initial guess on device: u_guess[b][k] = U[b][k+1], last row repeated
 * (north_star "shift kernel"; only the GUESS may be shifted). */
int mmpc_shift(MmpcHandle* h, int32_t B, const double* U, double* u_guess, void* stream);

/* Plant step without pybullet: x_next[b] = f_kinematics(x[b], u0[b])  interface_wholebody_qref.py:143 */
int mmpc_plant_step(MmpcHandle* h, int32_t B, const double* x, const double* u0, double* x_next,
                    void* stream);

/* Reference window on device: calcLocalRefTraj (interface_wholebody_qref.py:353-396).  For each instance the
 * nearest row i* of its global reference x_glob ([M][9], or [B][M][9] when shared_ref == 0) by Euclidean
 * distance over the state indices set in idx_mask (bit i = state i; the Interface uses {0,1} while moving :188
 * and {6,7,8} while manipulating :226), then x_ref[b] = rows [i*, i*+N] with the last row repeated (:385-389);
 * u_ref[b] = the same window of u_glob ([M-1][5] / [B][M-1][5]; NULL -> zeros, :266).  i_star may be NULL. */
int mmpc_window(MmpcHandle* h, int32_t B, int32_t M, int32_t idx_mask, int32_t shared_ref, const double* x,
                const double* x_glob, const double* u_glob, double* x_ref, double* u_ref, int32_t* i_star, void* stream);

/* ---- the callers either side of the solve, batched -------------------------- */

/* Batched inverse kinematics: ManipulatorPanda3DoF.inverse_transformation (robot_models/manipulator_3DoF.py:79-133),
 *     min (x(q) - xt)^2 + (z(q) - zt)^2   s.t.  q1 in [-pi/2, pi/2], q2 in [-3pi/4, 0], q3 in [0, 3pi/2]   (:123)
 * q_guess [B][3], target [B][3] = (x, y, z) in the arm frame with y == 0 (:99), q_out [B][3]; status[b] (may be
 * NULL) = 0, or 1 when no q reaches the target (the reference raises ValueError :124-125).  Device pointers. */
int mmpc_ik(MmpcHandle* h, int32_t B, const double* q_guess, const double* target, double* q_out, int32_t* status,
            void* stream);

/* Task flag of an episode: Interface.task_flag (interface_wholebody_qref.py:81, :146-228). */
enum { MMPC_TASK_MOVE = 0, MMPC_TASK_APPROACH = 1, MMPC_TASK_ROTATE = 2, MMPC_TASK_MOVE_FINISH = 3, MMPC_TASK_MANIPULATE = 4,
       MMPC_TASK_FINISHED = 5 /* 'manipulate finish', robot_status False :223 */, MMPC_TASK_IK_FAILED = 6 };

/* Per-episode arrays of mmpc_episode_update (device pointers, one leading batch axis). */
typedef struct MmpcEpisodeIO {
  const double* x;            /* [B][9]    current_state                                                (:134) */
  const double* pose_target;  /* [B][4]    global_pose_target x y z psi                                 (:20)  */
  double* traj;               /* [B][M][9] traj_ref of the current phase; rewritten by globalPlanManipulator (:277-297) */
  int32_t* traj_len;          /* [B]       rows of traj in use                                                  */
  int32_t* task;              /* [B]       MMPC_TASK_* (in/out)                                                 */
  uint8_t* flags;             /* [B]       solver flags; bit 0 is set when the episode enters 'approach' (:167) */
  int32_t* wset;              /* [B] out   weights of this step's solve: 0 = constructor defaults, 1 = 'rotate' set (:176-178),
                                           2 = 'manipulate' set (:212-215)                                      */
  int32_t* active;            /* [B] out   1 = solve and step this episode, 0 = finished / IK failed            */
  double* x_ref;              /* [B][N+1][9] out  local_traj_ref (calcLocalRefTraj :353-396 / calcLocalRefPose :398-410) */
  double* u_ref;              /* [B][N][5]   out  local_u_ref (identically zero in the reference :266, :296)    */
  double* local_pose_target;  /* [B][3] out at the manipulate hand-off (:206-210), may be NULL                   */
  int32_t* ik_status;         /* [B]    out at the manipulate hand-off, may be NULL                              */
} MmpcEpisodeIO;

/* Interface.stateMachineUpdate (interface_wholebody_qref.py:146-228) for B episodes: phase transitions, the local
 * reference of the step, the terminal-equality flag, the weight set, and at the move -> manipulate hand-off the
 * inverse kinematics and the joint-space plan with n_manip = t_manipulate / dt intervals.  M = rows allocated per
 * episode in io->traj (>= n_manip + 1). */
int mmpc_episode_update(MmpcHandle* h, int32_t B, int32_t M, int32_t n_manip, const MmpcEpisodeIO* io, void* stream);

/* Phases of the staged solver (one kernel each per round; COMPACT runs twice per round).  mmpc_phase_times fills MMPC_NPHASE entries. */
enum { MMPC_PHASE_COMPACT = 0, MMPC_PHASE_EVAL = 1, MMPC_PHASE_SOLVE = 2, MMPC_PHASE_STEP = 3, MMPC_PHASE_CTRL_STEP = 4,
       MMPC_PHASE_TRIAL = 5, MMPC_PHASE_CTRL_TRIAL = 6, MMPC_PHASE_INIT = 7,
       MMPC_PHASE_POSE = 8, /* MMPC_MODE_REFERENCE: forward kinematics + plane margins of every stage, once per evaluated point */
       MMPC_NPHASE = 9 };

/* Device-side timing of the phases (measurement support for bench.py's roofline).  With profiling on,
 * mmpc_solve brackets every launch of the staged solver with CUDA events on its stream and
 * synchronises at the end of the call; mmpc_phase_times returns, for the last mmpc_solve, the
 * accumulated device milliseconds and the number of launches per phase and the number of rounds. */
int mmpc_set_profile(MmpcHandle* h, int32_t on);
int mmpc_phase_times(const MmpcHandle* h, double* ms, int64_t* launches, int32_t* rounds);

/* Device bytes of the staged solver's per-instance state for B_max instances (allocated by the first
 * mmpc_solve; memory sizing for the 180 GB of one B200). */
int mmpc_workspace_bytes(const MmpcHandle* h, int64_t* bytes);

/* Number of kernel launches issued through this handle so far (bench.py's gpu_launches). */
int64_t mmpc_launch_count(const MmpcHandle* h);

/* Which execution strategy the last mmpc_solve of this handle took: MMPC_KERNEL_RESIDENT or MMPC_KERNEL_STAGED (any of the
 * staged variants); MMPC_KERNEL_AUTO before the first solve. */
int mmpc_last_solver(const MmpcHandle* h);

/* sizeof() of the three ABI structs, for binding self-checks. */
int mmpc_struct_sizes(int32_t* cfg_bytes, int32_t* in_bytes, int32_t* out_bytes);

/* Launch geometry chosen by mmpc_create: SM count, resident warps (= instances) per SM and the
 * dynamic shared memory one instance occupies. */
int mmpc_occupancy(const MmpcHandle* h, int32_t* sm_count, int32_t* blocks_per_sm, int32_t* smem_bytes);

/* FP64 FMA peak of `device` in TFLOP/s, measured by a register-resident DFMA micro-benchmark
 * (the roofline denominator of bench.py; MEASURED_PEAKS.json carries no FP64 figure). */
int mmpc_bench_fp64(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MMPC_H_ */
