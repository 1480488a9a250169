// mmpc_episode.cuh -- the callers either side of the solve, for B episodes at once:
//
//   ik_solve            ManipulatorPanda3DoF.inverse_transformation   robot_models/manipulator_3DoF.py:79-133
//   episode_update      Interface.stateMachineUpdate                  interface_wholebody_qref.py:146-228
//                       + calcLocalRefTraj :353-396, calcLocalRefPose :398-410, globalPlanManipulator :277-297,
//                       angleDiff controllers/mpc_wholebody_qref.py:92-117
//
// One thread per episode: the work is a few hundred flops and a scan over <= a few hundred reference rows, the
// point of doing it on the device is that the states, references and flags of all episodes never leave HBM
// between two solves.  The phase bodies are plain functions of (arrays, b) so that tests/emu can run them on the CPU.
#pragma once
#include <stdint.h>
#include "../../include/mmpc.h"
#include "mmpc_model.cuh"

namespace mmpc {

// The reference's IK is a 3-variable NLP handed to IPOPT (:121-123):
//     min (x(q) - xt)^2 + (z(q) - zt)^2   s.t.  q1 in [-pi/2, pi/2], q2 in [-3pi/4, 0], q3 in [0, 3pi/2]
// Two equations in three unknowns: the minimisers form a curve and which point IPOPT stops at depends on its path,
// so the answer is pinned only up to "a feasible q with zero residual"; this is a projected Levenberg-Marquardt
// iteration from the same start (the host class and oracle/ik.py run the same recurrence).
// Returns 0 when the residual is below 1e-5 (squared 1e-10), 1 otherwise (the reference raises ValueError).
__device__ inline int ik_solve(const double* q0, double xt, double zt, double* q_out) {
  const double PI = 3.14159265358979323846;
  const double lo[3] = {-PI / 2, -PI * 3 / 4, 0.0}, hi[3] = {PI / 2, 0.0, PI * 3 / 2};
  double q[3];
  for (int i = 0; i < 3; ++i) q[i] = fmin(fmax(q0[i], lo[i]), hi[i]);
  double lam = 1e-3;
  FK f; fk_eval(0.0, q[0], q[1], q[2], f);
  double r0 = (f.vr[0] + f.vr[1]) + f.vr[2] - xt, r1 = (f.vh[0] + f.vh[1]) + f.vh[2] - zt;
  for (int it = 0; it < 200; ++it) {
    // d(x,z)/dtheta_s = (h_s, -r_s); dtheta/dq = [[1,0,0],[1,-1,0],[1,-1,-1]]
    double J0[3], J1[3];
    J0[0] = (f.vh[0] + f.vh[1]) + f.vh[2]; J0[1] = -(f.vh[1] + f.vh[2]); J0[2] = -f.vh[2];
    J1[0] = -((f.vr[0] + f.vr[1]) + f.vr[2]); J1[1] = f.vr[1] + f.vr[2]; J1[2] = f.vr[2];
    double g[3]; bool fr[3]; double gn2 = 0;
    for (int i = 0; i < 3; ++i) {
      g[i] = J0[i] * r0 + J1[i] * r1;
      fr[i] = !((q[i] <= lo[i] && g[i] > 0) || (q[i] >= hi[i] && g[i] < 0));
      if (fr[i]) gn2 += g[i] * g[i];
    }
    if (sqrt(gn2) < 1e-14) break;
    // (J^T J + lam I) step = -g on the free coordinates (fixed ones: identity row), LDL^T
    double H[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        H[i][j] = (fr[i] && fr[j]) ? J0[i] * J0[j] + J1[i] * J1[j] + (i == j ? lam : 0.0) : (i == j ? 1.0 : 0.0);
    double b[3];
    for (int i = 0; i < 3; ++i) b[i] = fr[i] ? -g[i] : 0.0;
    const double d0 = H[0][0], l10 = H[1][0] / d0, l20 = H[2][0] / d0;
    const double d1 = H[1][1] - l10 * H[1][0], l21 = (H[2][1] - l20 * H[1][0]) / d1;
    const double d2 = H[2][2] - l20 * H[2][0] - l21 * l21 * d1;
    const double y0 = b[0], y1 = b[1] - l10 * y0, y2 = b[2] - l20 * y0 - l21 * y1;
    double st[3];
    st[2] = y2 / d2; st[1] = y1 / d1 - l21 * st[2]; st[0] = y0 / d0 - l10 * st[1] - l20 * st[2];
    double qn[3];
    for (int i = 0; i < 3; ++i) qn[i] = fmin(fmax(q[i] + st[i], lo[i]), hi[i]);
    FK fn; fk_eval(0.0, qn[0], qn[1], qn[2], fn);
    const double n0 = (fn.vr[0] + fn.vr[1]) + fn.vr[2] - xt, n1 = (fn.vh[0] + fn.vh[1]) + fn.vh[2] - zt;
    if (n0 * n0 + n1 * n1 < r0 * r0 + r1 * r1) {
      for (int i = 0; i < 3; ++i) q[i] = qn[i];
      f = fn; r0 = n0; r1 = n1; lam = fmax(lam / 3, 1e-12);
    } else {
      lam *= 4;
      if (lam > 1e8) break;
    }
  }
  for (int i = 0; i < 3; ++i) q_out[i] = q[i];
  return (r0 * r0 + r1 * r1 > 1e-10) ? 1 : 0;
}

struct EpisodeArgs {
  int B, N, M, n_manip;
  MmpcEpisodeIO io;
};

// calcLocalRefTraj (:353-396): nearest row (first minimum of the Euclidean distance over the masked states), rows
// [i*, i* + N] with the last row repeated; the reference's u_ref is identically zero (:266, :296)
__device__ inline void episode_window(const EpisodeArgs& A, int b, const double* x, int idx_mask) {
  const double* g = A.io.traj + (size_t)b * A.M * NX;
  const int len = A.io.traj_len[b];
  double best = 1e5; int ib = 0;  // min_distance = 1e5 (:365)
  for (int j = 0; j < len; ++j) {
    double d2 = 0;
    for (int i = 0; i < NX; ++i) if (idx_mask >> i & 1) { double e = x[i] - g[(size_t)j * NX + i]; d2 += e * e; }
    const double d = sqrt(d2);
    if (d < best) { best = d; ib = j; }
  }
  for (int k = 0; k <= A.N; ++k) {
    const int r = ib + k < len - 1 ? ib + k : len - 1;
    for (int i = 0; i < NX; ++i) A.io.x_ref[((size_t)b * (A.N + 1) + k) * NX + i] = g[(size_t)r * NX + i];
  }
}

// calcLocalRefPose (:398-410): the last reference row tiled, psi unwrapped towards the current heading
__device__ inline void episode_pose(const EpisodeArgs& A, int b, const double* x) {
  const double* last = A.io.traj + ((size_t)b * A.M + (A.io.traj_len[b] - 1)) * NX;
  const double psi = x[2] + angle_diff(last[2], x[2]);
  for (int k = 0; k <= A.N; ++k)
    for (int i = 0; i < NX; ++i) A.io.x_ref[((size_t)b * (A.N + 1) + k) * NX + i] = (i == 2) ? psi : last[i];
}

// stateMachineUpdate (:146-228) of episode b.  Task flags: MMPC_TASK_*.
__device__ inline void episode_update(const EpisodeArgs& A, int b) {
  const MmpcEpisodeIO& io = A.io;
  int flag = io.task[b];
  double x[NX];
  for (int i = 0; i < NX; ++i) x[i] = io.x[(size_t)b * NX + i];
  for (int k = 0; k < A.N; ++k)
    for (int j = 0; j < NU; ++j) io.u_ref[((size_t)b * A.N + k) * NU + j] = 0.0;
  if (flag >= MMPC_TASK_FINISHED) { io.active[b] = 0; return; }
  double* traj = io.traj + (size_t)b * A.M * NX;
  const double* gp = io.pose_target + (size_t)b * 4;
  if (flag == MMPC_TASK_MOVE || flag == MMPC_TASK_APPROACH) {
    const double* last = traj + (size_t)(io.traj_len[b] - 1) * NX;
    if (fabs(x[0] - last[0]) <= 2 && fabs(x[1] - last[1]) <= 2 && flag == MMPC_TASK_MOVE) {  // :153-166
      flag = MMPC_TASK_APPROACH;
      io.flags[b] |= 1;  // opti.subject_to(X[N,:2] == X_ref[N,:2]) :167 -- stays for the rest of the episode
    }
    const double ex = x[0] - last[0], ey = x[1] - last[1];
    if (sqrt(ex * ex + ey * ey) <= 0.2) flag = MMPC_TASK_ROTATE;  // :170-175 (weights: see wset below)
    else if (flag == MMPC_TASK_MOVE) episode_window(A, b, x, 0x3);  // calcLocalRefTraj([0,1]) :188
    else episode_pose(A, b, x);
  }
  if (flag == MMPC_TASK_ROTATE) {  // :192-197
    const double* last = traj + (size_t)(io.traj_len[b] - 1) * NX;
    const double PI = 3.14159265358979323846;
    const double ex = x[0] - last[0], ey = x[1] - last[1];
    if (fabs(angle_diff(x[2], last[2])) <= 0.5 * PI / 180 && sqrt(ex * ex + ey * ey) <= 0.01) flag = MMPC_TASK_MOVE_FINISH;
    else episode_pose(A, b, x);
  }
  if (flag == MMPC_TASK_MOVE_FINISH) {  // :204-216, globalPlanManipulator :277-297
    flag = MMPC_TASK_MANIPULATE;
    const double dx = gp[0] - x[0], dy = gp[1] - x[1];
    const double lt[3] = {sqrt(dx * dx + dy * dy) + 0.007, 0.0, gp[2] - (0.606 + 0.333)};
    if (io.local_pose_target) for (int i = 0; i < 3; ++i) io.local_pose_target[(size_t)b * 3 + i] = lt[i];
    double qt[3];
    const int rc = ik_solve(x + 6, lt[0], lt[2], qt);
    if (io.ik_status) io.ik_status[b] = rc;
    if (rc != 0) flag = MMPC_TASK_IK_FAILED;  // the reference raises ValueError :124-125
    else {
      // np.linspace(current_state, x_target, n + 1) with x_target[:6] = current_state[:6]: some steps are zero, so
      // NumPy forms y = (i / div) * delta + start for every column, and the last row is the stop value itself
      const int n = A.n_manip;
      for (int i = 0; i <= n; ++i)
        for (int c = 0; c < NX; ++c) {
          const double stop = c < 6 ? x[c] : qt[c - 6];
          traj[(size_t)i * NX + c] = (i == n) ? stop : ((double)i / (double)n) * (stop - x[c]) + x[c];
        }
      io.traj_len[b] = n + 1;
    }
  }
  if (flag == MMPC_TASK_MANIPULATE) {  // :219-226
    FK f; fk_eval(x[2], x[6], x[7], x[8], f);
    Point p; point_eval(x[0], x[1], f, BODY[5], p);  // world position of the end point
    const double e0 = p.P[0] - gp[0], e1 = p.P[1] - gp[1], e2 = p.P[2] - gp[2];
    if (sqrt(e0 * e0 + e1 * e1 + e2 * e2) <= 0.01) flag = MMPC_TASK_FINISHED;
    else episode_window(A, b, x, 0x1c0);  // calcLocalRefTraj([6,7,8]) :226
  }
  io.task[b] = flag;
  io.active[b] = flag < MMPC_TASK_FINISHED;
  // setWeight calls of the machine: rotate :176-178, manipulate :212-215; they persist until the next call
  io.wset[b] = flag == MMPC_TASK_MANIPULATE ? 2 : flag == MMPC_TASK_ROTATE ? 1 : 0;
}

#if !defined(MMPC_EMULATE) && !defined(MMPC_EMULATE_LANE)
__global__ void __launch_bounds__(128) ik_kernel(int B, const double* q_guess, const double* target, double* q_out, int32_t* status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int rc = ik_solve(q_guess + (size_t)b * 3, target[(size_t)b * 3 + 0], target[(size_t)b * 3 + 2], q_out + (size_t)b * 3);
  if (status) status[b] = rc;
}
__global__ void __launch_bounds__(128) episode_update_kernel(const __grid_constant__ EpisodeArgs A) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < A.B) episode_update(A, b);
}
#endif

}  // namespace mmpc
