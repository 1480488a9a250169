// emu_lane_runtime.h -- TEST INFRASTRUCTURE: lets the lane-per-instance solver source
// (csrc/mmpc_lane.cuh) compile with g++ and run one lane at a time on a CPU-only box.  A lane
// never talks to its neighbours, so warp votes are the identity and the barrier is a no-op.
// Never linked into the product library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <ucontext.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __noinline__
#define MMPC_LSTR 1

namespace mmpc {
constexpr unsigned FULL = 0xffffffffu;
inline int lane_id() { return 0; }
inline void sync_warp() {}
inline bool warp_any(bool p) { return p; }
inline bool warp_all(bool p) { return p; }
inline double shfl_xor(double v, int) { return v; }
inline int shfl_xor(int v, int) { return v; }
inline unsigned lane_next_instance(unsigned* counter) { return (*counter)++; }
inline double ldg(const double* p) { return *p; }
inline int ldg(const int* p) { return *p; }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline int min(int a, int b) { return a < b ? a : b; }

// ---- 16-lane team emulator (ucontext coroutines): used by the team phase of the staged solver ----
struct EmuTeam {
  ucontext_t ctx[16], main_ctx;
  int cur, done;
  double slot[16];
};
extern EmuTeam* g_team;
inline void emu_team_barrier() {
  EmuTeam* w = g_team;
  int me = w->cur, nx = (me + 1) & 15;
  w->cur = nx;
  swapcontext(&w->ctx[me], &w->ctx[nx]);
  w->cur = me;
}
inline double shfl16(double v, int src) {
  g_team->slot[g_team->cur] = v; emu_team_barrier();
  double r = g_team->slot[src & 15]; emu_team_barrier(); return r;
}
inline double shfl16_xor(double v, int m) { return shfl16(v, g_team->cur ^ m); }
inline void team_sync() { emu_team_barrier(); }
inline void async_copy8(double* dst, const double* src) { dst[0] = src[0]; }
inline void async_copy16(double* dst, const double* src) { dst[0] = src[0]; dst[1] = src[1]; }
inline void async_commit() {}
template <int PENDING> inline void async_wait() {}
inline void prefetch_l2(const void*) {}
inline void compiler_fence() {}
inline void sincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
}  // namespace mmpc
