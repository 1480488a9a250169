// emu_runtime.h -- TEST INFRASTRUCTURE: a 32-lane warp emulator (ucontext coroutines, one host
// thread) so that the solver kernel source (csrc/mmpc_solver.cuh) can be executed on a CPU-only
// box.  sync_warp() is a round-robin yield; shuffles go through a slot array bracketed by two
// barriers.  Never linked into the product library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <ucontext.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __noinline__

namespace mmpc {
constexpr unsigned FULL = 0xffffffffu;
struct EmuWarp {
  ucontext_t ctx[32], main_ctx;
  int cur, done;
  double slot_d[32];
  long long slot_i[32];
};
extern EmuWarp* g_emu;
inline int lane_id() { return g_emu->cur; }
inline void emu_barrier() {
  EmuWarp* w = g_emu;
  int me = w->cur, nx = (me + 1) & 31;
  w->cur = nx;
  swapcontext(&w->ctx[me], &w->ctx[nx]);
  w->cur = me;
}
inline void sync_warp() { emu_barrier(); }
inline double shfl(double v, int src) {
  g_emu->slot_d[lane_id()] = v; emu_barrier();
  double r = g_emu->slot_d[src & 31]; emu_barrier(); return r;
}
inline int shfl(int v, int src) {
  g_emu->slot_i[lane_id()] = v; emu_barrier();
  int r = (int)g_emu->slot_i[src & 31]; emu_barrier(); return r;
}
inline double shfl_xor(double v, int m) { return shfl(v, lane_id() ^ m); }
inline int shfl_xor(int v, int m) { return shfl(v, lane_id() ^ m); }
inline bool warp_any(bool p) {
  g_emu->slot_i[lane_id()] = p; emu_barrier();
  bool r = false; for (int i = 0; i < 32; ++i) r = r || g_emu->slot_i[i];
  emu_barrier(); return r;
}
inline bool warp_all(bool p) { return !warp_any(!p); }
inline unsigned next_instance(unsigned* counter) {
  if (lane_id() == 0) { g_emu->slot_i[0] = *counter; *counter += 1; }
  emu_barrier(); unsigned v = (unsigned)g_emu->slot_i[0]; emu_barrier(); return v;
}
inline double ldg(const double* p) { return *p; }
inline int ldg(const int* p) { return *p; }
inline void compiler_fence() {}
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline void sincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
}  // namespace mmpc
