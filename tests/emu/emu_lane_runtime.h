// emu_lane_runtime.h -- TEST INFRASTRUCTURE: lets the lane-per-instance solver source
// (csrc/mmpc_lane.cuh) compile with g++ and run one lane at a time on a CPU-only box.  A lane
// never talks to its neighbours, so warp votes are the identity and the barrier is a no-op.
// Never linked into the product library.
#pragma once
#include <math.h>
#include <stdint.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define MMPC_LSTR 1

namespace mmpc {
constexpr unsigned FULL = 0xffffffffu;
inline int lane_id() { return 0; }
inline void sync_warp() {}
inline bool warp_any(bool p) { return p; }
inline double shfl_xor(double v, int) { return v; }
inline int shfl_xor(int v, int) { return v; }
inline unsigned lane_next_instance(unsigned* counter) { return (*counter)++; }
inline double ldg(const double* p) { return *p; }
inline int ldg(const int* p) { return *p; }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline void sincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
}  // namespace mmpc
