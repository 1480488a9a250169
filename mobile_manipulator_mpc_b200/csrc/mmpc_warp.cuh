// mmpc_warp.cuh -- the handful of warp primitives the solver kernel uses.
//
// Under nvcc these are the CUDA intrinsics.  Under -DMMPC_EMULATE (tests/emu/, g++ only) the
// same names are provided by a 32-coroutine lane emulator so the *kernel source itself* can be
// executed and debugged on a CPU-only box; the emulator is test infrastructure and is never
// part of the shipped library (the product path has no CPU fallback).
#pragma once

#if defined(MMPC_EMULATE_LANE)
#include "emu_lane_runtime.h"  // tests/emu/emu_lane_runtime.h (one lane at a time, votes are the identity)
#elif defined(MMPC_EMULATE)
#include "emu_runtime.h"  // tests/emu/emu_runtime.h
#else
#include <cuda_runtime.h>

namespace mmpc {
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ void sync_warp() { __syncwarp(); }
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ int shfl_xor(int v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ bool warp_any(bool p) { return __any_sync(FULL, p); }
__device__ __forceinline__ unsigned next_instance(unsigned* counter) {
  unsigned v = 0;
  if (lane_id() == 0) v = atomicAdd(counter, 1u);
  return __shfl_sync(FULL, v, 0);
}
__device__ __forceinline__ unsigned lane_next_instance(unsigned* counter) { return atomicAdd(counter, 1u); }
__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ldg(const int* p) { return __ldg(p); }
}  // namespace mmpc
#endif

namespace mmpc {
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor(v, m));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor(v, m));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
}  // namespace mmpc
