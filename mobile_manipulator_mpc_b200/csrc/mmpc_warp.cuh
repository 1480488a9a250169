// mmpc_warp.cuh -- the handful of warp primitives the solver kernel uses.
//
// Under nvcc these are the CUDA intrinsics.  Under -DMMPC_EMULATE_LANE (tests/emu/, g++ only) the
// same names are provided by a lane emulator (one lane at a time, 16 coroutines for the team phase) so
// the *kernel source itself* can be executed and debugged on a CPU-only box; the emulator is test
// infrastructure and is never part of the shipped library (the product path has no CPU fallback).
#pragma once

#if defined(MMPC_EMULATE_LANE)
#include "emu_lane_runtime.h"  // tests/emu/emu_lane_runtime.h (one lane at a time, votes are the identity)
#else
#include <cuda_runtime.h>
#include <cuda_pipeline.h>

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
#ifdef MMPC_RESIDENT
__device__ __forceinline__ double ldg(const double* p) { return *p; }   // may point into shared memory: generic load
__device__ __forceinline__ int ldg(const int* p) { return *p; }
#else
__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ldg(const int* p) { return __ldg(p); }
#endif
// teams: the two 16-lane halves of a warp run in lock step (full-warp mask, width-16 shuffles): half-warp
// masks make the hardware issue every shuffle once per team (measured: 17.5 of 32 lanes active)
// (the halves of a double go through the non-volatile conversion intrinsics: CUDA's own __shfl_sync(double) splits and joins
// them with `asm volatile mov.b64`, which ptxas keeps as real moves -- 12 % of the team kernel's instructions)
#ifdef MMPC_SHFL_PLAIN
__device__ __forceinline__ double shfl16(double v, int src) { return __shfl_sync(FULL, v, src, 16); }
__device__ __forceinline__ double shfl16_xor(double v, int m) { return __shfl_xor_sync(FULL, v, m, 16); }
#else
__device__ __forceinline__ double shfl16(double v, int src) {
  const int lo = __shfl_sync(FULL, __double2loint(v), src, 16), hi = __shfl_sync(FULL, __double2hiint(v), src, 16);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl16_xor(double v, int m) {
  const int lo = __shfl_xor_sync(FULL, __double2loint(v), m, 16), hi = __shfl_xor_sync(FULL, __double2hiint(v), m, 16);
  return __hiloint2double(hi, lo);
}
#endif
__device__ __forceinline__ void team_sync() { __syncwarp(); }
__device__ __forceinline__ bool warp_all(bool p) { return __all_sync(FULL, p); }
#ifdef MMPC_RESIDENT
// resident build (mmpc_resident.cu): the workspace the phase bodies read IS shared memory, so the staging rings are filled
// by plain copies (cp.async needs a global source) and there is nothing to wait for
__device__ __forceinline__ void async_copy8(double* smem_dst, const double* src) { smem_dst[0] = src[0]; }
__device__ __forceinline__ void async_copy16(double* smem_dst, const double* src) { smem_dst[0] = src[0]; smem_dst[1] = src[1]; }
__device__ __forceinline__ void async_commit() {}
template <int PENDING> __device__ __forceinline__ void async_wait() {}
__device__ __forceinline__ void prefetch_l2(const void*) {}
#else
// cp.async: global -> shared without passing through registers (LDGSTS), grouped per commit
__device__ __forceinline__ void async_copy8(double* smem_dst, const double* src) { __pipeline_memcpy_async(smem_dst, src, 8); }
__device__ __forceinline__ void async_copy16(double* smem_dst, const double* src) { __pipeline_memcpy_async(smem_dst, src, 16); }
__device__ __forceinline__ void async_commit() { __pipeline_commit(); }
template <int PENDING> __device__ __forceinline__ void async_wait() { __pipeline_wait_prior(PENDING); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
// compiler-only fence: memory operations are not moved across it (no instruction is emitted)
__device__ __forceinline__ void compiler_fence() { asm volatile("" ::: "memory"); }
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
// reductions over the NL lanes that share one instance in the control kernels (NL = 32: the warp, fixed
// butterfly order; NL = 1: nothing to do)
template <int NL> __device__ __forceinline__ double lanes_sum(double v) { return NL == 1 ? v : warp_sum(v); }
template <int NL> __device__ __forceinline__ double lanes_min(double v) { return NL == 1 ? v : warp_min(v); }
template <int NL> __device__ __forceinline__ double lanes_max(double v) { return NL == 1 ? v : warp_max(v); }
template <int NL> __device__ __forceinline__ void lanes_sync() { if (NL > 1) sync_warp(); }
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
}  // namespace mmpc
