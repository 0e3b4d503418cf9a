// common.cuh — shared host/device plumbing for libb200det.so (error slot, launch checks, warp helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/b200det.h"

// thread-local error message (b200_last_error)
void b200_set_error(const char* fmt, ...);

#define B200_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      b200_set_error(__VA_ARGS__);         \
      return (code);                       \
    }                                      \
  } while (0)

#define B200_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      b200_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define B200_LAUNCH_CHECK()                                                                  \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      b200_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

static inline size_t b200_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// number of SMs of the current device (cached per thread)
int b200_sm_count();

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
#endif
