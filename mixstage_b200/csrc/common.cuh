// Shared device/host helpers for the mixstage_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/mixstage_b200.h"

#define MS_LAUNCH_CHECK()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

#define MS_CUDA(x)                                \
  do {                                            \
    cudaError_t e__ = (x);                        \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

static inline cudaStream_t ms_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// caller-owned parameter tensors are fp32 or fp64 (trainer.py:138 `.double()`)
__device__ __forceinline__ float ms_ldp(const void* p, int dt, int64_t i) {
  return dt == MS_F64 ? (float)reinterpret_cast<const double*>(p)[i] : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ double ms_ldp_d(const void* p, int dt, int64_t i) {
  return dt == MS_F64 ? reinterpret_cast<const double*>(p)[i] : (double)reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void ms_stp(void* p, int dt, int64_t i, double v) {
  if (dt == MS_F64) reinterpret_cast<double*>(p)[i] = v;
  else reinterpret_cast<float*>(p)[i] = (float)v;
}

__device__ __forceinline__ float ms_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double ms_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ms_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline int64_t ms_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
__device__ __forceinline__ int64_t ms_cdiv_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
