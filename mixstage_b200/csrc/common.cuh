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

// ---- bf16 tensor-core operand planes (hi plane; lo = bf16(v - hi) one plane stride later in split-bf16 mode)
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// hi plane always; lo plane (residual of the bf16 rounding) when fmt == MS_BF16X2.  i = element index (multiple of 4)
__device__ __forceinline__ void store_planes4(__nv_bfloat16* __restrict__ pl, int fmt, int64_t ps, int64_t i, float4 o) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(pl + i) = u;
  if (fmt == MS_BF16X2) {
    uint2 v;
    v.x = pack_bf162(o.x - __bfloat162float(h0.x), o.y - __bfloat162float(h0.y));
    v.y = pack_bf162(o.z - __bfloat162float(h1.x), o.w - __bfloat162float(h1.y));
    *reinterpret_cast<uint2*>(pl + ps + i) = v;
  }
}
__device__ __forceinline__ void store_planes1(__nv_bfloat16* __restrict__ pl, int fmt, int64_t ps, int64_t i, float o) {
  __nv_bfloat16 h = __float2bfloat16_rn(o);
  pl[i] = h;
  if (fmt == MS_BF16X2) pl[ps + i] = __float2bfloat16_rn(o - __bfloat162float(h));
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
