// Fused global-norm gradient clipping + Adam over FLAT parameter / gradient / moment buffers.
//
// Replaces, for one sub-network (generator or discriminator), the optimiser half of the reference's train step
// (src/model/trainer.py:1138-1146): torch.nn.utils.clip_grad_norm_(params, 1) followed by torch.optim.Adam.step()
// (lr 1e-4, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad: trainer.py:262-287).  The reference walks
// ~390 tensors with foreach kernels; here the parameters of a sub-network live in one contiguous buffer (their
// nn.Parameter objects are views into it), so the whole update is two HBM-bound passes:
//   1. sum of squares of the gradient buffer  (read n elements)
//   2. Adam update with the clip coefficient  (read p, g, m, v; write p, m, v  -> 7 * sizeof(T) bytes per element)
// The step counter lives on the device so the pair can be replayed inside a CUDA graph.
#include "common.cuh"

namespace {

// Sum of squares in a FIXED order (block partials, summed by the last block to finish): data-parallel replicas compute the
// clip coefficient from bit-identical gradients and must get bit-identical coefficients -- floating-point atomics in
// arrival order would let the parameters of the replicas drift apart by an ulp per step.
constexpr int SQNORM_MAX_BLOCKS = 2048;
__device__ double g_sqnorm_part[SQNORM_MAX_BLOCKS];
__device__ unsigned int g_sqnorm_ticket = 0;

__device__ __forceinline__ void sq_load4(const double* a, double (&x)[4]) {
  const double2 u = reinterpret_cast<const double2*>(a)[0], w = reinterpret_cast<const double2*>(a)[1];
  x[0] = u.x; x[1] = u.y; x[2] = w.x; x[3] = w.y;
}
__device__ __forceinline__ void sq_load4(const float* a, double (&x)[4]) {
  const float4 u = reinterpret_cast<const float4*>(a)[0];
  x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w;
}

template <typename T>
__global__ void __launch_bounds__(256) sqnorm_kernel(const T* __restrict__ g, int64_t n, double* __restrict__ acc,
                                                     long long* __restrict__ step) {
  __shared__ double part[8];
  __shared__ bool last;
  // four elements per thread and iteration (the order of the sum is still a function of the launch shape alone)
  double a = 0.0;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (((uintptr_t)g & 31) == 0) ? (n >> 2) : 0;
  for (int64_t i = tid; i < n4; i += nthr) {
    double x[4];
    sq_load4(g + 4 * i, x);
    a += (x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3]);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nthr) {
    double v = (double)g[i];
    a += v * v;
  }
  a = ms_warp_sum_d(a);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); i++) t += part[i];
    g_sqnorm_part[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&g_sqnorm_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    // one warp, fixed tree
    if (threadIdx.x < 32) {
      double t = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += __ldcg(&g_sqnorm_part[i]);
      t = ms_warp_sum_d(t);
      if (threadIdx.x == 0) {
        acc[0] = t;
        g_sqnorm_ticket = 0;
        if (step) step[0] += 1;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) clip_adam_kernel(T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m,
                                                        T* __restrict__ v, int64_t n, const double* __restrict__ sqnorm,
                                                        const long long* __restrict__ step, double lr, double b1, double b2,
                                                        double eps, double max_norm, const double* __restrict__ lr_dev) {
  if (lr_dev) lr = lr_dev[0];
  const double total = sqrt(sqnorm[0]);
  double coef = max_norm > 0.0 ? max_norm / (total + 1e-6) : 1.0;      // clip_grad_norm_: coef clamped to 1
  if (coef > 1.0) coef = 1.0;
  const double t = (double)step[0];
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const double step_size = lr / bc1, inv_bc2_sqrt = 1.0 / sqrt(bc2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double gi = (double)g[i] * coef;
    double mi = b1 * (double)m[i] + (1.0 - b1) * gi;
    double vi = b2 * (double)v[i] + (1.0 - b2) * gi * gi;
    double denom = sqrt(vi) * inv_bc2_sqrt + eps;
    m[i] = (T)mi;
    v[i] = (T)vi;
    p[i] = (T)((double)p[i] - step_size * mi / denom);
  }
}

// Same update, four elements per thread and iteration (16/32-byte loads), moments optionally stored in fp32 while the
// parameters and gradients stay fp64: 40 instead of 56 bytes per element.  The arithmetic is the scalar kernel's (fp64).
template <typename T> struct Vec4;
template <> struct Vec4<double> {
  static __device__ __forceinline__ void load(const double* a, double (&x)[4]) {
    const double2 u = reinterpret_cast<const double2*>(a)[0], w = reinterpret_cast<const double2*>(a)[1];
    x[0] = u.x; x[1] = u.y; x[2] = w.x; x[3] = w.y;
  }
  static __device__ __forceinline__ void store(double* a, const double (&x)[4]) {
    reinterpret_cast<double2*>(a)[0] = make_double2(x[0], x[1]);
    reinterpret_cast<double2*>(a)[1] = make_double2(x[2], x[3]);
  }
};
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* a, double (&x)[4]) {
    const float4 u = reinterpret_cast<const float4*>(a)[0];
    x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w;
  }
  static __device__ __forceinline__ void store(float* a, const double (&x)[4]) {
    reinterpret_cast<float4*>(a)[0] = make_float4((float)x[0], (float)x[1], (float)x[2], (float)x[3]);
  }
};

template <typename T, typename S>
__global__ void __launch_bounds__(256) clip_adam_vec_kernel(T* __restrict__ p, const T* __restrict__ g, S* __restrict__ m,
                                                            S* __restrict__ v, int64_t n, const double* __restrict__ sqnorm,
                                                            const long long* __restrict__ step, double lr, double b1, double b2,
                                                            double eps, double max_norm, const double* __restrict__ lr_dev) {
  if (lr_dev) lr = lr_dev[0];
  const double total = sqrt(sqnorm[0]);
  double coef = max_norm > 0.0 ? max_norm / (total + 1e-6) : 1.0;
  if (coef > 1.0) coef = 1.0;
  const double t = (double)step[0];
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const double step_size = lr / bc1, inv_bc2_sqrt = 1.0 / sqrt(bc2);
  const int64_t n4 = n >> 2;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += nthr) {
    double pv[4], gv[4], mv[4], vv[4];
    Vec4<T>::load(p + 4 * i, pv);
    Vec4<T>::load(g + 4 * i, gv);
    Vec4<S>::load(m + 4 * i, mv);
    Vec4<S>::load(v + 4 * i, vv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double gi = gv[k] * coef;
      mv[k] = b1 * mv[k] + (1.0 - b1) * gi;
      vv[k] = b2 * vv[k] + (1.0 - b2) * gi * gi;
      pv[k] = (double)(T)pv[k] - step_size * mv[k] / (sqrt(vv[k]) * inv_bc2_sqrt + eps);
    }
    Vec4<S>::store(m + 4 * i, mv);
    Vec4<S>::store(v + 4 * i, vv);
    Vec4<T>::store(p + 4 * i, pv);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nthr) {
    const double gi = (double)g[i] * coef;
    const double mi = b1 * (double)m[i] + (1.0 - b1) * gi;
    const double vi = b2 * (double)v[i] + (1.0 - b2) * gi * gi;
    m[i] = (S)mi;
    v[i] = (S)vi;
    p[i] = (T)((double)p[i] - step_size * mi / (sqrt(vi) * inv_bc2_sqrt + eps));
  }
}

inline int opt_blocks(int64_t n) {
  int64_t b = ms_cdiv(n, 256 * 4);
  int64_t cap = (int64_t)ms_num_sms() * 8;
  if (cap > SQNORM_MAX_BLOCKS) cap = SQNORM_MAX_BLOCKS;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int ms_grad_sqnorm(const void* g, int dt, int64_t n, double* acc, int64_t* step, void* stream) {
  if (!g || !acc || n < 1 || (dt != MS_F32 && dt != MS_F64)) return MS_EINVAL;
  if (dt == MS_F64)
    sqnorm_kernel<double><<<opt_blocks(n), 256, 0, ms_stream(stream)>>>((const double*)g, n, acc, (long long*)step);
  else
    sqnorm_kernel<float><<<opt_blocks(n), 256, 0, ms_stream(stream)>>>((const float*)g, n, acc, (long long*)step);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_clip_adam(void* p, const void* g, void* m, void* v, int dt, int64_t n, const double* sqnorm,
                            const int64_t* step, double lr, double beta1, double beta2, double eps, double max_norm,
                            const double* lr_dev, void* stream) {
  if (!p || !g || !m || !v || !sqnorm || !step || n < 1 || (dt != MS_F32 && dt != MS_F64)) return MS_EINVAL;
  if (dt == MS_F64)
    clip_adam_kernel<double><<<opt_blocks(n), 256, 0, ms_stream(stream)>>>((double*)p, (const double*)g, (double*)m, (double*)v, n,
                                                                           sqnorm, (const long long*)step, lr, beta1, beta2, eps, max_norm, lr_dev);
  else
    clip_adam_kernel<float><<<opt_blocks(n), 256, 0, ms_stream(stream)>>>((float*)p, (const float*)g, (float*)m, (float*)v, n, sqnorm,
                                                                          (const long long*)step, lr, beta1, beta2, eps, max_norm, lr_dev);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_clip_adam_mixed(void* p, const void* g, void* m, void* v, int dt, int state_dt, int64_t n, const double* sqnorm,
                                  const int64_t* step, double lr, double beta1, double beta2, double eps, double max_norm,
                                  const double* lr_dev, void* stream) {
  if (!p || !g || !m || !v || !sqnorm || !step || n < 1 || (dt != MS_F32 && dt != MS_F64)) return MS_EINVAL;
  if (state_dt != MS_F32 && state_dt != dt) return MS_EINVAL;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 31) return MS_EINVAL;
  const int blocks = opt_blocks(ms_cdiv(n, 4));
  cudaStream_t cs = ms_stream(stream);
  const long long* st = (const long long*)step;
  if (dt == MS_F64 && state_dt == MS_F64)
    clip_adam_vec_kernel<double, double><<<blocks, 256, 0, cs>>>((double*)p, (const double*)g, (double*)m, (double*)v, n, sqnorm, st,
                                                                lr, beta1, beta2, eps, max_norm, lr_dev);
  else if (dt == MS_F64)
    clip_adam_vec_kernel<double, float><<<blocks, 256, 0, cs>>>((double*)p, (const double*)g, (float*)m, (float*)v, n, sqnorm, st, lr,
                                                               beta1, beta2, eps, max_norm, lr_dev);
  else
    clip_adam_vec_kernel<float, float><<<blocks, 256, 0, cs>>>((float*)p, (const float*)g, (float*)m, (float*)v, n, sqnorm, st, lr,
                                                              beta1, beta2, eps, max_norm, lr_dev);
  MS_LAUNCH_CHECK();
  return 0;
}
