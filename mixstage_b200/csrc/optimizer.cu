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

template <typename T>
__global__ void __launch_bounds__(256) sqnorm_kernel(const T* __restrict__ g, int64_t n, double* __restrict__ acc,
                                                     long long* __restrict__ step) {
  __shared__ double part[8];
  __shared__ bool last;
  double a = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
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
