// HBM-bound kernels of the generator hot path: BatchNorm statistics / apply / backward
// fused with LeakyReLU and the UNet upsample+skip add, weight (un)packing, casts, bilinear
// time resize, style-embedding gather/scatter fused with the concat, softmax+CE+argmax,
// softmax-weighted cluster mixture, velocity, L1 reductions.
//
// All take channels-last fp32 activations ([rows, C], C contiguous); every warp reads
// contiguous 128-byte rows (float4 where C % 4 == 0), reductions accumulate in double and
// use one atomic per block and column.
#include "common.cuh"

namespace {

constexpr int EW_THREADS = 256;

inline int ew_blocks(int64_t work, int per_block = EW_THREADS) {
  int64_t b = ms_cdiv(work, per_block);
  int64_t cap = (int64_t)ms_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ casts / packing
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = (D)(float)s[i];
}
template <>
__global__ void cast_kernel<double, double>(const double* __restrict__ s, double* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
}
template <>
__global__ void cast_kernel<float, double>(const float* __restrict__ s, double* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (double)s[i];
}

// d[i] = (D)(s[i] * scale), vectorised where both pointers allow: the staging casts of the fp32 gradient exchange
// (train_step.py: fp64 flat gradients -> fp32 staging, scaled by 1/world; and back)
template <typename S, typename D>
__global__ void scale_cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n, double scale) {
  const int64_t n2 = n >> 1;
  const bool vec = ((reinterpret_cast<uintptr_t>(s) % (2 * sizeof(S))) | (reinterpret_cast<uintptr_t>(d) % (2 * sizeof(D)))) == 0;
  if (vec) {
    struct alignas(2 * sizeof(S)) S2 { S a, b; };
    struct alignas(2 * sizeof(D)) D2 { D a, b; };
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
      const S2 v = reinterpret_cast<const S2*>(s)[i];
      D2 o;
      o.a = (D)((double)v.a * scale);
      o.b = (D)((double)v.b * scale);
      reinterpret_cast<D2*>(d)[i] = o;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = (D)((double)s[n - 1] * scale);
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      d[i] = (D)((double)s[i] * scale);
  }
}

// w (Cout, Cin_g, kh, kw) -> wf[g][tap][c][n], wt[g][tap][n][c]
__global__ void pack_weight_kernel(const void* __restrict__ w, int pdt, int Cout, int Cin_g, int taps, int groups,
                                   float* __restrict__ wf, float* __restrict__ wt) {
  int Cout_g = Cout / groups;
  int64_t total = (int64_t)Cout * Cin_g * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps);
    int64_t t = i / taps;
    int c = (int)(t % Cin_g);
    int o = (int)(t / Cin_g);
    int g = o / Cout_g, n = o - g * Cout_g;
    float v = ms_ldp(w, pdt, i);
    if (wf) wf[(((int64_t)g * taps + tap) * Cin_g + c) * Cout_g + n] = v;
    if (wt) wt[(((int64_t)g * taps + tap) * Cout_g + n) * Cin_g + c] = v;
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwf, int Cout, int Cin_g, int taps, int groups,
                                    void* __restrict__ dw, int pdt, int accumulate) {
  int Cout_g = Cout / groups;
  int64_t total = (int64_t)Cout * Cin_g * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps);
    int64_t t = i / taps;
    int c = (int)(t % Cin_g);
    int o = (int)(t / Cin_g);
    int g = o / Cout_g, n = o - g * Cout_g;
    ms_stp(dw, pdt, i, (double)dwf[(((int64_t)g * taps + tap) * Cin_g + c) * Cout_g + n] + (accumulate ? ms_ldp_d(dw, pdt, i) : 0.0));
  }
}

__global__ void store_param_grad_kernel(const double* __restrict__ s, int n, void* __restrict__ d, int pdt, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ms_stp(d, pdt, i, s[i] + (accumulate ? ms_ldp_d(d, pdt, i) : 0.0));
}

// ------------------------------------------------------------------ column statistics
// block = 32 channels x 8 row lanes; grid = (C/32, row chunks)
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ x, int64_t rows, int C,
                                                        double* __restrict__ sum, double* __restrict__ sumsq) {
  __shared__ double s1[8][33], s2[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + tx;
  int64_t per = ms_cdiv_dev(rows, gridDim.y);
  int64_t r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  double a = 0.0, b = 0.0;
  if (c < C)
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float v = __ldg(x + r * C + c);
      a += v;
      b += (double)v * v;
    }
  s1[ty][tx] = a;
  s2[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; i++) { a += s1[i][tx]; b += s2[i][tx]; }
    atomicAdd(sum + c, a);
    if (sumsq) atomicAdd(sumsq + c, b);
  }
}

__device__ __forceinline__ void bn_finalize_one(int c, double sum, double sumsq, int64_t rows, const void* gamma, const void* beta,
                                                const void* cbias, void* rmean, void* rvar, int pdt, float momentum, float eps,
                                                float* scale, float* shift, float* mean_o, float* rstd_o) {
  double cb = cbias ? ms_ldp_d(cbias, pdt, c) : 0.0;
  double mean = sum / (double)rows;
  double var = sumsq / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  double unb = rows > 1 ? var * ((double)rows / (double)(rows - 1)) : var;
  double rm = ms_ldp_d(rmean, pdt, c), rv = ms_ldp_d(rvar, pdt, c);
  ms_stp(rmean, pdt, c, (1.0 - (double)momentum) * rm + (double)momentum * (mean + cb));
  ms_stp(rvar, pdt, c, (1.0 - (double)momentum) * rv + (double)momentum * unb);
  double rstd = 1.0 / sqrt(var + (double)eps);
  double g = ms_ldp_d(gamma, pdt, c), b = ms_ldp_d(beta, pdt, c);
  scale[c] = (float)(g * rstd);
  shift[c] = (float)(b - mean * g * rstd);
  mean_o[c] = (float)mean;
  rstd_o[c] = (float)rstd;
}

// column statistics and, in the block that finishes last, the training-mode finalize of every channel
// (one launch instead of statistics + finalize + counter increment)
__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(
    const float* __restrict__ x, int64_t rows, int C, double* __restrict__ sum, double* __restrict__ sumsq,
    unsigned int* __restrict__ ticket, const void* gamma, const void* beta, const void* cbias, void* rmean, void* rvar,
    long long* nbt, int pdt, float momentum, float eps, float* scale, float* shift, float* mean_o, float* rstd_o) {
  __shared__ double s1[8][33], s2[8][33];
  __shared__ unsigned int last;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + tx;
  int64_t per = ms_cdiv_dev(rows, gridDim.y);
  int64_t r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  double a = 0.0, b = 0.0;
  if (c < C)
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float v = __ldg(x + r * C + c);
      a += v;
      b += (double)v * v;
    }
  s1[ty][tx] = a;
  s2[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; i++) { a += s1[i][tx]; b += s2[i][tx]; }
    atomicAdd(sum + c, a);
    atomicAdd(sumsq + c, b);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x)
    bn_finalize_one(ch, __ldcg(sum + ch), __ldcg(sumsq + ch), rows, gamma, beta, cbias, rmean, rvar, pdt, momentum, eps, scale,
                    shift, mean_o, rstd_o);
  if (threadIdx.x == 0 && nbt) nbt[0] += 1;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, int64_t rows, int C,
                                   const void* gamma, const void* beta, const void* cbias, void* rmean, void* rvar, int pdt,
                                   int training, float momentum, float eps,
                                   float* scale, float* shift, float* mean_o, float* rstd_o) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  // cbias: bias of the producing convolution when the statistics were taken on the bias-free GEMM output
  double cb = cbias ? ms_ldp_d(cbias, pdt, c) : 0.0;
  if (training) {
    mean = sum[c] / (double)rows;
    var = sumsq[c] / (double)rows - mean * mean;
    if (var < 0.0) var = 0.0;
    double unb = rows > 1 ? var * ((double)rows / (double)(rows - 1)) : var;
    double rm = ms_ldp_d(rmean, pdt, c), rv = ms_ldp_d(rvar, pdt, c);
    ms_stp(rmean, pdt, c, (1.0 - (double)momentum) * rm + (double)momentum * (mean + cb));
    ms_stp(rvar, pdt, c, (1.0 - (double)momentum) * rv + (double)momentum * unb);
  } else {
    mean = ms_ldp_d(rmean, pdt, c) - cb;
    var = ms_ldp_d(rvar, pdt, c);
  }
  double rstd = 1.0 / sqrt(var + (double)eps);
  double g = ms_ldp_d(gamma, pdt, c), b = ms_ldp_d(beta, pdt, c);
  scale[c] = (float)(g * rstd);
  shift[c] = (float)(b - mean * g * rstd);
  mean_o[c] = (float)mean;
  rstd_o[c] = (float)rstd;
}

// ------------------------------------------------------------------ BN apply + LeakyReLU (+ upsample x2 + skip)
template <int VEC>
__global__ void bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                  float slope, int64_t rows_out, int C, float* __restrict__ y,
                                  const float* __restrict__ res, int up2, int L,
                                  __nv_bfloat16* __restrict__ planes, int pfmt, int64_t pstride) {
  int Cv = C / VEC;
  int64_t total = rows_out * Cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t ro = i / Cv;
    int cv = (int)(i - ro * Cv);
    int64_t ri = ro;
    if (up2) {
      int64_t b = ro / (2 * L);
      int l2 = (int)(ro - b * 2 * L);
      ri = b * L + (l2 >> 1);
    }
    if (VEC == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x + ri * C) + cv);
      float4 s = __ldg(reinterpret_cast<const float4*>(scale) + cv);
      float4 h = __ldg(reinterpret_cast<const float4*>(shift) + cv);
      float4 o;
      o.x = fmaf(v.x, s.x, h.x); o.y = fmaf(v.y, s.y, h.y); o.z = fmaf(v.z, s.z, h.z); o.w = fmaf(v.w, s.w, h.w);
      o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
      o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
      if (res) {
        float4 r = __ldg(reinterpret_cast<const float4*>(res + ro * C) + cv);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (y) reinterpret_cast<float4*>(y + ro * C)[cv] = o;
      if (planes) store_planes4(planes, pfmt, pstride, ro * C + 4 * cv, o);
    } else {
      float o = fmaf(__ldg(x + ri * C + cv), scale[cv], shift[cv]);
      o = o > 0.f ? o : o * slope;
      if (res) o += __ldg(res + ro * C + cv);
      if (y) y[ro * C + cv] = o;
      if (planes) store_planes1(planes, pfmt, pstride, ro * C + cv, o);
    }
  }
}

__device__ __forceinline__ float dy_at(const float* __restrict__ dy, int64_t r, int c, int C, int up2, int L) {
  if (!up2) return __ldg(dy + r * C + c);
  int64_t b = r / L;
  int l = (int)(r - b * L);
  int64_t ro = b * 2 * L + 2 * l;
  return __ldg(dy + ro * C + c) + __ldg(dy + (ro + 1) * C + c);
}

__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
    int64_t rows, int C, int up2, int L, double* __restrict__ dgamma, double* __restrict__ dbeta) {
  __shared__ double s1[8][33], s2[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + tx;
  int64_t per = ms_cdiv_dev(rows, gridDim.y);
  int64_t r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  double a = 0.0, b = 0.0;
  if (c < C) {
    float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float xv = __ldg(x + r * C + c);
      float z = fmaf(xv, sc, sh);
      float d = dy_at(dy, r, c, C, up2, L);
      float dz = z > 0.f ? d : d * slope;
      a += dz;
      b += (double)dz * (double)((xv - mu) * rs);
    }
  }
  s1[ty][tx] = a;
  s2[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; i++) { a += s1[i][tx]; b += s2[i][tx]; }
    atomicAdd(dbeta + c, a);
    atomicAdd(dgamma + c, b);
  }
}

// The same reduction for C = 64 / 128 / 256 without the upsample: a thread owns FOUR adjacent channels (16-byte loads of dy
// and x, a row's channels are contiguous across C/4 threads), two rows in flight per thread; 256 / (C/4) row lanes per CTA
// meet in shared memory, one fp64 atomic pair per channel and CTA.  (The one-channel-per-thread form above moves 4 bytes
// per load: 2.1 TB/s on the 64-channel first audio block, whose 2 x 134 MB at batch 128 end the backward pass.)
template <int G>          // float4 channel groups = C / 4
__global__ void __launch_bounds__(256) bn_act_bwd_reduce4_kernel(
    const float4* __restrict__ dy, const float4* __restrict__ x, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
    int64_t rows, double* __restrict__ dgamma, double* __restrict__ dbeta) {
  constexpr int LANES = 256 / G;
  __shared__ double s_a[LANES][G * 4 + 1], s_b[LANES][G * 4 + 1];
  const int cg = threadIdx.x % G, ty = threadIdx.x / G;
  const int64_t per = ms_cdiv_dev(rows, gridDim.x);
  const int64_t r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  const float4 sc = reinterpret_cast<const float4*>(scale)[cg], sh = reinterpret_cast<const float4*>(shift)[cg];
  const float4 mu = reinterpret_cast<const float4*>(mean)[cg], rs = reinterpret_cast<const float4*>(rstd)[cg];
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  auto add = [&](const float4& xv, const float4& d) {
    const float dz0 = fmaf(xv.x, sc.x, sh.x) > 0.f ? d.x : d.x * slope;
    const float dz1 = fmaf(xv.y, sc.y, sh.y) > 0.f ? d.y : d.y * slope;
    const float dz2 = fmaf(xv.z, sc.z, sh.z) > 0.f ? d.z : d.z * slope;
    const float dz3 = fmaf(xv.w, sc.w, sh.w) > 0.f ? d.w : d.w * slope;
    a[0] += dz0; a[1] += dz1; a[2] += dz2; a[3] += dz3;
    b[0] += (double)dz0 * (double)((xv.x - mu.x) * rs.x);
    b[1] += (double)dz1 * (double)((xv.y - mu.y) * rs.y);
    b[2] += (double)dz2 * (double)((xv.z - mu.z) * rs.z);
    b[3] += (double)dz3 * (double)((xv.w - mu.w) * rs.w);
  };
  int64_t r = r0 + ty;
  for (; r + LANES < r1; r += 2 * LANES) {
    const float4 x0 = __ldg(x + r * G + cg), x1 = __ldg(x + (r + LANES) * G + cg);
    const float4 d0 = __ldg(dy + r * G + cg), d1 = __ldg(dy + (r + LANES) * G + cg);
    add(x0, d0);
    add(x1, d1);
  }
  for (; r < r1; r += LANES) add(__ldg(x + r * G + cg), __ldg(dy + r * G + cg));
#pragma unroll
  for (int j = 0; j < 4; j++) { s_a[ty][cg * 4 + j] = a[j]; s_b[ty][cg * 4 + j] = b[j]; }
  __syncthreads();
  for (int c = threadIdx.x; c < G * 4; c += blockDim.x) {
    double ta = 0.0, tb = 0.0;
#pragma unroll 4
    for (int i = 0; i < LANES; i++) { ta += s_a[i][c]; tb += s_b[i][c]; }
    atomicAdd(dbeta + c, ta);
    atomicAdd(dgamma + c, tb);
  }
}

__global__ void bn_act_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
                                        int64_t rows, int C, int up2, int L, const double* __restrict__ dgamma,
                                        const double* __restrict__ dbeta, int training, float* __restrict__ dx,
                                        __nv_bfloat16* __restrict__ planes, int pfmt, int64_t pstride,
                                        void* __restrict__ ggamma, void* __restrict__ gbeta, int gdt) {
  int64_t total = rows * C;
  float inv = 1.f / (float)rows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C;
    int c = (int)(i - r * C);
    if (r == 0) {          // the thread of row 0 also accumulates the affine-parameter gradients of its channel
      if (ggamma) ms_stp(ggamma, gdt, c, ms_ldp_d(ggamma, gdt, c) + dgamma[c]);
      if (gbeta) ms_stp(gbeta, gdt, c, ms_ldp_d(gbeta, gdt, c) + dbeta[c]);
    }
    float xv = __ldg(x + i);
    float sc = scale[c];
    float z = fmaf(xv, sc, shift[c]);
    float d = dy_at(dy, r, c, C, up2, L);
    float dz = z > 0.f ? d : d * slope;
    float o;
    if (training) {
      float xh = (xv - mean[c]) * rstd[c];
      o = sc * (dz - (float)dbeta[c] * inv - xh * (float)dgamma[c] * inv);
    } else {
      o = sc * dz;
    }
    if (dx) dx[i] = o;
    if (planes) store_planes1(planes, pfmt, pstride, i, o);
  }
}

// 16-byte form of the apply pass (C % 4 == 0, no upsampling): a thread owns four channels of one row -- one index division per
// four elements, per-channel constants as float4, 16-byte loads and stores.  (The scalar form above ran at ~1.3 TB/s on the
// 134 MB tensors of audio_encoder.conv.0 at batch 128: 300 us per call.)
__global__ void bn_act_bwd_apply4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
                                         int64_t rows, int C4, const double* __restrict__ dgamma,
                                         const double* __restrict__ dbeta, int training, float4* __restrict__ dx,
                                         __nv_bfloat16* __restrict__ planes, int pfmt, int64_t pstride,
                                         void* __restrict__ ggamma, void* __restrict__ gbeta, int gdt) {
  const int64_t total = rows * C4;
  const float inv = 1.f / (float)rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (stride % C4 == 0) {
    // the thread's four channels never change along its grid-stride walk: constants in registers, a pure streaming loop
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c = 4 * (int)(i0 % C4);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, kb = mu, kg = mu;
    if (training) {
      mu = *reinterpret_cast<const float4*>(mean + c);
      rs = *reinterpret_cast<const float4*>(rstd + c);
      kb = make_float4((float)dbeta[c] * inv, (float)dbeta[c + 1] * inv, (float)dbeta[c + 2] * inv, (float)dbeta[c + 3] * inv);
      kg = make_float4((float)dgamma[c], (float)dgamma[c + 1], (float)dgamma[c + 2], (float)dgamma[c + 3]);
    }
    if (i0 < C4) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (ggamma) ms_stp(ggamma, gdt, c + j, ms_ldp_d(ggamma, gdt, c + j) + dgamma[c + j]);
        if (gbeta) ms_stp(gbeta, gdt, c + j, ms_ldp_d(gbeta, gdt, c + j) + dbeta[c + j]);
      }
    }
    for (int64_t i = i0; i < total; i += 2 * stride) {
      const int64_t i2 = i + stride;
      const bool two = i2 < total;
      const float4 xa = __ldg(x + i), da = __ldg(dy + i);
      float4 xb = xa, db = da;
      if (two) { xb = __ldg(x + i2); db = __ldg(dy + i2); }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (u == 1 && !two) break;
        const float4 xv = u ? xb : xa, d = u ? db : da;
        const float g0 = fmaf(xv.x, sc.x, sh.x) > 0.f ? d.x : d.x * slope, g1 = fmaf(xv.y, sc.y, sh.y) > 0.f ? d.y : d.y * slope;
        const float g2 = fmaf(xv.z, sc.z, sh.z) > 0.f ? d.z : d.z * slope, g3 = fmaf(xv.w, sc.w, sh.w) > 0.f ? d.w : d.w * slope;
        float4 o;
        if (training) {
          // (same expression order as the general form below: identical results)
          o.x = sc.x * (g0 - kb.x - ((xv.x - mu.x) * rs.x) * kg.x * inv);
          o.y = sc.y * (g1 - kb.y - ((xv.y - mu.y) * rs.y) * kg.y * inv);
          o.z = sc.z * (g2 - kb.z - ((xv.z - mu.z) * rs.z) * kg.z * inv);
          o.w = sc.w * (g3 - kb.w - ((xv.w - mu.w) * rs.w) * kg.w * inv);
        } else {
          o.x = sc.x * g0; o.y = sc.y * g1; o.z = sc.z * g2; o.w = sc.w * g3;
        }
        const int64_t io = u ? i2 : i;
        if (dx) dx[io] = o;
        if (planes) store_planes4(planes, pfmt, pstride, 4 * io, o);
      }
    }
    return;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C4;
    const int c = 4 * (int)(i - r * C4);
    if (r == 0) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (ggamma) ms_stp(ggamma, gdt, c + j, ms_ldp_d(ggamma, gdt, c + j) + dgamma[c + j]);
        if (gbeta) ms_stp(gbeta, gdt, c + j, ms_ldp_d(gbeta, gdt, c + j) + dbeta[c + j]);
      }
    }
    const float4 xv = __ldg(x + i), d = __ldg(dy + i);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    const float g0 = fmaf(xv.x, sc.x, sh.x) > 0.f ? d.x : d.x * slope, g1 = fmaf(xv.y, sc.y, sh.y) > 0.f ? d.y : d.y * slope;
    const float g2 = fmaf(xv.z, sc.z, sh.z) > 0.f ? d.z : d.z * slope, g3 = fmaf(xv.w, sc.w, sh.w) > 0.f ? d.w : d.w * slope;
    float4 o;
    if (training) {
      const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
      o.x = sc.x * (g0 - (float)dbeta[c] * inv - ((xv.x - mu.x) * rs.x) * (float)dgamma[c] * inv);
      o.y = sc.y * (g1 - (float)dbeta[c + 1] * inv - ((xv.y - mu.y) * rs.y) * (float)dgamma[c + 1] * inv);
      o.z = sc.z * (g2 - (float)dbeta[c + 2] * inv - ((xv.z - mu.z) * rs.z) * (float)dgamma[c + 2] * inv);
      o.w = sc.w * (g3 - (float)dbeta[c + 3] * inv - ((xv.w - mu.w) * rs.w) * (float)dgamma[c + 3] * inv);
    } else {
      o.x = sc.x * g0; o.y = sc.y * g1; o.z = sc.z * g2; o.w = sc.w * g3;
    }
    if (dx) dx[i] = o;
    if (planes) store_planes4(planes, pfmt, pstride, 4 * i, o);
  }
}

// x (rows, C) fp32 -> bf16 planes with row stride rs >= C (pad columns zero-filled)
__global__ void to_planes_kernel(const float* __restrict__ x, int64_t rows, int C, int rs, __nv_bfloat16* __restrict__ planes,
                                 int pfmt, int64_t pstride) {
  int64_t total = rows * rs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / rs;
    int c = (int)(i - r * rs);
    float v = c < C ? __ldg(x + r * C + c) : 0.f;
    store_planes1(planes, pfmt, pstride, i, v);
  }
}
__global__ void to_planes4_kernel(const float* __restrict__ x, int64_t n4, __nv_bfloat16* __restrict__ planes, int pfmt,
                                  int64_t pstride) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    store_planes4(planes, pfmt, pstride, 4 * i, __ldg(reinterpret_cast<const float4*>(x) + i));
}

// bf16 planes (row stride rs) -> x (rows, C) fp32: hi (+ lo)
__global__ void planes_to_f32_kernel(const __nv_bfloat16* __restrict__ planes, int pfmt, int64_t pstride, int64_t rows, int C,
                                     int rs, float* __restrict__ x) {
  int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C;
    int64_t j = r * rs + (i - r * C);
    float v = __bfloat162float(planes[j]);
    if (pfmt == MS_BF16X2) v += __bfloat162float(planes[pstride + j]);
    x[i] = v;
  }
}

__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float slope, int64_t n,
                                 float* __restrict__ dz, __nv_bfloat16* __restrict__ planes, int pfmt, int64_t pstride) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = __ldg(dy + i);
    float o = __ldg(y + i) > 0.f ? d : d * slope;
    if (dz) dz[i] = o;
    if (planes) store_planes1(planes, pfmt, pstride, i, o);
  }
}

// ------------------------------------------------------------------ bilinear (Hi,Wi) -> (T,1), align_corners=False
__device__ __forceinline__ void lin_src(int o, int in_size, int out_size, int& i0, int& i1, float& lam) {
  float scale = (float)in_size / (float)out_size;
  float src = ((float)o + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = src - (float)i0;
}

__global__ void bilinear_fwd_kernel(const float* __restrict__ x, int B, int Hi, int Wi, int C, int T, float* __restrict__ y) {
  int64_t total = (int64_t)B * T * C;
  int w0, w1;
  float lw;
  lin_src(0, Wi, 1, w0, w1, lw);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t2 = i / C;
    int t = (int)(t2 % T);
    int b = (int)(t2 / T);
    int h0, h1;
    float lh;
    lin_src(t, Hi, T, h0, h1, lh);
    const float* xb = x + (size_t)b * Hi * Wi * C;
    float v00 = __ldg(xb + ((size_t)h0 * Wi + w0) * C + c), v01 = __ldg(xb + ((size_t)h0 * Wi + w1) * C + c);
    float v10 = __ldg(xb + ((size_t)h1 * Wi + w0) * C + c), v11 = __ldg(xb + ((size_t)h1 * Wi + w1) * C + c);
    y[i] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
  }
}

// deterministic gather form of the adjoint: every input element sums its contributions
__global__ void bilinear_bwd_kernel(const float* __restrict__ dy, int B, int Hi, int Wi, int C, int T, float* __restrict__ dx) {
  int64_t total = (int64_t)B * Hi * Wi * C;
  int w0, w1;
  float lw;
  lin_src(0, Wi, 1, w0, w1, lw);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t2 = i / C;
    int w = (int)(t2 % Wi);
    t2 /= Wi;
    int h = (int)(t2 % Hi);
    int b = (int)(t2 / Hi);
    float ww = (w == w0 ? 1.f - lw : 0.f) + (w == w1 ? lw : 0.f);
    float acc = 0.f;
    if (ww != 0.f) {
      for (int t = 0; t < T; t++) {
        int h0, h1;
        float lh;
        lin_src(t, Hi, T, h0, h1, lh);
        float wh = (h == h0 ? 1.f - lh : 0.f) + (h == h1 ? lh : 0.f);
        if (wh != 0.f) acc += wh * __ldg(dy + ((size_t)b * T + t) * C + c);
      }
    }
    dx[i] = acc * ww;
  }
}

// ------------------------------------------------------------------ style embedding + concat
__global__ void style_concat_fwd_kernel(const float* __restrict__ x, int64_t rows, int C, const int64_t* __restrict__ idx,
                                        const float* __restrict__ soft, int rep, const void* __restrict__ emb, int pdt,
                                        int S, int sd, float* __restrict__ out) {
  int Co = C + sd;
  int64_t total = rows * Co;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / Co;
    int c = (int)(i - r * Co);
    float v;
    if (c < C) {
      v = __ldg(x + r * C + c);
    } else {
      int j = c - C;
      int64_t q = r / rep;
      if (idx) {
        int64_t s = idx[q];
        v = (s >= 0 && s < S) ? ms_ldp(emb, pdt, s * sd + j) : 0.f;
      } else {
        v = 0.f;
        for (int s = 0; s < S; s++) v = fmaf(__ldg(soft + q * S + s), ms_ldp(emb, pdt, (int64_t)s * sd + j), v);
      }
    }
    out[i] = v;
  }
}

// Warp-per-row form (C % 128 == 0, sd <= 32): a lane moves the row's content channels as float4 loads / 8-byte stores (output rows
// are C + sd floats: 8-byte aligned only), lanes 0..sd-1 fetch the style row (index gather or soft S x sd product) and the warp
// also emits the row as bf16 operand planes (row stride rs >= C + sd, padding zero-filled) for the tensor-core consumers -- the
// separate fp32 -> planes pass over the concatenated features is gone.
__global__ void __launch_bounds__(256) style_concat_rows_kernel(const float* __restrict__ x, int64_t rows, int C,
                                                                const int64_t* __restrict__ idx, const float* __restrict__ soft,
                                                                int rep, const void* __restrict__ emb, int pdt, int S, int sd,
                                                                float* __restrict__ out, __nv_bfloat16* __restrict__ planes, int pfmt,
                                                                int64_t pstride, int rs) {
  const int lane = threadIdx.x & 31;
  const int Co = C + sd;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float sv = 0.f;
    if (lane < sd) {
      const int64_t q = r / rep;
      if (idx) {
        const int64_t s_ = idx[q];
        sv = (s_ >= 0 && s_ < S) ? ms_ldp(emb, pdt, s_ * sd + lane) : 0.f;
      } else {
        for (int s_ = 0; s_ < S; s_++) sv = fmaf(__ldg(soft + q * S + s_), ms_ldp(emb, pdt, (int64_t)s_ * sd + lane), sv);
      }
    }
    const float* xr = x + r * C;
    float* orow = out + r * Co;
    for (int c0 = 4 * lane; c0 < C; c0 += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c0));
      *reinterpret_cast<float2*>(orow + c0) = make_float2(v.x, v.y);
      *reinterpret_cast<float2*>(orow + c0 + 2) = make_float2(v.z, v.w);
      if (planes) store_planes4(planes, pfmt, pstride, r * rs + c0, v);
    }
    if (lane < sd) orow[C + lane] = sv;
    if (planes && lane < rs - C) store_planes1(planes, pfmt, pstride, r * rs + C + lane, lane < sd ? sv : 0.f);
  }
}

// dx = dout[:, :C]
__global__ void slice_cols_kernel(const float* __restrict__ dout, int64_t rows, int C, int Co, float* __restrict__ dx) {
  int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C;
    int c = (int)(i - r * C);
    dx[i] = __ldg(dout + r * Co + c);
  }
}

// scatter-add of the style columns: a block stages a [S][sd] table in shared memory (one smem
// atomic per element, rows of one sequence share a style so contention is a broadcast add),
// then one global atomic per table entry and block.  'lin' mode adds soft^T * dstyle instead.
__global__ void __launch_bounds__(256) style_scatter_kernel(const float* __restrict__ dout, int64_t rows, int C,
                                                            const int64_t* __restrict__ idx, const float* __restrict__ soft,
                                                            int rep, int S, int sd, float* __restrict__ demb) {
  extern __shared__ float tab[];   // S*sd
  int Co = C + sd;
  for (int i = threadIdx.x; i < S * sd; i += blockDim.x) tab[i] = 0.f;
  __syncthreads();
  int64_t per = ms_cdiv_dev(rows, gridDim.x);
  int64_t r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // one warp per row: lanes over the sd style columns
  for (int64_t r = r0 + warp; r < r1; r += nw) {
    int64_t q = r / rep;
    for (int j = lane; j < sd; j += 32) {
      float d = __ldg(dout + r * Co + C + j);
      if (idx) {
        int64_t s = idx[q];
        if (s >= 0 && s < S) atomicAdd(&tab[s * sd + j], d);
      } else {
        for (int s = 0; s < S; s++) atomicAdd(&tab[s * sd + j], __ldg(soft + q * S + s) * d);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * sd; i += blockDim.x)
    if (tab[i] != 0.f) atomicAdd(demb + i, tab[i]);
}

// Warp-per-row backward of the concat in ONE pass over dout ('emb' mode, C % 128 == 0, sd <= 32): the content part of a row goes
// to dx (8-byte loads: rows of C + sd floats are 8-byte aligned; 16-byte stores), the style part is summed in the lanes over
// the warp's consecutive rows while they share a style row (rows of a sequence do) and leaves as sd shared-memory atomics
// per run, then one global atomic per table entry and block -- the scatter-add of the embedding gradient.
__global__ void __launch_bounds__(256) style_concat_rows_bwd_kernel(const float* __restrict__ dout, int64_t rows, int C,
                                                                    const int64_t* __restrict__ idx, int rep, int S, int sd,
                                                                    float* __restrict__ dx, float* __restrict__ demb) {
  extern __shared__ float tab[];   // S*sd
  const int Co = C + sd;
  for (int i = threadIdx.x; i < S * sd; i += blockDim.x) tab[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // contiguous chunk of rows per warp
  const int64_t warps_total = (int64_t)gridDim.x * nw, wid = (int64_t)blockIdx.x * nw + warp;
  const int64_t per = (rows + warps_total - 1) / warps_total;
  const int64_t r0 = wid * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  int64_t cur = -1;
  for (int64_t r = r0; r < r1; r++) {
    const float* dr = dout + r * Co;
    if (dx) {
      float* xr = dx + r * C;
      for (int c0 = 4 * lane; c0 < C; c0 += 128) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(dr + c0)), b = __ldg(reinterpret_cast<const float2*>(dr + c0 + 2));
        *reinterpret_cast<float4*>(xr + c0) = make_float4(a.x, a.y, b.x, b.y);
      }
    }
    if (demb) {
      const int64_t s_ = idx[r / rep];
      if (s_ != cur) {
        if (cur >= 0 && cur < S && lane < sd) atomicAdd(&tab[cur * sd + lane], acc);
        acc = 0.f;
        cur = s_;
      }
      if (lane < sd) acc += __ldg(dr + C + lane);
    }
  }
  if (demb && cur >= 0 && cur < S && lane < sd) atomicAdd(&tab[cur * sd + lane], acc);
  __syncthreads();
  if (demb)
    for (int i = threadIdx.x; i < S * sd; i += blockDim.x)
      if (tab[i] != 0.f) atomicAdd(demb + i, tab[i]);
}

// dsoft[q, s] = sum_{r in q} sum_j dstyle[r, j] * emb[s, j]
__global__ void style_dsoft_kernel(const float* __restrict__ dout, int64_t nq, int C, int rep, const void* __restrict__ emb,
                                   int pdt, int S, int sd, float* __restrict__ dsoft) {
  int Co = C + sd;
  int64_t total = nq * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = i / S;
    int s = (int)(i - q * S);
    float acc = 0.f;
    for (int t = 0; t < rep; t++) {
      const float* d = dout + (q * rep + t) * Co + C;
      for (int j = 0; j < sd; j++) acc = fmaf(__ldg(d + j), ms_ldp(emb, pdt, (int64_t)s * sd + j), acc);
    }
    dsoft[i] = acc;
  }
}

// ------------------------------------------------------------------ softmax + CE + argmax (K <= 64), one thread per row
constexpr int SM_MAXK = 64;

__global__ void softmax_ce_fwd_kernel(const float* __restrict__ score, int64_t rows, int K, const int64_t* __restrict__ target,
                                      int trep, float* __restrict__ soft, int64_t* __restrict__ amax, double* __restrict__ loss_sum) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double loss = 0.0;
  if (r < rows) {
    const float* s = score + r * K;
    float m = s[0];
    int am = 0;
    for (int k = 1; k < K; k++) {
      float v = s[k];
      if (v > m) { m = v; am = k; }           // first maximum wins, as torch.argmax on ties
    }
    float e[SM_MAXK];
    float z = 0.f;
    for (int k = 0; k < K; k++) { e[k] = expf(s[k] - m); z += e[k]; }
    float inv = 1.f / z;
    if (soft) for (int k = 0; k < K; k++) soft[r * K + k] = e[k] * inv;
    if (amax) amax[r] = am;
    if (target && loss_sum) {
      int64_t t = target[r / trep];
      if (t >= 0 && t < K) loss = (double)(logf(z) + m - s[t]);
    }
  }
  if (loss_sum) {
    loss = ms_warp_sum_d(loss);
    if ((threadIdx.x & 31) == 0 && loss != 0.0) atomicAdd(loss_sum, loss);
  }
}

__global__ void softmax_ce_bwd_kernel(const float* __restrict__ soft, int64_t rows, int K, const int64_t* __restrict__ target,
                                      int trep, const float* __restrict__ g_ce, const float* __restrict__ dsoft,
                                      float* __restrict__ dscore) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float g = (g_ce && target) ? g_ce[0] / (float)rows : 0.f;
  int64_t t = target ? target[r / trep] : -1;
  float dot = 0.f;
  if (dsoft) for (int k = 0; k < K; k++) dot = fmaf(soft[r * K + k], dsoft[r * K + k], dot);
  for (int k = 0; k < K; k++) {
    float p = soft[r * K + k];
    float v = g * (p - (k == t ? 1.f : 0.f));
    if (dsoft) v += p * (dsoft[r * K + k] - dot);
    dscore[r * K + k] = v;
  }
}

// ------------------------------------------------------------------ cluster mixture
__global__ void mixture_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w, int64_t rows, int K, int P,
                                   float* __restrict__ out) {
  int64_t total = rows * P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / P;
    int p = (int)(i - r * P);
    float acc = 0.f;
    for (int k = 0; k < K; k++) acc = fmaf(__ldg(w + r * K + k), __ldg(z + (r * K + k) * P + p), acc);
    out[i] = acc;
  }
}

// one warp per (row, k): dz = w * dout ; dw = <z, dout>
__global__ void mixture_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ z, const float* __restrict__ w,
                                   int64_t rows, int K, int P, float* __restrict__ dz, float* __restrict__ dw) {
  int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < rows * K; i += nwarps) {
    int64_t r = i / K;
    float wk = __ldg(w + i);
    float acc = 0.f;
    for (int p = lane; p < P; p += 32) {
      float d = __ldg(dout + r * P + p);
      acc = fmaf(__ldg(z + i * P + p), d, acc);
      dz[i * P + p] = wk * d;
    }
    acc = ms_warp_sum(acc);
    if (lane == 0) dw[i] = acc;
  }
}

__global__ void mean_rows_fwd_kernel(const float* __restrict__ x, int B, int L, int C, float* __restrict__ y) {
  int64_t total = (int64_t)B * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / C;
    int c = (int)(i - b * C);
    float acc = 0.f;
    for (int l = 0; l < L; l++) acc += __ldg(x + (b * L + l) * C + c);
    y[i] = acc / (float)L;
  }
}
__global__ void mean_rows_bwd_kernel(const float* __restrict__ dy, int B, int L, int C, float* __restrict__ dx) {
  int64_t total = (int64_t)B * L * C;
  float inv = 1.f / (float)L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t b = i / ((int64_t)L * C);
    dx[i] = __ldg(dy + b * C + c) * inv;
  }
}

// ------------------------------------------------------------------ velocity / L1
// 16-byte form (P % 4 == 0): a thread owns four features of one frame; the previous frame's four sit P floats earlier
__global__ void velocity_fwd4_kernel(const float4* __restrict__ x, int B, int T, int P4, float4* __restrict__ v) {
  const int64_t total = (int64_t)B * T * P4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)((i / P4) % T);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t) {
      const float4 a = __ldg(x + i), b = __ldg(x + i - P4);
      o = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    }
    v[i] = o;
  }
}
__global__ void velocity_bwd4_kernel(const float4* __restrict__ dv, int B, int T, int P4, float4* __restrict__ dx) {
  const int64_t total = (int64_t)B * T * P4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)((i / P4) % T);
    const float4 a = t >= 1 ? __ldg(dv + i) : zero;
    const float4 b = t + 1 < T ? __ldg(dv + i + P4) : zero;
    dx[i] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
  }
}
__global__ void velocity_fwd_kernel(const float* __restrict__ x, int B, int T, int P, float* __restrict__ v) {
  int64_t total = (int64_t)B * T * P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)((i / P) % T);
    v[i] = t == 0 ? 0.f : __ldg(x + i) - __ldg(x + i - P);
  }
}
__global__ void velocity_bwd_kernel(const float* __restrict__ dv, int B, int T, int P, float* __restrict__ dx) {
  int64_t total = (int64_t)B * T * P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)((i / P) % T);
    float a = t >= 1 ? __ldg(dv + i) : 0.f;
    float b = t + 1 < T ? __ldg(dv + i + P) : 0.f;
    dx[i] = a - b;
  }
}

__global__ void __launch_bounds__(256) l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float c, int64_t n,
                                                     double* __restrict__ loss_sum, float* __restrict__ sgn) {
  __shared__ double part[8];
  double acc = 0.0;
  int64_t start = 0;
  if (!sgn && ((((uintptr_t)a | (uintptr_t)b) & 15) == 0)) {
    // no sign tensor to write: 16-byte loads, fp32 partial of four elements before the fp64 accumulation
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i);
      const float4 bv = b ? __ldg(reinterpret_cast<const float4*>(b) + i) : make_float4(c, c, c, c);
      acc += (double)fabsf(av.x - bv.x) + (double)fabsf(av.y - bv.y) + (double)fabsf(av.z - bv.z) + (double)fabsf(av.w - bv.w);
    }
    start = n4 << 2;
  }
  for (int64_t i = start + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = __ldg(a + i) - (b ? __ldg(b + i) : c);
    acc += fabsf(d);
    if (sgn) sgn[i] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  }
  acc = ms_warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); i++) t += part[i];
    atomicAdd(loss_sum, t);
  }
}
// backward without a stored sign tensor: da = sign(a - b) * g / n recomputed from the operands (8 B read + 4 B written per
// element instead of 4 B written forward + 4 B read + 4 B written backward)
__global__ void l1_bwd_ab_kernel(const float* __restrict__ a, const float* __restrict__ b, float c, const float* __restrict__ g,
                                 int64_t n, float* __restrict__ da) {
  const float s = g[0] / (float)n;
  const int64_t n4 = n >> 2;
  const bool vec = (((uintptr_t)a | (uintptr_t)da | (uintptr_t)b) & 15) == 0;
  if (vec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i);
      const float4 bv = b ? __ldg(reinterpret_cast<const float4*>(b) + i) : make_float4(c, c, c, c);
      float4 o;
      o.x = av.x > bv.x ? s : (av.x < bv.x ? -s : 0.f); o.y = av.y > bv.y ? s : (av.y < bv.y ? -s : 0.f);
      o.z = av.z > bv.z ? s : (av.z < bv.z ? -s : 0.f); o.w = av.w > bv.w ? s : (av.w < bv.w ? -s : 0.f);
      reinterpret_cast<float4*>(da)[i] = o;
    }
  }
  for (int64_t i = (vec ? (n4 << 2) : 0) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __ldg(a + i) - (b ? __ldg(b + i) : c);
    da[i] = d > 0.f ? s : (d < 0.f ? -s : 0.f);
  }
}
__global__ void l1_bwd_kernel(const float* __restrict__ sgn, const float* __restrict__ g, int64_t n, float* __restrict__ da) {
  float s = g[0] / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    da[i] = __ldg(sgn + i) * s;
}
__global__ void scalar_finish_kernel(const double* in, double scale, float* out) { out[0] = (float)(in[0] * scale); }

// The train step's loss bookkeeping in one launch: report[i] = w_i * l_i (w_i = host weight x optional device-resident
// lambda), total = sum_i report[i]; and its backward, g_i = w_i * gtotal.
struct LossTerms {
  const float* l[MS_LOSS_MAX_TERMS];
  double w[MS_LOSS_MAX_TERMS];
  int lam[MS_LOSS_MAX_TERMS];
  int n;
};
__global__ void loss_combine_kernel(LossTerms t, const double* __restrict__ lam_dev, float* __restrict__ total,
                                    double* __restrict__ report) {
  double s = 0.0;
  for (int i = 0; i < t.n; i++) {
    const double w = t.w[i] * (t.lam[i] >= 0 ? lam_dev[t.lam[i]] : 1.0);
    const double v = w * (double)t.l[i][0];
    if (report) report[i] = v;
    s += v;
  }
  total[0] = (float)s;
}
__global__ void loss_combine_bwd_kernel(LossTerms t, const double* __restrict__ lam_dev, const float* __restrict__ gtotal,
                                        float* __restrict__ g) {
  const int i = threadIdx.x;
  if (i < t.n) g[i] = (float)(t.w[i] * (t.lam[i] >= 0 ? lam_dev[t.lam[i]] : 1.0) * (double)gtotal[0]);
}

dim3 col_grid(int64_t rows, int C) {
  int gx = (int)ms_cdiv(C, 32);
  int64_t want = ms_cdiv((int64_t)ms_num_sms() * 4, gx);
  int64_t maxy = ms_cdiv(rows, 64);
  int64_t gy = want < maxy ? want : maxy;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3((unsigned)gx, (unsigned)gy);
}

}  // namespace

#define ST ms_stream(stream)

extern "C" int ms_version(void) { return 100; }

extern "C" int ms_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

extern "C" int ms_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return MS_EINVAL;
  if (n == 0) return 0;
  int blocks = ew_blocks(n);
#define CASE(S, SD, D, DD) if (sdt == SD && ddt == DD) { cast_kernel<S, D><<<blocks, EW_THREADS, 0, ST>>>((const S*)src, (D*)dst, n); MS_LAUNCH_CHECK(); return 0; }
  CASE(float, MS_F32, float, MS_F32)
  CASE(float, MS_F32, double, MS_F64)
  CASE(double, MS_F64, float, MS_F32)
  CASE(double, MS_F64, double, MS_F64)
  CASE(float, MS_F32, __nv_bfloat16, MS_BF16)
  CASE(double, MS_F64, __nv_bfloat16, MS_BF16)
  CASE(__nv_bfloat16, MS_BF16, float, MS_F32)
  CASE(__nv_bfloat16, MS_BF16, double, MS_F64)
#undef CASE
  return MS_EINVAL;
}

extern "C" int ms_scale_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, double scale, void* stream) {
  if (!src || !dst || n < 0) return MS_EINVAL;
  if (n == 0) return 0;
  int blocks = ew_blocks((n + 1) / 2);
#define CASE(S, SD, D, DD) if (sdt == SD && ddt == DD) { scale_cast_kernel<S, D><<<blocks, EW_THREADS, 0, ST>>>((const S*)src, (D*)dst, n, scale); MS_LAUNCH_CHECK(); return 0; }
  CASE(float, MS_F32, float, MS_F32)
  CASE(float, MS_F32, double, MS_F64)
  CASE(double, MS_F64, float, MS_F32)
  CASE(double, MS_F64, double, MS_F64)
#undef CASE
  return MS_EINVAL;
}

extern "C" int ms_pack_conv_weight_f32(const void* w, int pdt, const ms_conv_desc* d, float* wf, float* wt, void* stream) {
  if (!w || !d || (!wf && !wt) || d->groups < 1 || d->Cin % d->groups || d->Cout % d->groups) return MS_EINVAL;
  int64_t total = (int64_t)d->Cout * (d->Cin / d->groups) * d->kh * d->kw;
  pack_weight_kernel<<<ew_blocks(total), EW_THREADS, 0, ST>>>(w, pdt, d->Cout, d->Cin / d->groups, d->kh * d->kw, d->groups, wf, wt);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_unpack_conv_wgrad(const float* dwf, const ms_conv_desc* d, void* dw, int pdt, int accumulate, void* stream) {
  if (!dwf || !d || !dw || d->groups < 1) return MS_EINVAL;
  int64_t total = (int64_t)d->Cout * (d->Cin / d->groups) * d->kh * d->kw;
  unpack_wgrad_kernel<<<ew_blocks(total), EW_THREADS, 0, ST>>>(dwf, d->Cout, d->Cin / d->groups, d->kh * d->kw, d->groups, dw, pdt, accumulate);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_store_param_grad(const double* src, int n, void* dst, int pdt, int accumulate, void* stream) {
  if (!src || !dst || n < 0) return MS_EINVAL;
  if (n == 0) return 0;
  store_param_grad_kernel<<<(n + 255) / 256, 256, 0, ST>>>(src, n, dst, pdt, accumulate);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_col_stats_f32(const float* x, int64_t rows, int C, double* sum, double* sumsq, void* stream) {
  if (!x || !sum || rows < 1 || C < 1) return MS_EINVAL;
  col_stats_kernel<<<col_grid(rows, C), 256, 0, ST>>>(x, rows, C, sum, sumsq);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_bn_finalize(const double* sum, const double* sumsq, int64_t rows, int C, const void* gamma, const void* beta,
                              const void* conv_bias, void* running_mean, void* running_var, int pdt, int training,
                              float momentum, float eps, float* scale, float* shift, float* mean, float* rstd, void* stream) {
  if (!gamma || !beta || !running_mean || !running_var || !scale || !shift || !mean || !rstd || C < 1) return MS_EINVAL;
  if (training && (!sum || !sumsq || rows < 1)) return MS_EINVAL;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(sum, sumsq, rows, C, gamma, beta, conv_bias, running_mean, running_var,
                                                      pdt, training, momentum, eps, scale, shift, mean, rstd);
  MS_LAUNCH_CHECK();
  return 0;
}

static bool planes_ok(const void* planes, int pfmt, int64_t pstride) {
  if (!planes) return true;
  if (pfmt != MS_BF16 && pfmt != MS_BF16X2) return false;
  if (pfmt == MS_BF16X2 && (pstride <= 0 || pstride % 8)) return false;
  return ((uintptr_t)planes & 15) == 0;
}

extern "C" int ms_bn_stats_finalize(const float* x, int64_t rows, int C, double* sum, double* sumsq, void* ticket,
                                    const void* gamma, const void* beta, const void* conv_bias, void* running_mean,
                                    void* running_var, int64_t* num_batches_tracked, int pdt, float momentum, float eps,
                                    float* scale, float* shift, float* mean, float* rstd, void* stream) {
  if (!x || !sum || !sumsq || !ticket || !gamma || !beta || !running_mean || !running_var || !scale || !shift || !mean || !rstd)
    return MS_EINVAL;
  if (rows < 1 || C < 1) return MS_EINVAL;
  bn_stats_finalize_kernel<<<col_grid(rows, C), 256, 0, ST>>>(x, rows, C, sum, sumsq, (unsigned int*)ticket, gamma, beta, conv_bias,
                                                              running_mean, running_var, (long long*)num_batches_tracked, pdt,
                                                              momentum, eps, scale, shift, mean, rstd);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_bn_act_fwd_f32(const float* x, const float* scale, const float* shift, float slope, int64_t rows, int C,
                                 float* y, const float* res, int up2, int rows_per_seq, void* planes, int pfmt,
                                 int64_t pstride, void* stream) {
  if (!x || !scale || !shift || (!y && !planes) || rows < 1 || C < 1) return MS_EINVAL;
  if (up2 && (rows_per_seq < 1 || rows % rows_per_seq)) return MS_EINVAL;
  if (!planes_ok(planes, pfmt, pstride)) return MS_EINVAL;
  int64_t rows_out = up2 ? rows * 2 : rows;
  __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
  bool vec = (C % 4 == 0) && ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)res) & 15) == 0);
  if (vec) bn_act_fwd_kernel<4><<<ew_blocks(rows_out * C / 4), EW_THREADS, 0, ST>>>(x, scale, shift, slope, rows_out, C, y, res, up2, rows_per_seq, pl, pfmt, pstride);
  else bn_act_fwd_kernel<1><<<ew_blocks(rows_out * C), EW_THREADS, 0, ST>>>(x, scale, shift, slope, rows_out, C, y, res, up2, rows_per_seq, pl, pfmt, pstride);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_to_planes(const float* x, int64_t rows, int C, int row_stride, void* planes, int pfmt, int64_t pstride,
                            void* stream) {
  if (!x || !planes || rows < 1 || C < 1 || row_stride < C) return MS_EINVAL;
  if (!planes_ok(planes, pfmt, pstride)) return MS_EINVAL;
  __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
  int64_t n = rows * C;
  if (row_stride == C && n % 4 == 0 && ((uintptr_t)x & 15) == 0)
    to_planes4_kernel<<<ew_blocks(n / 4), EW_THREADS, 0, ST>>>(x, n / 4, pl, pfmt, pstride);
  else
    to_planes_kernel<<<ew_blocks(rows * row_stride), EW_THREADS, 0, ST>>>(x, rows, C, row_stride, pl, pfmt, pstride);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_planes_to_f32(const void* planes, int pfmt, int64_t pstride, int64_t rows, int C, int row_stride, float* x,
                                void* stream) {
  if (!x || !planes || rows < 1 || C < 1 || row_stride < C) return MS_EINVAL;
  if (!planes_ok(planes, pfmt, pstride)) return MS_EINVAL;
  planes_to_f32_kernel<<<ew_blocks(rows * C), EW_THREADS, 0, ST>>>(reinterpret_cast<const __nv_bfloat16*>(planes), pfmt, pstride,
                                                                  rows, C, row_stride, x);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_bn_act_bwd_reduce_f32(const float* dy, const float* x, const float* scale, const float* shift,
                                        const float* mean, const float* rstd, float slope, int64_t rows, int C, int up2,
                                        int rows_per_seq, double* dgamma, double* dbeta, void* stream) {
  if (!dy || !x || !scale || !shift || !mean || !rstd || !dgamma || !dbeta || rows < 1 || C < 1) return MS_EINVAL;
  if (!up2 && (C == 64 || C == 128 || C == 256) && rows >= 4096 &&
      !(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)mean | (uintptr_t)rstd) & 15)) {
    int64_t blocks = ms_cdiv(rows, 64);
    if (blocks > (int64_t)ms_num_sms() * 6) blocks = (int64_t)ms_num_sms() * 6;
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    if (C == 64) bn_act_bwd_reduce4_kernel<16><<<(unsigned)blocks, 256, 0, ST>>>(d4, x4, scale, shift, mean, rstd, slope, rows, dgamma, dbeta);
    else if (C == 128) bn_act_bwd_reduce4_kernel<32><<<(unsigned)blocks, 256, 0, ST>>>(d4, x4, scale, shift, mean, rstd, slope, rows, dgamma, dbeta);
    else bn_act_bwd_reduce4_kernel<64><<<(unsigned)blocks, 256, 0, ST>>>(d4, x4, scale, shift, mean, rstd, slope, rows, dgamma, dbeta);
    MS_LAUNCH_CHECK();
    return 0;
  }
  bn_act_bwd_reduce_kernel<<<col_grid(rows, C), 256, 0, ST>>>(dy, x, scale, shift, mean, rstd, slope, rows, C, up2,
                                                              rows_per_seq, dgamma, dbeta);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_bn_act_bwd_apply_f32(const float* dy, const float* x, const float* scale, const float* shift,
                                       const float* mean, const float* rstd, float slope, int64_t rows, int C, int up2,
                                       int rows_per_seq, const double* dgamma, const double* dbeta, int training, float* dx,
                                       void* planes, int pfmt, int64_t pstride, void* grad_gamma, void* grad_beta, int gdt,
                                       void* stream) {
  if (!dy || !x || !scale || !shift || !mean || !rstd || (!dx && !planes) || rows < 1 || C < 1) return MS_EINVAL;
  if (training && (!dgamma || !dbeta)) return MS_EINVAL;
  if ((grad_gamma || grad_beta) && (!dgamma || !dbeta || (gdt != MS_F32 && gdt != MS_F64))) return MS_EINVAL;
  if (!planes_ok(planes, pfmt, pstride)) return MS_EINVAL;
  if (!up2 && C % 4 == 0 && !(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)mean |
                                (uintptr_t)rstd) & 15)) {
    bn_act_bwd_apply4_kernel<<<ew_blocks(rows * (C / 4)), EW_THREADS, 0, ST>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), scale, shift, mean, rstd, slope, rows, C / 4, dgamma,
        dbeta, training, reinterpret_cast<float4*>(dx), reinterpret_cast<__nv_bfloat16*>(planes), pfmt, pstride, grad_gamma, grad_beta,
        gdt);
    MS_LAUNCH_CHECK();
    return 0;
  }
  bn_act_bwd_apply_kernel<<<ew_blocks(rows * C), EW_THREADS, 0, ST>>>(dy, x, scale, shift, mean, rstd, slope, rows, C, up2,
                                                                      rows_per_seq, dgamma, dbeta, training, dx,
                                                                      reinterpret_cast<__nv_bfloat16*>(planes), pfmt, pstride,
                                                                      grad_gamma, grad_beta, gdt);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_lrelu_bwd_f32(const float* dy, const float* y, float slope, int64_t n, float* dz, void* planes, int pfmt,
                                int64_t pstride, void* stream) {
  if (!dy || !y || (!dz && !planes) || n < 1) return MS_EINVAL;
  if (!planes_ok(planes, pfmt, pstride)) return MS_EINVAL;
  lrelu_bwd_kernel<<<ew_blocks(n), EW_THREADS, 0, ST>>>(dy, y, slope, n, dz, reinterpret_cast<__nv_bfloat16*>(planes), pfmt, pstride);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_bilinear_to_T_fwd_f32(const float* x, int B, int Hi, int Wi, int C, int T, float* y, void* stream) {
  if (!x || !y || B < 1 || Hi < 1 || Wi < 1 || C < 1 || T < 1) return MS_EINVAL;
  bilinear_fwd_kernel<<<ew_blocks((int64_t)B * T * C), EW_THREADS, 0, ST>>>(x, B, Hi, Wi, C, T, y);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_bilinear_to_T_bwd_f32(const float* dy, int B, int Hi, int Wi, int C, int T, float* dx, void* stream) {
  if (!dy || !dx || B < 1 || Hi < 1 || Wi < 1 || C < 1 || T < 1) return MS_EINVAL;
  bilinear_bwd_kernel<<<ew_blocks((int64_t)B * Hi * Wi * C), EW_THREADS, 0, ST>>>(dy, B, Hi, Wi, C, T, dx);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_style_concat_fwd_f32(const float* x, int64_t rows, int C, const int64_t* idx, const float* soft, int rep,
                                       const void* emb, int pdt, int S, int sd, float* out, void* stream) {
  if (!x || !emb || !out || (!idx && !soft) || rows < 1 || rep < 1 || rows % rep || S < 1 || sd < 1) return MS_EINVAL;
  style_concat_fwd_kernel<<<ew_blocks(rows * (C + sd)), EW_THREADS, 0, ST>>>(x, rows, C, idx, soft, rep, emb, pdt, S, sd, out);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_style_concat_planes_fwd_f32(const float* x, int64_t rows, int C, const int64_t* idx, const float* soft, int rep,
                                              const void* emb, int pdt, int S, int sd, float* out, void* planes, int pfmt,
                                              int64_t pstride, int rs, void* stream) {
  if (!x || !emb || !out || (!idx && !soft) || rows < 1 || rep < 1 || rows % rep || S < 1 || sd < 1) return MS_EINVAL;
  if (C % 128 || sd > 32 || ((uintptr_t)x & 15) || ((uintptr_t)out & 7) || (C + sd) % 2) return MS_EINVAL;
  if (planes && (rs < C + sd || rs - C > 32 || rs % 4 || ((uintptr_t)planes & 7) || !planes_ok(planes, pfmt, pstride))) return MS_EINVAL;
  int64_t blocks = ms_cdiv(rows, 8);
  const int64_t cap = (int64_t)ms_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  style_concat_rows_kernel<<<(unsigned)blocks, 256, 0, ST>>>(x, rows, C, idx, soft, rep, emb, pdt, S, sd, out,
                                                             reinterpret_cast<__nv_bfloat16*>(planes), pfmt, pstride, rs);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_style_concat_bwd_f32(const float* dout, int64_t rows, int C, const int64_t* idx, const float* soft, int rep,
                                       const void* emb, int pdt, int S, int sd, float* dx, float* demb, float* dsoft,
                                       void* stream) {
  if (!dout || (!idx && !soft) || rows < 1 || rep < 1 || rows % rep || S < 1 || sd < 1) return MS_EINVAL;
  if (idx && !soft && C % 128 == 0 && sd <= 32 && (C + sd) % 2 == 0 && !(((uintptr_t)dout) & 7) && !(((uintptr_t)dx) & 15) &&
      (size_t)S * sd * sizeof(float) <= 40 * 1024) {
    int64_t blocks = ms_cdiv(rows, 8 * 2);          // two rows per warp at least (eight left 16 CTAs for the 1024 rows of batch 16)
    const int64_t cap = (int64_t)ms_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    style_concat_rows_bwd_kernel<<<(unsigned)blocks, 256, sizeof(float) * S * sd, ST>>>(dout, rows, C, idx, rep, S, sd, dx, demb);
    MS_LAUNCH_CHECK();
    return 0;
  }
  if (dx) {
    slice_cols_kernel<<<ew_blocks(rows * C), EW_THREADS, 0, ST>>>(dout, rows, C, C + sd, dx);
    MS_LAUNCH_CHECK();
  }
  if (demb) {
    int blocks = (int)ms_cdiv(rows, 256);
    if (blocks > ms_num_sms() * 2) blocks = ms_num_sms() * 2;
    style_scatter_kernel<<<blocks, 256, sizeof(float) * S * sd, ST>>>(dout, rows, C, idx, soft, rep, S, sd, demb);
    MS_LAUNCH_CHECK();
  }
  if (dsoft && soft) {
    style_dsoft_kernel<<<ew_blocks(rows / rep * S), EW_THREADS, 0, ST>>>(dout, rows / rep, C, rep, emb, pdt, S, sd, dsoft);
    MS_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int ms_softmax_ce_fwd_f32(const float* score, int64_t rows, int K, const int64_t* target, int trep, float* soft,
                                     int64_t* amax, double* loss_sum, void* stream) {
  if (!score || rows < 1 || K < 1 || K > SM_MAXK || trep < 1) return MS_EINVAL;
  softmax_ce_fwd_kernel<<<(unsigned)ms_cdiv(rows, 128), 128, 0, ST>>>(score, rows, K, target, trep, soft, amax, loss_sum);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_softmax_ce_bwd_f32(const float* soft, int64_t rows, int K, const int64_t* target, int trep, const float* g_ce,
                                     const float* dsoft, float* dscore, void* stream) {
  if (!soft || !dscore || rows < 1 || K < 1 || trep < 1) return MS_EINVAL;
  softmax_ce_bwd_kernel<<<(unsigned)ms_cdiv(rows, 128), 128, 0, ST>>>(soft, rows, K, target, trep, g_ce, dsoft, dscore);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_mixture_fwd_f32(const float* z, const float* w, int64_t rows, int K, int P, float* out, void* stream) {
  if (!z || !w || !out || rows < 1 || K < 1 || P < 1) return MS_EINVAL;
  mixture_fwd_kernel<<<ew_blocks(rows * P), EW_THREADS, 0, ST>>>(z, w, rows, K, P, out);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_mixture_bwd_f32(const float* dout, const float* z, const float* w, int64_t rows, int K, int P, float* dz,
                                  float* dw, void* stream) {
  if (!dout || !z || !w || !dz || !dw || rows < 1) return MS_EINVAL;
  mixture_bwd_kernel<<<ew_blocks(rows * K * 32), EW_THREADS, 0, ST>>>(dout, z, w, rows, K, P, dz, dw);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_mean_rows_fwd_f32(const float* x, int B, int L, int C, float* y, void* stream) {
  if (!x || !y || B < 1 || L < 1 || C < 1) return MS_EINVAL;
  mean_rows_fwd_kernel<<<ew_blocks((int64_t)B * C), EW_THREADS, 0, ST>>>(x, B, L, C, y);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_mean_rows_bwd_f32(const float* dy, int B, int L, int C, float* dx, void* stream) {
  if (!dy || !dx || B < 1 || L < 1 || C < 1) return MS_EINVAL;
  mean_rows_bwd_kernel<<<ew_blocks((int64_t)B * L * C), EW_THREADS, 0, ST>>>(dy, B, L, C, dx);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_velocity_fwd_f32(const float* x, int B, int T, int P, float* v, void* stream) {
  if (!x || !v || B < 1 || T < 1 || P < 1) return MS_EINVAL;
  if (P % 4 == 0 && !(((uintptr_t)x | (uintptr_t)v) & 15)) {
    velocity_fwd4_kernel<<<ew_blocks((int64_t)B * T * (P / 4)), EW_THREADS, 0, ST>>>(reinterpret_cast<const float4*>(x), B, T, P / 4,
                                                                                   reinterpret_cast<float4*>(v));
    MS_LAUNCH_CHECK();
    return 0;
  }
  velocity_fwd_kernel<<<ew_blocks((int64_t)B * T * P), EW_THREADS, 0, ST>>>(x, B, T, P, v);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_velocity_bwd_f32(const float* dv, int B, int T, int P, float* dx, void* stream) {
  if (!dv || !dx || B < 1 || T < 1 || P < 1) return MS_EINVAL;
  if (P % 4 == 0 && !(((uintptr_t)dv | (uintptr_t)dx) & 15)) {
    velocity_bwd4_kernel<<<ew_blocks((int64_t)B * T * (P / 4)), EW_THREADS, 0, ST>>>(reinterpret_cast<const float4*>(dv), B, T, P / 4,
                                                                                   reinterpret_cast<float4*>(dx));
    MS_LAUNCH_CHECK();
    return 0;
  }
  velocity_bwd_kernel<<<ew_blocks((int64_t)B * T * P), EW_THREADS, 0, ST>>>(dv, B, T, P, dx);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_l1_fwd_f32(const float* a, const float* b, float c, int64_t n, double* loss_sum, float* sgn, void* stream) {
  if (!a || !loss_sum || n < 1) return MS_EINVAL;
  int blocks = ew_blocks(n, 256 * 4);
  l1_fwd_kernel<<<blocks, 256, 0, ST>>>(a, b, c, n, loss_sum, sgn);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_l1_bwd_f32(const float* sgn, const float* g, int64_t n, float* da, void* stream) {
  if (!sgn || !g || !da || n < 1) return MS_EINVAL;
  l1_bwd_kernel<<<ew_blocks(n), EW_THREADS, 0, ST>>>(sgn, g, n, da);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_l1_bwd_ab_f32(const float* a, const float* b, float c, const float* g, int64_t n, float* da, void* stream) {
  if (!a || !g || !da || n < 1) return MS_EINVAL;
  l1_bwd_ab_kernel<<<ew_blocks((n + 3) / 4), EW_THREADS, 0, ST>>>(a, b, c, g, n, da);
  MS_LAUNCH_CHECK();
  return 0;
}
static int fill_terms(LossTerms& t, const float* const* losses, const double* weights, const int* lam_idx, int n,
                      const double* lam_dev) {
  if (!losses || !weights || !lam_idx || n < 1 || n > MS_LOSS_MAX_TERMS) return MS_EINVAL;
  t.n = n;
  for (int i = 0; i < MS_LOSS_MAX_TERMS; i++) {
    t.l[i] = i < n ? losses[i] : nullptr;
    t.w[i] = i < n ? weights[i] : 0.0;
    t.lam[i] = i < n ? lam_idx[i] : -1;
    if (i < n && (!t.l[i] || (t.lam[i] >= 0 && !lam_dev))) return MS_EINVAL;
  }
  return 0;
}
extern "C" int ms_loss_combine(const float* const* losses, const double* weights, const int* lam_idx, int n, const double* lam_dev,
                               float* total, double* report, void* stream) {
  LossTerms t;
  if (!total || fill_terms(t, losses, weights, lam_idx, n, lam_dev)) return MS_EINVAL;
  loss_combine_kernel<<<1, 1, 0, ST>>>(t, lam_dev, total, report);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_loss_combine_bwd(const float* gtotal, const double* weights, const int* lam_idx, int n, const double* lam_dev,
                                   float* g, void* stream) {
  LossTerms t;
  const float* dummy[MS_LOSS_MAX_TERMS];
  for (int i = 0; i < MS_LOSS_MAX_TERMS; i++) dummy[i] = gtotal;
  if (!gtotal || !g || fill_terms(t, dummy, weights, lam_idx, n, lam_dev)) return MS_EINVAL;
  loss_combine_bwd_kernel<<<1, 32, 0, ST>>>(t, lam_dev, gtotal, g);
  MS_LAUNCH_CHECK();
  return 0;
}
extern "C" int ms_scalar_finish(const double* in, double scale, float* out, void* stream) {
  if (!in || !out) return MS_EINVAL;
  scalar_finish_kernel<<<1, 1, 0, ST>>>(in, scale, out);
  MS_LAUNCH_CHECK();
  return 0;
}
