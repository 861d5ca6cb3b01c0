// Convolutions that are not GEMM-shaped: a handful of output channels (the cluster-classifier logits N = K clusters,
// the pose-style scores N = S speakers, the discriminator score N = 1) or a single input channel (the first
// audio-encoder layer: C_in = 1, 3x3, K = 9).  A 64x64 / 128xN tile kernel runs these on one or two CTAs; here the
// work is spread so that every warp streams contiguous channels-last rows:
//   * small-N forward : one warp per output pixel, lanes stride over (tap, channel), N accumulators per lane,
//                       butterfly reduction                                  (reads x once: HBM/L2-bound)
//   * small-N dgrad   : one thread per (input pixel, channel), loops over taps x N
//   * small-N wgrad   : one thread per (tap, channel) and pixel chunk, N accumulators, fp32 atomics per chunk
//   * C_in = 1 forward: one thread per (pixel, out channel); the k taps of x are warp-broadcast loads
//   * C_in = 1 wgrad  : one thread per (tap, out channel) and pixel chunk
// Reference call sites: layers.py:459 (ClusterClassify.logits), layers.py:276-287 (PoseStyleEncoder last block),
// speech2gesture.py:90 (D.logits), layers.py:167 (AudioEncoder conv 0).  Dispatched from ms_conv_*_f32.
#include "common.cuh"

namespace small {

struct P {
  int B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, Ho, Wo, taps;
};

constexpr int NMAX = 32;

__device__ __forceinline__ bool in_pixel(const P& p, int pos, int tap, long long& off) {
  int wo = pos % p.Wo;
  int t = pos / p.Wo;
  int ho = t % p.Ho;
  int b = t / p.Ho;
  int th = tap / p.kw, tw = tap - th * p.kw;
  int hi = ho * p.sh + th - p.ph, wi = wo * p.sw + tw - p.pw;
  if ((unsigned)hi >= (unsigned)p.H || (unsigned)wi >= (unsigned)p.W) return false;
  off = ((long long)(b * p.H + hi) * p.W + wi) * p.Cin;
  return true;
}

// ---- small-N forward.  wf: [tap][c][n] fp32 (a lane reads the N weights of its channel contiguously).
template <int N>
__global__ void __launch_bounds__(256) fwd_small_n(P p, const float* __restrict__ x, const float* __restrict__ wf,
                                                   const float* __restrict__ bias, float* __restrict__ y, int act, float slope,
                                                   int M) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int pos = warp; pos < M; pos += nwarps) {
    float acc[N];
#pragma unroll
    for (int n = 0; n < N; n++) acc[n] = 0.f;
    for (int tap = 0; tap < p.taps; tap++) {
      long long off;
      if (!in_pixel(p, pos, tap, off)) continue;
      const float* xr = x + off;
      const float* wr = wf + (size_t)tap * p.Cout * p.Cin;
      for (int c = lane; c < p.Cin; c += 32) {
        float xv = __ldg(xr + c);
        const float* wc = wr + (size_t)c * p.Cout;
#pragma unroll
        for (int n = 0; n < N; n++)
          if (n < p.Cout) acc[n] = fmaf(xv, __ldg(wc + n), acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < N; n++) acc[n] = ms_warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < N; n++)
        if (n < p.Cout) {
          float v = acc[n] + (bias ? bias[n] : 0.f);
          if (act) v = v > 0.f ? v : v * slope;
          y[(size_t)pos * p.Cout + n] = v;
        }
    }
  }
}

// ---- small-N dgrad.  wt: [tap][n][c].  One thread per (input pixel, channel).
__global__ void __launch_bounds__(256) dgrad_small_n(P p, const float* __restrict__ dy, const float* __restrict__ wt,
                                                     float* __restrict__ dx, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % p.Cin);
    long long ip = i / p.Cin;
    int w = (int)(ip % p.W);
    long long t = ip / p.W;
    int h = (int)(t % p.H);
    int b = (int)(t / p.H);
    float acc = 0.f;
    for (int tap = 0; tap < p.taps; tap++) {
      int th = tap / p.kw, tw = tap - th * p.kw;
      int hh = h + p.ph - th, ww = w + p.pw - tw;
      if (hh < 0 || ww < 0) continue;
      int ho = hh / p.sh, wo = ww / p.sw;
      if (ho * p.sh != hh || wo * p.sw != ww || ho >= p.Ho || wo >= p.Wo) continue;
      const float* d = dy + ((size_t)(b * p.Ho + ho) * p.Wo + wo) * p.Cout;
      const float* wr = wt + (size_t)tap * p.Cout * p.Cin + c;
      for (int n = 0; n < p.Cout; n++) acc = fmaf(__ldg(d + n), __ldg(wr + (size_t)n * p.Cin), acc);
    }
    dx[i] = acc;
  }
}

// ---- small-N wgrad.  dwf: [tap][c][n] (zeroed by the caller when chunks > 1).  grid = (c blocks, taps, pixel chunks)
template <int N>
__global__ void __launch_bounds__(128) wgrad_small_n(P p, const float* __restrict__ x, const float* __restrict__ dy,
                                                     float* __restrict__ dwf, int M, int chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int tap = blockIdx.y;
  if (c >= p.Cin) return;
  const int per = (M + chunks - 1) / chunks;
  const int p0 = blockIdx.z * per, p1 = min(M, p0 + per);
  float acc[N];
#pragma unroll
  for (int n = 0; n < N; n++) acc[n] = 0.f;
  for (int pos = p0; pos < p1; pos++) {
    long long off;
    if (!in_pixel(p, pos, tap, off)) continue;
    float xv = __ldg(x + off + c);
    const float* d = dy + (size_t)pos * p.Cout;
#pragma unroll
    for (int n = 0; n < N; n++)
      if (n < p.Cout) acc[n] = fmaf(xv, __ldg(d + n), acc[n]);
  }
  float* dst = dwf + ((size_t)tap * p.Cin + c) * p.Cout;
#pragma unroll
  for (int n = 0; n < N; n++)
    if (n < p.Cout) {
      if (chunks > 1) atomicAdd(dst + n, acc[n]);
      else dst[n] = acc[n];
    }
}

// ---- C_in = 1 forward.  wf: [tap][0][n].  One thread per (pixel, n), n fastest.
__global__ void __launch_bounds__(256) fwd_cin1(P p, const float* __restrict__ x, const float* __restrict__ wf,
                                                const float* __restrict__ bias, float* __restrict__ y, int act, float slope,
                                                long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i % p.Cout);
    int pos = (int)(i / p.Cout);
    float acc = bias ? bias[n] : 0.f;
    for (int tap = 0; tap < p.taps; tap++) {
      long long off;
      if (in_pixel(p, pos, tap, off)) acc = fmaf(__ldg(x + off), __ldg(wf + (size_t)tap * p.Cout + n), acc);
    }
    if (act) acc = acc > 0.f ? acc : acc * slope;
    y[i] = acc;
  }
}

// ---- C_in = 1 wgrad.  dwf: [tap][0][n], zeroed by the caller.  block = taps*Cout threads (<= 1024), grid = pixel chunks
__global__ void wgrad_cin1(P p, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dwf, int M) {
  const int n = threadIdx.x % p.Cout, tap = threadIdx.x / p.Cout;
  const int per = (M + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(M, p0 + per);
  float acc = 0.f;
  for (int pos = p0; pos < p1; pos++) {
    long long off;
    if (in_pixel(p, pos, tap, off)) acc = fmaf(__ldg(x + off), __ldg(dy + (size_t)pos * p.Cout + n), acc);
  }
  atomicAdd(dwf + (size_t)tap * p.Cout + n, acc);
}

P make(const ms_conv_desc* d) {
  P p;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.kh = d->kh; p.kw = d->kw; p.sh = d->sh; p.sw = d->sw; p.ph = d->ph; p.pw = d->pw;
  p.Ho = d->Ho; p.Wo = d->Wo; p.taps = d->kh * d->kw;
  return p;
}

inline int blocks_for(long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  long long cap = (long long)ms_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace small

// Returns 1 when the launch was taken here, 0 when the caller should use the tiled kernel, < 0 / cudaError on failure.
int ms_small_conv_fwd(const float* x, const float* wf, const float* bias, float* y, const ms_conv_desc* d, int act,
                      float slope, cudaStream_t st) {
  using namespace small;
  if (d->groups != 1) return 0;
  P p = make(d);
  const int M = d->B * d->Ho * d->Wo;
  if (d->Cin == 1 && d->Cout <= 1024) {
    long long total = (long long)M * d->Cout;
    fwd_cin1<<<blocks_for(total, 256), 256, 0, st>>>(p, x, wf, bias, y, act, slope, total);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (d->Cout <= NMAX) {
    int blocks = blocks_for((long long)M * 32, 256);
    if (d->Cout <= 8) fwd_small_n<8><<<blocks, 256, 0, st>>>(p, x, wf, bias, y, act, slope, M);
    else fwd_small_n<NMAX><<<blocks, 256, 0, st>>>(p, x, wf, bias, y, act, slope, M);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  return 0;
}

int ms_small_conv_dgrad(const float* dy, const float* wt, float* dx, const ms_conv_desc* d, cudaStream_t st) {
  using namespace small;
  if (d->groups != 1 || d->Cout > NMAX || d->Cin == 1) return 0;
  P p = make(d);
  long long total = (long long)d->B * d->H * d->W * d->Cin;
  dgrad_small_n<<<blocks_for(total, 256), 256, 0, st>>>(p, dy, wt, dx, total);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ms_small_conv_wgrad(const float* x, const float* dy, float* dwf, const ms_conv_desc* d, cudaStream_t st) {
  using namespace small;
  if (d->groups != 1) return 0;
  P p = make(d);
  const int M = d->B * d->Ho * d->Wo;
  if (d->Cin == 1 && p.taps * d->Cout <= 1024) {
    if (cudaMemsetAsync(dwf, 0, sizeof(float) * (size_t)p.taps * d->Cout, st) != cudaSuccess) return -1;
    int chunks = (M + 511) / 512;
    if (chunks > ms_num_sms() * 2) chunks = ms_num_sms() * 2;
    wgrad_cin1<<<chunks, p.taps * d->Cout, 0, st>>>(p, x, dy, dwf, M);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (d->Cout <= NMAX && d->Cin > 1) {
    int cb = (d->Cin + 127) / 128;
    int chunks = (ms_num_sms() + cb * p.taps - 1) / (cb * p.taps);
    int maxc = (M + 63) / 64;
    if (chunks > maxc) chunks = maxc;
    if (chunks < 1) chunks = 1;
    if (chunks > 1 && cudaMemsetAsync(dwf, 0, sizeof(float) * (size_t)p.taps * d->Cin * d->Cout, st) != cudaSuccess) return -1;
    dim3 grid((unsigned)cb, (unsigned)p.taps, (unsigned)chunks);
    if (d->Cout <= 8) wgrad_small_n<8><<<grid, 128, 0, st>>>(p, x, dy, dwf, M, chunks);
    else wgrad_small_n<NMAX><<<grid, 128, 0, st>>>(p, x, dy, dwf, M, chunks);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  return 0;
}
