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
      // four channel steps in flight per lane (few pixels, long K: the one-at-a-time loop was L2-latency bound, ~13 us for the
      // 16-pixel style scorer)
      for (int c0 = lane; c0 < p.Cin; c0 += 128) {
        float xv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) xv[u] = (c0 + 32 * u) < p.Cin ? __ldg(xr + c0 + 32 * u) : 0.f;
        float wv[4][N];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const float* wc = wr + (size_t)(c0 + 32 * u) * p.Cout;
#pragma unroll
          for (int n = 0; n < N; n++) wv[u][n] = (n < p.Cout && (c0 + 32 * u) < p.Cin) ? __ldg(wc + n) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int n = 0; n < N; n++) acc[n] = fmaf(xv[u], wv[u][n], acc[n]);
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

// ---- C_in = 1 forward, 8 output channels per thread (Cout % 8 == 0): the tap addresses are computed once per 8 outputs
// and every thread stores 32 (fp32) / 16 (bf16 plane) contiguous bytes, consecutive threads consecutive channels.
// FUSED = 0: y = conv + bias [LeakyReLU]  (training forward, z of the block)
// FUSED = 1: v = LeakyReLU(conv * scale[n] + shift[n]) (eval-mode BatchNorm folded, conv bias inside shift), written as fp32
//            (y nullable) and/or bf16 operand planes (hi, lo = bf16(v - hi) at + pstride; nullable) for the next layer.
template <int FUSED>
__global__ void __launch_bounds__(256) fwd_cin1_v8(P p, const float* __restrict__ x, const float* __restrict__ wf,
                                                   const float* __restrict__ bias, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, float* __restrict__ y,
                                                   __nv_bfloat16* __restrict__ planes, int pfmt, long long pstride, int act,
                                                   float slope, long long total8) {
  const int c8 = p.Cout >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int n0 = (int)(i % c8) * 8;
    const int pos = (int)(i / c8);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int tap = 0; tap < p.taps; tap++) {
      long long off;
      if (!in_pixel(p, pos, tap, off)) continue;
      const float xv = __ldg(x + off);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)tap * p.Cout + n0));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)tap * p.Cout + n0) + 1);
      acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
      acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float v = acc[j];
      if (FUSED) {
        v = fmaf(v, __ldg(scale + n0 + j), __ldg(shift + n0 + j));
        v = v > 0.f ? v : v * slope;
      } else {
        if (bias) v += __ldg(bias + n0 + j);
        if (act) v = v > 0.f ? v : v * slope;
      }
      acc[j] = v;
    }
    const long long o = (long long)pos * p.Cout + n0;
    if (y) {
      float4* d4 = reinterpret_cast<float4*>(y + o);
      d4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      d4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (FUSED && planes) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h2);
        acc[2 * j] -= __bfloat162float(h2.x);
        acc[2 * j + 1] -= __bfloat162float(h2.y);
      }
      *reinterpret_cast<uint4*>(planes + o) = make_uint4(w[0], w[1], w[2], w[3]);
      if (pfmt == MS_BF16X2) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
          w[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(planes + pstride + o) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
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

// ---- C_in = 1 wgrad, row-wise: a block owns whole output rows (b, ho), so the tap's input row and its validity are
// resolved once per row and the inner loop over wo is load-load-fma.  Same thread mapping (tap, n) and atomics as above.
__global__ void wgrad_cin1_rows(P p, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dwf,
                                int rows_total) {
  const int n = threadIdx.x % p.Cout, tap = threadIdx.x / p.Cout;
  const int th = tap / p.kw, tw = tap - th * p.kw;
  const int per = (rows_total + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows_total, r0 + per);
  float acc = 0.f;
  for (int row = r0; row < r1; row++) {
    const int b = row / p.Ho, ho = row - b * p.Ho;
    const int hi = ho * p.sh + th - p.ph;
    if ((unsigned)hi >= (unsigned)p.H) continue;
    const float* xr = x + ((size_t)b * p.H + hi) * p.W;
    const float* dr = dy + (size_t)row * p.Wo * p.Cout + n;
    // eight independent load pairs in flight per thread (the serial load-load-fma form was L2-latency bound: 80 us at batch 16)
    for (int wo = 0; wo < p.Wo; wo += 8) {
      float xv[8], dv[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int w = wo + j;
        const int wi = w * p.sw + tw - p.pw;
        const bool ok = w < p.Wo && (unsigned)wi < (unsigned)p.W;
        xv[j] = ok ? __ldg(xr + wi) : 0.f;
        dv[j] = ok ? __ldg(dr + (size_t)w * p.Cout) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) acc = fmaf(xv[j], dv[j], acc);
    }
  }
  atomicAdd(dwf + (size_t)tap * p.Cout + n, acc);
}

// ---- C_in = 1, 3x3, stride 1, pad 1, 64 output channels (audio_encoder.conv.0): dW[tap][n] = sum over pixels of
// x[pixel + tap] * dz[pixel][n].  A "stream" of 16 threads walks 16-pixel segments of output rows; thread cg of a stream
// owns four channels: per pixel ONE 16-byte load of dz (the 16 threads read the pixel's 256 contiguous bytes), three
// broadcast loads of the input column entering the sliding 3x3 window, 36 FMAs into register accumulators.  dz is read
// once (the row-wise kernel above reads it once per tap and spends ~10 instructions per FMA: 49 us at batch 16, 350 us at
// batch 128 -- on the critical path at the very end of backward).  Streams are combined in shared memory: 576 atomics
// per CTA.
constexpr int CIN1_SEG = 16;
__global__ void __launch_bounds__(256) wgrad_cin1_k3c64(int B, int H, int W, const float* __restrict__ x,
                                                        const float* __restrict__ dy, float* __restrict__ dwf) {
  __shared__ float s_part[8][16][36];
  const int cg = threadIdx.x & 15, stream = threadIdx.x >> 4;
  const int segs = (W + CIN1_SEG - 1) / CIN1_SEG;
  const long long units = (long long)B * H * segs;
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[t][j] = 0.f;
  for (long long u = (long long)blockIdx.x * 16 + stream; u < units; u += (long long)gridDim.x * 16) {
    const int row = (int)(u / segs), seg = (int)(u - (long long)row * segs);
    const int h = row % H;
    const int w0 = seg * CIN1_SEG, w1 = min(W, w0 + CIN1_SEG);
    const float* xr = x + (size_t)row * W;                      // input row h of image b (C_in = 1: (B, H, W))
    const bool up = h > 0, dn = h + 1 < H;
    const float* dr = dy + ((size_t)row * W) * 64 + 4 * cg;
    float xw[3][3];                                             // [input row h-1..h+1][column w-1..w+1]
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const bool okr = r == 1 || (r == 0 ? up : dn);
      const float* xp = xr + (r - 1) * W;
      xw[r][1] = (okr && w0 > 0) ? __ldg(xp + w0 - 1) : 0.f;
      xw[r][2] = okr ? __ldg(xp + w0) : 0.f;
    }
#pragma unroll 4
    for (int w = w0; w < w1; w++) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dr + (size_t)w * 64));
      const bool okc = w + 1 < W;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const bool okr = r == 1 || (r == 0 ? up : dn);
        xw[r][0] = xw[r][1];
        xw[r][1] = xw[r][2];
        xw[r][2] = (okr && okc) ? __ldg(xr + (r - 1) * W + w + 1) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float xv = xw[r][c];
          acc[r * 3 + c][0] = fmaf(xv, d.x, acc[r * 3 + c][0]);
          acc[r * 3 + c][1] = fmaf(xv, d.y, acc[r * 3 + c][1]);
          acc[r * 3 + c][2] = fmaf(xv, d.z, acc[r * 3 + c][2]);
          acc[r * 3 + c][3] = fmaf(xv, d.w, acc[r * 3 + c][3]);
        }
    }
  }
  // the two streams of a warp, then the eight warps
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 16) {
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
      for (int j = 0; j < 4; j++) s_part[warp][lane][t * 4 + j] = acc[t][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
    const int t = i >> 6, n = i & 63;
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) v += s_part[wv][n >> 2][t * 4 + (n & 3)];
    atomicAdd(dwf + (size_t)t * 64 + n, v);
  }
}

P make(const ms_conv_desc* d) {
  P p;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.kh = d->kh; p.kw = d->kw; p.sh = d->sh; p.sw = d->sw; p.ph = d->ph; p.pw = d->pw;
  p.Ho = d->Ho; p.Wo = d->Wo; p.taps = d->kh * d->kw;
  return p;
}


// ---- C_in = 1, 3x3, stride 1, pad 1 (audio_encoder.conv.0), row-wise: a thread owns 8 output channels of TWO adjacent
// output pixels; its 72 weights and 16 epilogue constants live in registers for the whole launch, the 3x4 input window is
// read once per pixel pair (L1 hits), and consecutive threads store consecutive 16/32-byte pieces of a pixel's channel
// vector.  ~14 instructions per output instead of ~33 for fwd_cin1_v8 (tap addressing + weight reloads per pixel).
// Requires 128 % (Cout/8) == 0 so that a thread's channel group never changes (128-thread blocks: ~136 registers per thread).
template <int FUSED>
__global__ void __launch_bounds__(128) fwd_cin1_k3(int B, int H, int W, int Cout, const float* __restrict__ x,
                                                   const float* __restrict__ wf, const float* __restrict__ bias,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   float* __restrict__ y, __nv_bfloat16* __restrict__ planes, int pfmt,
                                                   long long pstride, int act, float slope) {
  const int cgs = Cout >> 3;
  const int cg = threadIdx.x % cgs, n0 = cg * 8;
  const int wp0 = threadIdx.x / cgs, wp_step = blockDim.x / cgs;
  const int wpairs = (W + 1) >> 1;
  float wreg[9][8], sc[8], sh[8];
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(wf + (size_t)t * Cout + n0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(wf + (size_t)t * Cout + n0) + 1);
    wreg[t][0] = a.x; wreg[t][1] = a.y; wreg[t][2] = a.z; wreg[t][3] = a.w;
    wreg[t][4] = b.x; wreg[t][5] = b.y; wreg[t][6] = b.z; wreg[t][7] = b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (FUSED) { sc[j] = __ldg(scale + n0 + j); sh[j] = __ldg(shift + n0 + j); }
    else { sc[j] = 1.f; sh[j] = bias ? __ldg(bias + n0 + j) : 0.f; }
  }
  const float sl = (FUSED || act) ? slope : 1.f;
  const bool lrelu_max = sl >= 0.f && sl <= 1.f;          // LeakyReLU as max(v, slope * v)
  const int rows = B * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int h = row % H;
    const float* xr = x + (size_t)row * W;
    const bool up = h > 0, dn = h + 1 < H;
    for (int wp = wp0; wp < wpairs; wp += wp_step) {
      const int w = 2 * wp;
      float xv[3][4];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const bool rok = r == 0 ? up : (r == 2 ? dn : true);
        const float* xp = xr + (r - 1) * W + w;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int ww = w + c - 1;
          xv[r][c] = (rok && ww >= 0 && ww < W) ? __ldg(xp + c - 1) : 0.f;
        }
      }
      float a0[8], a1[8];
#pragma unroll
      for (int j = 0; j < 8; j++) { a0[j] = 0.f; a1[j] = 0.f; }
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int j = 0; j < 8; j++) {
            a0[j] = fmaf(xv[r][c], wreg[r * 3 + c][j], a0[j]);
            a1[j] = fmaf(xv[r][c + 1], wreg[r * 3 + c][j], a1[j]);
          }
#pragma unroll
      for (int px = 0; px < 2; px++) {
        if (w + px >= W) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float t = fmaf(px == 0 ? a0[j] : a1[j], sc[j], sh[j]);
          v[j] = lrelu_max ? fmaxf(t, t * sl) : (t > 0.f ? t : t * sl);
        }
        const long long o = ((long long)row * W + w + px) * Cout + n0;
        if (y) {
          float4* d4 = reinterpret_cast<float4*>(y + o);
          d4[0] = make_float4(v[0], v[1], v[2], v[3]);
          d4[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (FUSED && planes) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            v[2 * j] -= __bfloat162float(h2.x);
            v[2 * j + 1] -= __bfloat162float(h2.y);
          }
          *reinterpret_cast<uint4*>(planes + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (pfmt == MS_BF16X2) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(planes + pstride + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
  }
}

inline bool cin1_k3_ok(const ms_conv_desc* d) {
  const int cgs = d->Cout / 8;
  return d->Cin == 1 && d->groups == 1 && d->kh == 3 && d->kw == 3 && d->sh == 1 && d->sw == 1 && d->ph == 1 && d->pw == 1 &&
         d->Cout % 8 == 0 && cgs >= 1 && cgs <= 128 && 128 % cgs == 0 && d->Ho == d->H && d->Wo == d->W;
}

// ---- small-N forward for 1x1 convolutions over 256 channels (ClusterClassify.logits, layers.py:459): a lane owns 8
// consecutive channels, whose N x 8 weights stay in registers; per pixel it reads 32 contiguous bytes (the warp one 1 KB
// row), does 8N FMAs and joins the butterfly reduction.  wf: [0][c][n].
template <int N>
__global__ void __launch_bounds__(256) fwd_small_n_k1c256(int M, int Cout, const float* __restrict__ x,
                                                          const float* __restrict__ wf, const float* __restrict__ bias,
                                                          float* __restrict__ y, int act, float slope) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // the 256 x N weights reach the registers through shared memory: one coalesced pass per CTA (every lane fetching its own
  // 64 values from global memory cost 32 sectors per load instruction -- 67 us at 8192 pixels, one pixel per warp)
  __shared__ __align__(16) float s_w[32 * (8 * N + 4)];
  for (int i = threadIdx.x; i < 256 * N; i += blockDim.x) {
    const int c = i / N, n = i - c * N;
    s_w[(c >> 3) * (8 * N + 4) + (c & 7) * N + n] = n < Cout ? __ldg(wf + (size_t)c * Cout + n) : 0.f;
  }
  __syncthreads();
  float wreg[8][N];
#pragma unroll
  for (int c = 0; c < 8; c++)
#pragma unroll
    for (int n = 0; n < N; n++) wreg[c][n] = s_w[lane * (8 * N + 4) + c * N + n];
  for (int pos = warp; pos < M; pos += nwarps) {
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)pos * 256 + lane * 8);
    const float4 u = __ldg(xp), v = __ldg(xp + 1);
    const float xv[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
    float acc[N];
#pragma unroll
    for (int n = 0; n < N; n++) acc[n] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++)
#pragma unroll
      for (int n = 0; n < N; n++) acc[n] = fmaf(xv[c], wreg[c][n], acc[n]);
#pragma unroll
    for (int n = 0; n < N; n++) acc[n] = ms_warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < N; n++)
        if (n < Cout) {
          float t = acc[n] + (bias ? bias[n] : 0.f);
          if (act) t = t > 0.f ? t : t * slope;
          y[(size_t)pos * Cout + n] = t;
        }
    }
  }
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
  if (cin1_k3_ok(d) && (((uintptr_t)wf | (uintptr_t)y) & 15) == 0) {
    int blocks = d->B * d->H;
    if (blocks > ms_num_sms() * 24) blocks = ms_num_sms() * 24;
    fwd_cin1_k3<0><<<blocks, 128, 0, st>>>(d->B, d->H, d->W, d->Cout, x, wf, bias, nullptr, nullptr, y, nullptr, 0, 0, act, slope);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (d->Cin == 1 && d->Cout % 8 == 0 && (((uintptr_t)wf | (uintptr_t)y) & 15) == 0) {
    long long total8 = (long long)M * (d->Cout / 8);
    fwd_cin1_v8<0><<<blocks_for(total8, 256), 256, 0, st>>>(p, x, wf, bias, nullptr, nullptr, y, nullptr, 0, 0, act, slope, total8);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (d->Cin == 1 && d->Cout <= 1024) {
    long long total = (long long)M * d->Cout;
    fwd_cin1<<<blocks_for(total, 256), 256, 0, st>>>(p, x, wf, bias, y, act, slope, total);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (d->Cout <= 8 && d->Cin == 256 && p.taps == 1 && d->sh == 1 && d->sw == 1 && d->ph == 0 && d->pw == 0 &&
      ((uintptr_t)x & 15) == 0) {
    // >= 4 pixels per warp: the weight staging is paid once per CTA
    int blocks = (M + 31) / 32;
    if (blocks > ms_num_sms() * 4) blocks = ms_num_sms() * 4;
    fwd_small_n_k1c256<8><<<blocks, 256, 0, st>>>(M, d->Cout, x, wf, bias, y, act, slope);
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

extern "C" int ms_conv_cin1_bnact(const float* x, const float* wf, const float* scale, const float* shift, float slope,
                                  const ms_conv_desc* d, float* y, void* planes, int pfmt, int64_t pstride, void* stream) {
  using namespace small;
  if (!x || !wf || !scale || !shift || !d || (!y && !planes)) return MS_EINVAL;
  if (d->Cin != 1 || d->groups != 1 || d->Cout % 8) return MS_EINVAL;
  if ((((uintptr_t)wf | (uintptr_t)y | (uintptr_t)planes) & 15) != 0) return MS_EINVAL;
  if (planes && pfmt != MS_BF16 && pfmt != MS_BF16X2) return MS_EINVAL;
  if (planes && pfmt == MS_BF16X2 && (pstride <= 0 || pstride % 8)) return MS_EINVAL;
  if (cin1_k3_ok(d) && slope >= 0.f) {
    int blocks = d->B * d->H;
    if (blocks > ms_num_sms() * 24) blocks = ms_num_sms() * 24;
    fwd_cin1_k3<1><<<blocks, 128, 0, ms_stream(stream)>>>(d->B, d->H, d->W, d->Cout, x, wf, nullptr, scale, shift, y,
                                                          reinterpret_cast<__nv_bfloat16*>(planes), pfmt, pstride, 1, slope);
    MS_LAUNCH_CHECK();
    return 0;
  }
  P p = make(d);
  const long long total8 = (long long)d->B * d->Ho * d->Wo * (d->Cout / 8);
  fwd_cin1_v8<1><<<blocks_for(total8, 256), 256, 0, ms_stream(stream)>>>(p, x, wf, nullptr, scale, shift, y,
                                                                        reinterpret_cast<__nv_bfloat16*>(planes), pfmt, pstride, 1,
                                                                        slope, total8);
  MS_LAUNCH_CHECK();
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
    if (d->kh == 3 && d->kw == 3 && d->sh == 1 && d->sw == 1 && d->ph == 1 && d->pw == 1 && d->Cout == 64 &&
        !(((uintptr_t)dy) & 15)) {
      const long long units = (long long)d->B * d->H * ((d->W + CIN1_SEG - 1) / CIN1_SEG);
      long long blocks = (units + 15) / 16;
      if (blocks > (long long)ms_num_sms() * 6) blocks = (long long)ms_num_sms() * 6;
      wgrad_cin1_k3c64<<<(unsigned)blocks, 256, 0, st>>>(d->B, d->H, d->W, x, dy, dwf);
      return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    const int rows_total = d->B * d->Ho;
    int chunks = rows_total;
    if (chunks > ms_num_sms() * 8) chunks = ms_num_sms() * 8;
    wgrad_cin1_rows<<<chunks, p.taps * d->Cout, 0, st>>>(p, x, dy, dwf, rows_total);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (d->Cout <= NMAX && d->Cin > 1) {
    int cb = (d->Cin + 127) / 128;
    // every thread walks its chunk's pixels serially with dependent L2 loads (~0.4 us each): short chunks (>= 8 pixels) over
    // ~four waves of CTAs instead of 64-pixel chunks (25 us for the 1024-pixel classifier logits at batch 16)
    int chunks = (4 * ms_num_sms() + cb * p.taps - 1) / (cb * p.taps);
    int maxc = (M + 7) / 8;
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
