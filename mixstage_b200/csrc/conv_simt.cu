// fp32 implicit-GEMM convolution on CUDA cores: forward, input-gradient and
// weight-gradient of a grouped NHWC conv2d (conv1d = H=1).  This is the exact-fp32
// path ("precision=fp32") and the fallback for the shapes the tcgen05 path does not
// take (C_in=1, N<16, L<8 ...).  One 64x64x16 tile per CTA, 4x4 outputs per thread.
//
// Replaces: aten::conv1d/conv2d + convolution_backward reached from
// ConvNormRelu.forward (reference layers.py:78), Speech2Gesture_D (speech2gesture.py:92-100),
// the grouped 1x1 `logits` conv (joint_late_cluster_soft_style.py:83,193).
#include "common.cuh"

// csrc/conv_small.cu: layers that are not GEMM-shaped (N <= 32 output channels, or C_in = 1).  1 = launched there.
int ms_small_conv_fwd(const float* x, const float* wf, const float* bias, float* y, const ms_conv_desc* d, int act,
                      float slope, cudaStream_t st);
int ms_small_conv_dgrad(const float* dy, const float* wt, float* dx, const ms_conv_desc* d, cudaStream_t st);
int ms_small_conv_wgrad(const float* x, const float* dy, float* dwf, const ms_conv_desc* d, cudaStream_t st);

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvP {
  int B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo;
  int Cin_g, Cout_g, taps;
};

enum { FWD = 0, DGRAD = 1, WGRAD = 2 };

// x element feeding output position `pos` (= (b*Ho+ho)*Wo+wo) through tap `tap`, channel c of group g
__device__ __forceinline__ float fetch_x(const ConvP& p, const float* __restrict__ X, int g, int pos, int tap, int c) {
  int wo = pos % p.Wo;
  int t = pos / p.Wo;
  int ho = t % p.Ho;
  int b = t / p.Ho;
  int th = tap / p.kw, tw = tap - th * p.kw;
  int hi = ho * p.sh + th - p.ph;
  int wi = wo * p.sw + tw - p.pw;
  if ((unsigned)hi >= (unsigned)p.H || (unsigned)wi >= (unsigned)p.W) return 0.f;
  return __ldg(X + ((size_t)(b * p.H + hi) * p.W + wi) * p.Cin + g * p.Cin_g + c);
}

// dy element that input position `ipos` (= (b*H+h)*W+w) receives through tap `tap`, out-channel n of group g
__device__ __forceinline__ float fetch_dy(const ConvP& p, const float* __restrict__ DY, int g, int ipos, int tap, int n) {
  int w = ipos % p.W;
  int t = ipos / p.W;
  int h = t % p.H;
  int b = t / p.H;
  int th = tap / p.kw, tw = tap - th * p.kw;
  int hh = h + p.ph - th;
  int ww = w + p.pw - tw;
  if (hh < 0 || ww < 0) return 0.f;
  int ho = hh / p.sh, wo = ww / p.sw;
  if (ho * p.sh != hh || wo * p.sw != ww || ho >= p.Ho || wo >= p.Wo) return 0.f;
  return __ldg(DY + ((size_t)(b * p.Ho + ho) * p.Wo + wo) * p.Cout + g * p.Cout_g + n);
}

template <int MODE>
__global__ void __launch_bounds__(NT) conv_gemm_simt(ConvP p, const float* __restrict__ A, const float* __restrict__ Bm,
                                                      const float* __restrict__ bias, float* __restrict__ C,
                                                      int act, float slope, int splitk, int M, int N, int K) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int g = blockIdx.z / splitk, ks = blockIdx.z % splitk;
  // K range of this split (multiple of BK)
  int kchunk = (int)((((int64_t)K + splitk - 1) / splitk + BK - 1) / BK) * BK;
  int kbeg = ks * kchunk, kend = min(K, kbeg + kchunk);
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  const int chan_k = (MODE == FWD) ? p.Cin_g : p.Cout_g;   // channels per tap along K (FWD/DGRAD)

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * NT;
      int r, kk;
      if (MODE == WGRAD) { r = e % BM; kk = e / BM; } else { r = e / BK; kk = e % BK; }
      int m = m0 + r, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < kend) {
        if (MODE == FWD) { int tap = k / chan_k; v = fetch_x(p, A, g, m, tap, k - tap * chan_k); }
        else if (MODE == DGRAD) { int tap = k / chan_k; v = fetch_dy(p, A, g, m, tap, k - tap * chan_k); }
        else { int tap = m / p.Cin_g; v = fetch_x(p, A, g, k, tap, m - tap * p.Cin_g); }
      }
      As[kk][r] = v;
    }
    // ---- B tile (always N-contiguous)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * NT;
      int kk = e / BN, c = e % BN;
      int k = k0 + kk, n = n0 + c;
      float v = 0.f;
      if (k < kend && n < N) {
        if (MODE == WGRAD) v = __ldg(Bm + (size_t)k * p.Cout + g * p.Cout_g + n);   // dy[pos, g*N+n]
        else v = __ldg(Bm + ((size_t)g * K + k) * N + n);                            // packed weight [g][k][n]
      }
      Bs[kk][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (MODE == FWD) {
        if (bias) v += bias[g * N + n];
        if (act) v = v > 0.f ? v : v * slope;
        C[(size_t)m * p.Cout + g * N + n] = v;
      } else if (MODE == DGRAD) {
        C[(size_t)m * p.Cin + g * N + n] = v;
      } else {
        float* dst = C + ((size_t)g * M + m) * N + n;
        if (splitk > 1) atomicAdd(dst, v); else *dst = v;
      }
    }
  }
}

ConvP make_p(const ms_conv_desc* d) {
  ConvP p;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.kh = d->kh; p.kw = d->kw; p.sh = d->sh; p.sw = d->sw; p.ph = d->ph; p.pw = d->pw;
  p.groups = d->groups; p.Ho = d->Ho; p.Wo = d->Wo;
  p.Cin_g = d->Cin / d->groups; p.Cout_g = d->Cout / d->groups; p.taps = d->kh * d->kw;
  return p;
}

bool valid(const ms_conv_desc* d) {
  if (!d || d->groups < 1 || d->Cin % d->groups || d->Cout % d->groups) return false;
  if (d->B < 1 || d->H < 1 || d->W < 1 || d->kh < 1 || d->kw < 1 || d->sh < 1 || d->sw < 1) return false;
  if (d->Ho != (d->H + 2 * d->ph - d->kh) / d->sh + 1) return false;
  if (d->Wo != (d->W + 2 * d->pw - d->kw) / d->sw + 1) return false;
  if (d->Ho < 1 || d->Wo < 1) return false;
  if ((int64_t)d->B * d->H * d->W >= (1ll << 31) || (int64_t)d->B * d->Ho * d->Wo >= (1ll << 31)) return false;
  return true;
}

}  // namespace

extern "C" int ms_conv_fwd_f32(const float* x, const float* wf, const float* bias, float* y,
                               const ms_conv_desc* d, int act, float slope, void* stream) {
  if (!valid(d) || !x || !wf || !y) return MS_EINVAL;
  if (int r = ms_small_conv_fwd(x, wf, bias, y, d, act, slope, ms_stream(stream))) return r == 1 ? 0 : MS_EINVAL;
  ConvP p = make_p(d);
  int M = d->B * d->Ho * d->Wo, N = p.Cout_g, K = p.taps * p.Cin_g;
  dim3 grid((unsigned)ms_cdiv(M, BM), (unsigned)ms_cdiv(N, BN), (unsigned)d->groups);
  conv_gemm_simt<FWD><<<grid, NT, 0, ms_stream(stream)>>>(p, x, wf, bias, y, act, slope, 1, M, N, K);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_conv_dgrad_f32(const float* dy, const float* wt, float* dx, const ms_conv_desc* d, void* stream) {
  if (!valid(d) || !dy || !wt || !dx) return MS_EINVAL;
  if (int r = ms_small_conv_dgrad(dy, wt, dx, d, ms_stream(stream))) return r == 1 ? 0 : MS_EINVAL;
  ConvP p = make_p(d);
  int M = d->B * d->H * d->W, N = p.Cin_g, K = p.taps * p.Cout_g;
  dim3 grid((unsigned)ms_cdiv(M, BM), (unsigned)ms_cdiv(N, BN), (unsigned)d->groups);
  conv_gemm_simt<DGRAD><<<grid, NT, 0, ms_stream(stream)>>>(p, dy, wt, nullptr, dx, 0, 1.f, 1, M, N, K);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_conv_wgrad_f32(const float* x, const float* dy, float* dwf, const ms_conv_desc* d, void* stream) {
  if (!valid(d) || !x || !dy || !dwf) return MS_EINVAL;
  if (int r = ms_small_conv_wgrad(x, dy, dwf, d, ms_stream(stream))) return r == 1 ? 0 : MS_EINVAL;
  ConvP p = make_p(d);
  int M = p.taps * p.Cin_g, N = p.Cout_g, K = d->B * d->Ho * d->Wo;
  int64_t tiles = ms_cdiv(M, BM) * ms_cdiv(N, BN) * d->groups;
  int splitk = (int)ms_cdiv(2 * ms_num_sms(), tiles);
  int maxsplit = (int)ms_cdiv(K, 4 * BK);
  if (splitk > maxsplit) splitk = maxsplit;
  if (splitk < 1) splitk = 1;
  if ((int64_t)d->groups * splitk > 65535) splitk = 65535 / d->groups;
  if (splitk > 1) MS_CUDA(cudaMemsetAsync(dwf, 0, sizeof(float) * (size_t)d->groups * M * N, ms_stream(stream)));
  dim3 grid((unsigned)ms_cdiv(M, BM), (unsigned)ms_cdiv(N, BN), (unsigned)(d->groups * splitk));
  conv_gemm_simt<WGRAD><<<grid, NT, 0, ms_stream(stream)>>>(p, x, dy, nullptr, dwf, 0, 1.f, splitk, M, N, K);
  MS_LAUNCH_CHECK();
  return 0;
}
