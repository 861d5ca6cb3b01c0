// The step immediately BEFORE the generator hot path, on the GPU: per batch the reference runs, on the
// host in fp64 (src/model/trainer.py:1290-1308),
//     labels = KMeans.predict(RemoveJoints(pose_raw))            src/data/transform.py:352-410, 463-510
//     y      = RemoveJoints(ZNorm(pose_raw))                     src/data/transform.py:221-226
// and then copies fp64 tensors to the device.  ms_pose_prepare does both in ONE pass over the raw pose batch:
//   * RemoveJoints: the (B,T,2,J) view loses the masked joints -> a column gather `cols` (P = 2*(J - |mask|) kept columns)
//   * ZNorm.znorm: std = sqrt(var * (var >= 0)), std == 0 -> eps, y = (x - mean) / std
//   * KMeans.get_feats: 'pose' x, 'velocity' x[t]-x[t-1] (0 at t=0), 'speed' sqrt(vx^2+vy^2) over the two coordinate halves,
//     'acceleration' v[t]-v[t-1]; concatenated in the order given
//   * KMeans.predict: squared distance to the K centres, first index of the minimum (torch.min tie rule), or the soft labels
//     softmax(-mse / mean(mse))
// One warp per frame, fp64 arithmetic like the reference; the K x D centres sit in shared memory.  HBM-bound: 8*Pr bytes read
// (the t-1 / t-2 rows come from L1/L2) + 8*P + 8 written per frame.
#include "common.cuh"

namespace {

constexpr int PREP_WARPS = 4;
constexpr int MAX_FEATS = 4;

struct PrepParams {
  int B, T, Pr, P, K, D, nfeats;
  int feats[MAX_FEATS];       // 1 pose, 2 velocity, 3 speed, 4 acceleration
  double eps;
};

__global__ void __launch_bounds__(PREP_WARPS * 32)
pose_prepare_kernel(const double* __restrict__ x, const double* __restrict__ mean, const double* __restrict__ var,
                    const int* __restrict__ cols, const double* __restrict__ centers, double* __restrict__ y,
                    long long* __restrict__ labels, double* __restrict__ soft, PrepParams p) {
  extern __shared__ double sm[];
  double* s_cent = sm;                                   // K * D
  double* s_warp = sm + (size_t)p.K * p.D;               // per warp: x[P], v[P], a[P], f[D], mse[K]
  const int per_warp = 3 * p.P + p.D + p.K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sx = s_warp + (size_t)warp * per_warp;
  double* sv = sx + p.P;
  double* sa = sv + p.P;
  double* sf = sa + p.P;
  double* smse = sf + p.D;
  if (centers)
    for (int i = threadIdx.x; i < p.K * p.D; i += blockDim.x) s_cent[i] = centers[i];
  __syncthreads();
  const long long frames = (long long)p.B * p.T;
  for (long long fr = (long long)blockIdx.x * PREP_WARPS + warp; fr < frames; fr += (long long)gridDim.x * PREP_WARPS) {
    const int t = (int)(fr % p.T);
    const double* xr = x + fr * p.Pr;
    for (int c = lane; c < p.P; c += 32) {
      const int col = cols[c];
      const double x0 = xr[col];
      const double x1 = t > 0 ? xr[col - p.Pr] : 0.0;
      const double x2 = t > 1 ? xr[col - 2 * p.Pr] : 0.0;
      const double v0 = t > 0 ? x0 - x1 : 0.0;                      // pose_v[:, 1:] = x[:, 1:] - x[:, :-1]; pose_v[:, 0] = 0
      const double v1 = t > 1 ? x1 - x2 : 0.0;
      sx[c] = x0;
      sv[c] = v0;
      sa[c] = t > 0 ? v0 - v1 : 0.0;                                // pose_a[:, 1:] = pose_v[:, 1:] - pose_v[:, :-1]
      if (y) {
        const double vr = var[col];
        double sd = sqrt(vr >= 0.0 ? vr : 0.0);                     // (muvar[1] * mask_std) ** 0.5
        if (sd == 0.0) sd = p.eps;
        y[fr * p.P + c] = (x0 - mean[col]) / sd;
      }
    }
    if (!labels && !soft) continue;
    __syncwarp();
    int off = 0;
    for (int i = 0; i < p.nfeats; i++) {
      const int f = p.feats[i];
      if (f == 3) {
        const int h = p.P / 2;
        for (int j = lane; j < h; j += 32) sf[off + j] = sqrt(sv[j] * sv[j] + sv[j + h] * sv[j + h]);
        off += h;
      } else {
        const double* src = f == 1 ? sx : (f == 2 ? sv : sa);
        for (int c = lane; c < p.P; c += 32) sf[off + c] = src[c];
        off += p.P;
      }
    }
    __syncwarp();
    for (int k = 0; k < p.K; k++) {
      const double* ck = s_cent + (size_t)k * p.D;
      double acc = 0.0;
      for (int d = lane; d < p.D; d += 32) {
        const double df = ck[d] - sf[d];
        acc += df * df;
      }
      acc = ms_warp_sum_d(acc);
      if (lane == 0) smse[k] = acc;
    }
    __syncwarp();
    if (lane == 0) {
      int best = 0;
      double bv = smse[0], tot = smse[0];
      for (int k = 1; k < p.K; k++) {
        tot += smse[k];
        if (smse[k] < bv) { bv = smse[k]; best = k; }               // first index of the minimum
      }
      if (labels) labels[fr] = best;
      if (soft) {
        // softmax(-mse / mse.mean(-1)) (transform.py:403)
        const double m = tot / p.K;
        double mx = -smse[0] / m;
        for (int k = 1; k < p.K; k++) mx = fmax(mx, -smse[k] / m);
        double z = 0.0;
        for (int k = 0; k < p.K; k++) z += exp(-smse[k] / m - mx);
        for (int k = 0; k < p.K; k++) soft[fr * p.K + k] = exp(-smse[k] / m - mx) / z;
      }
    }
    __syncwarp();
  }
}

// ZNorm.inv_znorm (transform.py:228-229): x * var**0.5 + mean, elementwise over the last dimension
__global__ void inv_znorm_kernel(const double* __restrict__ x, const double* __restrict__ mean, const double* __restrict__ var,
                                 long long rows, int C, double* __restrict__ out) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    out[i] = __dadd_rn(__dmul_rn(x[i], sqrt(var[c])), mean[c]);      // separate multiply and add, as torch evaluates it (no FMA)
  }
}

}  // namespace

extern "C" int ms_pose_prepare(const double* x, const double* mean, const double* var, const int32_t* cols, const double* centers,
                               int B, int T, int Pr, int P, int K, const int32_t* feats_host, int nfeats, double eps, double* y,
                               int64_t* labels, double* soft, void* stream) {
  if (!x || !cols || B < 1 || T < 1 || Pr < 1 || P < 1 || P > Pr) return MS_EINVAL;
  if (y && (!mean || !var)) return MS_EINVAL;
  const bool want_labels = labels || soft;
  PrepParams p;
  p.B = B; p.T = T; p.Pr = Pr; p.P = P; p.K = want_labels ? K : 0; p.D = 0; p.nfeats = 0; p.eps = eps;
  for (int i = 0; i < MAX_FEATS; i++) p.feats[i] = 0;
  if (want_labels) {
    if (!centers || !feats_host || K < 1 || nfeats < 1 || nfeats > MAX_FEATS) return MS_EINVAL;
    for (int i = 0; i < nfeats; i++) {
      const int f = feats_host[i];
      if (f < 1 || f > 4) return MS_EINVAL;
      if (f == 3 && (P & 1)) return MS_EINVAL;
      p.feats[i] = f;
      p.D += f == 3 ? P / 2 : P;
    }
    p.nfeats = nfeats;
  }
  const size_t smem = sizeof(double) * ((size_t)p.K * p.D + (size_t)PREP_WARPS * (3 * P + p.D + p.K));
  if (smem > 200 * 1024) return MS_EINVAL;
  if (smem > 48 * 1024) MS_CUDA(cudaFuncSetAttribute(pose_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = ((long long)B * T + PREP_WARPS - 1) / PREP_WARPS;
  const long long cap = (long long)ms_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pose_prepare_kernel<<<(unsigned)blocks, PREP_WARPS * 32, smem, ms_stream(stream)>>>(
      x, mean, var, cols, want_labels ? centers : nullptr, y, reinterpret_cast<long long*>(labels), soft, p);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_inv_znorm(const double* x, const double* mean, const double* var, int64_t rows, int C, double* out, void* stream) {
  if (!x || !mean || !var || !out || rows < 1 || C < 1) return MS_EINVAL;
  long long blocks = (rows * C + 255) / 256;
  const long long cap = (long long)ms_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  inv_znorm_kernel<<<(unsigned)blocks, 256, 0, ms_stream(stream)>>>(x, mean, var, rows, C, out);
  MS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Evaluation metrics on the device: what TrainerBase.calculate_metrics (src/model/trainer.py:865-907)
// computes per batch on the host after a D2H copy -- L1 and VelL1 on the normalised full-width poses
// (src/evaluation/metrics.py:94-131), PCK on the un-normalised, root-centred frames (metrics.py:247-303) -- in ONE pass
// over (y_cap, y_gt); only 2 doubles + 2*J counters leave the GPU.  One warp per frame, lanes over joints.
//   acc[0] = sum |y - gt| over kept joints, acc[1] = sum |vel(y) - vel(gt)| over kept joints (t >= 1)
//   cnt[a * J + j] = number of frames with dist_j < alpha_a * max(h, w)   (exact integer counts)
// ---------------------------------------------------------------------------------------------
namespace {

struct MetricParams {
  int B, T, J, nalpha;
  double alpha[4];
};

__global__ void __launch_bounds__(128)
pose_metrics_kernel(const double* __restrict__ y, const double* __restrict__ gt, const double* __restrict__ mean,
                    const double* __restrict__ var, const unsigned char* __restrict__ keep, double* __restrict__ acc,
                    unsigned long long* __restrict__ cnt, MetricParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = 2 * p.J;
  const long long frames = (long long)p.B * p.T;
  double l1 = 0.0, vl = 0.0;
  for (long long fr = (long long)blockIdx.x * 4 + warp; fr < frames; fr += (long long)gridDim.x * 4) {
    const int t = (int)(fr % p.T);
    const double* yr = y + fr * W;
    const double* gr = gt + fr * W;
    double gx_min = 1e300, gx_max = -1e300, gy_min = 1e300, gy_max = -1e300;
    for (int j = lane; j < p.J; j += 32) {
      const double yx = yr[j], yy = yr[p.J + j], gx = gr[j], gy = gr[p.J + j];
      if (keep[j]) {
        l1 += fabs(yx - gx) + fabs(yy - gy);
        if (t > 0) {
          vl += fabs((yx - yr[j - W]) - (gx - gr[j - W])) + fabs((yy - yr[p.J + j - W]) - (gy - gr[p.J + j - W]));
        }
      }
      // PCK threshold box of the ground truth: un-normalised (inv_znorm), root joint at (0,0)
      const double ugx = j == 0 ? 0.0 : __dadd_rn(__dmul_rn(gx, sqrt(var[j])), mean[j]);
      const double ugy = j == 0 ? 0.0 : __dadd_rn(__dmul_rn(gy, sqrt(var[p.J + j])), mean[p.J + j]);
      gx_min = fmin(gx_min, ugx); gx_max = fmax(gx_max, ugx);
      gy_min = fmin(gy_min, ugy); gy_max = fmax(gy_max, ugy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gx_min = fmin(gx_min, __shfl_xor_sync(0xffffffffu, gx_min, o));
      gx_max = fmax(gx_max, __shfl_xor_sync(0xffffffffu, gx_max, o));
      gy_min = fmin(gy_min, __shfl_xor_sync(0xffffffffu, gy_min, o));
      gy_max = fmax(gy_max, __shfl_xor_sync(0xffffffffu, gy_max, o));
    }
    const double box = fmax(gx_max - gx_min, gy_max - gy_min);      // max(h, w) (metrics.py:276-279)
    for (int j = lane; j < p.J; j += 32) {
      double dx = 0.0, dy = 0.0;
      if (j != 0) {
        const double sx = sqrt(var[j]), sy = sqrt(var[p.J + j]);
        dx = __dadd_rn(__dmul_rn(yr[j], sx), mean[j]) - __dadd_rn(__dmul_rn(gr[j], sx), mean[j]);
        dy = __dadd_rn(__dmul_rn(yr[p.J + j], sy), mean[p.J + j]) - __dadd_rn(__dmul_rn(gr[p.J + j], sy), mean[p.J + j]);
      }
      const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      for (int a = 0; a < p.nalpha; a++)
        if (dist < p.alpha[a] * box) atomicAdd(cnt + (size_t)a * p.J + j, 1ull);
    }
  }
  l1 = ms_warp_sum_d(l1);
  vl = ms_warp_sum_d(vl);
  if (lane == 0) {
    atomicAdd(acc, l1);
    atomicAdd(acc + 1, vl);
  }
}

}  // namespace

extern "C" int ms_pose_metrics(const double* y, const double* gt, const double* mean, const double* var, const uint8_t* keep,
                               int B, int T, int J, const double* alphas_host, int nalpha, double* acc, uint64_t* cnt, void* stream) {
  if (!y || !gt || !mean || !var || !keep || !acc || !cnt || !alphas_host) return MS_EINVAL;
  if (B < 1 || T < 1 || J < 1 || nalpha < 1 || nalpha > 4) return MS_EINVAL;
  MetricParams p;
  p.B = B; p.T = T; p.J = J; p.nalpha = nalpha;
  for (int i = 0; i < 4; i++) p.alpha[i] = i < nalpha ? alphas_host[i] : 0.0;
  MS_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), ms_stream(stream)));
  MS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint64_t) * (size_t)nalpha * J, ms_stream(stream)));
  long long blocks = ((long long)B * T + 3) / 4;
  const long long cap = (long long)ms_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pose_metrics_kernel<<<(unsigned)blocks, 128, 0, ms_stream(stream)>>>(y, gt, mean, var, keep, acc,
                                                                       reinterpret_cast<unsigned long long*>(cnt), p);
  MS_LAUNCH_CHECK();
  return 0;
}
