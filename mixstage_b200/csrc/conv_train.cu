// Fused TRAINING blocks of ConvNormRelu (reference src/model/layers.py:32-78): one launch per block and direction.
//
// A training-mode block is  z = conv(x)  ->  batch statistics of z  ->  y = LeakyReLU(BN(z)) [upsample x2 + skip]
// and needs a grid-wide reduction in the middle, so the round-1 path ran it as three dependent kernels forward
// (GEMM, statistics + finalize, normalise) and three backward (reduce, apply, input-gradient GEMM), each 3-25 us at
// batch 16: the step was ~300 dependent launches deep.  Here one PERSISTENT launch of at most one CTA per SM walks the
// phases with device-wide barriers between them (all CTAs are co-resident: cooperative launch, grid <= #SMs):
//
//   forward   [GEMM tiles: TMA -> tcgen05.mma -> TMEM -> z (split-K slices combine with red.global.add.v4.f32)]
//             | barrier | per-channel sum / sum-of-squares of z (fp64 atomics) | barrier |
//             finalize (scale/shift/mean/rstd, running statistics, batch counter) + normalise + LeakyReLU (+ UNet
//             upsample x2 + skip) -> fp32 activation and the next GEMM's bf16 operand planes (hi [, lo])
//   backward  per-channel reductions of dy*act' and dy*act'*xhat | barrier | dz = BN-backward(dy) -> bf16 operand planes,
//             affine-parameter gradients into the flat gradient buffer | barrier |
//             [input-gradient GEMM tiles reading those planes through TMA -> dx]
//
// The GEMM phase is the one-tile-per-work-item tcgen05 pipeline of conv_tc.cu (same descriptors, same tap tables, same
// split-bf16 passes) run as a loop over (tile, k-slice) work items with one TMEM accumulator; the element-wise phases are
// the arithmetic of elementwise.cu's BatchNorm kernels, statement for statement, so results agree with the unfused path
// to the order of the fp32 split-K reductions.  z stays in L2 between the phases at the batch sizes this path serves.
//
// Also here: the weight gradient accumulated in place (red.global.add.v4.f32 into ONE persistent fp32 accumulator per
// weight, instead of one partial per pixel slice summed by a second kernel) and the table-driven kernel that converts
// every accumulator of a sub-network into its flat gradient buffer in one launch.
#include <cstring>
#include "tc_common.cuh"

namespace {

constexpr int TB_THREADS = 256;       // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: TMEM epilogue; all 8: element-wise phases
constexpr int TB_STAGES = 4;
constexpr uint32_t TB_B_STAGE_BYTES = 256 * BLOCK_K * 2;                 // room for the widest weight tile (32 KB)
constexpr uint32_t TB_STAGE_BYTES = A_STAGE_BYTES + TB_B_STAGE_BYTES;      // 48 KB
constexpr uint32_t TB_RING_BYTES = TB_STAGES * TB_STAGE_BYTES;            // 192 KB, reused as scratch by the element-wise phases
constexpr uint32_t GRID_SPIN_LIMIT = 1u << 24;

// BatchNorm side of a block
struct BnParams {
  int C, pdt, training;
  float momentum, eps, slope;
  const void* gamma;
  const void* beta;
  const void* cbias;          // conv bias (the GEMM output excludes it): enters the running mean only
  void* rmean;
  void* rvar;
  long long* nbt;
  double* sums;               // [2][C] zero-filled: sum, sum of squares (forward) / dgamma, dbeta (backward)
  float* ss;                  // [4][C]: scale, shift, mean, rstd (written forward, read backward)
};

struct FwdIO {
  float* z;                   // (rows, C) fp32 GEMM output; zero-filled by the caller when split_k > 1
  float* y;                   // nullable (rows_out, C) fp32 activation
  __nv_bfloat16* planes;      // nullable operand planes of the activation
  int pfmt;
  long long pstride;
  const float* res;           // up2: skip tensor laid out like y (fp32), or
  const __nv_bfloat16* res_pl;   // ... as operand planes
  int res_fmt;
  long long res_ps;
  int up2, L;                 // L = GEMM rows per sequence (1-D)
  long long rows;             // GEMM rows
  unsigned int* sync;         // zero-filled barrier counter
};

struct BwdIO {
  const float* dy;            // (rows_out, C)
  const float* z;             // (rows, C)
  __nv_bfloat16* dzp;         // operand planes of dz, row stride C
  int pfmt;
  long long pstride;
  int up2, L;
  long long rows;
  void* ggamma;               // nullable: += dgamma / dbeta in dtype gdt
  void* gbeta;
  int gdt;
  float* dx;                  // nullable: input gradient (zero-filled by the caller when split_k > 1)
  int has_gemm;
  unsigned int* sync;
};

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int spins = 0, v;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > GRID_SPIN_LIMIT) __trap();       // a CTA that never arrives must not hang the GPU
      __nanosleep(40);
    }
    __threadfence();
  }
  __syncthreads();
}

// Pipeline state that survives from one GEMM phase to the next inside a launch (ring slot / parity per role thread,
// accumulator hand-over count)
struct PipeState {
  int s;
  uint32_t ph;
  uint32_t li;
};

struct GemmSmem {
  uint8_t* a;
  uint8_t* b;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;
  uint64_t* tempty;
};

// One GEMM phase: work items (tile, k-slice) it = blockIdx.x, blockIdx.x + gridDim.x, ...; result into `out` (fp32) by plain
// 16-byte stores (split_k == 1) or vector reductions (split_k > 1, `out` zero-filled).  No bias, no activation.
__device__ __forceinline__ void gemm_phase(const CUtensorMap* map_a, const CUtensorMap* map_w, const CUtensorMap* map_a_lo,
                                           const CUtensorMap* map_w_lo, const IgemmParams& p, float* __restrict__ out,
                                           const GemmSmem& sm, uint32_t tmem_base, PipeState& st) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
  const int ny = p.n_tiles_per_class * p.num_classes;
  const int tiles = p.tiles_w * p.tiles_h * p.tiles_b * ny;
  const int items = tiles * p.split_k;
  const int num_k_total = p.ntaps * p.cchunks * p.npass;
  const int k_per = (num_k_total + p.split_k - 1) / p.split_k;
  if (warp == 0) {
    if (lane == 0) {
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int slice = it % p.split_k, tile = it / p.split_k;
        const int y = tile % ny;
        int mt = tile / ny;
        const int tw = mt % p.tiles_w; mt /= p.tiles_w;
        const int th = mt % p.tiles_h; mt /= p.tiles_h;
        const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = mt * p.box_b;
        const int cls = y / p.n_tiles_per_class;
        const int n0 = (y - cls * p.n_tiles_per_class) * p.block_n;
        const int tap_base = p.shared_taps ? 0 : cls * p.ntaps;
        const int chan_base = p.a_chan_base[cls];
        const int wrow = cls * p.class_n + n0;
        const int k_beg = slice * k_per;
        const int k_end = min(num_k_total, k_beg + k_per);
        for (int kg = k_beg; kg < k_end; kg++) {
          mbar_wait(&sm.empty[st.s], st.ph ^ 1u);
          const int kk = kg / p.npass, pass = kg - kk * p.npass;          // split-bf16: hi*hi, hi*lo, lo*hi
          const int tap = kk / p.cchunks, cc = kk - tap * p.cchunks;
          const short* t = p.taps[tap_base + tap];
          mbar_expect_tx(&sm.full[st.s], A_STAGE_BYTES + b_stage_bytes);
          tma_load_5d(pass == 2 ? map_a_lo : map_a, &sm.full[st.s], sm.a + (size_t)st.s * A_STAGE_BYTES,
                      chan_base + t[0] + cc * BLOCK_K, w0 + t[1], t[2], h0 + t[3], b0);
          tma_load_2d(pass == 1 ? map_w_lo : map_w, &sm.full[st.s], sm.b + (size_t)st.s * TB_B_STAGE_BYTES, kk * BLOCK_K, wrow);
          if (++st.s == TB_STAGES) { st.s = 0; st.ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int slice = it % p.split_k;
        const int k_beg = slice * k_per;
        const int num_k = min(num_k_total, k_beg + k_per) - k_beg;
        mbar_wait(sm.tempty, (st.li & 1u) ^ 1u);         // the epilogue warps drained the previous item's accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < num_k; ks++) {
          mbar_wait(&sm.full[st.s], st.ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(sm.a + (size_t)st.s * A_STAGE_BYTES));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(sm.b + (size_t)st.s * TB_B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++)
            umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit(&sm.empty[st.s]);
          if (++st.s == TB_STAGES) { st.s = 0; st.ph ^= 1u; }
        }
        umma_commit(sm.tfull);
        st.li++;
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // tile row == TMEM lane
    const int wi = r % p.box_w;
    const int hi = (r / p.box_w) % p.box_h;
    const int bi = r / (p.box_w * p.box_h);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int tile = it / p.split_k;
      const int y = tile % ny;
      int mt = tile / ny;
      const int tw = mt % p.tiles_w; mt /= p.tiles_w;
      const int th = mt % p.tiles_h; mt /= p.tiles_h;
      const int cls = y / p.n_tiles_per_class;
      const int n0 = (y - cls * p.n_tiles_per_class) * p.block_n;
      const int ow = tw * p.box_w + wi, oh = th * p.box_h + hi, ob = mt * p.box_b + bi;
      const bool valid = ow < p.out_w && oh < p.out_h && ob < p.out_b;
      float* dst = out + (long long)ob * p.os_b + (long long)oh * p.os_h + (long long)ow * p.os_w + p.out_off[cls] + n0;
      mbar_wait(sm.tfull, st.li & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_wait_ld16(v);
        if (valid && (n0 + c0) < p.class_n) {
          if (p.split_k > 1) {
#pragma unroll
            for (int j = 0; j < 4; j++)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j), "f"(__uint_as_float(v[4 * j])),
                           "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
              *reinterpret_cast<float4*>(dst + c0 + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.tempty);
      st.li++;
    }
  }
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// rows [r0, r1) of this CTA out of `rows`
__device__ __forceinline__ void cta_rows(long long rows, long long& r0, long long& r1) {
  r0 = rows * (long long)blockIdx.x / (long long)gridDim.x;
  r1 = rows * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
}

// Column reduction helper: every thread owns one 4-channel group cg and one row lane rl; eight fp64 partials per thread
// are combined over the row lanes through `scratch` and added to dst_a[4cg..] / dst_b[4cg..] with fp64 atomics.
__device__ __forceinline__ void reduce_lanes_and_add(double (&acc)[8], int cg, int rl, int RL, int ncg, bool active, double* scratch,
                                                     double* __restrict__ dst_a, double* __restrict__ dst_b) {
  if (RL > 1) {
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; j++) scratch[((size_t)rl * ncg + cg) * 8 + j] = acc[j];
    }
    __syncthreads();
    if (active && rl == 0) {
      for (int l = 1; l < RL; l++) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += scratch[((size_t)l * ncg + cg) * 8 + j];
      }
    }
    __syncthreads();
  }
  if (active && rl == 0) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      atomicAdd(dst_a + 4 * cg + j, acc[j]);
      atomicAdd(dst_b + 4 * cg + j, acc[4 + j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward block
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS, 1)
conv_block_train_fwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                            const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_w_lo,
                            const __grid_constant__ IgemmParams p, const __grid_constant__ BnParams bn,
                            const __grid_constant__ FwdIO io) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[TB_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TB_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar, tempty_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.block_n) tmem_cols <<= 1;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (p.npass > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_lo)) : "memory");
    }
    for (int s = 0; s < TB_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar, 1);
    mbar_init(&tempty_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  GemmSmem sm;
  sm.a = smem; sm.b = smem + TB_STAGES * A_STAGE_BYTES;
  sm.full = full_bar; sm.empty = empty_bar; sm.tfull = &tfull_bar; sm.tempty = &tempty_bar;
  PipeState st;
  st.s = 0; st.ph = 0; st.li = 0;

  // ---- phase 1: z = conv(x) on the tensor cores
  gemm_phase(&map_a, &map_w, &map_a_lo, &map_w_lo, p, io.z, sm, tmem_base, st);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  grid_barrier(io.sync, gridDim.x);

  // ---- phase 2: per-channel sum and sum of squares of z over this CTA's rows (elementwise.cu: bn_stats_finalize_kernel)
  const int C = bn.C, ncg = C >> 2, T = TB_THREADS, t = threadIdx.x;
  const float* __restrict__ z = io.z;
  double* scratch = reinterpret_cast<double*>(smem);
  if (bn.training) {
    long long r0, r1;
    cta_rows(io.rows, r0, r1);
    if (r1 > r0) {                                         // uniform per CTA
      if (ncg >= T) {
        for (int cg = t; cg < ncg; cg += T) {
          double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          for (long long r = r0; r < r1; r++) {
            const float4 v = ldcg4(z + r * C + 4 * cg);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            acc[4] += (double)v.x * v.x; acc[5] += (double)v.y * v.y; acc[6] += (double)v.z * v.z; acc[7] += (double)v.w * v.w;
          }
          reduce_lanes_and_add(acc, cg, 0, 1, ncg, true, scratch, bn.sums, bn.sums + C);
        }
      } else {
        const int RL = T / ncg, rl = t / ncg, cg = t - rl * ncg;
        const bool active = rl < RL;
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (active)
          for (long long r = r0 + rl; r < r1; r += RL) {
            const float4 v = ldcg4(z + r * C + 4 * cg);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            acc[4] += (double)v.x * v.x; acc[5] += (double)v.y * v.y; acc[6] += (double)v.z * v.z; acc[7] += (double)v.w * v.w;
          }
        reduce_lanes_and_add(acc, cg, rl, RL, ncg, active, scratch, bn.sums, bn.sums + C);
      }
    }
  }
  if (bn.training) grid_barrier(io.sync, 2u * gridDim.x);

  // ---- phase 3: finalize (every CTA for itself; CTA 0 publishes) + normalise + LeakyReLU (+ upsample x2 + skip)
  float* s_scale = reinterpret_cast<float*>(smem);
  float* s_shift = s_scale + C;
  if (!bn.training) {                                     // inference: BatchNorm folded by the host-side finalize
    for (int c = t; c < C; c += T) {
      s_scale[c] = __ldg(bn.ss + c);
      s_shift[c] = __ldg(bn.ss + C + c);
    }
  } else
  for (int c = t; c < C; c += T) {
    const double sum = __ldcg(bn.sums + c), sumsq = __ldcg(bn.sums + C + c);
    const double mean = sum / (double)io.rows;
    double var = sumsq / (double)io.rows - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)bn.eps);
    const double g = ms_ldp_d(bn.gamma, bn.pdt, c), b = ms_ldp_d(bn.beta, bn.pdt, c);
    const float sc = (float)(g * rstd), sh = (float)(b - mean * g * rstd);
    s_scale[c] = sc;
    s_shift[c] = sh;
    if (blockIdx.x == 0) {
      const double cb = bn.cbias ? ms_ldp_d(bn.cbias, bn.pdt, c) : 0.0;
      const double unb = io.rows > 1 ? var * ((double)io.rows / (double)(io.rows - 1)) : var;
      const double rm = ms_ldp_d(bn.rmean, bn.pdt, c), rv = ms_ldp_d(bn.rvar, bn.pdt, c);
      ms_stp(bn.rmean, bn.pdt, c, (1.0 - (double)bn.momentum) * rm + (double)bn.momentum * (mean + cb));
      ms_stp(bn.rvar, bn.pdt, c, (1.0 - (double)bn.momentum) * rv + (double)bn.momentum * unb);
      bn.ss[c] = sc;
      bn.ss[C + c] = sh;
      bn.ss[2 * C + c] = (float)mean;
      bn.ss[3 * C + c] = (float)rstd;
    }
  }
  if (bn.training && blockIdx.x == 0 && t == 0 && bn.nbt) bn.nbt[0] += 1;
  __syncthreads();
  {
    const long long rows_out = io.up2 ? 2 * io.rows : io.rows;
    long long o0, o1;
    cta_rows(rows_out, o0, o1);
    const long long total = (o1 - o0) * ncg;
    const float slope = bn.slope;
    for (long long i = t; i < total; i += T) {
      const long long ro = o0 + i / ncg;
      const int cg = (int)(i % ncg);
      long long ri = ro;
      if (io.up2) {
        const long long b = ro / (2 * io.L);
        const int l2 = (int)(ro - b * 2 * io.L);
        ri = b * io.L + (l2 >> 1);
      }
      const float4 v = ldcg4(z + ri * C + 4 * cg);
      const float4 s = *reinterpret_cast<const float4*>(s_scale + 4 * cg);
      const float4 h = *reinterpret_cast<const float4*>(s_shift + 4 * cg);
      float4 o;
      o.x = fmaf(v.x, s.x, h.x); o.y = fmaf(v.y, s.y, h.y); o.z = fmaf(v.z, s.z, h.z); o.w = fmaf(v.w, s.w, h.w);
      o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
      o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
      if (io.res) {
        const float4 r = ldcg4(io.res + ro * C + 4 * cg);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      } else if (io.res_pl) {
        for (int pl = 0; pl < (io.res_fmt == MS_BF16X2 ? 2 : 1); pl++) {
          const uint2 u = __ldcg(reinterpret_cast<const uint2*>(io.res_pl + pl * io.res_ps + ro * C + 4 * cg));
          const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&u.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
          o.x += __bfloat162float(h0.x); o.y += __bfloat162float(h0.y); o.z += __bfloat162float(h1.x); o.w += __bfloat162float(h1.y);
        }
      }
      if (io.y) *reinterpret_cast<float4*>(io.y + ro * C + 4 * cg) = o;
      if (io.planes) store_planes4(io.planes, io.pfmt, io.pstride, ro * C + 4 * cg, o);
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward block
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 dy_at4(const float* __restrict__ dy, long long r, int cg, int C, int up2, int L) {
  if (!up2) return ldcg4(dy + r * C + 4 * cg);
  const long long b = r / L;
  const int l = (int)(r - b * L);
  const long long ro = b * 2 * L + 2 * l;
  const float4 a = ldcg4(dy + ro * C + 4 * cg), c = ldcg4(dy + (ro + 1) * C + 4 * cg);
  return make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
}

__global__ void __launch_bounds__(TB_THREADS, 1)
conv_block_train_bwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                            const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_w_lo,
                            const __grid_constant__ IgemmParams p, const __grid_constant__ BnParams bn,
                            const __grid_constant__ BwdIO io) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[TB_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TB_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar, tempty_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  if (io.has_gemm) {
    while (tmem_cols < (uint32_t)p.block_n) tmem_cols <<= 1;
    if (warp == 0 && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
      if (p.npass > 1) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_lo)) : "memory");
      }
      for (int s = 0; s < TB_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(&tfull_bar, 1);
      mbar_init(&tempty_bar, 4);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = io.has_gemm ? tmem_base_smem : 0u;

  const int C = bn.C, ncg = C >> 2, T = TB_THREADS, t = threadIdx.x;
  const float* __restrict__ z = io.z;
  const float* __restrict__ dy = io.dy;
  const float slope = bn.slope;
  // per-channel constants of the forward pass -> shared memory (scale, shift, mean, rstd)
  float* s_sc = reinterpret_cast<float*>(smem);
  float* s_sh = s_sc + C;
  float* s_mu = s_sh + C;
  float* s_rs = s_mu + C;
  float* s_dg = s_rs + C;
  float* s_db = s_dg + C;
  double* scratch = reinterpret_cast<double*>(s_db + C);          // 6*C floats = 24*C bytes: 8-byte aligned
  for (int c = t; c < C; c += T) {
    s_sc[c] = __ldcg(bn.ss + c);
    s_sh[c] = __ldcg(bn.ss + C + c);
    s_mu[c] = __ldcg(bn.ss + 2 * C + c);
    s_rs[c] = __ldcg(bn.ss + 3 * C + c);
  }
  __syncthreads();
  long long r0, r1;
  cta_rows(io.rows, r0, r1);

  // ---- phase 1: dbeta = sum dz, dgamma = sum dz * xhat with dz = dy * act'(z)   (elementwise.cu: bn_act_bwd_reduce_kernel)
  if (r1 > r0) {
    if (ncg >= T) {
      for (int cg = t; cg < ncg; cg += T) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * cg), sh = *reinterpret_cast<const float4*>(s_sh + 4 * cg);
        const float4 mu = *reinterpret_cast<const float4*>(s_mu + 4 * cg), rs = *reinterpret_cast<const float4*>(s_rs + 4 * cg);
        for (long long r = r0; r < r1; r++) {
          const float4 xv = ldcg4(z + r * C + 4 * cg);
          const float4 d = dy_at4(dy, r, cg, C, io.up2, io.L);
          const float g0 = fmaf(xv.x, sc.x, sh.x) > 0.f ? d.x : d.x * slope, g1 = fmaf(xv.y, sc.y, sh.y) > 0.f ? d.y : d.y * slope;
          const float g2 = fmaf(xv.z, sc.z, sh.z) > 0.f ? d.z : d.z * slope, g3 = fmaf(xv.w, sc.w, sh.w) > 0.f ? d.w : d.w * slope;
          acc[0] += (double)g0 * (double)((xv.x - mu.x) * rs.x); acc[1] += (double)g1 * (double)((xv.y - mu.y) * rs.y);
          acc[2] += (double)g2 * (double)((xv.z - mu.z) * rs.z); acc[3] += (double)g3 * (double)((xv.w - mu.w) * rs.w);
          acc[4] += g0; acc[5] += g1; acc[6] += g2; acc[7] += g3;
        }
        reduce_lanes_and_add(acc, cg, 0, 1, ncg, true, scratch, bn.sums, bn.sums + C);
      }
    } else {
      const int RL = T / ncg, rl = t / ncg, cg = t - rl * ncg;
      const bool active = rl < RL;
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (active) {
        const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * cg), sh = *reinterpret_cast<const float4*>(s_sh + 4 * cg);
        const float4 mu = *reinterpret_cast<const float4*>(s_mu + 4 * cg), rs = *reinterpret_cast<const float4*>(s_rs + 4 * cg);
        for (long long r = r0 + rl; r < r1; r += RL) {
          const float4 xv = ldcg4(z + r * C + 4 * cg);
          const float4 d = dy_at4(dy, r, cg, C, io.up2, io.L);
          const float g0 = fmaf(xv.x, sc.x, sh.x) > 0.f ? d.x : d.x * slope, g1 = fmaf(xv.y, sc.y, sh.y) > 0.f ? d.y : d.y * slope;
          const float g2 = fmaf(xv.z, sc.z, sh.z) > 0.f ? d.z : d.z * slope, g3 = fmaf(xv.w, sc.w, sh.w) > 0.f ? d.w : d.w * slope;
          acc[0] += (double)g0 * (double)((xv.x - mu.x) * rs.x); acc[1] += (double)g1 * (double)((xv.y - mu.y) * rs.y);
          acc[2] += (double)g2 * (double)((xv.z - mu.z) * rs.z); acc[3] += (double)g3 * (double)((xv.w - mu.w) * rs.w);
          acc[4] += g0; acc[5] += g1; acc[6] += g2; acc[7] += g3;
        }
      }
      reduce_lanes_and_add(acc, cg, rl, RL, ncg, active, scratch, bn.sums, bn.sums + C);
    }
  }
  grid_barrier(io.sync, gridDim.x);

  // ---- phase 2: dz = scale * (g - dbeta/N - xhat * dgamma/N) -> operand planes; affine gradients (elementwise.cu: bn_act_bwd_apply_kernel)
  for (int c = t; c < C; c += T) {
    const double dg = __ldcg(bn.sums + c), db = __ldcg(bn.sums + C + c);
    s_dg[c] = (float)dg;
    s_db[c] = (float)db;
    if (blockIdx.x == 0) {
      if (io.ggamma) ms_stp(io.ggamma, io.gdt, c, ms_ldp_d(io.ggamma, io.gdt, c) + dg);
      if (io.gbeta) ms_stp(io.gbeta, io.gdt, c, ms_ldp_d(io.gbeta, io.gdt, c) + db);
    }
  }
  __syncthreads();
  {
    const float inv = 1.f / (float)io.rows;
    const long long total = (r1 - r0) * ncg;
    for (long long i = t; i < total; i += T) {
      const long long r = r0 + i / ncg;
      const int cg = (int)(i % ncg);
      const float4 xv = ldcg4(z + r * C + 4 * cg);
      const float4 d = dy_at4(dy, r, cg, C, io.up2, io.L);
      const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * cg), sh = *reinterpret_cast<const float4*>(s_sh + 4 * cg);
      const float4 mu = *reinterpret_cast<const float4*>(s_mu + 4 * cg), rs = *reinterpret_cast<const float4*>(s_rs + 4 * cg);
      const float4 dgv = *reinterpret_cast<const float4*>(s_dg + 4 * cg), dbv = *reinterpret_cast<const float4*>(s_db + 4 * cg);
      const float g0 = fmaf(xv.x, sc.x, sh.x) > 0.f ? d.x : d.x * slope, g1 = fmaf(xv.y, sc.y, sh.y) > 0.f ? d.y : d.y * slope;
      const float g2 = fmaf(xv.z, sc.z, sh.z) > 0.f ? d.z : d.z * slope, g3 = fmaf(xv.w, sc.w, sh.w) > 0.f ? d.w : d.w * slope;
      float4 o;
      if (bn.training) {
        o.x = sc.x * (g0 - dbv.x * inv - ((xv.x - mu.x) * rs.x) * dgv.x * inv);
        o.y = sc.y * (g1 - dbv.y * inv - ((xv.y - mu.y) * rs.y) * dgv.y * inv);
        o.z = sc.z * (g2 - dbv.z * inv - ((xv.z - mu.z) * rs.z) * dgv.z * inv);
        o.w = sc.w * (g3 - dbv.w * inv - ((xv.w - mu.w) * rs.w) * dgv.w * inv);
      } else {
        o.x = sc.x * g0; o.y = sc.y * g1; o.z = sc.z * g2; o.w = sc.w * g3;
      }
      store_planes4(io.dzp, io.pfmt, io.pstride, r * C + 4 * cg, o);
    }
  }
  if (!io.has_gemm) return;
  // the planes just written (generic proxy) are read by other CTAs' TMA loads (async proxy) in the next phase
  asm volatile("fence.proxy.async;" ::: "memory");
  grid_barrier(io.sync, 2u * gridDim.x);
  asm volatile("fence.proxy.async;" ::: "memory");

  // ---- phase 3: dx = conv^T(dz) on the tensor cores
  GemmSmem sm;
  sm.a = smem; sm.b = smem + TB_STAGES * A_STAGE_BYTES;
  sm.full = full_bar; sm.empty = empty_bar; sm.tfull = &tfull_bar; sm.tempty = &tempty_bar;
  PipeState st;
  st.s = 0; st.ph = 0; st.li = 0;
  gemm_phase(&map_a, &map_w, &map_a_lo, &map_w_lo, p, io.dx, sm, tmem_base, st);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient, accumulated in place
// ------------------------------------------------------------------------------------------------------------------
// Same tiling as wgrad_tc_kernel (conv_tc.cu): a CTA owns a (class, tap, 128 x c_tile) tile of dWp and a slice of the pixel
// rows.  Every slice adds its tile into the SAME fp32 accumulator with 16-byte vector reductions, so the x18 partial
// traffic of the workspace scheme (and the kernel that summed it) is gone; the accumulator is zero-filled once per step.
struct WgradAccParams {
  int ntaps, cchunks, shared_taps, num_classes, class_n;
  int box_w, box_h, box_b, tiles_w, tiles_h, tiles_b;
  int n_tiles, c_tiles, kpad, split, npass;
  int c_tile;
  int a_chan_base[MS_IGEMM_MAX_CLASSES];
  int z_chan_base[MS_IGEMM_MAX_CLASSES];
  short taps[MS_IGEMM_MAX_TAPS][4];
};
constexpr int WGA_ROWS = 64;
constexpr uint32_t WGA_CHUNK_BYTES = WGA_ROWS * 64 * 2;
constexpr int WGA_STAGES = 4;

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc2(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;      // LBO: next 64-channel chunk
  d |= (uint64_t)(1024 >> 4) << 32;      // SBO: next group of 8 pixel rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_acc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z,
                 const __grid_constant__ CUtensorMap map_x_lo, const __grid_constant__ CUtensorMap map_z_lo,
                 const __grid_constant__ WgradAccParams p, float* __restrict__ acc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[WGA_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WGA_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = blockIdx.z / p.ntaps, tap = blockIdx.z - cls * p.ntaps;
  const int nt = blockIdx.y / p.c_tiles, ct = blockIdx.y - nt * p.c_tiles;
  const int n0 = nt * 128, c0 = ct * p.c_tile;
  const int nc = min(p.c_tile, p.kpad - c0);
  const int xchunks = nc / 64;
  const uint32_t stage_bytes = (2 + xchunks) * WGA_CHUNK_BYTES;
  const int total_rt = p.tiles_w * p.tiles_h * p.tiles_b;
  const int per = (total_rt + p.split - 1) / p.split;
  const int rt_beg = blockIdx.x * per, rt_end = min(total_rt, rt_beg + per);
  const int num_k = max(0, rt_end - rt_beg) * p.npass;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + WGA_STAGES * 2 * WGA_CHUNK_BYTES;
  uint32_t tmem_cols = 64;
  while (tmem_cols < (uint32_t)nc) tmem_cols <<= 1;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_z)) : "memory");
    for (int s = 0; s < WGA_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  if (num_k > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const short* t = p.taps[(p.shared_taps ? 0 : cls * p.ntaps) + tap];
        const int xc = p.a_chan_base[cls] + t[0] + c0;
        const int zc = p.z_chan_base[cls] + n0;
        for (int ks = 0; ks < num_k; ks++) {
          const int s = ks % WGA_STAGES;
          const uint32_t ph = (uint32_t)(ks / WGA_STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int pass = ks % p.npass;              // split-bf16: x_hi*z_hi, x_hi*z_lo, x_lo*z_hi
          int rt = rt_beg + ks / p.npass;
          const CUtensorMap* mz = pass == 1 ? &map_z_lo : &map_z;
          const CUtensorMap* mx = pass == 2 ? &map_x_lo : &map_x;
          const int tw = rt % p.tiles_w; rt /= p.tiles_w;
          const int th = rt % p.tiles_h; rt /= p.tiles_h;
          const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = rt * p.box_b;
          mbar_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* sa = smem_a + (size_t)s * 2 * WGA_CHUNK_BYTES;
          uint8_t* sb = smem_b + (size_t)s * 4 * WGA_CHUNK_BYTES;
          tma_load_5d(mz, &full_bar[s], sa, zc, w0, 0, h0, b0);
          tma_load_5d(mz, &full_bar[s], sa + WGA_CHUNK_BYTES, zc + 64, w0, 0, h0, b0);
          for (int i = 0; i < xchunks; i++)
            tma_load_5d(mx, &full_bar[s], sb + (size_t)i * WGA_CHUNK_BYTES, xc + 64 * i, w0 + t[1], t[2], h0 + t[3], b0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < num_k; ks++) {
          const int s = ks % WGA_STAGES;
          const uint32_t ph = (uint32_t)(ks / WGA_STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_mnmajor_sw128_desc2(smem_u32(smem_a + (size_t)s * 2 * WGA_CHUNK_BYTES));
          const uint64_t db = make_mnmajor_sw128_desc2(smem_u32(smem_b + (size_t)s * 4 * WGA_CHUNK_BYTES));
#pragma unroll
          for (int k = 0; k < WGA_ROWS / UMMA_K; k++)
            umma_bf16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const bool valid = (n0 + r) < p.class_n;
      float* dst_row = acc + ((size_t)(cls * p.class_n + n0 + r) * p.ntaps + tap) * p.kpad + c0;
      mbar_wait(&tmem_full_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int cc = 0; cc < nc; cc += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)cc, v);
        tmem_wait_ld16(v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; j++)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + cc + 4 * j), "f"(__uint_as_float(v[4 * j])),
                         "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                         : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// every accumulator of a sub-network -> its parameter-gradient buffer (dw += unpack(acc)), one launch
__global__ void unpack_wgrad_multi_kernel(const ms_wgrad_entry* __restrict__ table) {
  const ms_wgrad_entry& e = table[blockIdx.y];
  const long long total = (long long)e.Cout * e.Cin_g * e.taps;
  const float* __restrict__ acc = reinterpret_cast<const float*>(e.acc);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % e.taps);
    const long long t2 = i / e.taps;
    const int c = (int)(t2 % e.Cin_g);
    const long long o = t2 / e.Cin_g;
    const double v = (double)acc[(o * e.taps + tap) * e.kpad + c];
    ms_stp(e.dw, e.pdt, i, v + (e.accumulate ? ms_ldp_d(e.dw, e.pdt, i) : 0.0));
  }
}

static int launch_coop(const void* fn, dim3 grid, size_t smem, cudaStream_t cs, void** args) {
  // cooperative launch: the grid barriers need every CTA resident; MS_TRAIN_COOP=0 falls back to a plain launch
  static int coop = -1;
  if (coop < 0) {
    const char* e = getenv("MS_TRAIN_COOP");
    coop = (e && e[0] == '0') ? 0 : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.gridDim = grid; cfg.blockDim = dim3(TB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = cs;
  cfg.attrs = at; cfg.numAttrs = coop ? 1 : 0;
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  return (int)e;
}

static int fill_bn(BnParams& b, const ms_block_bn* s) {
  if (!s || s->C < 16 || s->C % 4 || !s->gamma || !s->beta || !s->sums || !s->ss) return MS_EINVAL;
  if (s->pdt != MS_F32 && s->pdt != MS_F64) return MS_EINVAL;
  b.C = s->C; b.pdt = s->pdt; b.training = s->training; b.momentum = s->momentum; b.eps = s->eps; b.slope = s->slope;
  b.gamma = s->gamma; b.beta = s->beta; b.cbias = s->conv_bias; b.rmean = s->running_mean; b.rvar = s->running_var;
  b.nbt = reinterpret_cast<long long*>(s->num_batches_tracked); b.sums = s->sums; b.ss = s->ss;
  return 0;
}

}  // namespace

extern "C" int ms_conv_block_train_fwd(const ms_igemm_desc* d, const void* a, const void* w, float* z, const ms_block_bn* bn,
                                       float* y, void* planes, int pfmt, int64_t pstride, const float* res,
                                       const void* res_planes, int res_pfmt, int64_t res_pstride, int up2, void* sync,
                                       void* stream) {
  if (!d || !a || !w || !z || !bn || !sync || (!y && !planes)) return MS_EINVAL;
  if (d->out_dtype != MS_F32 || d->epilogue != 0) return MS_EINVAL;
  if (bn->training && (!bn->running_mean || !bn->running_var)) return MS_EINVAL;
  if (res_planes && (((uintptr_t)res_planes & 7) || (res_pfmt != MS_BF16 && res_pfmt != MS_BF16X2) ||
                     (res_pfmt == MS_BF16X2 && (res_pstride <= 0 || (res_pstride * 2) % 8))))
    return MS_EINVAL;
  if (((uintptr_t)z & 15) || ((uintptr_t)y & 15) || ((uintptr_t)planes & 15) || ((uintptr_t)res & 15)) return MS_EINVAL;
  if (planes && pfmt != MS_BF16 && pfmt != MS_BF16X2) return MS_EINVAL;
  if (planes && pfmt == MS_BF16X2 && (pstride <= 0 || (pstride * 2) % 8)) return MS_EINVAL;
  if (up2 && ((!res && !res_planes) || d->out_dims[1] != 1)) return MS_EINVAL;
  CUtensorMap maps[4];
  IgemmParams p;
  int rc = igemm_prepare(d, a, w, d->block_n, maps, &p);
  if (rc) return rc;
  BnParams b;
  rc = fill_bn(b, bn);
  if (rc) return rc;
  const long long rows = (long long)d->out_dims[0] * d->out_dims[1] * d->out_dims[2];
  const int C = d->num_classes * d->class_n;
  if (C != b.C) return MS_EINVAL;
  // z must be a dense (rows, C) matrix: the element-wise phases index it that way
  if (d->out_strides[0] != C || d->out_strides[1] != (int64_t)C * d->out_dims[0] ||
      d->out_strides[2] != (int64_t)C * d->out_dims[0] * d->out_dims[1])
    return MS_EINVAL;
  for (int i = 0; i < d->num_classes; i++)
    if (d->out_off[i] != (int64_t)i * d->class_n) return MS_EINVAL;
  if ((size_t)C * 8 + 8 * 8 * TB_THREADS > TB_RING_BYTES) return MS_EINVAL;
  FwdIO io;
  io.z = z; io.y = y; io.planes = reinterpret_cast<__nv_bfloat16*>(planes); io.pfmt = pfmt; io.pstride = pstride;
  io.res = up2 ? res : nullptr; io.up2 = up2 ? 1 : 0; io.L = d->out_dims[0]; io.rows = rows;
  io.res_pl = (up2 && !res) ? reinterpret_cast<const __nv_bfloat16*>(res_planes) : nullptr; io.res_fmt = res_pfmt; io.res_ps = res_pstride;
  io.sync = reinterpret_cast<unsigned int*>(sync);
  const long long items = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles_per_class * p.num_classes * p.split_k;
  if (items < 1 || items > 0x7fffffffLL) return MS_EINVAL;
  const int sms = ms_num_sms();
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  const size_t smem = TB_RING_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(conv_block_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  void* args[] = {&maps[0], &maps[1], &maps[2], &maps[3], &p, &b, &io};
  rc = launch_coop(reinterpret_cast<const void*>(conv_block_train_fwd_kernel), dim3(grid), smem, ms_stream(stream), args);
  if (rc) return rc;
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_conv_block_train_bwd(const ms_igemm_desc* dg, const float* dy, const float* z, const ms_block_bn* bn,
                                       int64_t rows, int up2, int rows_per_seq, void* dz_planes, int pfmt, int64_t pstride,
                                       void* grad_gamma, void* grad_beta, int gdt, const void* wt, float* dx, void* sync,
                                       void* stream) {
  if (!dy || !z || !bn || !dz_planes || !sync || rows < 1) return MS_EINVAL;
  if (pfmt != MS_BF16 && pfmt != MS_BF16X2) return MS_EINVAL;
  if (pfmt == MS_BF16X2 && (pstride <= 0 || (pstride * 2) % 8)) return MS_EINVAL;
  if (((uintptr_t)dy & 15) || ((uintptr_t)z & 15) || ((uintptr_t)dz_planes & 15) || ((uintptr_t)dx & 15)) return MS_EINVAL;
  if (gdt != MS_F32 && gdt != MS_F64) return MS_EINVAL;
  if (up2 && rows_per_seq < 1) return MS_EINVAL;
  BnParams b;
  int rc = fill_bn(b, bn);
  if (rc) return rc;
  if ((size_t)b.C * 24 + 8 * 8 * TB_THREADS > TB_RING_BYTES) return MS_EINVAL;
  CUtensorMap maps[4];
  IgemmParams p;
  BwdIO io;
  io.dy = dy; io.z = z; io.dzp = reinterpret_cast<__nv_bfloat16*>(dz_planes); io.pfmt = pfmt; io.pstride = pstride;
  io.up2 = up2 ? 1 : 0; io.L = rows_per_seq; io.rows = rows; io.ggamma = grad_gamma; io.gbeta = grad_beta; io.gdt = gdt;
  io.dx = dx; io.has_gemm = dg ? 1 : 0; io.sync = reinterpret_cast<unsigned int*>(sync);
  const int sms = ms_num_sms();
  long long want;
  if (dg) {
    if (!wt || !dx || dg->out_dtype != MS_F32 || dg->epilogue != 0) return MS_EINVAL;
    rc = igemm_prepare(dg, dz_planes, wt, dg->block_n, maps, &p);
    if (rc) return rc;
    want = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles_per_class * p.num_classes * p.split_k;
    if (want < 1 || want > 0x7fffffffLL) return MS_EINVAL;
  } else {
    memset(&p, 0, sizeof(p));
    memset(maps, 0, sizeof(maps));
    p.block_n = 32;
    // element-wise only: enough CTAs to spread the rows, no more than one per SM
    want = (rows * b.C + 16383) / 16384;
    if (want < 1) want = 1;
  }
  const unsigned grid = (unsigned)(want < sms ? want : sms);
  const size_t smem = TB_RING_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(conv_block_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  void* args[] = {&maps[0], &maps[1], &maps[2], &maps[3], &p, &b, &io};
  rc = launch_coop(reinterpret_cast<const void*>(conv_block_train_bwd_kernel), dim3(grid), smem, ms_stream(stream), args);
  if (rc) return rc;
  MS_LAUNCH_CHECK();
  return 0;
}

static int encode_5d_acc(EncodeTiledFn enc, CUtensorMap* m, const void* base, const int32_t* dims, const int64_t* strides_el,
                         const int* box) {
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; i++) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; }
  for (int i = 1; i < 5; i++) {
    st[i - 1] = (cuuint64_t)strides_el[i] * 2;
    if (st[i - 1] % 16) return MS_EINVAL;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MS_EINVAL;
}

extern "C" int ms_wgrad_bf16_acc(const ms_igemm_desc* d, const void* x, const void* dz, float* acc, void* stream) {
  if (!d || !x || !dz || !acc) return MS_EINVAL;
  if (d->num_classes < 1 || d->num_classes > MS_IGEMM_MAX_CLASSES || d->ntaps < 1 || d->cchunks < 1) return MS_EINVAL;
  if (((uintptr_t)x & 15) || ((uintptr_t)dz & 15) || ((uintptr_t)acc & 15)) return MS_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return MS_ENOTSUP;
  WgradAccParams p;
  p.ntaps = d->ntaps; p.cchunks = d->cchunks; p.shared_taps = d->shared_taps;
  p.num_classes = d->num_classes; p.class_n = d->class_n;
  int bw = d->box[1], bh = d->box[3], bb = d->box[4];
  if (bb > 1) bb /= 2; else if (bh > 1) bh /= 2; else bw /= 2;
  if (bw * bh * bb != WGA_ROWS) return MS_EINVAL;
  p.box_w = bw; p.box_h = bh; p.box_b = bb;
  const int Wo = d->out_dims[0], Ho = d->out_dims[1], Bo = d->out_dims[2];
  p.tiles_w = (Wo + bw - 1) / bw; p.tiles_h = (Ho + bh - 1) / bh; p.tiles_b = (Bo + bb - 1) / bb;
  p.kpad = d->cchunks * BLOCK_K;
  p.n_tiles = (d->class_n + 127) / 128;
  p.c_tile = d->wgrad_c_tile > 0 ? d->wgrad_c_tile : 256;
  if (p.c_tile % 64 || p.c_tile > 256) return MS_EINVAL;
  p.c_tiles = (p.kpad + p.c_tile - 1) / p.c_tile;
  for (int i = 0; i < MS_IGEMM_MAX_CLASSES; i++) { p.a_chan_base[i] = d->a_chan_base[i]; p.z_chan_base[i] = (int)d->out_off[i]; }
  for (int i = 0; i < MS_IGEMM_MAX_TAPS; i++)
    for (int j = 0; j < 4; j++) p.taps[i][j] = d->taps[i][j];
  const long long total_rt = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
  long long split = d->split_k > 1 ? d->split_k : 1;
  if (split > total_rt) split = total_rt;
  const long long per = (total_rt + split - 1) / split;
  split = (total_rt + per - 1) / per;
  p.split = (int)split;
  if (d->planes != 1 && d->planes != 2) return MS_EINVAL;
  p.npass = d->planes == 2 ? 3 : 1;
  if (d->planes == 2 && (d->a_plane_stride <= 0 || d->out_plane_stride <= 0 || (d->a_plane_stride * 2) % 16 || (d->out_plane_stride * 2) % 16))
    return MS_EINVAL;
  CUtensorMap map_x, map_z, map_x_lo, map_z_lo;
  int box[5] = {64, bw, 1, bh, bb};
  const int32_t zdims[5] = {(int32_t)d->out_strides[0], Wo, 1, Ho, Bo};
  const int64_t zstr[5] = {1, d->out_strides[0], d->out_strides[1], d->out_strides[1], d->out_strides[2]};
  int rc = encode_5d_acc(enc, &map_x, x, d->a_dims, d->a_strides, box);
  if (rc) return rc;
  rc = encode_5d_acc(enc, &map_z, dz, zdims, zstr, box);
  if (rc) return rc;
  if (d->planes == 2) {
    rc = encode_5d_acc(enc, &map_x_lo, reinterpret_cast<const __nv_bfloat16*>(x) + d->a_plane_stride, d->a_dims, d->a_strides, box);
    if (rc) return rc;
    rc = encode_5d_acc(enc, &map_z_lo, reinterpret_cast<const __nv_bfloat16*>(dz) + d->out_plane_stride, zdims, zstr, box);
    if (rc) return rc;
  } else {
    map_x_lo = map_x; map_z_lo = map_z;
  }
  const size_t smem = (size_t)WGA_STAGES * 6 * WGA_CHUNK_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(wgrad_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  dim3 grid((unsigned)p.split, (unsigned)(p.n_tiles * p.c_tiles), (unsigned)(d->num_classes * d->ntaps));
  wgrad_acc_kernel<<<grid, NUM_THREADS, smem, ms_stream(stream)>>>(map_x, map_z, map_x_lo, map_z_lo, p, acc);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_unpack_wgrad_multi(const ms_wgrad_entry* table_dev, int n_entries, int blocks_per_entry, void* stream) {
  if (!table_dev || n_entries < 1 || n_entries > 65535) return MS_EINVAL;
  if (blocks_per_entry < 1) blocks_per_entry = 32;
  dim3 grid((unsigned)blocks_per_entry, (unsigned)n_entries);
  unpack_wgrad_multi_kernel<<<grid, 256, 0, ms_stream(stream)>>>(table_dev);
  MS_LAUNCH_CHECK();
  return 0;
}
