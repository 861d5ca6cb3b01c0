// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands fed by TMA.
//
//   out[row, n] = sum_{tap} sum_{c} A[row shifted by tap, c] * W[n, tap, c]       (bf16 x bf16 -> fp32)
//
// * A is a channels-last bf16 activation tensor seen through a 5-D TMA tensor map
//   (channel, w-like, h-parity, h-like, batch).  Each k-step loads one 128-row x 64-channel box
//   whose start coordinate is the tile origin plus the tap's shift; rows that fall outside the
//   tensor (the convolution's zero padding, ragged batch tails, channel tails) are zero-filled
//   by the TMA unit, so im2col is never materialised and padding costs nothing.  Stride-2
//   convolutions use a space-to-depth *view* of the same memory ((B,L,C) == (B,L/2,2C)), so the
//   strided taps are plain boxes as well.
// * W is the weight re-tiled once per parameter version to [n][tap][c] (K-major bf16).
// * 128 x block_n fp32 accumulator in TMEM; UMMA 128 x block_n x 16, cta_group::1.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld, bias / BN scale-shift / LeakyReLU, fp32 or bf16 stores).  4-stage smem ring with
//   full/empty mbarriers; tcgen05.commit releases stages and signals the epilogue.
// * "Classes" generalise groups and stride parities: every class owns class_n output columns, a
//   channel base in A, an output offset and (optionally) its own tap table.  Grouped sub-decoders
//   are classes with a shared tap table; the input gradient of a stride-2 convolution is a
//   stride-1 problem with one class per output parity.
//
// The same kernel serves forward and input-gradient (descriptors differ); see ms_igemm_desc.
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ IgemmParams p, const float* __restrict__ bias, const float* __restrict__ scale,
                const float* __restrict__ shift, void* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
  uint8_t* smem_a = smem;
  const int STAGES = p.stages;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates
  int mt = blockIdx.x;
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  const int tb = mt;
  const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = tb * p.box_b;
  const int cls = blockIdx.y / p.n_tiles_per_class;
  const int nt = blockIdx.y - cls * p.n_tiles_per_class;
  const int n0 = nt * p.block_n;                       // column offset inside the class
  // k-steps of this CTA: all of them, or one contiguous slice when the tile is split over gridDim.z
  const int num_k_total = p.ntaps * p.cchunks * p.npass;
  const int k_per = (num_k_total + p.split_k - 1) / p.split_k;
  const int k_beg = (int)blockIdx.z * k_per;
  const int num_k = max(0, min(num_k_total, k_beg + k_per) - k_beg);

  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.block_n) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (p.npass > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_lo)) : "memory");
    }
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int tap_base = p.shared_taps ? 0 : cls * p.ntaps;
      const int chan_base = p.a_chan_base[cls];
      const int wrow = cls * p.class_n + n0;
      for (int ks = 0; ks < num_k; ks++) {
        const int s = ks % STAGES;
        const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        // split-bf16: three passes per (tap, channel chunk): hi*hi, hi*lo, lo*hi
        const int kg = k_beg + ks;
        const int kk = kg / p.npass, pass = kg - kk * p.npass;
        const int tap = kk / p.cchunks, cc = kk - tap * p.cchunks;
        const short* t = p.taps[tap_base + tap];
        mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + b_stage_bytes);
        tma_load_5d(pass == 2 ? &map_a_lo : &map_a, &full_bar[s], smem_a + (size_t)s * A_STAGE_BYTES,
                    chan_base + t[0] + cc * BLOCK_K, w0 + t[1], t[2], h0 + t[3], b0);
        tma_load_2d(pass == 1 ? &map_w_lo : &map_w, &full_bar[s], smem_b + (size_t)s * b_stage_bytes, kk * BLOCK_K, wrow);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = block_n, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int ks = 0; ks < num_k; ks++) {
        const int s = ks % STAGES;
        const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + (size_t)s * A_STAGE_BYTES));
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + (size_t)s * b_stage_bytes));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
          // advance 16 bf16 = 32 B inside the swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);            // stage reusable once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);             // accumulator complete
    }
  } else {
    // ================= epilogue: 4 warps, one TMEM lane quadrant each =================
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // tile row == TMEM lane
    const int wi = r % p.box_w;
    const int hi = (r / p.box_w) % p.box_h;
    const int bi = r / (p.box_w * p.box_h);
    const int ow = w0 + wi, oh = h0 + hi, ob = b0 + bi;
    const bool valid = ow < p.out_w && oh < p.out_h && ob < p.out_b;
    const long long row_off = (long long)ob * p.os_b + (long long)oh * p.os_h + (long long)ow * p.os_w + p.out_off[cls] + n0;
    const int ncol = cls * p.class_n + n0;             // global column (bias / scale index)
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool lead = blockIdx.z == 0;                  // split-K: the first slice adds the bias
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (valid && (n0 + c0) < p.class_n && p.split_k > 1) {
        float* dst = reinterpret_cast<float*>(out) + row_off + c0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float a = __uint_as_float(v[4 * j]), b = __uint_as_float(v[4 * j + 1]), c = __uint_as_float(v[4 * j + 2]),
                d = __uint_as_float(v[4 * j + 3]);
          if (bias && lead) {
            a += __ldg(bias + ncol + c0 + 4 * j); b += __ldg(bias + ncol + c0 + 4 * j + 1);
            c += __ldg(bias + ncol + c0 + 4 * j + 2); d += __ldg(bias + ncol + c0 + 4 * j + 3);
          }
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
        }
      } else if (valid && (n0 + c0) < p.class_n) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          float x = __uint_as_float(v[j]);
          if (p.epilogue == 1) {
            x = fmaf(x, __ldg(scale + ncol + c0 + j), __ldg(shift + ncol + c0 + j));
            x = x > 0.f ? x : x * p.slope;
          } else {
            if (bias) x += __ldg(bias + ncol + c0 + j);
            if (p.epilogue == 2) x = x > 0.f ? x : x * p.slope;
          }
          f[j] = x;
        }
        const int nrep = p.up2 ? 2 : 1;
        for (int j2 = 0; j2 < nrep; j2++) {
          long long off = row_off + c0;
          float g[16];
#pragma unroll
          for (int j = 0; j < 16; j++) g[j] = f[j];
          if (p.up2) {
            // UNet1D decoder step (layers.py:151): y[b, 2w + j2, :] = act(bn(z))[b, w, :] + residual[b, 2w + j2, :]
            off = (long long)ob * p.os_b * 2 + (long long)(2 * ow + j2) * p.os_w + p.out_off[cls] + n0 + c0;
            for (int pl = 0; pl < p.res_planes; pl++) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.res + (long long)pl * p.res_pstride + off);
              const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
              const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
                g[2 * j] += __bfloat162float(h2.x);
                g[2 * j + 1] += __bfloat162float(h2.y);
              }
            }
          }
          if (p.out_f32 || p.out_dtype == MS_F32) {
            float4* dst = reinterpret_cast<float4*>((p.out_dtype == MS_F32 ? reinterpret_cast<float*>(out) : p.out_f32) + off);
#pragma unroll
            for (int j = 0; j < 4; j++) dst[j] = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
          }
          if (p.out_dtype != MS_F32) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + off);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(g[2 * j], g[2 * j + 1]);
              w[j] = *reinterpret_cast<uint32_t*>(&h2);
              if (p.out_dtype == MS_BF16X2) {          // residuals for the lo plane
                g[2 * j] -= __bfloat162float(h2.x);
                g[2 * j + 1] -= __bfloat162float(h2.y);
              }
            }
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            if (p.out_dtype == MS_BF16X2) {
              dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + p.out_plane_stride + off);
#pragma unroll
              for (int j = 0; j < 8; j++) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(g[2 * j], g[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h2);
              }
              dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
              dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// Persistent form of the same GEMM (every launch that is not split-K).
//
// The one-tile-per-CTA kernel above serialises  mainloop -> epilogue  inside a CTA; ncu (profiles/r01_igemm_infer_*)
// showed the epilogue (TMEM -> registers -> scale/shift -> global stores, with the per-column constants fetched from
// global memory on the critical path) taking ~2x the mainloop, i.e. the tensor pipe idle two thirds of the time.
// Here one CTA per SM walks a static round-robin list of tiles with
//   * TWO accumulator buffers in TMEM (2 x block_n columns): the MMA warp fills buffer (i+1)&1 while the epilogue
//     warps drain buffer i&1 (tmem_full / tmem_empty mbarriers),
//   * the smem operand ring running straight across tile boundaries (the TMA producer never drains),
//   * 8 epilogue warps (two per TMEM lane quadrant, each taking half of the tile's columns),
//   * the tile's per-column constants (BN scale/shift or bias) staged in shared memory BEFORE the accumulator is ready.
// Epilogue extras (ms_igemm_bf16_mix): multiply every row by its soft cluster weight (row_w_mode 1) so that the grouped
// 1x1 `logits` convolution that follows becomes ONE dense GEMM over K*256 channels whose accumulator IS the mixture
// sum_k w_k * logits_k (jlcss.py:106-115,190-194) -- the per-cluster outputs (B,T,K*P) are never written to HBM;
// row_w_mode 2 adds the mixed bias sum_k w_k * b_k in that GEMM's epilogue.
// ---------------------------------------------------------------------------------------------
constexpr int PERSIST_THREADS = 320;     // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int EPI_THREADS = 256;
constexpr int MIX_MAX_K = 16, MIX_MAX_N = 128;


__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load delivered to the same shared-memory offset (and the same mbarrier offset) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

struct TileCoord {
  int w0, h0, b0, cls, n0;
};
// tile group g of a cluster of CSZ CTAs = CSZ consecutive 128-row tiles of the same (class, column tile); rank picks the row tile
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int g, int ny, int csz, int rank) {
  TileCoord t;
  const int y = g % ny;
  int mt = (g / ny) * csz + rank;
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  t.w0 = tw * p.box_w; t.h0 = th * p.box_h; t.b0 = mt * p.box_b;
  t.cls = y / p.n_tiles_per_class;
  t.n0 = (y - t.cls * p.n_tiles_per_class) * p.block_n;
  return t;
}

// 16 consecutive output columns of one row: activation / row weight / residual / stores
__device__ __forceinline__ void epilogue_store16(const IgemmParams& p, void* __restrict__ out, const float (&f)[16],
                                                 long long off, long long off_up, int nrep) {
  for (int j2 = 0; j2 < nrep; j2++) {
    float g[16];
#pragma unroll
    for (int j = 0; j < 16; j++) g[j] = f[j];
    long long o = off;
    if (p.up2) {
      // UNet1D decoder step (layers.py:151): y[b, 2w + j2, :] = act(bn(z))[b, w, :] + residual[b, 2w + j2, :]
      o = off_up + (long long)j2 * p.os_w;
      for (int pl = 0; pl < p.res_planes; pl++) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + (long long)pl * p.res_pstride + o);
        const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
          g[2 * j] += __bfloat162float(h2.x);
          g[2 * j + 1] += __bfloat162float(h2.y);
        }
      }
    }
    if (p.out_f32 || p.out_dtype == MS_F32) {
      float4* dst = reinterpret_cast<float4*>((p.out_dtype == MS_F32 ? reinterpret_cast<float*>(out) : p.out_f32) + o);
#pragma unroll
      for (int j = 0; j < 4; j++) dst[j] = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
    }
    if (p.out_dtype != MS_F32) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o);
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(g[2 * j], g[2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h2);
        if (p.out_dtype == MS_BF16X2) {          // residuals for the lo plane
          g[2 * j] -= __bfloat162float(h2.x);
          g[2 * j + 1] -= __bfloat162float(h2.y);
        }
      }
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      if (p.out_dtype == MS_BF16X2) {
        dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + p.out_plane_stride + o);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(g[2 * j], g[2 * j + 1]);
          w[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
}

// Lean epilogue bodies for the layouts that carry almost all of the traffic (VARIANT != 0): no residual / upsample, no
// extra fp32 copy; per-column constants read as float4 from shared memory, LeakyReLU as max(x, slope * x) (0 <= slope <= 1).
//   1: bf16 planes out     2: fp32 out     3: bf16 planes out, every row scaled by its cluster weight (row_w_mode 1)
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                             uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
               "r"(g), "r"(h)
               : "memory");
}

template <int VARIANT>
__device__ __forceinline__ void epilogue_chunk_fast(const uint32_t* cur, const float* __restrict__ sc,
                                                    const float* __restrict__ sh, float slope, float rw, void* __restrict__ dst) {
  float f[16];
#pragma unroll
  for (int j4 = 0; j4 < 4; j4++) {
    const float4 a = *reinterpret_cast<const float4*>(sc + 4 * j4);
    const float4 b = *reinterpret_cast<const float4*>(sh + 4 * j4);
    f[4 * j4 + 0] = fmaf(__uint_as_float(cur[4 * j4 + 0]), a.x, b.x);
    f[4 * j4 + 1] = fmaf(__uint_as_float(cur[4 * j4 + 1]), a.y, b.y);
    f[4 * j4 + 2] = fmaf(__uint_as_float(cur[4 * j4 + 2]), a.z, b.z);
    f[4 * j4 + 3] = fmaf(__uint_as_float(cur[4 * j4 + 3]), a.w, b.w);
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
    f[j] = fmaxf(f[j], f[j] * slope);
    if (VARIANT == 3) f[j] *= rw;
  }
  // 256-bit stores (sm_100 STG.256): every lane fills whole 32-byte sectors of its own output row
  if (VARIANT == 2) {
    st_global_v8(dst, __float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]),
                 __float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
    st_global_v8(reinterpret_cast<uint8_t*>(dst) + 32, __float_as_uint(f[8]), __float_as_uint(f[9]), __float_as_uint(f[10]),
                 __float_as_uint(f[11]), __float_as_uint(f[12]), __float_as_uint(f[13]), __float_as_uint(f[14]), __float_as_uint(f[15]));
  } else {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    st_global_v8(dst, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
  }
}

// One warp's share of a tile on the lean path: chunks [c_beg, c_lim) of 16 columns of its 32 rows.  TMEM is read 32 columns
// at a time (tcgen05.ld ... .x32), the next load in flight while the current 32 columns are scaled, activated, packed and
// stored; an odd trailing chunk takes a 16-column load.  `release` hands the accumulator buffer back to the MMA warp once the
// last load has landed.  (Measured alternative, profiles/r01_igemm_epilogue_experiments.txt: pulling all 128 columns out
// first and releasing before the math was 3 % SLOWER -- 168 registers, spills, stores bunched at the end of the tile.)
#define MS_TIE16(a, o)                                                                                                          \
  asm volatile("" : "+r"(a[o + 0]), "+r"(a[o + 1]), "+r"(a[o + 2]), "+r"(a[o + 3]), "+r"(a[o + 4]), "+r"(a[o + 5]), "+r"(a[o + 6]), \
               "+r"(a[o + 7]), "+r"(a[o + 8]), "+r"(a[o + 9]), "+r"(a[o + 10]), "+r"(a[o + 11]), "+r"(a[o + 12]), "+r"(a[o + 13]), \
               "+r"(a[o + 14]), "+r"(a[o + 15])::"memory")
// wait for the outstanding tcgen05.ld; the loaded registers pass THROUGH statements placed after it, so no use can move above
#define MS_WAIT_LD32(a)                                              \
  do {                                                               \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");     \
    MS_TIE16(a, 0);                                                  \
    MS_TIE16(a, 16);                                                 \
  } while (0)
template <int VARIANT, bool X32 = true, typename Release>
__device__ __forceinline__ void lean_epilogue_tile(uint32_t taddr, int c_beg, int c_lim, bool valid, const float* __restrict__ sc,
                                                   const float* __restrict__ sh, float slope, float rw, uint8_t* __restrict__ dst,
                                                   int dbg, Release release) {
  constexpr int CB = 16 * (VARIANT == 2 ? 4 : 2);          // output bytes per 16-column chunk
  const bool st = valid && !(dbg & 1);
  const int nch = c_lim - c_beg;
  if constexpr (!X32) {
    // 16 columns per load, two buffers (the 16-epilogue-warp configuration has 112 registers per thread)
    uint32_t va[16], vb[16];
    if (nch > 0) tmem_ld16(taddr + (uint32_t)(c_beg * 16), va);
    else release();
    for (int i = 0; i < nch; i += 2) {
      const int c = c_beg + i;
      tmem_wait_ld16(va);
      if (i + 1 < nch) tmem_ld16(taddr + (uint32_t)((c + 1) * 16), vb);
      else release();
      if (st) epilogue_chunk_fast<VARIANT>(va, sc + c * 16, sh + c * 16, slope, rw, dst);
      dst += CB;
      if (i + 1 < nch) {
        tmem_wait_ld16(vb);
        if (i + 2 < nch) tmem_ld16(taddr + (uint32_t)((c + 2) * 16), va);
        else release();
        if (st) epilogue_chunk_fast<VARIANT>(vb, sc + c * 16 + 16, sh + c * 16 + 16, slope, rw, dst);
        dst += CB;
      }
    }
    return;
  }
  const int n2 = nch >> 1;                                  // 32-column units
  uint32_t va[32], vb[32];
  if (n2 > 0) tmem_ld32(taddr + (uint32_t)(c_beg * 16), va);
  for (int u = 0; u < n2; u += 2) {
    const int c = c_beg + 2 * u;
    MS_WAIT_LD32(va);
    if (u + 1 < n2) tmem_ld32(taddr + (uint32_t)((c + 2) * 16), vb);
    else if (!(nch & 1)) release();
    if (st) {
      epilogue_chunk_fast<VARIANT>(va, sc + c * 16, sh + c * 16, slope, rw, dst);
      epilogue_chunk_fast<VARIANT>(va + 16, sc + c * 16 + 16, sh + c * 16 + 16, slope, rw, dst + CB);
    }
    dst += 2 * CB;
    if (u + 1 < n2) {
      MS_WAIT_LD32(vb);
      if (u + 2 < n2) tmem_ld32(taddr + (uint32_t)((c + 4) * 16), va);
      else if (!(nch & 1)) release();
      if (st) {
        epilogue_chunk_fast<VARIANT>(vb, sc + c * 16 + 32, sh + c * 16 + 32, slope, rw, dst);
        epilogue_chunk_fast<VARIANT>(vb + 16, sc + c * 16 + 48, sh + c * 16 + 48, slope, rw, dst + CB);
      }
      dst += 2 * CB;
    }
  }
  if (nch & 1) {
    const int c = c_lim - 1;
    uint32_t vt[16];
    tmem_ld16(taddr + (uint32_t)(c * 16), vt);
    tmem_wait_ld16(vt);
    release();
    if (st) epilogue_chunk_fast<VARIANT>(vt, sc + c * 16, sh + c * 16, slope, rw, dst);
  } else if (n2 == 0) {
    release();
  }
}

// CSZ > 1: thread-block clusters of CSZ CTAs along M.  Every CTA loads its own 128 activation rows and 1/CSZ of the weight
// tile, multicasting that slice into all CSZ shared memories: the L2 read traffic per k-step drops from 16 + 32 KB to
// 16 + 32/CSZ KB per CTA.  (Opt-in, MS_IGEMM_CLUSTER: it turned out not to be the bound, see igemm_cluster_size.)
// A stage is reusable once the MMAs of ALL CSZ CTAs have read it: the empty barriers count CSZ multicast commits.
template <int VARIANT, int CSZ>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
igemm_tc_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_w_lo,
                        const __grid_constant__ IgemmParams p, const float* __restrict__ bias, const float* __restrict__ scale,
                        const float* __restrict__ shift, void* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
  const int STAGES = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[2][256];
  __shared__ __align__(16) float s_shift[2][256];
  __shared__ float s_mixb[VARIANT == 0 ? MIX_MAX_K * MIX_MAX_N : 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.ntaps * p.cchunks * p.npass;
  const int ny = p.n_tiles_per_class * p.num_classes;
  const int crank = CSZ > 1 ? (int)cluster_ctarank() : 0;
  const int cid = (int)blockIdx.x / CSZ, ncl = (int)gridDim.x / CSZ;
  const int total_tiles = ((p.tiles_w * p.tiles_h * p.tiles_b + CSZ - 1) / CSZ) * ny;      // tile groups
  constexpr uint16_t cmask = (uint16_t)((1u << CSZ) - 1u);
  uint32_t cols_per_buf = 32;
  while (cols_per_buf < (uint32_t)p.block_n) cols_per_buf <<= 1;
  const uint32_t tmem_cols = 2 * cols_per_buf;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (p.npass > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_lo)) : "memory");
    }
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CSZ); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], EPI_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();         // peers' barriers are initialised before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================= TMA producer: the ring never drains between tiles =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t b_slice_bytes = b_stage_bytes / CSZ;
      const int b_slice_rows = p.block_n / CSZ;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const TileCoord t = decode_tile(p, tile, ny, CSZ, crank);
        const int tap_base = p.shared_taps ? 0 : t.cls * p.ntaps;
        const int chan_base = p.a_chan_base[t.cls];
        const int wrow = t.cls * p.class_n + t.n0;
        for (int kg = 0; kg < num_k; kg++) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int kk = kg / p.npass, pass = kg - kk * p.npass;          // split-bf16: hi*hi, hi*lo, lo*hi
          const int tap = kk / p.cchunks, cc = kk - tap * p.cchunks;
          const short* tp = p.taps[tap_base + tap];
          if (p.dbg & 2) {               // timing experiment: no operand loads, the MMAs run on whatever is in smem
            mbar_arrive(&full_bar[s]);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + b_stage_bytes);
          tma_load_5d(pass == 2 ? &map_a_lo : &map_a, &full_bar[s], smem_a + (size_t)s * A_STAGE_BYTES,
                      chan_base + tp[0] + cc * BLOCK_K, t.w0 + tp[1], tp[2], t.h0 + tp[3], t.b0);
          if (CSZ == 1)
            tma_load_2d(pass == 1 ? &map_w_lo : &map_w, &full_bar[s], smem_b + (size_t)s * b_stage_bytes, kk * BLOCK_K, wrow);
          else     // this CTA's slice of the weight tile, delivered to every CTA of the cluster
            tma_load_2d_mc(pass == 1 ? &map_w_lo : &map_w, &full_bar[s], smem_b + (size_t)s * b_stage_bytes + (size_t)crank * b_slice_bytes,
                           kk * BLOCK_K, wrow + crank * b_slice_rows, cmask);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: accumulator buffer lt & 1 =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl, lt++) {
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);     // epilogue drained this buffer (tile lt - 2)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)buf * cols_per_buf;
        for (int ks = 0; ks < num_k; ks++) {
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + (size_t)s * A_STAGE_BYTES));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + (size_t)s * b_stage_bytes));
          if (!(p.dbg & 4)) {            // timing experiment (bit 2): loads only, no MMAs
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++)
              umma_bf16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
          }
          if (CSZ == 1) umma_commit(&empty_bar[s]);
          else umma_commit_mc(&empty_bar[s], cmask);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ================= epilogue: 8 warps; TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =================
    const int et = (int)threadIdx.x - 64;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                       // tile row == TMEM lane
    const int wi = r % p.box_w;
    const int hi = (r / p.box_w) % p.box_h;
    const int bi = r / (p.box_w * p.box_h);
    const int chunks = p.block_n >> 4;
    const int chunks_h = (chunks + 1) >> 1;
    const int c_beg = half * chunks_h, c_end = min(chunks, c_beg + chunks_h);
    const bool act = p.epilogue != 0 && p.slope != 1.f;
    const int nrep = p.up2 ? 2 : 1;
    const int ncols_total = p.num_classes * p.class_n;
    if (VARIANT == 0 && p.row_w_mode == 2) {
      for (int i = et; i < p.mix_k * MIX_MAX_N; i += EPI_THREADS) {
        const int k = i / MIX_MAX_N, n = i - k * MIX_MAX_N;
        s_mixb[i] = (bias && n < ncols_total) ? __ldg(bias + (long long)k * ncols_total + n) : 0.f;
      }
    }
    const float slope_eff = act ? p.slope : 1.f;
    int lt = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl, lt++) {
      const int buf = lt & 1;
      const TileCoord t = decode_tile(p, tile, ny, CSZ, crank);
      const int ncol = t.cls * p.class_n + t.n0;             // global column (bias / scale index)
      // per-column constants of this tile -> shared memory, while the MMAs are still running
      for (int i = et; i < p.block_n; i += EPI_THREADS) {
        float sc = 1.f, sh = 0.f;
        if (t.n0 + i < p.class_n) {
          if (p.epilogue == 1) { sc = __ldg(scale + ncol + i); sh = __ldg(shift + ncol + i); }
          else if (bias && p.row_w_mode != 2) sh = __ldg(bias + ncol + i);
        }
        s_scale[buf][i] = sc;
        s_shift[buf][i] = sh;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      const int ow = t.w0 + wi, oh = t.h0 + hi, ob = t.b0 + bi;
      const bool valid = ow < p.out_w && oh < p.out_h && ob < p.out_b;
      const long long row_off = (long long)ob * p.os_b + (long long)oh * p.os_h + (long long)ow * p.os_w + p.out_off[t.cls] + t.n0;
      float rw = 1.f;
      if constexpr (VARIANT != 0) {
        if (VARIANT == 3 && valid) rw = __ldg(p.row_w + ((long long)(ob * p.out_h + oh) * p.out_w + ow) * p.row_w_stride + t.cls);
        // columns of this tile that exist (class_n is a multiple of 16): chunks past them are skipped
        const int c_lim = min(c_end, (p.class_n - t.n0) >> 4);
        uint8_t* dst = reinterpret_cast<uint8_t*>(out) + (row_off + c_beg * 16) * (VARIANT == 2 ? 4 : 2);
        mbar_wait(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * cols_per_buf;
        uint64_t* ebar = &tmem_empty_bar[buf];
        lean_epilogue_tile<VARIANT>(taddr, c_beg, max(c_beg, c_lim), valid, s_scale[buf], s_shift[buf], slope_eff, rw, dst, p.dbg, [&]() {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(ebar);
        });
      } else {
      const long long row_off_up = (long long)ob * p.os_b * 2 + (long long)(2 * ow) * p.os_w + p.out_off[t.cls] + t.n0;
      float mw[MIX_MAX_K];
      if (p.row_w_mode != 0 && valid) {
        const float* wr = p.row_w + ((long long)(ob * p.out_h + oh) * p.out_w + ow) * p.row_w_stride;
        if (p.row_w_mode == 1) rw = __ldg(wr + t.cls);
        else {
#pragma unroll
          for (int k = 0; k < MIX_MAX_K; k++) mw[k] = k < p.mix_k ? __ldg(wr + k) : 0.f;
        }
      }
      mbar_wait(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * cols_per_buf;
      uint32_t va[16], vb[16];
      if (c_beg < c_end) tmem_ld16(taddr + (uint32_t)(c_beg * 16), va);
      for (int c = c_beg; c < c_end; c += 2) {
        // chunk c (in va), prefetching chunk c + 1 into vb; then chunk c + 1, prefetching c + 2 into va
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
          const int cc = c + sub;
          if (cc >= c_end) break;
          uint32_t (&cur)[16] = sub == 0 ? va : vb;
          uint32_t (&nxt)[16] = sub == 0 ? vb : va;
          tmem_wait_ld16(cur);
          if (cc + 1 < c_end) tmem_ld16(taddr + (uint32_t)((cc + 1) * 16), nxt);
          const int c0 = cc * 16;
          if (valid && (t.n0 + c0) < p.class_n) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = fmaf(__uint_as_float(cur[j]), s_scale[buf][c0 + j], s_shift[buf][c0 + j]);
            if (p.row_w_mode == 2) {
#pragma unroll
              for (int k = 0; k < MIX_MAX_K; k++) {
                if (k < p.mix_k) {
#pragma unroll
                  for (int j = 0; j < 16; j++) f[j] = fmaf(mw[k], s_mixb[k * MIX_MAX_N + t.n0 + c0 + j], f[j]);
                }
              }
            }
            if (act) {
#pragma unroll
              for (int j = 0; j < 16; j++) f[j] = f[j] > 0.f ? f[j] : f[j] * p.slope;
            }
            if (p.row_w_mode == 1) {
#pragma unroll
              for (int j = 0; j < 16; j++) f[j] *= rw;
            }
            epilogue_store16(p, out, f, row_off + c0, row_off_up + c0, nrep);
          }
        }
      }
      }
      if constexpr (VARIANT == 0) {
        // all tcgen05.ld of this buffer have completed (wait::ld above): hand it back to the MMA warp
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();         // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair form (tcgen05 cta_group::2): two CTAs on the SMs of one TPC compute a 256 x block_n tile together.
//
// A 128 x 256 tile stages 16 KB (A) + 32 KB (W) per 64-deep k-step = 96 B/clk per SM at full tensor rate, and the
// single-CTA kernel was observed taking in ~43 B/clk.  In a pair each CTA loads its own 128 activation rows and HALF of the
// weight tile (block_n/2 rows); the tensor cores read the other half from the peer's shared memory: 32 KB per k-step per
// SM = 64 B/clk, and a 6-deep instead of a 4-deep ring in the same shared memory.  Measured (profiles/
// r01_igemm_epilogue_experiments.txt): 36.8 -> 33.4 us on the M=65536 N=256 K=768 layers, 67 % tensor-active on the long-K
// audio layer; the K=768 layers stay near 47 % (wave quantisation + per-launch prologue, see the log's addendum).
//
//   * one UMMA = 256 x block_n x 16, issued by the leader CTA (cluster rank 0) only; accumulator rows 0..127 live in the
//     leader's TMEM, rows 128..255 in the peer's, at the same column offset (tcgen05.alloc.cta_group::2 in both CTAs)
//   * both producers signal the LEADER's full barrier (peer: TMA with .cta_group::2 + a remote arrive); the leader's
//     tcgen05.commit.cta_group::2 multicasts to both CTAs' empty / accumulator-full barriers
//   * each CTA's 8 epilogue warps drain that CTA's 128 rows (lean epilogue bodies only: VARIANT 1..3) and arrive on the
//     leader's accumulator-empty barrier (peer: remote arrive)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (cta-scope) semantics as CUTLASS's ClusterBarrier::arrive(cta_id): ".release.cluster" compiles to a GPU-scope
  // MEMBAR per arrive, which throttled the peer's producer to half speed; TMA data is tracked by complete_tx, the TMEM
  // hand-over by tcgen05.fence::before_thread_sync
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// EW = epilogue warps per TMEM lane quadrant: 2 (8 warps, 320 threads) or 4 (16 warps, 576 threads, 112 registers each).
template <int VARIANT, int EW>
__global__ void __launch_bounds__(64 + 128 * EW, 1)
igemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_w_lo,
                     const __grid_constant__ IgemmParams p, const float* __restrict__ bias, const float* __restrict__ scale,
                     const float* __restrict__ shift, void* __restrict__ out) {
  static_assert(VARIANT >= 1 && VARIANT <= 3, "pair kernel: lean epilogues only");
  constexpr int EPI_T = 128 * EW;                 // epilogue threads
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t b_half_bytes = (uint32_t)(p.block_n / 2) * BLOCK_K * 2;     // this CTA's half of the weight tile
  const int STAGES = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];      // used in the leader only
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];         // used in the leader only
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[2][256];
  __shared__ __align__(16) float s_shift[2][256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.ntaps * p.cchunks * p.npass;
  const int ny = p.n_tiles_per_class * p.num_classes;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  const int cid = (int)blockIdx.x / 2, ncl = (int)gridDim.x / 2;
  const int total_tiles = ((p.tiles_w * p.tiles_h * p.tiles_b + 1) / 2) * ny;      // pair tiles
  uint32_t cols_per_buf = 32;
  while (cols_per_buf < (uint32_t)p.block_n) cols_per_buf <<= 1;
  const uint32_t tmem_cols = 2 * cols_per_buf;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (p.npass > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_lo)) : "memory");
    }
    // full: the leader's arrive.expect_tx + the peer's remote arrive; empty / accumulator-full: one multicast commit;
    // accumulator-empty: the epilogue warps of both CTAs
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 2 * (EPI_T / 32)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own A rows + own half of W; completion lands on the leader's barrier
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int b_half_rows = p.block_n / 2;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const TileCoord t = decode_tile(p, tile, ny, 2, crank);
        const int tap_base = p.shared_taps ? 0 : t.cls * p.ntaps;
        const int chan_base = p.a_chan_base[t.cls];
        const int wrow = t.cls * p.class_n + t.n0 + crank * b_half_rows;
        for (int kg = 0; kg < num_k; kg++) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int kk = kg / p.npass, pass = kg - kk * p.npass;          // split-bf16: hi*hi, hi*lo, lo*hi
          const int tap = kk / p.cchunks, cc = kk - tap * p.cchunks;
          const short* tp = p.taps[tap_base + tap];
          const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * (A_STAGE_BYTES + b_half_bytes));
          tma_load_5d_pair(pass == 2 ? &map_a_lo : &map_a, lbar, smem_a + (size_t)s * A_STAGE_BYTES,
                           chan_base + tp[0] + cc * BLOCK_K, t.w0 + tp[1], tp[2], t.h0 + tp[3], t.b0);
          tma_load_2d_pair(pass == 1 ? &map_w_lo : &map_w, lbar, smem_b + (size_t)s * b_half_bytes, kk * BLOCK_K, wrow);
          if (!leader) mbar_arrive_cluster(lbar);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: leader only =================
    if (lane == 0 && leader) {
      // D=f32, A=B=bf16, K-major, N = block_n, M = 256 (pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl, lt++) {
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)buf * cols_per_buf;
        for (int ks = 0; ks < num_k; ks++) {
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + (size_t)s * A_STAGE_BYTES));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + (size_t)s * b_half_bytes));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++)
            umma_bf16_pair(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit_pair(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ================= epilogue (both CTAs): this CTA's 128 rows =================
    const int et = (int)threadIdx.x - 64;
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;                 // which share of the tile's columns
    const int r = q * 32 + lane;
    const int wi = r % p.box_w;
    const int hi = (r / p.box_w) % p.box_h;
    const int bi = r / (p.box_w * p.box_h);
    const int chunks = p.block_n >> 4;
    const int chunks_h = (chunks + EW - 1) / EW;
    const int c_beg = min(chunks, part * chunks_h), c_end = min(chunks, c_beg + chunks_h);
    const bool act = p.epilogue != 0 && p.slope != 1.f;
    const float slope_eff = act ? p.slope : 1.f;
    int lt = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl, lt++) {
      const int buf = lt & 1;
      const TileCoord t = decode_tile(p, tile, ny, 2, crank);
      const int ncol = t.cls * p.class_n + t.n0;
      for (int i = et; i < p.block_n; i += EPI_T) {
        float sc = 1.f, sh = 0.f;
        if (t.n0 + i < p.class_n) {
          if (p.epilogue == 1) { sc = __ldg(scale + ncol + i); sh = __ldg(shift + ncol + i); }
          else if (bias) sh = __ldg(bias + ncol + i);
        }
        s_scale[buf][i] = sc;
        s_shift[buf][i] = sh;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_T) : "memory");
      const int ow = t.w0 + wi, oh = t.h0 + hi, ob = t.b0 + bi;
      const bool valid = ow < p.out_w && oh < p.out_h && ob < p.out_b;
      const long long row_off = (long long)ob * p.os_b + (long long)oh * p.os_h + (long long)ow * p.os_w + p.out_off[t.cls] + t.n0;
      float rw = 1.f;
      if (VARIANT == 3 && valid) rw = __ldg(p.row_w + ((long long)(ob * p.out_h + oh) * p.out_w + ow) * p.row_w_stride + t.cls);
      const int c_lim = min(c_end, (p.class_n - t.n0) >> 4);
      uint8_t* dst = reinterpret_cast<uint8_t*>(out) + (row_off + c_beg * 16) * (VARIANT == 2 ? 4 : 2);
      mbar_wait(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * cols_per_buf;
      const uint32_t ebar = mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0);     // the leader's MMA warp waits on it
      lean_epilogue_tile<VARIANT, EW == 2>(taddr, c_beg, max(c_beg, c_lim), valid, s_scale[buf], s_shift[buf], slope_eff, rw, dst, p.dbg, [&]() {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ebar);
      });
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient:  dWp[q*class_n + n][t][c] = sum_rows dZ[row, off[q] + n] * X[row shifted by tap t, base[q] + c]
// Both operands are activations whose contraction index (the pixel row) is the slow index in
// memory, so they are fed as MN-major UMMA operands: a TMA box of 64 pixel rows x 64 channels
// lands as 64 swizzled 128-byte rows = the canonical MN-major SWIZZLE_128B layout (8-row atoms,
// SBO = 1024 B between 8-pixel groups, LBO = 8192 B between 64-channel chunks).  One CTA owns a
// (class, tap, 128 x Nc) tile of dWp and a slice of the pixel rows (split-K); partial tiles are
// combined with fp32 reductions in L2.
// ---------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 64;                       // pixel rows per k-step
constexpr uint32_t WG_CHUNK_BYTES = WG_ROWS * 64 * 2;   // 8 KB: 64 rows x 64 channels bf16
constexpr int WG_STAGES = 4;

struct WgradParams {
  int ntaps, cchunks, shared_taps, num_classes, class_n;
  int box_w, box_h, box_b, tiles_w, tiles_h, tiles_b;
  int n_tiles, c_tiles, kpad, split, npass;
  int c_tile;                            // x-channel columns per CTA tile: 64, 128, 192 or 256
  long long wp_numel;
  int a_chan_base[MS_IGEMM_MAX_CLASSES];
  int z_chan_base[MS_IGEMM_MAX_CLASSES];
  short taps[MS_IGEMM_MAX_TAPS][4];
};

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;      // LBO: next 64-channel chunk
  d |= (uint64_t)(1024 >> 4) << 32;      // SBO: next group of 8 pixel rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z,
                const __grid_constant__ CUtensorMap map_x_lo, const __grid_constant__ CUtensorMap map_z_lo,
                const __grid_constant__ WgradParams p, float* __restrict__ dwp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[WG_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int cls = blockIdx.z / p.ntaps, tap = blockIdx.z - cls * p.ntaps;
  const int nt = blockIdx.y / p.c_tiles, ct = blockIdx.y - nt * p.c_tiles;
  const int n0 = nt * 128, c0 = ct * p.c_tile;
  const int nc = min(p.c_tile, p.kpad - c0);            // columns of this tile (multiple of 64)
  const int xchunks = nc / 64;
  const uint32_t stage_bytes = (2 + xchunks) * WG_CHUNK_BYTES;
  const int total_rt = p.tiles_w * p.tiles_h * p.tiles_b;
  const int per = (total_rt + p.split - 1) / p.split;
  const int rt_beg = blockIdx.x * per, rt_end = min(total_rt, rt_beg + per);
  const int num_k = max(0, rt_end - rt_beg) * p.npass;
  uint8_t* smem_a = smem;                                // [stage][2 chunks]
  uint8_t* smem_b = smem + WG_STAGES * 2 * WG_CHUNK_BYTES;   // [stage][4 chunks]
  uint32_t tmem_cols = 64;
  while (tmem_cols < (uint32_t)nc) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_z)) : "memory");
    for (int s = 0; s < WG_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
          const int s = ks % WG_STAGES;
          const uint32_t ph = (uint32_t)(ks / WG_STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int pass = ks % p.npass;              // split-bf16: x_hi*z_hi, x_hi*z_lo, x_lo*z_hi
          int rt = rt_beg + ks / p.npass;
          const CUtensorMap* mz = pass == 1 ? &map_z_lo : &map_z;
          const CUtensorMap* mx = pass == 2 ? &map_x_lo : &map_x;
          const int tw = rt % p.tiles_w; rt /= p.tiles_w;
          const int th = rt % p.tiles_h; rt /= p.tiles_h;
          const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = rt * p.box_b;
          mbar_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* sa = smem_a + (size_t)s * 2 * WG_CHUNK_BYTES;
          uint8_t* sb = smem_b + (size_t)s * 4 * WG_CHUNK_BYTES;
          tma_load_5d(mz, &full_bar[s], sa, zc, w0, 0, h0, b0);
          tma_load_5d(mz, &full_bar[s], sa + WG_CHUNK_BYTES, zc + 64, w0, 0, h0, b0);
          for (int i = 0; i < xchunks; i++)
            tma_load_5d(mx, &full_bar[s], sb + (size_t)i * WG_CHUNK_BYTES, xc + 64 * i, w0 + t[1], t[2], h0 + t[3], b0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // D=f32, A=B=bf16, both MN-major (bits 15/16), N = nc, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < num_k; ks++) {
          const int s = ks % WG_STAGES;
          const uint32_t ph = (uint32_t)(ks / WG_STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_mnmajor_sw128_desc(smem_u32(smem_a + (size_t)s * 2 * WG_CHUNK_BYTES));
          const uint64_t db = make_mnmajor_sw128_desc(smem_u32(smem_b + (size_t)s * 4 * WG_CHUNK_BYTES));
#pragma unroll
          for (int k = 0; k < WG_ROWS / UMMA_K; k++) {
            // 16 pixel rows = 2 atoms of 1024 B: +2048 B = +128 in the (addr >> 4) field
            umma_bf16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;                     // n index inside the tile
      const bool valid = (n0 + r) < p.class_n;
      // split-K partials: slice blockIdx.x writes its own copy of dWp (summed by ms_unpack_igemm_wgrad)
      float* dst_row = dwp + (size_t)blockIdx.x * p.wp_numel + ((size_t)(cls * p.class_n + n0 + r) * p.ntaps + tap) * p.kpad + c0;
      mbar_wait(&tmem_full_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int cc = 0; cc < nc; cc += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)cc, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          float4* d4 = reinterpret_cast<float4*>(dst_row + cc);
#pragma unroll
          for (int j = 0; j < 4; j++)
            d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                __uint_as_float(v[4 * j + 3]));
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

// dWp[row][t][c] fp32 -> dw (Cout, Cin_g, taps_total) in dtype pdt (inverse of the forward re-tiling)
__global__ void unpack_igemm_wgrad_kernel(const float* __restrict__ dwp, int Cout, int Cin_g, int taps_total, int ntaps, int kpad,
                                          void* __restrict__ dw, int pdt, int nsplit, long long wp_numel, int accumulate) {
  const long long total = (long long)Cout * Cin_g * taps_total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps_total);
    long long t2 = i / taps_total;
    int c = (int)(t2 % Cin_g);
    long long o = t2 / Cin_g;
    const float* src = dwp + (o * ntaps + tap) * kpad + c;
    double acc = accumulate ? ms_ldp_d(dw, pdt, i) : 0.0;
    for (int sidx = 0; sidx < nsplit; sidx++) acc += (double)src[(long long)sidx * wp_numel];
    ms_stp(dw, pdt, i, acc);
  }
}

struct PackParams {
  int Cout, Cin_g, taps_total, groups, mode, num_classes, class_n, ntaps, kpad;
  short srctap[MS_IGEMM_MAX_TAPS];
};

__global__ void pack_igemm_weight_kernel(const void* __restrict__ w, int pdt, PackParams q, __nv_bfloat16* __restrict__ wp,
                                         __nv_bfloat16* __restrict__ wp_lo) {
  const long long total = (long long)q.num_classes * q.class_n * q.ntaps * q.kpad;
  const int Cout_g = q.Cout / q.groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kc = (int)(i % q.kpad);
    long long t2 = i / q.kpad;
    int t = (int)(t2 % q.ntaps);
    long long row = t2 / q.ntaps;
    int cls = (int)(row / q.class_n), r = (int)(row - (long long)cls * q.class_n);
    float v = 0.f;
    if (q.mode == 0) {
      long long o = row;                    // global output channel
      if (o < q.Cout && kc < q.Cin_g) v = ms_ldp(w, pdt, (o * q.Cin_g + kc) * q.taps_total + q.srctap[t]);
    } else {
      int g = q.groups > 1 ? cls : 0;
      if (kc < Cout_g && r < q.Cin_g)
        v = ms_ldp(w, pdt, ((long long)(g * Cout_g + kc) * q.Cin_g + r) * q.taps_total + q.srctap[cls * q.ntaps + t]);
    }
    __nv_bfloat16 h = __float2bfloat16(v);
    wp[i] = h;
    if (wp_lo) wp_lo[i] = __float2bfloat16(v - __bfloat162float(h));
  }
}

// Table-driven form: ONE launch re-tiles every weight of a sub-network (the train step refreshes all packed copies right after
// the optimiser update instead of ~80 separate launches).  Work units of ~PACK_UNIT packed elements are spread over the
// entries in proportion to their size; a thread converts two adjacent k positions and stores bf16 pairs.
constexpr int PACK_UNIT = 16384;
constexpr int PACK_UNIT_MIN = 2048;
constexpr int PACK_MAX_ENTRIES = 256;
// sum of a table in shared memory by warp 0 (the caller synchronised after filling it); every thread gets the result
__device__ __forceinline__ long long pack_table_sum(const int* __restrict__ v, int n, long long* s_out) {
  if (threadIdx.x < 32) {
    long long a = 0;
    for (int i = threadIdx.x; i < n; i += 32) a += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (threadIdx.x == 0) *s_out = a;
  }
  __syncthreads();
  const long long r = *s_out;
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(256) pack_igemm_weight_multi_kernel(const ms_pack_entry* __restrict__ table, int n_entries) {
  __shared__ int s_units[PACK_MAX_ENTRIES];
  __shared__ long long s_sum;
  const int tid = threadIdx.x;
  for (int i = tid; i < n_entries; i += blockDim.x)
    s_units[i] = table[i].num_classes * table[i].class_n * table[i].ntaps * table[i].kpad;      // packed elements (< 2^31)
  __syncthreads();
  // unit size: ~half a unit per CTA, between PACK_UNIT_MIN and PACK_UNIT elements (small tables spread over more CTAs)
  const long long elems = pack_table_sum(s_units, n_entries, &s_sum);
  long long unit_ll = 2 * elems / gridDim.x;
  const unsigned unit = (unsigned)(unit_ll < PACK_UNIT_MIN ? PACK_UNIT_MIN : (unit_ll > PACK_UNIT ? PACK_UNIT : unit_ll));
  for (int i = tid; i < n_entries; i += blockDim.x) s_units[i] = (int)(((unsigned)s_units[i] + unit - 1) / unit);
  __syncthreads();
  const int total_units = (int)pack_table_sum(s_units, n_entries, &s_sum);
  for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
    int ei = 0, local = u;
    while (ei < n_entries && local >= s_units[ei]) { local -= s_units[ei]; ei++; }
    const ms_pack_entry& e = table[ei];
    const int nu = s_units[ei];
    // Eight adjacent k positions per thread: one index decomposition, eight strided loads, one 16-byte store per plane.
    // The entry's fields are copied to registers first (the stores go through pointers read from the table, so the
    // compiler would otherwise re-read every field after every store); 32-bit index arithmetic (a packed weight has far
    // fewer than 2^31 elements).
    const int Cout = e.Cout, Cin_g = e.Cin_g, taps_total = e.taps_total, mode = e.mode, pdt = e.pdt;
    const int Cout_g = Cout / e.groups, grouped = e.groups > 1;
    const unsigned kpad8 = (unsigned)e.kpad >> 3, ntaps = (unsigned)e.ntaps, class_n = (unsigned)e.class_n;      // kpad % 64 == 0
    const long long octs = ((long long)e.num_classes * e.class_n * e.ntaps * e.kpad) >> 3;
    const long long q0 = octs * local / nu, q1 = octs * (local + 1) / nu;
    const void* __restrict__ w = e.w;
    uint4* __restrict__ wp = reinterpret_cast<uint4*>(e.wp);
    uint4* __restrict__ wp_lo = reinterpret_cast<uint4*>(e.wp_lo);
    const short* __restrict__ srctap = e.srctap;
    if (((uintptr_t)wp | (uintptr_t)wp_lo) & 15) __trap();          // 16-byte stores
    for (long long qi = q0 + tid; qi < q1; qi += blockDim.x) {
      const unsigned qu = (unsigned)qi;
      const unsigned t2 = qu / kpad8;
      const int kc = (int)(qu - t2 * kpad8) << 3;
      const unsigned rowu = t2 / ntaps;
      const int t = (int)(t2 - rowu * ntaps);
      long long base = 0, stride = 0;
      int nvalid = 0;
      if (mode == 0) {
        if ((int)rowu < Cout && kc < Cin_g) {
          base = ((long long)rowu * Cin_g + kc) * taps_total + __ldg(&srctap[t]);
          stride = taps_total;
          nvalid = min(8, Cin_g - kc);
        }
      } else {
        const int cls = (int)(rowu / class_n), r = (int)(rowu - (unsigned)cls * class_n);
        if (kc < Cout_g && r < Cin_g) {
          base = ((long long)((grouped ? cls : 0) * Cout_g + kc) * Cin_g + r) * taps_total + __ldg(&srctap[cls * (int)ntaps + t]);
          stride = (long long)Cin_g * taps_total;
          nvalid = min(8, Cout_g - kc);
        }
      }
      float v[8];
      if (pdt == MS_F64) {
        const double* __restrict__ wd = reinterpret_cast<const double*>(w) + base;
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = j < nvalid ? (float)__ldg(wd + j * stride) : 0.f;
      } else {
        const float* __restrict__ wf = reinterpret_cast<const float*>(w) + base;
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = j < nvalid ? __ldg(wf + j * stride) : 0.f;
      }
      __nv_bfloat162 h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        l[j] = __floats2bfloat162_rn(v[2 * j] - __bfloat162float(h[j].x), v[2 * j + 1] - __bfloat162float(h[j].y));
      }
      wp[qi] = *reinterpret_cast<const uint4*>(h);
      if (wp_lo) wp_lo[qi] = *reinterpret_cast<const uint4*>(l);
    }
  }
}

}  // namespace

extern "C" int ms_pack_igemm_weight_multi(const ms_pack_entry* table_dev, int n_entries, int blocks, void* stream) {
  if (!table_dev || n_entries < 1 || n_entries > PACK_MAX_ENTRIES) return MS_EINVAL;
  if (blocks < 1) blocks = 8 * ms_num_sms();
  pack_igemm_weight_multi_kernel<<<dim3((unsigned)blocks), 256, 0, ms_stream(stream)>>>(table_dev, n_entries);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_pack_igemm_weight_bf16(const void* w, int pdt, int Cout, int Cin_g, int taps_total, int groups, int mode,
                                         int num_classes, int class_n, int ntaps, int kpad, const int16_t* srctap_host,
                                         void* wp, void* wp_lo, void* stream) {
  if (!w || !wp || !srctap_host || groups < 1 || num_classes < 1 || ntaps < 1) return MS_EINVAL;
  const int nsrc = mode == 0 ? ntaps : num_classes * ntaps;
  if (nsrc > MS_IGEMM_MAX_TAPS || kpad % 64) return MS_EINVAL;
  PackParams q;
  q.Cout = Cout; q.Cin_g = Cin_g; q.taps_total = taps_total; q.groups = groups; q.mode = mode;
  q.num_classes = num_classes; q.class_n = class_n; q.ntaps = ntaps; q.kpad = kpad;
  for (int i = 0; i < MS_IGEMM_MAX_TAPS; i++) q.srctap[i] = i < nsrc ? srctap_host[i] : 0;
  const long long total = (long long)num_classes * class_n * ntaps * kpad;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_igemm_weight_kernel<<<(unsigned)blocks, 256, 0, ms_stream(stream)>>>(w, pdt, q, reinterpret_cast<__nv_bfloat16*>(wp),
                                                                               reinterpret_cast<__nv_bfloat16*>(wp_lo));
  MS_LAUNCH_CHECK();
  return 0;
}

struct IgemmFused {
  float* out_f32;
  const void* res;
  int res_planes;
  long long res_pstride;
  int up2;
  const float* row_w;
  int row_w_stride, row_w_mode, mix_k;
};

// MS_IGEMM_LEGACY=1 routes every launch to the one-tile-per-CTA kernel (A/B timing, debugging)
static bool igemm_legacy() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MS_IGEMM_LEGACY");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// MS_IGEMM_GENERIC_EPILOGUE=1 keeps the persistent kernel on its generic epilogue body (A/B timing, debugging)
static bool igemm_generic_epilogue() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MS_IGEMM_GENERIC_EPILOGUE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// MS_IGEMM_CLUSTER=1|2|4 forces the cluster size where the geometry allows it (A/B timing)
static int igemm_cluster_size(long long m_tiles, long long ny, int block_n) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MS_IGEMM_CLUSTER");
    forced = e ? atoi(e) : 0;
  }
  // Default 1: measured on B200 (profiles/r01_cluster_multicast_ab.txt) the multicast does not pay: L2 slice reads are not
  // the bound of these launches (nor, as the CTA-pair kernel later showed, is the per-SM operand ingest).
  (void)ny;
  int c = 1;
  if (forced == 1 || forced == 2 || forced == 4) c = forced;
  while (c > 1 && (block_n % (8 * c) || m_tiles < c)) c >>= 1;     // weight slices are whole 8-row swizzle atoms
  return c;
}

// CTA pairs (igemm_tc_pair_kernel) for GEMMs with enough 128-row tiles to keep every SM pair busy; MS_IGEMM_PAIR=0 disables,
// MS_IGEMM_PAIR=2 forces them wherever the geometry allows (tests)
static bool igemm_use_pair(long long m_tiles, long long ny, int block_n) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("MS_IGEMM_PAIR");
    mode = e ? atoi(e) : 1;
  }
  if (mode == 0 || block_n % 32 || block_n < 32 || m_tiles < 2) return false;
  if (mode == 2) return true;
  return m_tiles * ny >= 2LL * ms_num_sms();
}

// MS_IGEMM_EPI16=1: 16 epilogue warps in the pair kernel (experiment; default 8)
static bool igemm_epi16() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MS_IGEMM_EPI16");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int V, int EW>
static int launch_pair(long long groups, size_t smem, cudaStream_t cs, const CUtensorMap& map_a, const CUtensorMap& map_w,
                       const CUtensorMap& map_a_lo, const CUtensorMap& map_w_lo, const IgemmParams& p, const float* bias,
                       const float* scale, const float* shift, void* out) {
  static int max_pairs = -1;
  const int dyn = 227 * 1024 - 7 * 1024 + 1024;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(64 + 128 * EW); cfg.stream = cs;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (max_pairs < 0) {
    MS_CUDA(cudaFuncSetAttribute(igemm_tc_pair_kernel<V, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int n = ms_num_sms() / 2;
    cfg.gridDim = dim3((unsigned)(ms_num_sms() / 2 * 2)); cfg.dynamicSmemBytes = dyn;
    if (cudaOccupancyMaxActiveClusters(&n, igemm_tc_pair_kernel<V, EW>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 1;
    }
    if (n > ms_num_sms() / 2) n = ms_num_sms() / 2;
    max_pairs = n;
  }
  const long long ncl = groups < max_pairs ? groups : max_pairs;
  cfg.gridDim = dim3((unsigned)(ncl * 2));
  cfg.dynamicSmemBytes = smem;
  MS_CUDA(cudaLaunchKernelEx(&cfg, igemm_tc_pair_kernel<V, EW>, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out));
  return 0;
}

template <int V, int C>
static int launch_persist(long long groups, size_t smem, cudaStream_t cs, const CUtensorMap& map_a, const CUtensorMap& map_w,
                          const CUtensorMap& map_a_lo, const CUtensorMap& map_w_lo, const IgemmParams& p, const float* bias,
                          const float* scale, const float* shift, void* out) {
  static int max_clusters = -1;
  const int dyn = 227 * 1024 - 15 * 1024 + 1024;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(PERSIST_THREADS); cfg.stream = cs;
  cfg.attrs = at; cfg.numAttrs = C > 1 ? 1 : 0;
  if (max_clusters < 0) {
    MS_CUDA(cudaFuncSetAttribute(igemm_tc_persist_kernel<V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int n = ms_num_sms() / C;
    if (C > 1) {
      cfg.gridDim = dim3((unsigned)(ms_num_sms() / C * C)); cfg.dynamicSmemBytes = dyn;
      if (cudaOccupancyMaxActiveClusters(&n, igemm_tc_persist_kernel<V, C>, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 1;
      }
      if (n > ms_num_sms() / C) n = ms_num_sms() / C;
    }
    max_clusters = n;
  }
  const long long ncl = groups < max_clusters ? groups : max_clusters;
  cfg.gridDim = dim3((unsigned)(ncl * C));
  cfg.dynamicSmemBytes = smem;
  MS_CUDA(cudaLaunchKernelEx(&cfg, igemm_tc_persist_kernel<V, C>, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out));
  return 0;
}

static int igemm_launch(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                        const float* shift, void* out, const IgemmFused* fx, void* stream) {
  if (!d || !a || !w || !out) return MS_EINVAL;
  if (d->num_classes < 1 || d->num_classes > MS_IGEMM_MAX_CLASSES) return MS_EINVAL;
  if (d->ntaps < 1 || d->cchunks < 1) return MS_EINVAL;
  const int tap_rows = d->shared_taps ? d->ntaps : d->ntaps * d->num_classes;
  if (tap_rows > MS_IGEMM_MAX_TAPS) return MS_EINVAL;
  if (d->block_n < 16 || d->block_n > 256 || d->block_n % 16) return MS_EINVAL;
  if (d->class_n % 16 || d->class_n < 16) return MS_EINVAL;
  if (d->box[0] != BLOCK_K || d->box[2] != 1 || d->box[1] * d->box[3] * d->box[4] != BLOCK_M) return MS_EINVAL;
  if (d->epilogue == 1 && (!scale || !shift)) return MS_EINVAL;
  if (d->epilogue < 0 || d->epilogue > 2) return MS_EINVAL;
  if (d->out_dtype != MS_F32 && d->out_dtype != MS_BF16 && d->out_dtype != MS_BF16X2) return MS_EINVAL;
  if (d->out_dtype == MS_BF16X2 && (d->out_plane_stride <= 0 || (d->out_plane_stride * 2) % 16)) return MS_EINVAL;
  if (((uintptr_t)a & 15) || ((uintptr_t)w & 15) || ((uintptr_t)out & 15)) return MS_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return MS_ENOTSUP;

  // epilogue body of the persistent kernels: lean variants (see epilogue_chunk_fast) where the layout allows, else generic
  int variant = 0;
  {
    const int esz0 = d->out_dtype == MS_F32 ? 4 : 2;
    const float slope_eff = (d->epilogue != 0 && d->slope != 1.f) ? d->slope : 1.f;
    bool al32 = ((uintptr_t)out & 31) == 0 && (d->out_strides[0] * esz0) % 32 == 0 && (d->out_strides[1] * esz0) % 32 == 0 &&
                (d->out_strides[2] * esz0) % 32 == 0;
    for (int i = 0; i < d->num_classes; i++) al32 = al32 && (d->out_off[i] * esz0) % 32 == 0;
    const bool up2 = fx && fx->up2;
    const bool extra_f32 = fx && fx->out_f32 && d->out_dtype != MS_F32;
    const int rwm = fx ? fx->row_w_mode : 0;
    if (al32 && !up2 && !extra_f32 && slope_eff >= 0.f && slope_eff <= 1.f) {
      if (d->out_dtype == MS_BF16 && rwm == 0) variant = 1;
      else if (d->out_dtype == MS_F32 && rwm == 0) variant = 2;
      else if (d->out_dtype == MS_BF16 && rwm == 1) variant = 3;
    }
    if (igemm_generic_epilogue()) variant = 0;
  }
  // thread-block cluster size of the persistent kernel (weight-tile multicast, opt-in) / CTA pairs (cta_group::2)
  int csz = 1;
  bool pair = false;
  if (d->split_k <= 1 && !igemm_legacy()) {
    const long long mt = (long long)((d->out_dims[0] + d->box[1] - 1) / d->box[1]) * ((d->out_dims[1] + d->box[3] - 1) / d->box[3]) *
                         ((d->out_dims[2] + d->box[4] - 1) / d->box[4]);
    const long long ny = (long long)((d->class_n + d->block_n - 1) / d->block_n) * d->num_classes;
    pair = variant != 0 && igemm_use_pair(mt, ny, d->block_n);
    csz = pair ? 2 : igemm_cluster_size(mt, ny, d->block_n);
  }

  const int planes = d->planes == 2 ? 2 : 1;
  if (d->planes != 1 && d->planes != 2) return MS_EINVAL;
  if (planes == 2 && ((d->a_plane_stride * 2) % 16 || (d->w_plane_stride * 2) % 16 || d->a_plane_stride <= 0 || d->w_plane_stride <= 0))
    return MS_EINVAL;
  CUtensorMap map_a, map_w, map_a_lo, map_w_lo;
  for (int pl = 0; pl < planes; pl++) {
    const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(a) + (pl ? d->a_plane_stride : 0);
    const __nv_bfloat16* wq = reinterpret_cast<const __nv_bfloat16*>(w) + (pl ? d->w_plane_stride : 0);
    {
      cuuint64_t dims[5], strides[4];
      cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
      for (int i = 0; i < 5; i++) { dims[i] = (cuuint64_t)d->a_dims[i]; box[i] = (cuuint32_t)d->box[i]; }
      for (int i = 1; i < 5; i++) {
        strides[i - 1] = (cuuint64_t)d->a_strides[i] * 2;
        if (strides[i - 1] % 16) return MS_EINVAL;
      }
      CUresult r = enc(pl ? &map_a_lo : &map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(ap), dims, strides,
                       box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return MS_EINVAL;
    }
    {
      const long long ktot = (long long)d->ntaps * d->cchunks * BLOCK_K;
      cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)((long long)d->num_classes * d->class_n)};
      cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
      cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(d->block_n / csz)}, es[2] = {1, 1};
      CUresult r = enc(pl ? &map_w_lo : &map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wq), dims, strides,
                       box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return MS_EINVAL;
    }
  }
  if (planes == 1) { map_a_lo = map_a; map_w_lo = map_w; }
  IgemmParams p;
  p.npass = planes == 2 ? 3 : 1;
  p.out_plane_stride = d->out_plane_stride;
  p.ntaps = d->ntaps; p.cchunks = d->cchunks; p.shared_taps = d->shared_taps;
  p.num_classes = d->num_classes; p.class_n = d->class_n; p.block_n = d->block_n;
  p.n_tiles_per_class = (d->class_n + d->block_n - 1) / d->block_n;
  p.box_w = d->box[1]; p.box_h = d->box[3]; p.box_b = d->box[4];
  p.out_w = d->out_dims[0]; p.out_h = d->out_dims[1]; p.out_b = d->out_dims[2];
  p.tiles_w = (p.out_w + p.box_w - 1) / p.box_w;
  p.tiles_h = (p.out_h + p.box_h - 1) / p.box_h;
  p.tiles_b = (p.out_b + p.box_b - 1) / p.box_b;
  p.os_w = d->out_strides[0]; p.os_h = d->out_strides[1]; p.os_b = d->out_strides[2];
  for (int i = 0; i < MS_IGEMM_MAX_CLASSES; i++) { p.a_chan_base[i] = d->a_chan_base[i]; p.out_off[i] = d->out_off[i]; }
  for (int i = 0; i < MS_IGEMM_MAX_TAPS; i++)
    for (int j = 0; j < 4; j++) p.taps[i][j] = d->taps[i][j];
  p.out_dtype = d->out_dtype; p.epilogue = d->epilogue; p.slope = d->slope;
  p.out_f32 = nullptr; p.res = nullptr; p.res_pstride = 0; p.res_planes = 0; p.up2 = 0;
  p.row_w = nullptr; p.row_w_stride = 0; p.row_w_mode = 0; p.mix_k = 0;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("MS_IGEMM_DBG");
      dbg = e ? atoi(e) : 0;
    }
    p.dbg = dbg;
  }
  if (fx) {
    if (d->split_k > 1) return MS_EINVAL;
    if (fx->row_w_mode) {
      if (!fx->row_w || fx->up2 || fx->row_w_stride < 1) return MS_EINVAL;
      if (fx->row_w_mode == 1) {
        if (fx->row_w_stride < d->num_classes) return MS_EINVAL;
      } else if (fx->row_w_mode == 2) {
        // mixed bias: bias is [mix_k][num_classes*class_n]; one tile covers every column
        if (fx->mix_k < 1 || fx->mix_k > MIX_MAX_K || fx->row_w_stride < fx->mix_k) return MS_EINVAL;
        if (d->num_classes * d->class_n > MIX_MAX_N || d->epilogue != 0) return MS_EINVAL;
      } else {
        return MS_EINVAL;
      }
      p.row_w = fx->row_w; p.row_w_stride = fx->row_w_stride; p.row_w_mode = fx->row_w_mode; p.mix_k = fx->mix_k;
    }
    if (fx->out_f32 && ((uintptr_t)fx->out_f32 & 15)) return MS_EINVAL;
    p.out_f32 = d->out_dtype == MS_F32 ? nullptr : fx->out_f32;
    if (fx->up2) {
      // 1-D only: out_strides describe ONE GEMM row per (b, w); the written tensor has 2*out_w rows per sequence
      if (d->out_dims[1] != 1 || !fx->res || (fx->res_planes != 1 && fx->res_planes != 2) || ((uintptr_t)fx->res & 15)) return MS_EINVAL;
      if (fx->res_planes == 2 && (fx->res_pstride <= 0 || (fx->res_pstride * 2) % 16)) return MS_EINVAL;
      p.up2 = 1; p.res = reinterpret_cast<const __nv_bfloat16*>(fx->res); p.res_planes = fx->res_planes; p.res_pstride = fx->res_pstride;
    }
  }
  // vector stores need 16-byte aligned rows
  const int esz = d->out_dtype == MS_F32 ? 4 : 2;
  if ((p.os_w * esz) % 16 || (p.os_h * esz) % 16 || (p.os_b * esz) % 16) return MS_EINVAL;
  for (int i = 0; i < d->num_classes; i++)
    if ((p.out_off[i] * esz) % 16) return MS_EINVAL;

  const size_t stage_bytes = A_STAGE_BYTES + (size_t)d->block_n * BLOCK_K * 2;
  const int num_k_total = d->ntaps * d->cchunks * p.npass;
  // split-K: slices of k-steps reduced with fp32 vector reductions into a zero-filled fp32 output
  int split = d->split_k > 1 ? d->split_k : 1;
  if (split > 1) {
    if (d->out_dtype != MS_F32 || d->epilogue != 0 || d->out_numel <= 0) return MS_EINVAL;
    if (split > num_k_total) split = num_k_total;
    const int per = (num_k_total + split - 1) / split;
    split = (num_k_total + per - 1) / per;            // every slice owns at least one k-step
  }
  p.split_k = split;
  const long long tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles_per_class * d->num_classes;
  if (tiles < 1 || tiles > 0x7fffffffLL) return MS_EINVAL;
  if (split == 1 && !(igemm_legacy() && p.row_w_mode == 0)) {
    // ---- persistent kernel: one CTA per SM, double-buffered TMEM accumulators, ring across tiles
    int stages = (int)((227 * 1024 - 15 * 1024) / stage_bytes);      // static smem: constants + mixed bias + barriers
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return MS_EINVAL;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    if (pair) {
      // CTA pairs: each CTA stages 16 KB of A + its half of the weight tile per k-step
      const size_t pstage = A_STAGE_BYTES + (size_t)(d->block_n / 2) * BLOCK_K * 2;
      int pst = (int)((227 * 1024 - 7 * 1024) / pstage);
      if (pst > MAX_STAGES) pst = MAX_STAGES;
      p.stages = pst;
      const long long pgroups = (((long long)p.tiles_w * p.tiles_h * p.tiles_b + 1) / 2) * p.n_tiles_per_class * d->num_classes;
      const size_t psmem = (size_t)pst * pstage + 1024;
      cudaStream_t pcs = ms_stream(stream);
      if (igemm_epi16()) {
        switch (variant) {
          case 1: return launch_pair<1, 4>(pgroups, psmem, pcs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
          case 2: return launch_pair<2, 4>(pgroups, psmem, pcs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
          case 3: return launch_pair<3, 4>(pgroups, psmem, pcs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
          default: return MS_EINVAL;
        }
      }
      switch (variant) {
        case 1: return launch_pair<1, 2>(pgroups, psmem, pcs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
        case 2: return launch_pair<2, 2>(pgroups, psmem, pcs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
        case 3: return launch_pair<3, 2>(pgroups, psmem, pcs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
        default: return MS_EINVAL;
      }
    }
    cudaStream_t cs = ms_stream(stream);
    const long long groups = (((long long)p.tiles_w * p.tiles_h * p.tiles_b + csz - 1) / csz) * p.n_tiles_per_class * d->num_classes;
    int rc = MS_EINVAL;
#define MS_PERSIST_CASE(V, C) \
  case (V) * 8 + (C): rc = launch_persist<V, C>(groups, smem, cs, map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out); break;
    switch (variant * 8 + csz) {
      MS_PERSIST_CASE(0, 1) MS_PERSIST_CASE(0, 2) MS_PERSIST_CASE(0, 4)
      MS_PERSIST_CASE(1, 1) MS_PERSIST_CASE(1, 2) MS_PERSIST_CASE(1, 4)
      MS_PERSIST_CASE(2, 1) MS_PERSIST_CASE(2, 2) MS_PERSIST_CASE(2, 4)
      MS_PERSIST_CASE(3, 1) MS_PERSIST_CASE(3, 2) MS_PERSIST_CASE(3, 4)
      default: break;
    }
#undef MS_PERSIST_CASE
    return rc;
  }
  if (p.row_w_mode != 0) return MS_EINVAL;
  const int k_per_cta = (num_k_total + split - 1) / split;
  int stages = (int)((227 * 1024 - 4096) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > k_per_cta) stages = k_per_cta < 2 ? 2 : k_per_cta;
  {
    // more CTAs than SMs: keep the ring under half of the shared memory so that two CTAs are co-resident
    const long long ctas = tiles * split;
    const int half = (int)((113 * 1024 - 2048) / stage_bytes);
    if (ctas > ms_num_sms() && half >= 3 && stages > half) stages = half;
  }
  if (stages < 2) return MS_EINVAL;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  if (split > 1) MS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)d->out_numel, ms_stream(stream)));
  dim3 grid((unsigned)(p.tiles_w * p.tiles_h * p.tiles_b), (unsigned)(p.n_tiles_per_class * d->num_classes), (unsigned)split);
  igemm_tc_kernel<<<grid, NUM_THREADS, smem, ms_stream(stream)>>>(map_a, map_w, map_a_lo, map_w_lo, p, bias, scale, shift, out);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_igemm_bf16(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                             const float* shift, void* out, void* stream) {
  return igemm_launch(d, a, w, bias, scale, shift, out, nullptr, stream);
}

extern "C" int ms_igemm_bf16_fused(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                                   const float* shift, void* out, float* out_f32, const void* res, int res_planes,
                                   int64_t res_pstride, int up2, void* stream) {
  IgemmFused fx;
  fx.out_f32 = out_f32; fx.res = res; fx.res_planes = res_planes; fx.res_pstride = res_pstride; fx.up2 = up2;
  fx.row_w = nullptr; fx.row_w_stride = 0; fx.row_w_mode = 0; fx.mix_k = 0;
  return igemm_launch(d, a, w, bias, scale, shift, out, &fx, stream);
}

extern "C" int ms_igemm_bf16_mix(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                                 const float* shift, void* out, float* out_f32, const float* row_w, int row_w_stride,
                                 int row_w_mode, int mix_k, void* stream) {
  IgemmFused fx;
  fx.out_f32 = out_f32; fx.res = nullptr; fx.res_planes = 0; fx.res_pstride = 0; fx.up2 = 0;
  fx.row_w = row_w; fx.row_w_stride = row_w_stride; fx.row_w_mode = row_w_mode; fx.mix_k = mix_k;
  return igemm_launch(d, a, w, bias, scale, shift, out, &fx, stream);
}

static int encode_5d(EncodeTiledFn enc, CUtensorMap* m, const void* base, const int32_t* dims, const int64_t* strides_el,
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

extern "C" int ms_wgrad_bf16(const ms_igemm_desc* d, const void* x, const void* dz, float* dwp, void* stream) {
  if (!d || !x || !dz || !dwp) return MS_EINVAL;
  if (d->num_classes < 1 || d->num_classes > MS_IGEMM_MAX_CLASSES || d->ntaps < 1 || d->cchunks < 1) return MS_EINVAL;
  if (((uintptr_t)x & 15) || ((uintptr_t)dz & 15) || ((uintptr_t)dwp & 15)) return MS_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return MS_ENOTSUP;
  WgradParams p;
  p.ntaps = d->ntaps; p.cchunks = d->cchunks; p.shared_taps = d->shared_taps;
  p.num_classes = d->num_classes; p.class_n = d->class_n;
  // 64-row boxes: halve the outermost non-unit dimension of the forward (128-row) box
  int bw = d->box[1], bh = d->box[3], bb = d->box[4];
  if (bb > 1) bb /= 2; else if (bh > 1) bh /= 2; else bw /= 2;
  if (bw * bh * bb != WG_ROWS) return MS_EINVAL;
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
  split = (total_rt + per - 1) / per;                 // every slice owns at least one row tile
  if (split != (d->split_k > 1 ? d->split_k : 1)) return MS_EINVAL;   // the caller sized dwp for exactly split_k partials
  p.split = (int)split;
  p.wp_numel = (long long)d->num_classes * d->class_n * d->ntaps * p.kpad;
  if (d->planes != 1 && d->planes != 2) return MS_EINVAL;
  p.npass = d->planes == 2 ? 3 : 1;
  if (d->planes == 2 && (d->a_plane_stride <= 0 || d->out_plane_stride <= 0 || (d->a_plane_stride * 2) % 16 || (d->out_plane_stride * 2) % 16))
    return MS_EINVAL;

  CUtensorMap map_x, map_z, map_x_lo, map_z_lo;
  int box[5] = {64, bw, 1, bh, bb};
  const int32_t zdims[5] = {(int32_t)d->out_strides[0], Wo, 1, Ho, Bo};
  const int64_t zstr[5] = {1, d->out_strides[0], d->out_strides[1], d->out_strides[1], d->out_strides[2]};
  int rc = encode_5d(enc, &map_x, x, d->a_dims, d->a_strides, box);
  if (rc) return rc;
  rc = encode_5d(enc, &map_z, dz, zdims, zstr, box);
  if (rc) return rc;
  if (d->planes == 2) {
    rc = encode_5d(enc, &map_x_lo, reinterpret_cast<const __nv_bfloat16*>(x) + d->a_plane_stride, d->a_dims, d->a_strides, box);
    if (rc) return rc;
    rc = encode_5d(enc, &map_z_lo, reinterpret_cast<const __nv_bfloat16*>(dz) + d->out_plane_stride, zdims, zstr, box);
    if (rc) return rc;
  } else {
    map_x_lo = map_x; map_z_lo = map_z;
  }

  const size_t smem = (size_t)WG_STAGES * 6 * WG_CHUNK_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  dim3 grid((unsigned)p.split, (unsigned)(p.n_tiles * p.c_tiles), (unsigned)(d->num_classes * d->ntaps));
  wgrad_tc_kernel<<<grid, NUM_THREADS, smem, ms_stream(stream)>>>(map_x, map_z, map_x_lo, map_z_lo, p, dwp);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_unpack_igemm_wgrad(const float* dwp, int Cout, int Cin_g, int taps_total, int ntaps, int kpad, void* dw,
                                     int pdt, int nsplit, int accumulate, void* stream) {
  if (!dwp || !dw || ntaps != taps_total || nsplit < 1) return MS_EINVAL;
  const long long total = (long long)Cout * Cin_g * taps_total;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  unpack_igemm_wgrad_kernel<<<(unsigned)blocks, 256, 0, ms_stream(stream)>>>(dwp, Cout, Cin_g, taps_total, ntaps, kpad, dw, pdt, nsplit,
                                                                             (long long)Cout * ntaps * kpad, accumulate);
  MS_LAUNCH_CHECK();
  return 0;
}
