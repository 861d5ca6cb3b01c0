// Fused TRAINING blocks of ConvNormRelu (reference src/model/layers.py:32-78): one launch per block and direction.
//
// A training-mode block is  z = conv(x)  ->  batch statistics of z  ->  y = LeakyReLU(BN(z)) [upsample x2 + skip]
// and needs a grid-wide reduction in the middle.  One PERSISTENT launch of at most one CTA per SM walks the phases with
// device-wide barriers between them (all CTAs co-resident: cooperative launch, grid <= #SMs).
//
//   forward, full-K form (split_k == 1; every layer whose tiles keep the machine busy on their own)
//             [GEMM tiles: TMA -> tcgen05.mma -> TMEM; the epilogue warps write z AND reduce the per-channel sum / sum of
//              squares of their 32 rows straight from the accumulator registers (shuffle butterfly, fp64 atomics per warp)]
//             | ONE barrier | finalize the tile's own channels, normalise + LeakyReLU (+ UNet upsample x2 + skip) reading
//             the accumulators that are STILL IN TMEM (up to 512 columns of resident tiles per CTA; larger layers re-read z)
//             -> fp32 activation and the next GEMM's bf16 operand planes (hi [, lo])
//   forward, split-K form (few tiles, long K: audio encoder tail, UNet bottleneck)
//             [k-slices over the whole machine, red.global.add.v4.f32 into z] | barrier | statistics over (64-channel x row
//             block) slabs, one fp64 atomic per channel and slab | barrier | finalize + normalise from z
//   backward  per-channel reductions of dy*act' and dy*act'*xhat over slabs | barrier | dz = BN-backward(dy) -> bf16 operand
//             planes, affine-parameter gradients into the flat gradient buffer | barrier |
//             [input-gradient GEMM tiles reading those planes through TMA -> dx]
//
// Split-bf16 operands (hi + lo planes): ONE k-step stages A_hi, A_lo, W_hi, W_lo once and issues hi*hi + hi*lo + lo*hi from
// the same shared-memory tiles -- 4 tile loads per 3 MMA passes instead of the 6 of pass-by-pass staging; the L2 -> SM
// operand traffic (~6300 B/clk chip-wide, B300_MICROARCH.md) is what bounds these small-batch GEMMs.
//
// Also here: the weight gradient accumulated in place (red.global.add.v4.f32 into ONE persistent fp32 accumulator per
// weight) and the table-driven kernel that converts every accumulator of a sub-network into its flat gradient buffer.
#include <cstring>
#include "tc_common.cuh"

namespace {

constexpr int TB_THREADS = 256;       // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: TMEM epilogue; all 8: element-wise phases
constexpr uint32_t TB_RING_BYTES = 196608;      // operand ring (192 KB); reused as scratch by the element-wise phases
constexpr int TB_MAX_STAGES = 8;
constexpr int TB_MAX_SLOTS = 16;      // TMEM accumulator slots per CTA (resident tiles)
constexpr uint32_t GRID_SPIN_LIMIT = 1u << 22;

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

// Execution shape of a GEMM phase (chosen by the host wrapper)
struct GemmCfg {
  int stages;                 // ring depth
  uint32_t stage_bytes;       // one k-step: A planes (16 KB each) + W planes (block_n * 128 B each)
  uint32_t w_plane_bytes;
  int nslots;                 // TMEM accumulator slots: 2 (streaming) or the tiles of a CTA (resident)
  int resident;               // accumulators are kept (never handed back) for the normalise pass
  int stats;                  // epilogue reduces per-channel sum / sum of squares of the valid rows
  int split_k;                // k-slices per tile (> 1: vector reductions into a zero-filled output)
};

struct FwdIO {
  float* z;                   // (rows, C) fp32 GEMM output (nullable in inference form); zero-filled by the caller when split_k > 1
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
  int dbg;
};

struct BwdIO {
  const float* dy;            // (rows_out, C)
  const float* dy2;           // nullable second addend of the incoming gradient (a skip connection's contribution)
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
  int dbg;
};

// One block of a chain launch: tensor maps (A, W, A lo, W lo), GEMM geometry, execution shape, BatchNorm side, tensors
struct FwdLayer {
  CUtensorMap map_a, map_w, map_a_lo, map_w_lo;
  IgemmParams p;
  GemmCfg g;
  BnParams bn;
  FwdIO io;
};
struct BwdLayer {
  CUtensorMap map_a, map_w, map_a_lo, map_w_lo;
  IgemmParams p;
  GemmCfg g;
  BnParams bn;
  BwdIO io;
};
constexpr int CHAIN_MAX = MS_CHAIN_MAX;
struct FwdChain {
  int n, dbg;
  unsigned int* sync;
  FwdLayer L[CHAIN_MAX];
};
struct BwdChain {
  int n, dbg;
  unsigned int* sync;
  BwdLayer L[CHAIN_MAX];
};
static_assert(sizeof(FwdChain) <= 32000 && sizeof(BwdChain) <= 32000, "chain parameter block exceeds the kernel parameter space");

// Where a bounded spin gave up (pipeline deadlock): written to MAPPED HOST memory right before the trap, so the host can
// still read it from a dead context (ms_debug_trap_info).  [0] site, [1] CTA, [2] thread, [3] block of the chain, [4..5] extra
__device__ int* g_trap_info = nullptr;
__device__ int g_trap_layer = 0;
__device__ __noinline__ void trap_report(int site, int x0, int x1) {
  int* q = g_trap_info;
  if (q) {
    if (atomicCAS(reinterpret_cast<unsigned int*>(q), 0u, (unsigned int)site) == 0u) {
      q[1] = (int)blockIdx.x; q[2] = (int)threadIdx.x; q[3] = g_trap_layer; q[4] = x0; q[5] = x1;
      __threadfence_system();
    }
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait_at(uint64_t* bar, uint32_t parity, int site, int x0 = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) trap_report(site, x0, (int)parity);
  }
}

// MS_PHASE_TS=1 (timing experiments only): CTA 0 stamps %globaltimer at the phase boundaries of the last fused launch
__device__ unsigned long long g_phase_ts[8 * MS_CHAIN_MAX];
// dbg = 0: off; else 1 + 8 * (layer index in the chain)
__device__ __forceinline__ void phase_ts(int dbg, int i) {
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_phase_ts[(dbg - 1) + i] = t;
  }
}

// Device-wide barrier on a monotonically increasing counter: bar.sync orders the CTA's writes before thread 0's release
// (cumulativity), thread 0 arrives with a release reduction and polls with acquire loads.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target, int dbg = 0, int ts = 0) {
  __syncthreads();
  phase_ts(dbg, ts);
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned int spins = 0, v;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > GRID_SPIN_LIMIT) trap_report(6, (int)v, (int)target);       // a CTA that never arrives must not hang the GPU
    }
  }
  __syncthreads();
}

// Phase parity of every pipeline barrier as each role thread has consumed it (bit s = uses so far of barrier s, mod 2).
// The barriers are initialised once per launch; the bits travel from block to block of a chain, so nothing is ever
// re-initialised (re-initialising between blocks raced with arrivals still in flight: observed deadlocks).
struct PipeBits {
  uint32_t e, f, te, tf;       // empty (producer), full (MMA), accumulator free (MMA), accumulator full (epilogue)
};

struct GemmSmem {
  uint8_t* ring;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;            // [TB_MAX_SLOTS]
  uint64_t* tempty;           // [TB_MAX_SLOTS]
};

struct TileAt {
  int tw, th, mt, cls, n0;
};
__device__ __forceinline__ TileAt tile_at(const IgemmParams& p, int tile) {
  const int ny = p.n_tiles_per_class * p.num_classes;
  TileAt a;
  const int y = tile % ny;
  int mt = tile / ny;
  a.tw = mt % p.tiles_w; mt /= p.tiles_w;
  a.th = mt % p.tiles_h; mt /= p.tiles_h;
  a.mt = mt;
  a.cls = y / p.n_tiles_per_class;
  a.n0 = (y - a.cls * p.n_tiles_per_class) * p.block_n;
  return a;
}

__device__ __forceinline__ void tmem_ld32x(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + 16u, v + 16);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
               "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])::"memory");
  asm volatile("" : "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
               "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory");
}

// Column sums over the 32 lanes of a warp for 32 columns held one row per lane: five exchange-and-halve steps (31 shuffles);
// lane l ends up with the total of column l.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; j++) {
      const float send = upper ? v[j] : v[j + off];
      const float keep = upper ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// One GEMM phase: work items (tile, k-slice) it = blockIdx.x, blockIdx.x + gridDim.x, ...; fp32 result into `out` (nullable)
// by 16-byte stores (split_k == 1) or vector reductions (split_k > 1, `out` zero-filled).  No bias, no activation.
// g.stats: the epilogue also adds every channel's sum / sum of squares over the valid rows into sums[0..C) / sums[C..2C).
__device__ __forceinline__ void gemm_phase(const CUtensorMap* map_a, const CUtensorMap* map_w, const CUtensorMap* map_a_lo,
                                           const CUtensorMap* map_w_lo, const IgemmParams& p, const GemmCfg& g,
                                           float* __restrict__ out, double* __restrict__ sums, int C, const GemmSmem& sm,
                                           uint32_t tmem_base, PipeBits& pb, float* stat_part = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ny = p.n_tiles_per_class * p.num_classes;
  const int tiles = p.tiles_w * p.tiles_h * p.tiles_b * ny;
  const int items = tiles * g.split_k;
  const int num_k_total = p.ntaps * p.cchunks;
  const int k_per = (num_k_total + g.split_k - 1) / g.split_k;
  const bool split = p.npass > 1;
  const uint32_t a_bytes = split ? 2u * A_STAGE_BYTES : A_STAGE_BYTES;
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int slice = it % g.split_k;
        const TileAt a = tile_at(p, it / g.split_k);
        const int w0 = a.tw * p.box_w, h0 = a.th * p.box_h, b0 = a.mt * p.box_b;
        const int tap_base = p.shared_taps ? 0 : a.cls * p.ntaps;
        const int chan_base = p.a_chan_base[a.cls];
        const int wrow = a.cls * p.class_n + a.n0;
        const int k_beg = slice * k_per;
        const int k_end = min(num_k_total, k_beg + k_per);
        for (int kk = k_beg; kk < k_end; kk++) {
          mbar_wait_at(&sm.empty[s], ((pb.e >> s) & 1u) ^ 1u, 1, kk);     // the release of this stage's previous use
          pb.e ^= 1u << s;
          const int tap = kk / p.cchunks, cc = kk - tap * p.cchunks;
          const short* t = p.taps[tap_base + tap];
          uint8_t* dst = sm.ring + (size_t)s * g.stage_bytes;
          mbar_expect_tx(&sm.full[s], g.stage_bytes);
          const int c0 = chan_base + t[0] + cc * BLOCK_K, c1 = w0 + t[1], c2 = t[2], c3 = h0 + t[3];
          tma_load_5d(map_a, &sm.full[s], dst, c0, c1, c2, c3, b0);
          tma_load_2d(map_w, &sm.full[s], dst + a_bytes, kk * BLOCK_K, wrow);
          if (split) {
            tma_load_5d(map_a_lo, &sm.full[s], dst + A_STAGE_BYTES, c0, c1, c2, c3, b0);
            tma_load_2d(map_w_lo, &sm.full[s], dst + a_bytes + g.w_plane_bytes, kk * BLOCK_K, wrow);
          }
          if (++s == g.stages) s = 0;
        }
      }
      // drain: the MMA thread's commits on the `empty` barriers of the last k-steps have nobody waiting for them; they must
      // have landed before this phase ends, so that the phase bits every role carries from block to block of a chain
      // (the barriers are initialised ONCE per launch) stay in step with the barriers
      for (int i = 0; i < TB_MAX_STAGES; i++) mbar_wait_at(&sm.empty[i], ((pb.e >> i) & 1u) ^ 1u, 5, i);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      int s = 0;
      uint32_t li = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int slice = it % g.split_k;
        const int k_beg = slice * k_per;
        const int num_k = min(num_k_total, k_beg + k_per) - k_beg;
        const int slot = (int)(li % (uint32_t)g.nslots);
        if (!g.resident) {
          mbar_wait_at(&sm.tempty[slot], ((pb.te >> slot) & 1u) ^ 1u, 2, slot);   // the epilogue warps drained this slot's previous accumulator
          pb.te ^= 1u << slot;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(slot * p.block_n);
        for (int ks = 0; ks < num_k; ks++) {
          mbar_wait_at(&sm.full[s], (pb.f >> s) & 1u, 3, ks);
          pb.f ^= 1u << s;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(sm.ring + (size_t)s * g.stage_bytes);
          const uint64_t da = make_kmajor_sw128_desc(base);
          const uint64_t db = make_kmajor_sw128_desc(base + a_bytes);
          if (split) {
            const uint64_t da_lo = make_kmajor_sw128_desc(base + A_STAGE_BYTES);
            const uint64_t db_lo = make_kmajor_sw128_desc(base + a_bytes + g.w_plane_bytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
              umma_bf16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
              umma_bf16(tacc, da + (uint64_t)(k * 2), db_lo + (uint64_t)(k * 2), idesc, 1u);
              umma_bf16(tacc, da_lo + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++)
              umma_bf16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(&sm.empty[s]);
          if (++s == g.stages) s = 0;
        }
        umma_commit(&sm.tfull[slot]);
        li++;
      }
      // drain (streaming accumulators): the epilogue's hand-back of the last use of every slot
      if (!g.resident)
        for (int i = 0; i < g.nslots; i++) mbar_wait_at(&sm.tempty[i], ((pb.te >> i) & 1u) ^ 1u, 7, i);
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // tile row == TMEM lane
    const int wi = r % p.box_w;
    const int hi = (r / p.box_w) % p.box_h;
    const int bi = r / (p.box_w * p.box_h);
    uint32_t li = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const TileAt a = tile_at(p, it / g.split_k);
      const int ow = a.tw * p.box_w + wi, oh = a.th * p.box_h + hi, ob = a.mt * p.box_b + bi;
      const bool valid = ow < p.out_w && oh < p.out_h && ob < p.out_b;
      const long long col0 = p.out_off[a.cls] + a.n0;
      float* dst = out ? out + (long long)ob * p.os_b + (long long)oh * p.os_h + (long long)ow * p.os_w + col0 : nullptr;
      const int slot = (int)(li % (uint32_t)g.nslots);
      mbar_wait_at(&sm.tfull[slot], (pb.tf >> slot) & 1u, 4, slot);
      pb.tf ^= 1u << slot;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * p.block_n);
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        const int rem = min(p.class_n - a.n0, p.block_n) - c0;     // columns left in this tile (uniform; multiple of 16)
        if (rem <= 0) break;
        if (rem < 32) {
          // trailing 16 columns (e.g. the 272-channel input gradient of the style-concatenated features); never with stats
          uint32_t u[16];
          tmem_ld16(taddr + (uint32_t)c0, u);
          tmem_wait_ld16(u);
          if (valid && dst) {
            if (g.split_k > 1) {
#pragma unroll
              for (int j = 0; j < 4; j++)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j), "f"(__uint_as_float(u[4 * j])),
                             "f"(__uint_as_float(u[4 * j + 1])), "f"(__uint_as_float(u[4 * j + 2])), "f"(__uint_as_float(u[4 * j + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 4; j++)
                *reinterpret_cast<float4*>(dst + c0 + 4 * j) = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]),
                                                                           __uint_as_float(u[4 * j + 2]), __uint_as_float(u[4 * j + 3]));
            }
          }
          break;
        }
        uint32_t v[32];
        tmem_ld32x(taddr + (uint32_t)c0, v);
        if (valid && dst) {
          if (g.split_k > 1) {
#pragma unroll
            for (int j = 0; j < 8; j++)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j), "f"(__uint_as_float(v[4 * j])),
                           "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++)
              *reinterpret_cast<float4*>(dst + c0 + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
        if (g.stats) {
          float f[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; j++) {
            f[j] = valid ? __uint_as_float(v[j]) : 0.f;
            s2[j] = f[j] * f[j];
          }
          const float cs = warp_colsum32(f, lane);
          const float cq = warp_colsum32(s2, lane);
          if (stat_part) {
            // resident tiles: the four epilogue warps' partials meet in shared memory, ONE atomic per channel and CTA later
            float* sp = stat_part + ((size_t)(slot * p.block_n + c0 + lane) * 4 + q) * 2;
            sp[0] = cs;
            sp[1] = cq;
          } else {
            const long long c = col0 + c0 + lane;
            atomicAdd(sums + c, (double)cs);
            atomicAdd(sums + C + c, (double)cq);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (!g.resident) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tempty[slot]);
      }
      li++;
    }
  }
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// rows [r0, r1) of this CTA out of `rows`
__device__ __forceinline__ void cta_rows(long long rows, long long& r0, long long& r1) {
  r0 = rows * (long long)blockIdx.x / (long long)gridDim.x;
  r1 = rows * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
}

// Slab decomposition of a (rows, C) matrix for the column reductions: 64-channel blocks x row blocks, one slab per CTA and
// pass; 16 float4 channel groups x 16 row lanes of threads.  A slab's column totals cost ONE fp64 atomic per channel, and
// an address sees (row blocks) of them instead of one per CTA.
struct Slabs {
  int ncb, nrb, total;
};
__device__ __forceinline__ Slabs slabs_of(int C, long long rows) {
  Slabs s;
  s.ncb = ((C >> 2) + 15) >> 4;
  int nrb = (int)gridDim.x / s.ncb;
  if (nrb < 1) nrb = 1;
  if ((long long)nrb > (rows + 15) / 16) nrb = (int)((rows + 15) / 16);
  s.nrb = nrb;
  s.total = s.ncb * nrb;
  return s;
}

struct SlabAt {
  int cbi, cg;                // channel block, this thread's float4 channel group (global)
  bool ok;                    // the group exists (4 * cg < C)
  long long r0, r1;           // rows of the slab
  int rbi;
};
__device__ __forceinline__ SlabAt slab_at(const Slabs& s, int sl, int C, long long rows) {
  SlabAt a;
  a.cbi = sl % s.ncb;
  a.rbi = sl / s.ncb;
  a.cg = a.cbi * 16 + (threadIdx.x & 15);
  a.ok = a.cg * 4 < C;
  a.r0 = rows * (long long)a.rbi / (long long)s.nrb;
  a.r1 = rows * (long long)(a.rbi + 1) / (long long)s.nrb;
  return a;
}

// Combine eight per-thread fp64 partials (acc[0..3] -> dst_a, acc[4..7] -> dst_b, four channels each) over the 16 row lanes
// of a slab and add them to global memory: one atomic per channel and slab.  scratch: [8 warps][16 groups][8] doubles.
__device__ __forceinline__ void slab_reduce_add(double (&acc)[8], const SlabAt& a, int C, double* scratch,
                                                double* __restrict__ dst_a, double* __restrict__ dst_b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);     // the warp's two row lanes
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 8; j++) scratch[(warp * 16 + lane) * 8 + j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c_l = threadIdx.x >> 3, j = threadIdx.x & 7;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < TB_THREADS / 32; w++) v += scratch[(w * 16 + c_l) * 8 + j];
    const int ch = (a.cbi * 16 + c_l) * 4 + (j & 3);
    if (ch < C) atomicAdd((j < 4 ? dst_a : dst_b) + ch, v);
  }
  __syncthreads();
}

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

// one output row segment of 32 channels: y (fp32, nullable) and operand planes (nullable) at element index e
__device__ __forceinline__ void store_row32(float* __restrict__ y, __nv_bfloat16* __restrict__ planes, int pfmt, long long ps,
                                            long long e, const float (&o)[32]) {
  if (y) {
#pragma unroll
    for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(y + e + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  if (planes) {
    uint32_t h[16];
    float lo[32];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      __nv_bfloat162 b = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
      h[j] = *reinterpret_cast<uint32_t*>(&b);
      lo[2 * j] = o[2 * j] - __bfloat162float(b.x);
      lo[2 * j + 1] = o[2 * j + 1] - __bfloat162float(b.y);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(planes + e + 8 * j) = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    if (pfmt == MS_BF16X2) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        *reinterpret_cast<uint4*>(planes + ps + e + 8 * j) =
            make_uint4(pack_bf162(lo[8 * j], lo[8 * j + 1]), pack_bf162(lo[8 * j + 2], lo[8 * j + 3]),
                       pack_bf162(lo[8 * j + 4], lo[8 * j + 5]), pack_bf162(lo[8 * j + 6], lo[8 * j + 7]));
    }
  }
}

// skip tensor (fp32 or operand planes) added to 32 channels at element index e
__device__ __forceinline__ void add_res32(const FwdIO& io, long long e, float (&o)[32]) {
  if (io.res) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float4 r = ldcg4(io.res + e + 4 * j);
      o[4 * j] += r.x; o[4 * j + 1] += r.y; o[4 * j + 2] += r.z; o[4 * j + 3] += r.w;
    }
  } else if (io.res_pl) {
    for (int pl = 0; pl < (io.res_fmt == MS_BF16X2 ? 2 : 1); pl++) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(io.res_pl + pl * io.res_ps + e + 8 * j));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
          o[8 * j + 2 * i] += __bfloat162float(b.x);
          o[8 * j + 2 * i + 1] += __bfloat162float(b.y);
        }
      }
    }
  }
}

// per-channel finalize of training-mode BatchNorm in two parts: the parameter loads do not depend on the statistics and are
// issued BEFORE the device barrier (bn_pre); after it only the two sums are fetched.  `publish`: running statistics and the
// [scale, shift, mean, rstd] table the backward reads
struct BnPre {
  double g, b, cb, rm, rv;
};
__device__ __forceinline__ BnPre bn_pre(const BnParams& bn, int c, bool publish) {
  BnPre q;
  q.g = ms_ldp_d(bn.gamma, bn.pdt, c);
  q.b = ms_ldp_d(bn.beta, bn.pdt, c);
  q.cb = 0.0; q.rm = 0.0; q.rv = 0.0;
  if (publish) {
    if (bn.cbias) q.cb = ms_ldp_d(bn.cbias, bn.pdt, c);
    q.rm = ms_ldp_d(bn.rmean, bn.pdt, c);
    q.rv = ms_ldp_d(bn.rvar, bn.pdt, c);
  }
  return q;
}
__device__ __forceinline__ void bn_finish(const BnParams& bn, const BnPre& q, long long rows, int c, bool publish, float& sc, float& sh) {
  const int C = bn.C;
  const double sum = __ldcg(bn.sums + c), sumsq = __ldcg(bn.sums + C + c);
  const double mean = sum / (double)rows;
  double var = sumsq / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)bn.eps);
  sc = (float)(q.g * rstd);
  sh = (float)(q.b - mean * q.g * rstd);
  if (publish) {
    const double unb = rows > 1 ? var * ((double)rows / (double)(rows - 1)) : var;
    ms_stp(bn.rmean, bn.pdt, c, (1.0 - (double)bn.momentum) * q.rm + (double)bn.momentum * (mean + q.cb));
    ms_stp(bn.rvar, bn.pdt, c, (1.0 - (double)bn.momentum) * q.rv + (double)bn.momentum * unb);
    bn.ss[c] = sc;
    bn.ss[C + c] = sh;
    bn.ss[2 * C + c] = (float)mean;
    bn.ss[3 * C + c] = (float)rstd;
  }
}
__device__ __forceinline__ void bn_finalize_channel(const BnParams& bn, long long rows, int c, bool publish, float& sc, float& sh) {
  const BnPre q = bn_pre(bn, c, publish);
  bn_finish(bn, q, rows, c, publish, sc, sh);
}

// inference: BatchNorm over the running statistics folded to (scale, shift), the conv bias inside the shift (the GEMM output
// excludes it) -- the arithmetic of ms_bn_finalize's inference form, per channel, inside the launch
__device__ __forceinline__ void bn_fold_channel(const BnParams& bn, int c, float& sc, float& sh) {
  const double g = ms_ldp_d(bn.gamma, bn.pdt, c), b = ms_ldp_d(bn.beta, bn.pdt, c);
  const double rm = ms_ldp_d(bn.rmean, bn.pdt, c), rv = ms_ldp_d(bn.rvar, bn.pdt, c);
  const double cb = bn.cbias ? ms_ldp_d(bn.cbias, bn.pdt, c) : 0.0;
  const double rstd = 1.0 / sqrt(rv + (double)bn.eps);
  sc = (float)(g * rstd);
  sh = (float)(b + (cb - rm) * g * rstd);
}

// ------------------------------------------------------------------------------------------------------------------
// forward block
// ------------------------------------------------------------------------------------------------------------------
// One forward block inside a chain launch.  Barriers are already initialised and TMEM allocated by the caller; bar_target is
// the running target of the device-wide barrier counter (all CTAs walk the same sequence of barriers).
__device__ __forceinline__ void fwd_layer(const FwdLayer& L, uint8_t* smem, const GemmSmem& sm, double* red_scratch,
                                          float* stat_part, uint32_t tmem_base, unsigned int* sync, unsigned int& bar_target,
                                          PipeBits& pb, int dbg) {
  const CUtensorMap& map_a = L.map_a;
  const CUtensorMap& map_w = L.map_w;
  const CUtensorMap& map_a_lo = L.map_a_lo;
  const CUtensorMap& map_w_lo = L.map_w_lo;
  const IgemmParams& p = L.p;
  const GemmCfg& g = L.g;
  const BnParams& bn = L.bn;
  const FwdIO& io = L.io;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = bn.C, T = TB_THREADS, t = threadIdx.x;
  const float slope = bn.slope;
  const bool train = bn.training == 1;     // 0: inference, folded constants in bn.ss; 2: inference, folded here from the running statistics

  // ---- phase 1: z = conv(x) on the tensor cores (+ statistics from the accumulators in the full-K training form)
  phase_ts(dbg, 1);
  const bool smem_stats = g.resident && g.stats;
  gemm_phase(&map_a, &map_w, &map_a_lo, &map_w_lo, p, g, io.z, bn.sums, C, sm, tmem_base, pb, smem_stats ? stat_part : nullptr);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // resident form: at most 512 accumulator columns per CTA = two (slot, column) entries per thread.  Their channel, the
  // BatchNorm parameters (independent of the statistics: fetched BEFORE the barrier) and the CTA-level statistics atomics
  int n_my = 0;
  int my_c[2] = {-1, -1};
  bool my_pub[2] = {false, false};
  BnPre my_pre[2];
  if (g.resident) {
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles_per_class * p.num_classes;
    for (int it = blockIdx.x; it < tiles; it += gridDim.x) n_my++;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int idx = t + k * T;
      if (idx < n_my * p.block_n) {
        const int slot = idx / p.block_n, j = idx - slot * p.block_n;
        const TileAt a = tile_at(p, blockIdx.x + slot * gridDim.x);
        if (a.n0 + j < p.class_n) {
          my_c[k] = (int)p.out_off[a.cls] + a.n0 + j;
          my_pub[k] = a.tw == 0 && a.th == 0 && a.mt == 0;
          if (train) my_pre[k] = bn_pre(bn, my_c[k], my_pub[k]);
        }
      }
    }
    if (smem_stats) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (my_c[k] >= 0) {
          const float* sp = stat_part + (size_t)(t + k * T) * 8;
          const double cs = (double)sp[0] + (double)sp[2] + (double)sp[4] + (double)sp[6];
          const double cq = (double)sp[1] + (double)sp[3] + (double)sp[5] + (double)sp[7];
          atomicAdd(bn.sums + my_c[k], cs);
          atomicAdd(bn.sums + C + my_c[k], cq);
        }
      }
    }
  }
  const bool need_bar1 = train || g.split_k > 1;      // inference, full-K: every CTA normalises its own tiles at once
  if (need_bar1) {
    bar_target += gridDim.x;
    grid_barrier(sync, bar_target, dbg, 2);
  } else {
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  phase_ts(dbg, 3);

  // ---- phase 2 (split-K training form only): per-channel sum and sum of squares of z over slabs
  if (train && !g.stats) {
    const Slabs sl = slabs_of(C, io.rows);
    const float* __restrict__ z = io.z;
    for (int s = blockIdx.x; s < sl.total; s += gridDim.x) {
      const SlabAt a = slab_at(sl, s, C, io.rows);
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (a.ok) {
        long long r = a.r0 + (t >> 4);
        for (; r + 16 < a.r1; r += 32) {
          const float4 v = ldcg4(z + r * C + 4 * a.cg), w = ldcg4(z + (r + 16) * C + 4 * a.cg);
          acc[0] += (double)v.x + (double)w.x; acc[1] += (double)v.y + (double)w.y;
          acc[2] += (double)v.z + (double)w.z; acc[3] += (double)v.w + (double)w.w;
          acc[4] += (double)v.x * v.x + (double)w.x * w.x; acc[5] += (double)v.y * v.y + (double)w.y * w.y;
          acc[6] += (double)v.z * v.z + (double)w.z * w.z; acc[7] += (double)v.w * v.w + (double)w.w * w.w;
        }
        for (; r < a.r1; r += 16) {
          const float4 v = ldcg4(z + r * C + 4 * a.cg);
          acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
          acc[4] += (double)v.x * v.x; acc[5] += (double)v.y * v.y; acc[6] += (double)v.z * v.z; acc[7] += (double)v.w * v.w;
        }
      }
      slab_reduce_add(acc, a, C, red_scratch, bn.sums, bn.sums + C);
    }
    bar_target += gridDim.x;
    grid_barrier(sync, bar_target, dbg, 4);
  }
  phase_ts(dbg, 5);

  // ---- phase 3: finalize + normalise + LeakyReLU (+ upsample x2 + skip)
  if (train && blockIdx.x == 0 && t == 0 && bn.nbt) bn.nbt[0] += 1;
  float* s_scale = reinterpret_cast<float*>(smem);          // the ring is idle from here on
  if (g.resident) {
    // the CTA's own tiles, accumulators still in TMEM: constants of each slot's columns, then TMEM -> y / planes
    float* s_shift = s_scale + 512;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (my_c[k] >= 0) {
        const int c = my_c[k];
        float sc, sh;
        if (train) bn_finish(bn, my_pre[k], io.rows, c, my_pub[k], sc, sh);
        else if (bn.training == 2) bn_fold_channel(bn, c, sc, sh);
        else { sc = __ldg(bn.ss + c); sh = __ldg(bn.ss + C + c); }
        s_scale[t + k * T] = sc;
        s_shift[t + k * T] = sh;
      }
    }
    __syncthreads();
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const int wi = r % p.box_w;
    const int hi = (r / p.box_w) % p.box_h;
    const int bi = r / (p.box_w * p.box_h);
    for (int slot = 0; slot < n_my; slot++) {
      const TileAt a = tile_at(p, blockIdx.x + slot * gridDim.x);
      const int ow = a.tw * p.box_w + wi, oh = a.th * p.box_h + hi, ob = a.mt * p.box_b + bi;
      const bool valid = ow < p.out_w && oh < p.out_h && ob < p.out_b;
      const long long zrow = ((long long)ob * p.out_h + oh) * p.out_w + ow;
      const long long col0 = p.out_off[a.cls] + a.n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * p.block_n);
      for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
        if (a.n0 + c0 >= p.class_n) break;
        uint32_t v[32];
        tmem_ld32x(taddr + (uint32_t)c0, v);
        if (!valid) continue;
        float o[32];
        const float* sc = s_scale + slot * p.block_n + c0;
        const float* sh = s_shift + slot * p.block_n + c0;
#pragma unroll
        for (int j = 0; j < 32; j++) o[j] = lrelu(fmaf(__uint_as_float(v[j]), sc[j], sh[j]), slope);
        if (!io.up2) {
          store_row32(io.y, io.planes, io.pfmt, io.pstride, zrow * C + col0 + c0, o);
        } else {
          const long long b = zrow / io.L;
          const long long ro = b * 2 * io.L + 2 * (zrow - b * io.L);
#pragma unroll 1
          for (int rr = 0; rr < 2; rr++) {
            float o2[32];
#pragma unroll
            for (int j = 0; j < 32; j++) o2[j] = o[j];
            const long long e = (ro + rr) * C + col0 + c0;
            add_res32(io, e, o2);
            store_row32(io.y, io.planes, io.pfmt, io.pstride, e, o2);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // streaming form: constants of all channels, then this CTA's share of the rows re-read from z
    float* s_shift = s_scale + C;
    for (int c = t; c < C; c += T) {
      float sc, sh;
      if (train) bn_finalize_channel(bn, io.rows, c, blockIdx.x == 0, sc, sh);
      else if (bn.training == 2) bn_fold_channel(bn, c, sc, sh);
      else { sc = __ldg(bn.ss + c); sh = __ldg(bn.ss + C + c); }
      s_scale[c] = sc;
      s_shift[c] = sh;
    }
    __syncthreads();
    const float* __restrict__ z = io.z;
    const int ncg = C >> 2;
    const long long rows_out = io.up2 ? 2 * io.rows : io.rows;
    long long o0, o1;
    cta_rows(rows_out, o0, o1);
    const long long total = (o1 - o0) * ncg;
    for (long long i0 = t; i0 < total; i0 += 4 * T) {
      float4 v[4];
      long long ro[4];
      int cg[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long long i = i0 + (long long)u * T;
        if (i < total) {
          ro[u] = o0 + i / ncg;
          cg[u] = (int)(i % ncg);
          long long ri = ro[u];
          if (io.up2) {
            const long long b = ro[u] / (2 * io.L);
            ri = b * io.L + ((ro[u] - b * 2 * io.L) >> 1);
          }
          v[u] = ldcg4(z + ri * C + 4 * cg[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long long i = i0 + (long long)u * T;
        if (i < total) {
          const float4 s = *reinterpret_cast<const float4*>(s_scale + 4 * cg[u]);
          const float4 h = *reinterpret_cast<const float4*>(s_shift + 4 * cg[u]);
          float4 o;
          o.x = lrelu(fmaf(v[u].x, s.x, h.x), slope); o.y = lrelu(fmaf(v[u].y, s.y, h.y), slope);
          o.z = lrelu(fmaf(v[u].z, s.z, h.z), slope); o.w = lrelu(fmaf(v[u].w, s.w, h.w), slope);
          const long long e = ro[u] * C + 4 * cg[u];
          if (io.res) {
            const float4 r4 = ldcg4(io.res + e);
            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
          } else if (io.res_pl) {
            for (int pl = 0; pl < (io.res_fmt == MS_BF16X2 ? 2 : 1); pl++) {
              const uint2 w2 = __ldcg(reinterpret_cast<const uint2*>(io.res_pl + pl * io.res_ps + e));
              const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&w2.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&w2.y);
              o.x += __bfloat162float(h0.x); o.y += __bfloat162float(h0.y); o.z += __bfloat162float(h1.x); o.w += __bfloat162float(h1.y);
            }
          }
          if (io.y) *reinterpret_cast<float4*>(io.y + e) = o;
          if (io.planes) store_planes4(io.planes, io.pfmt, io.pstride, e, o);
        }
      }
    }
  }
  (void)warp; (void)lane;
  __syncthreads();
  phase_ts(dbg, 6);
}

// initialise every pipeline barrier of a chain launch, once (thread 0; callers sync after it)
__device__ __forceinline__ void chain_init_barriers(const GemmSmem& sm) {
  if (threadIdx.x != 0) return;
  for (int s = 0; s < TB_MAX_STAGES; s++) {
    mbar_init(&sm.full[s], 1);
    mbar_init(&sm.empty[s], 1);
  }
  for (int s = 0; s < TB_MAX_SLOTS; s++) {
    mbar_init(&sm.tfull[s], 1);
    mbar_init(&sm.tempty[s], 4);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// A chain of forward blocks in ONE launch: block i+1 reads the operand planes block i wrote (a device-wide barrier with
// generic -> async proxy fences in between); TMEM (512 columns) is allocated once.
__global__ void __launch_bounds__(TB_THREADS, 1) conv_chain_fwd_kernel(const __grid_constant__ FwdChain ch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[TB_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TB_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[TB_MAX_SLOTS];
  __shared__ __align__(8) uint64_t tempty_bar[TB_MAX_SLOTS];
  __shared__ __align__(16) double red_scratch[8 * 16 * 8];
  __shared__ __align__(16) float stat_part[512 * 4 * 2];      // [accumulator column][epilogue warp][sum, sum of squares]
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  GemmSmem sm;
  sm.ring = smem; sm.full = full_bar; sm.empty = empty_bar; sm.tfull = tfull_bar; sm.tempty = tempty_bar;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  unsigned int bar_target = 0;
  uint32_t tmem_base = 0;
  PipeBits pb;
  pb.e = pb.f = pb.te = pb.tf = 0u;
  for (int li = 0; li < ch.n; li++) {
    const FwdLayer& L = ch.L[li];
    if (threadIdx.x == 0 && blockIdx.x == 0) g_trap_layer = li;
    phase_ts(ch.dbg ? 1 + 8 * li : 0, 0);
    if (li > 0) {
      // this block's TMA loads (async proxy, any CTA) read the planes the previous block's normalise pass wrote
      asm volatile("fence.proxy.async;" ::: "memory");
      bar_target += gridDim.x;
      grid_barrier(ch.sync, bar_target);
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (threadIdx.x == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_w)) : "memory");
      if (L.p.npass > 1) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_a_lo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_w_lo)) : "memory");
      }
    }
    if (li == 0) chain_init_barriers(sm);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = tmem_base_smem;
    fwd_layer(L, smem, sm, red_scratch, stat_part, tmem_base, ch.sync, bar_target, pb, ch.dbg ? 1 + 8 * li : 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward block
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 dy_at4(const float* __restrict__ dy, const float* __restrict__ dy2, long long r, int cg, int C,
                                         int up2, int L) {
  if (!up2) {
    float4 a = ldcg4(dy + r * C + 4 * cg);
    if (dy2) {
      const float4 b = ldcg4(dy2 + r * C + 4 * cg);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
  }
  const long long b = r / L;
  const int l = (int)(r - b * L);
  const long long ro = b * 2 * L + 2 * l;
  float4 a = ldcg4(dy + ro * C + 4 * cg);
  const float4 c = ldcg4(dy + (ro + 1) * C + 4 * cg);
  a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
  if (dy2) {
    const float4 e = ldcg4(dy2 + ro * C + 4 * cg), f = ldcg4(dy2 + (ro + 1) * C + 4 * cg);
    a.x += e.x + f.x; a.y += e.y + f.y; a.z += e.z + f.z; a.w += e.w + f.w;
  }
  return a;
}

__device__ __forceinline__ void bwd_layer(const BwdLayer& L, uint8_t* smem, const GemmSmem& sm, double* red_scratch,
                                          uint32_t tmem_base, unsigned int* sync, unsigned int& bar_target, PipeBits& pb, int dbg) {
  const CUtensorMap& map_a = L.map_a;
  const CUtensorMap& map_w = L.map_w;
  const CUtensorMap& map_a_lo = L.map_a_lo;
  const CUtensorMap& map_w_lo = L.map_w_lo;
  const IgemmParams& p = L.p;
  const GemmCfg& g = L.g;
  const BnParams& bn = L.bn;
  const BwdIO& io = L.io;

  const int C = bn.C, t = threadIdx.x;
  const float* __restrict__ z = io.z;
  const float* __restrict__ dy = io.dy;
  const float slope = bn.slope;
  const Slabs sl = slabs_of(C, io.rows);
  phase_ts(dbg, 1);

  // ---- phase 1: dbeta = sum g, dgamma = sum g * xhat with g = dy * act'(z)   (slab reductions)
  // Fast form (one slab per CTA, <= 4 rows per thread: every small-batch layer): g and xhat stay in registers across the
  // barrier, so the apply phase re-reads nothing but the two column totals.
  constexpr int RMAX = 4;
  const bool fast = sl.total <= (int)gridDim.x && (io.rows / sl.nrb + 1) <= 16 * RMAX;
  float4 kg[RMAX], kx[RMAX];
  float4 ksc = make_float4(0.f, 0.f, 0.f, 0.f);
  SlabAt fa;
  fa.ok = false; fa.r0 = fa.r1 = 0; fa.cg = 0; fa.cbi = 0; fa.rbi = 0;
  if (fast) {
    if ((int)blockIdx.x < sl.total) {
      fa = slab_at(sl, blockIdx.x, C, io.rows);
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (fa.ok) {
        ksc = __ldg(reinterpret_cast<const float4*>(bn.ss + 4 * fa.cg));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(bn.ss + C + 4 * fa.cg));
        const float4 mu = __ldg(reinterpret_cast<const float4*>(bn.ss + 2 * C + 4 * fa.cg)), rs = __ldg(reinterpret_cast<const float4*>(bn.ss + 3 * C + 4 * fa.cg));
#pragma unroll
        for (int k = 0; k < RMAX; k++) {
          const long long r = fa.r0 + (t >> 4) + 16 * k;
          if (r < fa.r1) {
            const float4 xv = ldcg4(z + r * C + 4 * fa.cg);
            const float4 d = dy_at4(dy, io.dy2, r, fa.cg, C, io.up2, io.L);
            float4 gq, xh;
            gq.x = fmaf(xv.x, ksc.x, sh.x) > 0.f ? d.x : d.x * slope; gq.y = fmaf(xv.y, ksc.y, sh.y) > 0.f ? d.y : d.y * slope;
            gq.z = fmaf(xv.z, ksc.z, sh.z) > 0.f ? d.z : d.z * slope; gq.w = fmaf(xv.w, ksc.w, sh.w) > 0.f ? d.w : d.w * slope;
            xh.x = (xv.x - mu.x) * rs.x; xh.y = (xv.y - mu.y) * rs.y; xh.z = (xv.z - mu.z) * rs.z; xh.w = (xv.w - mu.w) * rs.w;
            kg[k] = gq; kx[k] = xh;
            acc[0] += (double)gq.x * (double)xh.x; acc[1] += (double)gq.y * (double)xh.y;
            acc[2] += (double)gq.z * (double)xh.z; acc[3] += (double)gq.w * (double)xh.w;
            acc[4] += gq.x; acc[5] += gq.y; acc[6] += gq.z; acc[7] += gq.w;
          }
        }
      }
      if (bn.training) slab_reduce_add(acc, fa, C, red_scratch, bn.sums, bn.sums + C);
    }
  } else if (bn.training) {
    for (int s = blockIdx.x; s < sl.total; s += gridDim.x) {
      const SlabAt a = slab_at(sl, s, C, io.rows);
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (a.ok) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(bn.ss + 4 * a.cg)), sh = __ldg(reinterpret_cast<const float4*>(bn.ss + C + 4 * a.cg));
        const float4 mu = __ldg(reinterpret_cast<const float4*>(bn.ss + 2 * C + 4 * a.cg)), rs = __ldg(reinterpret_cast<const float4*>(bn.ss + 3 * C + 4 * a.cg));
        // four rows in flight per thread (large batches: this pass is HBM-bound, one row at a time left it latency-bound)
        for (long long r = a.r0 + (t >> 4); r < a.r1; r += 64) {
          float4 xv[4], d[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const long long ru = r + 16 * u;
            if (ru < a.r1) {
              xv[u] = ldcg4(z + ru * C + 4 * a.cg);
              d[u] = dy_at4(dy, io.dy2, ru, a.cg, C, io.up2, io.L);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (r + 16 * u < a.r1) {
              const float g0 = fmaf(xv[u].x, sc.x, sh.x) > 0.f ? d[u].x : d[u].x * slope, g1 = fmaf(xv[u].y, sc.y, sh.y) > 0.f ? d[u].y : d[u].y * slope;
              const float g2 = fmaf(xv[u].z, sc.z, sh.z) > 0.f ? d[u].z : d[u].z * slope, g3 = fmaf(xv[u].w, sc.w, sh.w) > 0.f ? d[u].w : d[u].w * slope;
              acc[0] += (double)g0 * (double)((xv[u].x - mu.x) * rs.x); acc[1] += (double)g1 * (double)((xv[u].y - mu.y) * rs.y);
              acc[2] += (double)g2 * (double)((xv[u].z - mu.z) * rs.z); acc[3] += (double)g3 * (double)((xv[u].w - mu.w) * rs.w);
              acc[4] += g0; acc[5] += g1; acc[6] += g2; acc[7] += g3;
            }
          }
        }
      }
      slab_reduce_add(acc, a, C, red_scratch, bn.sums, bn.sums + C);
    }
  }
  bar_target += gridDim.x;
  grid_barrier(sync, bar_target, dbg, 2);
  phase_ts(dbg, 3);

  // ---- phase 2: dz = scale * (g - dbeta/N - xhat * dgamma/N) -> operand planes; affine gradients
  {
    const float inv = 1.f / (float)io.rows;
    const int s_beg = fast ? ((int)blockIdx.x < sl.total ? (int)blockIdx.x : sl.total) : (int)blockIdx.x;
    const int s_step = fast ? sl.total : (int)gridDim.x;
    for (int s = s_beg; s < sl.total; s += s_step) {
      const SlabAt a = fast ? fa : slab_at(sl, s, C, io.rows);
      if (!a.ok) continue;
      float4 dgv = make_float4(0.f, 0.f, 0.f, 0.f), dbv = dgv;
      if (bn.training) {
        const double d0 = __ldcg(bn.sums + 4 * a.cg), d1 = __ldcg(bn.sums + 4 * a.cg + 1), d2 = __ldcg(bn.sums + 4 * a.cg + 2), d3 = __ldcg(bn.sums + 4 * a.cg + 3);
        const double b0 = __ldcg(bn.sums + C + 4 * a.cg), b1 = __ldcg(bn.sums + C + 4 * a.cg + 1), b2 = __ldcg(bn.sums + C + 4 * a.cg + 2), b3 = __ldcg(bn.sums + C + 4 * a.cg + 3);
        dgv = make_float4((float)d0, (float)d1, (float)d2, (float)d3);
        dbv = make_float4((float)b0, (float)b1, (float)b2, (float)b3);
        if (a.rbi == 0 && (t >> 4) == 0) {
          const double dgs[4] = {d0, d1, d2, d3}, dbs[4] = {b0, b1, b2, b3};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (io.ggamma) ms_stp(io.ggamma, io.gdt, 4 * a.cg + j, ms_ldp_d(io.ggamma, io.gdt, 4 * a.cg + j) + dgs[j]);
            if (io.gbeta) ms_stp(io.gbeta, io.gdt, 4 * a.cg + j, ms_ldp_d(io.gbeta, io.gdt, 4 * a.cg + j) + dbs[j]);
          }
        }
      }
      if (fast) {
#pragma unroll
        for (int k = 0; k < RMAX; k++) {
          const long long r = a.r0 + (t >> 4) + 16 * k;
          if (r < a.r1) {
            float4 o;
            if (bn.training) {
              o.x = ksc.x * (kg[k].x - dbv.x * inv - kx[k].x * dgv.x * inv);
              o.y = ksc.y * (kg[k].y - dbv.y * inv - kx[k].y * dgv.y * inv);
              o.z = ksc.z * (kg[k].z - dbv.z * inv - kx[k].z * dgv.z * inv);
              o.w = ksc.w * (kg[k].w - dbv.w * inv - kx[k].w * dgv.w * inv);
            } else {
              o.x = ksc.x * kg[k].x; o.y = ksc.y * kg[k].y; o.z = ksc.z * kg[k].z; o.w = ksc.w * kg[k].w;
            }
            store_planes4(io.dzp, io.pfmt, io.pstride, r * C + 4 * a.cg, o);
          }
        }
        continue;
      }
      const float4 sc = __ldg(reinterpret_cast<const float4*>(bn.ss + 4 * a.cg)), sh = __ldg(reinterpret_cast<const float4*>(bn.ss + C + 4 * a.cg));
      const float4 mu = __ldg(reinterpret_cast<const float4*>(bn.ss + 2 * C + 4 * a.cg)), rs = __ldg(reinterpret_cast<const float4*>(bn.ss + 3 * C + 4 * a.cg));
      for (long long r = a.r0 + (t >> 4); r < a.r1; r += 64) {
        float4 xq[4], dq[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const long long ru = r + 16 * u;
          if (ru < a.r1) {
            xq[u] = ldcg4(z + ru * C + 4 * a.cg);
            dq[u] = dy_at4(dy, io.dy2, ru, a.cg, C, io.up2, io.L);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const long long ru = r + 16 * u;
          if (ru < a.r1) {
            const float4 xv = xq[u], d = dq[u];
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
            store_planes4(io.dzp, io.pfmt, io.pstride, ru * C + 4 * a.cg, o);
          }
        }
      }
    }
  }
  if (!io.has_gemm) {
    __syncthreads();
    return;
  }
  // the planes just written (generic proxy) are read by other CTAs' TMA loads (async proxy) in the next phase
  asm volatile("fence.proxy.async;" ::: "memory");
  bar_target += gridDim.x;
  grid_barrier(sync, bar_target, dbg, 4);
  asm volatile("fence.proxy.async;" ::: "memory");
  phase_ts(dbg, 5);

  // ---- phase 3: dx = conv^T(dz) on the tensor cores
  gemm_phase(&map_a, &map_w, &map_a_lo, &map_w_lo, p, g, io.dx, nullptr, 0, sm, tmem_base, pb);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  phase_ts(dbg, 6);
}

// A chain of backward blocks in ONE launch, last block of the forward chain first: block i's incoming gradient is the input
// gradient block i+1 just produced (plus, for a UNet skip source, the gradient of the block that consumed the skip).
__global__ void __launch_bounds__(TB_THREADS, 1) conv_chain_bwd_kernel(const __grid_constant__ BwdChain ch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[TB_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TB_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[TB_MAX_SLOTS];
  __shared__ __align__(8) uint64_t tempty_bar[TB_MAX_SLOTS];
  __shared__ __align__(16) double red_scratch[8 * 16 * 8];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  GemmSmem sm;
  sm.ring = smem; sm.full = full_bar; sm.empty = empty_bar; sm.tfull = tfull_bar; sm.tempty = tempty_bar;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  unsigned int bar_target = 0;
  uint32_t tmem_base = 0;
  PipeBits pb;
  pb.e = pb.f = pb.te = pb.tf = 0u;
  for (int li = 0; li < ch.n; li++) {
    const BwdLayer& L = ch.L[li];
    if (threadIdx.x == 0 && blockIdx.x == 0) g_trap_layer = 100 + li;
    phase_ts(ch.dbg ? 1 + 8 * li : 0, 0);
    if (li > 0) {
      // the previous block's input gradient (this block's dy) must be complete and visible
      bar_target += gridDim.x;
      grid_barrier(ch.sync, bar_target);
    }
    if (threadIdx.x == 0) {
      if (L.io.has_gemm) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_w)) : "memory");
        if (L.p.npass > 1) {
          asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_a_lo)) : "memory");
          asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&L.map_w_lo)) : "memory");
        }
      }
    }
    if (li == 0) chain_init_barriers(sm);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = tmem_base_smem;
    bwd_layer(L, smem, sm, red_scratch, tmem_base, ch.sync, bar_target, pb, ch.dbg ? 1 + 8 * li : 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

__device__ __forceinline__ void wgrad_acc_body(const CUtensorMap& map_x, const CUtensorMap& map_z, const CUtensorMap& map_x_lo,
                                               const CUtensorMap& map_z_lo, const WgradAccParams& p, float* __restrict__ acc,
                                               const int bx, const int by, const int bz) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[WGA_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WGA_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = bz / p.ntaps, tap = bz - cls * p.ntaps;
  const int nt = by / p.c_tiles, ct = by - nt * p.c_tiles;
  const int n0 = nt * 128, c0 = ct * p.c_tile;
  const int nc = min(p.c_tile, p.kpad - c0);
  const int xchunks = nc / 64;
  // One k-step = one tile of 64 pixel rows.  Split-bf16 stages dz_hi, dz_lo, x_hi, x_lo ONCE and issues the three passes from
  // them (4 tile loads per 3 passes instead of 6: this GEMM streams its operands from L2 and is bound by that traffic);
  // stage = [dz hi (2 chunks)][dz lo (2)][x hi (xchunks)][x lo (xchunks)], ring depth = what fits in 192 KB.
  const bool split = p.npass > 1;
  const int stage_chunks = split ? (4 + 2 * xchunks) : (2 + xchunks);
  const uint32_t stage_bytes = (uint32_t)stage_chunks * WGA_CHUNK_BYTES;
  int nstages = (WGA_STAGES * 6) / stage_chunks;
  if (nstages > WGA_STAGES) nstages = WGA_STAGES;
  const int total_rt = p.tiles_w * p.tiles_h * p.tiles_b;
  const int per = (total_rt + p.split - 1) / p.split;
  const int rt_beg = bx * per, rt_end = min(total_rt, rt_beg + per);
  const int num_k = max(0, rt_end - rt_beg);
  const uint32_t zlo_off = 2 * WGA_CHUNK_BYTES;
  const uint32_t x_off = (split ? 4 : 2) * WGA_CHUNK_BYTES;
  const uint32_t xlo_off = x_off + (uint32_t)xchunks * WGA_CHUNK_BYTES;
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
          const int s = ks % nstages;
          const uint32_t ph = (uint32_t)(ks / nstages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          int rt = rt_beg + ks;
          const int tw = rt % p.tiles_w; rt /= p.tiles_w;
          const int th = rt % p.tiles_h; rt /= p.tiles_h;
          const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = rt * p.box_b;
          mbar_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* st_ = smem + (size_t)s * stage_bytes;
          tma_load_5d(&map_z, &full_bar[s], st_, zc, w0, 0, h0, b0);
          tma_load_5d(&map_z, &full_bar[s], st_ + WGA_CHUNK_BYTES, zc + 64, w0, 0, h0, b0);
          for (int i = 0; i < xchunks; i++)
            tma_load_5d(&map_x, &full_bar[s], st_ + x_off + (size_t)i * WGA_CHUNK_BYTES, xc + 64 * i, w0 + t[1], t[2], h0 + t[3], b0);
          if (split) {
            tma_load_5d(&map_z_lo, &full_bar[s], st_ + zlo_off, zc, w0, 0, h0, b0);
            tma_load_5d(&map_z_lo, &full_bar[s], st_ + zlo_off + WGA_CHUNK_BYTES, zc + 64, w0, 0, h0, b0);
            for (int i = 0; i < xchunks; i++)
              tma_load_5d(&map_x_lo, &full_bar[s], st_ + xlo_off + (size_t)i * WGA_CHUNK_BYTES, xc + 64 * i, w0 + t[1], t[2], h0 + t[3], b0);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < num_k; ks++) {
          const int s = ks % nstages;
          const uint32_t ph = (uint32_t)(ks / nstages) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da = make_mnmajor_sw128_desc2(base);
          const uint64_t db = make_mnmajor_sw128_desc2(base + x_off);
#pragma unroll
          for (int k = 0; k < WGA_ROWS / UMMA_K; k++)
            umma_bf16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (ks | k) != 0 ? 1u : 0u);
          if (split) {
            // same pass order as pass-by-pass staging: x_hi * dz_hi, x_hi * dz_lo, x_lo * dz_hi
            const uint64_t da_lo = make_mnmajor_sw128_desc2(base + zlo_off);
            const uint64_t db_lo = make_mnmajor_sw128_desc2(base + xlo_off);
#pragma unroll
            for (int k = 0; k < WGA_ROWS / UMMA_K; k++)
              umma_bf16(tmem_base, da_lo + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, 1u);
#pragma unroll
            for (int k = 0; k < WGA_ROWS / UMMA_K; k++)
              umma_bf16(tmem_base, da + (uint64_t)(k * 128), db_lo + (uint64_t)(k * 128), idesc, 1u);
          }
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

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_acc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z,
                 const __grid_constant__ CUtensorMap map_x_lo, const __grid_constant__ CUtensorMap map_z_lo,
                 const __grid_constant__ WgradAccParams p, float* __restrict__ acc) {
  wgrad_acc_body(map_x, map_z, map_x_lo, map_z_lo, p, acc, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
}

// The weight gradients of every block of a chain in ONE launch: CTA ranges per layer, the same tile program inside.
struct WgradLayer {
  CUtensorMap map_x, map_z, map_x_lo, map_z_lo;
  WgradAccParams p;
  float* acc;
  int gx, gy, gz, first;
};
struct WgradMulti {
  int n;
  WgradLayer L[CHAIN_MAX];
};
static_assert(sizeof(WgradMulti) <= 32000, "multi-wgrad parameter block exceeds the kernel parameter space");
__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad_acc_multi_kernel(const __grid_constant__ WgradMulti m) {
  int li = 0;
  while (li + 1 < m.n && (int)blockIdx.x >= m.L[li + 1].first) li++;
  const WgradLayer& L = m.L[li];
  const int local = (int)blockIdx.x - L.first;
  const int bx = local % L.gx, by = (local / L.gx) % L.gy, bz = local / (L.gx * L.gy);
  wgrad_acc_body(L.map_x, L.map_z, L.map_x_lo, L.map_z_lo, L.p, L.acc, bx, by, bz);
}

// every accumulator of a sub-network -> its parameter-gradient buffer (dw += unpack(acc)), one launch.
// Work is cut into units of ~MULTI_UNIT elements spread over the entries in proportion to their size (every CTA derives
// the same unit -> (entry, output-channel range) map from the table), and every output channel goes through shared memory:
// its accumulator row [tap][kpad] is read contiguously, its gradient row [c][tap] written contiguously.
constexpr int MULTI_UNIT = 16384;
constexpr int MULTI_UNIT_MIN = 2048;
constexpr int MULTI_MAX_ENTRIES = 256;
constexpr int UNPACK_SMEM_FLOATS = 10240;      // 40 KB: rows of up to 48 taps x 192 padded channels; longer rows go direct

// Sum of a table in shared memory by warp 0 (the caller synchronised after filling it); every thread gets the result.
__device__ __forceinline__ long long multi_table_sum(const int* __restrict__ v, int n, long long* s_out) {
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

// unit u -> entry index and the unit's position inside the entry; *nu = units of that entry.  All threads get the same answer.
__device__ __forceinline__ int multi_find(const int* __restrict__ s_units, int n_entries, int u, int* local, int* nu) {
  int e = 0;
  while (e < n_entries && u >= s_units[e]) { u -= s_units[e]; e++; }
  *local = u;
  *nu = e < n_entries ? s_units[e] : 1;
  return e;
}

__global__ void __launch_bounds__(256) unpack_wgrad_multi_kernel(const ms_wgrad_entry* __restrict__ table, int n_entries) {
  __shared__ int s_units[MULTI_MAX_ENTRIES];
  __shared__ long long s_sum;
  __shared__ __align__(16) float s_row[UNPACK_SMEM_FLOATS];
  const int t = threadIdx.x;
  for (int i = t; i < n_entries; i += blockDim.x) s_units[i] = table[i].Cout * table[i].Cin_g * table[i].taps;      // elements (< 2^31)
  __syncthreads();
  // unit size: ~half a unit per CTA, between MULTI_UNIT_MIN and MULTI_UNIT elements (a small table -- the discriminator's
  // 0.4 M elements -- would otherwise keep 25 CTAs busy for eight serial passes each while 1159 idle)
  const long long elems = multi_table_sum(s_units, n_entries, &s_sum);
  long long unit_ll = 2 * elems / gridDim.x;
  const unsigned unit = (unsigned)(unit_ll < MULTI_UNIT_MIN ? MULTI_UNIT_MIN : (unit_ll > MULTI_UNIT ? MULTI_UNIT : unit_ll));
  for (int i = t; i < n_entries; i += blockDim.x) s_units[i] = (int)(((unsigned)s_units[i] + unit - 1) / unit);
  __syncthreads();
  const int total = (int)multi_table_sum(s_units, n_entries, &s_sum);
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    int local, nu;
    const int ei = multi_find(s_units, n_entries, u, &local, &nu);
    const ms_wgrad_entry e = table[ei];
    const int o0 = (int)((long long)e.Cout * local / nu), o1 = (int)((long long)e.Cout * (local + 1) / nu);
    const float* __restrict__ acc = reinterpret_cast<const float*>(e.acc);
    const int row_in = e.taps * e.kpad, row_out = e.Cin_g * e.taps;
    if (row_in <= UNPACK_SMEM_FLOATS) {
      // as many output channels per pass as fit in shared memory: contiguous reads, contiguous writes, two syncs per pass
      const int R = UNPACK_SMEM_FLOATS / row_in;
      for (int ob0 = o0; ob0 < o1; ob0 += R) {
        const int nr = min(R, o1 - ob0);
        const float* src = acc + (long long)ob0 * row_in;
        __syncthreads();
        for (int i = 4 * t; i < nr * row_in; i += 4 * blockDim.x) *reinterpret_cast<float4*>(s_row + i) = *reinterpret_cast<const float4*>(src + i);
        __syncthreads();
        const long long ob = (long long)ob0 * row_out;
        if (e.pdt == MS_F64 && !e.accumulate && !(row_out & 1) && !((uintptr_t)e.dw & 15)) {
          // store-only fp64 sink: two adjacent (channel, tap) positions per thread and 16-byte stores, the channel index
          // and the row by multiplies (i < 2^14, row_out < 2^13, taps <= 64: exact)
          const unsigned inv = (unsigned)((0x100000000ULL + (unsigned)e.taps - 1) / (unsigned)e.taps);
          const unsigned inv_ro = (unsigned)((0x100000000ULL + (unsigned)row_out - 1) / (unsigned)row_out);
          double* __restrict__ dw = reinterpret_cast<double*>(e.dw) + ob;
          for (int i = 2 * t; i < nr * row_out; i += 2 * blockDim.x) {      // i, row_out even: a pair never straddles two rows
            const int rr = (int)__umulhi((unsigned)i, inv_ro);
            const int j = i - rr * row_out;
            const float* sr = s_row + rr * row_in;
            const int c0 = e.taps == 1 ? j : (int)__umulhi((unsigned)j, inv);      // (taps == 1: the multiplier would be 2^32)
            const int tap0 = j - c0 * e.taps;
            int c1 = c0, tap1 = tap0 + 1;
            if (tap1 == e.taps) { tap1 = 0; c1++; }
            *reinterpret_cast<double2*>(dw + i) = make_double2((double)sr[tap0 * e.kpad + c0], (double)sr[tap1 * e.kpad + c1]);
          }
        } else {
          for (int i = t; i < nr * row_out; i += blockDim.x) {
            const int rr = i / row_out, j = i - rr * row_out;
            const int c = j / e.taps, tap = j - c * e.taps;
            const double v = (double)s_row[rr * row_in + tap * e.kpad + c];
            ms_stp(e.dw, e.pdt, ob + i, v + (e.accumulate ? ms_ldp_d(e.dw, e.pdt, ob + i) : 0.0));
          }
        }
      }
    } else {
      for (int o = o0; o < o1; o++) {
        const float* src = acc + (long long)o * row_in;
        const long long ob = (long long)o * row_out;
        for (int i = t; i < row_out; i += blockDim.x) {
          const int c = i / e.taps, tap = i - c * e.taps;
          const double v = (double)src[tap * e.kpad + c];
          ms_stp(e.dw, e.pdt, ob + i, v + (e.accumulate ? ms_ldp_d(e.dw, e.pdt, ob + i) : 0.0));
        }
      }
    }
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

// SM budget of the cooperative chain launches (0 = the whole device).  Data-parallel training leaves a few SMs to the NCCL
// kernels of the overlapped gradient exchange: a chain launch that holds every SM would serialise them behind itself.
static int g_sm_budget = 0;
static int chain_sms() {
  const int n = ms_num_sms();
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}

static int phase_dbg() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MS_PHASE_TS");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

static int fill_bn(BnParams& b, const ms_block_bn* s) {
  if (!s || s->C < 16 || s->C % 4 || !s->gamma || !s->beta) return MS_EINVAL;
  if (s->training == 1 && (!s->sums || !s->ss)) return MS_EINVAL;
  if (s->training == 0 && !s->ss) return MS_EINVAL;
  if (s->pdt != MS_F32 && s->pdt != MS_F64) return MS_EINVAL;
  b.C = s->C; b.pdt = s->pdt; b.training = s->training; b.momentum = s->momentum; b.eps = s->eps; b.slope = s->slope;
  b.gamma = s->gamma; b.beta = s->beta; b.cbias = s->conv_bias; b.rmean = s->running_mean; b.rvar = s->running_var;
  b.nbt = reinterpret_cast<long long*>(s->num_batches_tracked); b.sums = s->sums; b.ss = s->ss;
  return 0;
}

// k-slices of a GEMM phase normalised so that every slice owns at least one k-step; work items = tiles x slices
static int gemm_split(const ms_igemm_desc* d) {
  const int num_k = d->ntaps * d->cchunks;
  int split = d->split_k > 1 ? d->split_k : 1;
  if (split > num_k) split = num_k;
  const int per = (num_k + split - 1) / split;
  return (num_k + per - 1) / per;
}
static long long gemm_tiles(const IgemmParams& p) {
  return (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles_per_class * p.num_classes;
}

// Execution shape of a GEMM phase on a launch of `grid` CTAs: ring depth from the stage size, accumulator slots.
// want_resident: keep all tiles of a CTA in TMEM when they fit (forward blocks).
static int gemm_cfg(const ms_igemm_desc* d, const IgemmParams& p, unsigned grid, bool stats, bool want_resident, GemmCfg* gp) {
  GemmCfg& g = *gp;
  if ((stats || want_resident) && (p.block_n % 32 || d->class_n % 32)) return MS_EINVAL;
  if (p.block_n % 16) return MS_EINVAL;
  const int planes = p.npass > 1 ? 2 : 1;
  g.w_plane_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
  g.stage_bytes = (uint32_t)planes * (A_STAGE_BYTES + g.w_plane_bytes);
  g.stages = (int)(TB_RING_BYTES / g.stage_bytes);
  if (g.stages > TB_MAX_STAGES) g.stages = TB_MAX_STAGES;
  if (g.stages < 2) return MS_EINVAL;
  g.split_k = gemm_split(d);
  const long long tiles = gemm_tiles(p);
  if (tiles < 1 || tiles * g.split_k > 0x7fffffffLL) return MS_EINVAL;
  g.stats = (stats && g.split_k == 1) ? 1 : 0;
  g.resident = 0;
  g.nslots = 2;
  if (want_resident && g.split_k == 1) {
    const long long per_cta = (tiles + grid - 1) / grid;
    if (per_cta <= TB_MAX_SLOTS && per_cta * p.block_n <= 512) {
      g.resident = 1;
      g.nslots = (int)per_cta;
    }
  }
  if (2 * p.block_n > 512 && !g.resident) return MS_EINVAL;
  return 0;
}

static int fill_fwd_layer(const ms_chain_fwd_layer* e, FwdLayer* L, long long* want) {
  const ms_igemm_desc* d = e->d;
  if (!d || !e->a || !e->w || !e->bn || (!e->y && !e->planes)) return MS_EINVAL;
  if (d->out_dtype != MS_F32 || d->epilogue != 0) return MS_EINVAL;
  if (e->bn->training && (!e->bn->running_mean || !e->bn->running_var)) return MS_EINVAL;
  if (e->bn->training == 1 && !e->z) return MS_EINVAL;
  if (!e->z && d->split_k > 1) return MS_EINVAL;
  if (e->res_planes && (((uintptr_t)e->res_planes & 15) || (e->res_pfmt != MS_BF16 && e->res_pfmt != MS_BF16X2) ||
                        (e->res_pfmt == MS_BF16X2 && (e->res_pstride <= 0 || (e->res_pstride * 2) % 16))))
    return MS_EINVAL;
  if (((uintptr_t)e->z & 15) || ((uintptr_t)e->y & 15) || ((uintptr_t)e->planes & 15) || ((uintptr_t)e->res & 15)) return MS_EINVAL;
  if (e->planes && e->pfmt != MS_BF16 && e->pfmt != MS_BF16X2) return MS_EINVAL;
  if (e->planes && e->pfmt == MS_BF16X2 && (e->pstride <= 0 || (e->pstride * 2) % 16)) return MS_EINVAL;
  if (e->up2 && ((!e->res && !e->res_planes) || d->out_dims[1] != 1)) return MS_EINVAL;
  CUtensorMap maps[4];
  int rc = igemm_prepare(d, e->a, e->w, d->block_n, maps, &L->p);
  if (rc) return rc;
  L->map_a = maps[0]; L->map_w = maps[1]; L->map_a_lo = maps[2]; L->map_w_lo = maps[3];
  rc = fill_bn(L->bn, e->bn);
  if (rc) return rc;
  const long long rows = (long long)d->out_dims[0] * d->out_dims[1] * d->out_dims[2];
  const int C = d->num_classes * d->class_n;
  if (C != L->bn.C || C % 32) return MS_EINVAL;
  // z / y are dense (rows, C) matrices: the element-wise phases index them that way
  if (d->out_strides[0] != C || d->out_strides[1] != (int64_t)C * d->out_dims[0] ||
      d->out_strides[2] != (int64_t)C * d->out_dims[0] * d->out_dims[1])
    return MS_EINVAL;
  for (int i = 0; i < d->num_classes; i++)
    if (d->out_off[i] != (int64_t)i * d->class_n) return MS_EINVAL;
  if ((size_t)C * 8 > TB_RING_BYTES) return MS_EINVAL;
  FwdIO& io = L->io;
  io.z = e->z; io.y = e->y; io.planes = reinterpret_cast<__nv_bfloat16*>(e->planes); io.pfmt = e->pfmt; io.pstride = e->pstride;
  io.res = e->up2 ? e->res : nullptr; io.up2 = e->up2 ? 1 : 0; io.L = d->out_dims[0]; io.rows = rows;
  io.res_pl = (e->up2 && !e->res) ? reinterpret_cast<const __nv_bfloat16*>(e->res_planes) : nullptr;
  io.res_fmt = e->res_pfmt; io.res_ps = e->res_pstride;
  io.sync = nullptr; io.dbg = 0;
  *want = gemm_tiles(L->p) * gemm_split(d);
  return 0;
}

static FwdChain g_fwd_chain;      // host staging of the parameter blocks (one launch at a time per process: the C-ABI is
static BwdChain g_bwd_chain;      // called from the single Python thread that owns the device)
static WgradMulti g_wgrad_multi;

}  // namespace

static int* g_trap_host = nullptr;
static int trap_buffer_init() {
  if (g_trap_host) return 0;
  int* h = nullptr;
  MS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped));
  memset(h, 0, 64);
  int* d = nullptr;
  MS_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0));
  MS_CUDA(cudaMemcpyToSymbol(g_trap_info, &d, sizeof(d)));
  g_trap_host = h;
  return 0;
}

extern "C" int ms_set_chain_sm_budget(int sms) {
  if (sms < 0) return MS_EINVAL;
  g_sm_budget = sms;
  return 0;
}

extern "C" int ms_debug_trap_info(int* out8) {
  if (!out8) return MS_EINVAL;
  for (int i = 0; i < 8; i++) out8[i] = g_trap_host ? g_trap_host[i] : 0;
  return 0;
}

extern "C" int ms_debug_phase_ts(unsigned long long* out16) {
  if (!out16) return MS_EINVAL;
  MS_CUDA(cudaMemcpyFromSymbol(out16, g_phase_ts, sizeof(unsigned long long) * 8 * MS_CHAIN_MAX));
  return 0;
}

extern "C" int ms_conv_chain_fwd(const ms_chain_fwd_layer* layers, int n, void* sync, void* stream) {
  if (!layers || n < 1 || n > CHAIN_MAX || !sync) return MS_EINVAL;
  FwdChain& ch = g_fwd_chain;
  if (phase_dbg()) { int rc0 = trap_buffer_init(); if (rc0) return rc0; }
  ch.n = n; ch.dbg = phase_dbg(); ch.sync = reinterpret_cast<unsigned int*>(sync);
  const int sms = chain_sms();
  long long want = 1;
  for (int i = 0; i < n; i++) {
    long long w = 0;
    int rc = fill_fwd_layer(&layers[i], &ch.L[i], &w);
    if (rc) return rc;
    if (w > want) want = w;
  }
  const unsigned grid = (unsigned)(want < sms ? want : sms);
  for (int i = 0; i < n; i++) {
    int rc = gemm_cfg(layers[i].d, ch.L[i].p, grid, ch.L[i].bn.training == 1, true, &ch.L[i].g);
    if (rc) return rc;
    if (ch.L[i].bn.training != 1 && !ch.L[i].g.resident && !ch.L[i].io.z) return MS_EINVAL;   // the streaming normalise pass re-reads z
  }
  const size_t smem = TB_RING_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(conv_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  void* args[] = {&ch};
  int rc = launch_coop(reinterpret_cast<const void*>(conv_chain_fwd_kernel), dim3(grid), smem, ms_stream(stream), args);
  if (rc) return rc;
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_conv_block_train_fwd(const ms_igemm_desc* d, const void* a, const void* w, float* z, const ms_block_bn* bn,
                                       float* y, void* planes, int pfmt, int64_t pstride, const float* res,
                                       const void* res_planes, int res_pfmt, int64_t res_pstride, int up2, void* sync,
                                       void* stream) {
  ms_chain_fwd_layer e;
  memset(&e, 0, sizeof(e));
  e.d = d; e.a = a; e.w = w; e.z = z; e.bn = bn; e.y = y; e.planes = planes; e.pfmt = pfmt; e.pstride = pstride;
  e.res = res; e.res_planes = res_planes; e.res_pfmt = res_pfmt; e.res_pstride = res_pstride; e.up2 = up2;
  return ms_conv_chain_fwd(&e, 1, sync, stream);
}

extern "C" int ms_conv_chain_bwd(const ms_chain_bwd_layer* layers, int n, void* sync, void* stream) {
  if (!layers || n < 1 || n > CHAIN_MAX || !sync) return MS_EINVAL;
  BwdChain& ch = g_bwd_chain;
  if (phase_dbg()) { int rc0 = trap_buffer_init(); if (rc0) return rc0; }
  ch.n = n; ch.dbg = phase_dbg(); ch.sync = reinterpret_cast<unsigned int*>(sync);
  const int sms = chain_sms();
  long long want = 1;
  for (int i = 0; i < n; i++) {
    const ms_chain_bwd_layer* e = &layers[i];
    BwdLayer& L = ch.L[i];
    if (!e->dy || !e->z || !e->bn || !e->dz_planes || e->rows < 1) return MS_EINVAL;
    if (e->pfmt != MS_BF16 && e->pfmt != MS_BF16X2) return MS_EINVAL;
    if (e->pfmt == MS_BF16X2 && (e->pstride <= 0 || (e->pstride * 2) % 8)) return MS_EINVAL;
    if (((uintptr_t)e->dy & 15) || ((uintptr_t)e->dy2 & 15) || ((uintptr_t)e->z & 15) || ((uintptr_t)e->dz_planes & 15) ||
        ((uintptr_t)e->dx & 15))
      return MS_EINVAL;
    if (e->gdt != MS_F32 && e->gdt != MS_F64) return MS_EINVAL;
    if (e->up2 && e->rows_per_seq < 1) return MS_EINVAL;
    int rc = fill_bn(L.bn, e->bn);
    if (rc) return rc;
    BwdIO& io = L.io;
    io.dy = e->dy; io.dy2 = e->dy2; io.z = e->z; io.dzp = reinterpret_cast<__nv_bfloat16*>(e->dz_planes); io.pfmt = e->pfmt;
    io.pstride = e->pstride; io.up2 = e->up2 ? 1 : 0; io.L = e->rows_per_seq; io.rows = e->rows; io.ggamma = e->grad_gamma;
    io.gbeta = e->grad_beta; io.gdt = e->gdt; io.dx = e->dx; io.has_gemm = e->dg ? 1 : 0; io.sync = nullptr; io.dbg = 0;
    // element-wise phases: one slab (64 channels x >= 16 rows) per CTA and pass
    long long w = (long long)((L.bn.C / 4 + 15) / 16) * ((e->rows + 15) / 16);
    if (e->dg) {
      if (!e->wt || !e->dx || e->dg->out_dtype != MS_F32 || e->dg->epilogue != 0) return MS_EINVAL;
      CUtensorMap maps[4];
      rc = igemm_prepare(e->dg, e->dz_planes, e->wt, e->dg->block_n, maps, &L.p);
      if (rc) return rc;
      L.map_a = maps[0]; L.map_w = maps[1]; L.map_a_lo = maps[2]; L.map_w_lo = maps[3];
      const long long items = gemm_tiles(L.p) * gemm_split(e->dg);
      if (items > w) w = items;
    } else {
      memset(&L.p, 0, sizeof(L.p));
      memset(&L.map_a, 0, 4 * sizeof(CUtensorMap));
      L.p.block_n = 32;
    }
    if (w > want) want = w;
  }
  const unsigned grid = (unsigned)(want < sms ? want : sms);
  for (int i = 0; i < n; i++) {
    if (layers[i].dg) {
      int rc = gemm_cfg(layers[i].dg, ch.L[i].p, grid, false, false, &ch.L[i].g);
      if (rc) return rc;
    } else {
      memset(&ch.L[i].g, 0, sizeof(GemmCfg));
      ch.L[i].g.nslots = 1; ch.L[i].g.stages = 2; ch.L[i].g.split_k = 1;
    }
  }
  const size_t smem = TB_RING_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(conv_chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  void* args[] = {&ch};
  int rc = launch_coop(reinterpret_cast<const void*>(conv_chain_bwd_kernel), dim3(grid), smem, ms_stream(stream), args);
  if (rc) return rc;
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_conv_block_train_bwd(const ms_igemm_desc* dg, const float* dy, const float* z, const ms_block_bn* bn,
                                       int64_t rows, int up2, int rows_per_seq, void* dz_planes, int pfmt, int64_t pstride,
                                       void* grad_gamma, void* grad_beta, int gdt, const void* wt, float* dx, void* sync,
                                       void* stream) {
  ms_chain_bwd_layer e;
  memset(&e, 0, sizeof(e));
  e.dg = dg; e.dy = dy; e.dy2 = nullptr; e.z = z; e.bn = bn; e.rows = rows; e.up2 = up2; e.rows_per_seq = rows_per_seq;
  e.dz_planes = dz_planes; e.pfmt = pfmt; e.pstride = pstride; e.grad_gamma = grad_gamma; e.grad_beta = grad_beta; e.gdt = gdt;
  e.wt = wt; e.dx = dx;
  return ms_conv_chain_bwd(&e, 1, sync, stream);
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

// tensor maps, tile program and CTA grid (gx pixel slices, gy output x input channel tiles, gz class x tap) of one weight gradient
static int fill_wgrad_layer(const ms_igemm_desc* d, const void* x, const void* dz, float* acc, WgradLayer* L) {
  if (!d || !x || !dz || !acc) return MS_EINVAL;
  if (d->num_classes < 1 || d->num_classes > MS_IGEMM_MAX_CLASSES || d->ntaps < 1 || d->cchunks < 1) return MS_EINVAL;
  if (((uintptr_t)x & 15) || ((uintptr_t)dz & 15) || ((uintptr_t)acc & 15)) return MS_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return MS_ENOTSUP;
  WgradAccParams& p = L->p;
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
  int box[5] = {64, bw, 1, bh, bb};
  const int32_t zdims[5] = {(int32_t)d->out_strides[0], Wo, 1, Ho, Bo};
  const int64_t zstr[5] = {1, d->out_strides[0], d->out_strides[1], d->out_strides[1], d->out_strides[2]};
  int rc = encode_5d_acc(enc, &L->map_x, x, d->a_dims, d->a_strides, box);
  if (rc) return rc;
  rc = encode_5d_acc(enc, &L->map_z, dz, zdims, zstr, box);
  if (rc) return rc;
  if (d->planes == 2) {
    rc = encode_5d_acc(enc, &L->map_x_lo, reinterpret_cast<const __nv_bfloat16*>(x) + d->a_plane_stride, d->a_dims, d->a_strides, box);
    if (rc) return rc;
    rc = encode_5d_acc(enc, &L->map_z_lo, reinterpret_cast<const __nv_bfloat16*>(dz) + d->out_plane_stride, zdims, zstr, box);
    if (rc) return rc;
  } else {
    L->map_x_lo = L->map_x; L->map_z_lo = L->map_z;
  }
  L->acc = acc;
  L->gx = p.split; L->gy = p.n_tiles * p.c_tiles; L->gz = d->num_classes * d->ntaps; L->first = 0;
  return 0;
}

static const size_t WGA_SMEM = (size_t)WGA_STAGES * 6 * WGA_CHUNK_BYTES + 1024;

extern "C" int ms_wgrad_bf16_acc(const ms_igemm_desc* d, const void* x, const void* dz, float* acc, void* stream) {
  WgradLayer& L = g_wgrad_multi.L[0];
  int rc = fill_wgrad_layer(d, x, dz, acc, &L);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(wgrad_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  dim3 grid((unsigned)L.gx, (unsigned)L.gy, (unsigned)L.gz);
  wgrad_acc_kernel<<<grid, NUM_THREADS, WGA_SMEM, ms_stream(stream)>>>(L.map_x, L.map_z, L.map_x_lo, L.map_z_lo, L.p, acc);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_wgrad_bf16_acc_multi(const ms_wgrad_item* items, int n, void* stream) {
  if (!items || n < 1 || n > CHAIN_MAX) return MS_EINVAL;
  WgradMulti& m = g_wgrad_multi;
  m.n = n;
  long long total = 0;
  for (int i = 0; i < n; i++) {
    int rc = fill_wgrad_layer(items[i].d, items[i].x, items[i].dz, items[i].acc, &m.L[i]);
    if (rc) return rc;
    m.L[i].first = (int)total;
    total += (long long)m.L[i].gx * m.L[i].gy * m.L[i].gz;
    if (total > 0x7fffffffLL) return MS_EINVAL;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MS_CUDA(cudaFuncSetAttribute(wgrad_acc_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  wgrad_acc_multi_kernel<<<dim3((unsigned)total), NUM_THREADS, WGA_SMEM, ms_stream(stream)>>>(m);
  MS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ms_unpack_wgrad_multi(const ms_wgrad_entry* table_dev, int n_entries, int blocks, void* stream) {
  if (!table_dev || n_entries < 1 || n_entries > MULTI_MAX_ENTRIES) return MS_EINVAL;
  if (blocks < 1) blocks = 8 * ms_num_sms();
  unpack_wgrad_multi_kernel<<<dim3((unsigned)blocks), 256, 0, ms_stream(stream)>>>(table_dev, n_entries);
  MS_LAUNCH_CHECK();
  return 0;
}
