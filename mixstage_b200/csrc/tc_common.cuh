// Device / host helpers shared by the tcgen05 kernels (conv_tc.cu: implicit-GEMM forward / input-gradient / weight-gradient
// launches; conv_train.cu: the fused training blocks): PTX wrappers for mbarriers, TMA, UMMA and TMEM loads, the kernel-side
// parameter block of an implicit GEMM, and the host-side tensor-map encoder lookup.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 12;        // smem ring depth is chosen per launch: as many stages as fit in 227 KB
constexpr int NUM_THREADS = 192;
constexpr uint32_t A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr uint32_t SPIN_LIMIT = 1u << 22;   // bounded mbarrier spins: trap instead of hanging the GPU

struct IgemmParams {
  int ntaps, cchunks, shared_taps;
  int num_classes, class_n, block_n, n_tiles_per_class;
  int box_w, box_h, box_b;
  int tiles_w, tiles_h, tiles_b;
  int out_w, out_h, out_b;
  long long os_w, os_h, os_b;           // output element strides
  int a_chan_base[MS_IGEMM_MAX_CLASSES];
  long long out_off[MS_IGEMM_MAX_CLASSES];
  short taps[MS_IGEMM_MAX_TAPS][4];     // chan_off, d_w, d_par, d_h
  int out_dtype, epilogue;
  float slope;
  int npass;                            // 1: bf16 operands; 3: split-bf16 (hi*hi + hi*lo + lo*hi)
  long long out_plane_stride;           // MS_BF16X2 output: elements between the hi and lo planes
  int stages;                           // smem ring depth (2..MAX_STAGES)
  int split_k;                          // > 1: gridDim.z CTAs share the k-steps of a tile, fp32 vector reductions into out
  // fused inference epilogue (ms_igemm_bf16_fused): extra fp32 copy of the result, UNet upsample(x2) + residual
  float* out_f32;                       // nullable: the result also as fp32 (same element offsets as `out`)
  const __nv_bfloat16* res;             // up2: residual as bf16 planes laid out like the (2x longer) output
  long long res_pstride;                // elements between the residual's hi and lo planes
  int res_planes;                       // 1 or 2
  int up2;                              // every GEMM row (b, w) produces output rows (b, 2w) and (b, 2w+1)
  // cluster mixture folded into the sub-decoder GEMMs (ms_igemm_bf16_mix; persistent kernel only)
  const float* row_w;                   // nullable: soft cluster weights [rows][row_w_stride]
  int row_w_stride;
  int row_w_mode;                       // 1: result *= row_w[row][class]; 2: result += sum_k row_w[row][k] * bias[k*N + n]
  int mix_k;                            // mode 2: number of clusters K (<= 16)
  int dbg;                              // MS_IGEMM_DBG (timing experiments only): bit 0 = lean epilogue computes but does not store
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128-byte swizzle: 8-row atoms of 1024 B, SBO = 1024 B, LBO unused (=1), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// wait for the outstanding tcgen05.ld; the loaded registers pass THROUGH the statement so that no use can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}


// ---- host: tensor maps and the kernel-side parameter block of one implicit-GEMM launch (forward or input gradient)
// maps[0..3] = A, W, A lo plane, W lo plane (lo = hi when the operands are plain bf16); W boxes are w_box_rows rows tall.
static int igemm_prepare(const ms_igemm_desc* d, const void* a, const void* w, int w_box_rows, CUtensorMap* maps, IgemmParams* pp) {
  if (!d || !a || !w) return MS_EINVAL;
  if (d->num_classes < 1 || d->num_classes > MS_IGEMM_MAX_CLASSES) return MS_EINVAL;
  if (d->ntaps < 1 || d->cchunks < 1) return MS_EINVAL;
  const int tap_rows = d->shared_taps ? d->ntaps : d->ntaps * d->num_classes;
  if (tap_rows > MS_IGEMM_MAX_TAPS) return MS_EINVAL;
  if (d->block_n < 16 || d->block_n > 256 || d->block_n % 16) return MS_EINVAL;
  if (d->class_n % 16 || d->class_n < 16) return MS_EINVAL;
  if (d->box[0] != BLOCK_K || d->box[2] != 1 || d->box[1] * d->box[3] * d->box[4] != BLOCK_M) return MS_EINVAL;
  if (((uintptr_t)a & 15) || ((uintptr_t)w & 15)) return MS_EINVAL;
  if (d->planes != 1 && d->planes != 2) return MS_EINVAL;
  const int planes = d->planes;
  if (planes == 2 && ((d->a_plane_stride * 2) % 16 || (d->w_plane_stride * 2) % 16 || d->a_plane_stride <= 0 || d->w_plane_stride <= 0))
    return MS_EINVAL;
  EncodeTiledFn enc = get_encode();
  if (!enc) return MS_ENOTSUP;
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
      CUresult r = enc(&maps[pl ? 2 : 0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(ap), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return MS_EINVAL;
    }
    {
      const long long ktot = (long long)d->ntaps * d->cchunks * BLOCK_K;
      cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)((long long)d->num_classes * d->class_n)};
      cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
      cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)w_box_rows}, es[2] = {1, 1};
      CUresult r = enc(&maps[pl ? 3 : 1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wq), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return MS_EINVAL;
    }
  }
  if (planes == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }
  IgemmParams& p = *pp;
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
  p.row_w = nullptr; p.row_w_stride = 0; p.row_w_mode = 0; p.mix_k = 0; p.dbg = 0;
  p.stages = 0;
  const int num_k_total = d->ntaps * d->cchunks * p.npass;
  int split = d->split_k > 1 ? d->split_k : 1;
  if (split > num_k_total) split = num_k_total;
  const int per = (num_k_total + split - 1) / split;
  p.split_k = (num_k_total + per - 1) / per;            // every slice owns at least one k-step
  // fp32 vector stores / reductions need 16-byte aligned rows
  if ((p.os_w * 4) % 16 || (p.os_h * 4) % 16 || (p.os_b * 4) % 16) return MS_EINVAL;
  for (int i = 0; i < d->num_classes; i++)
    if ((p.out_off[i] * 4) % 16) return MS_EINVAL;
  return 0;
}

}  // namespace
