/*
 * mixstage_b200 C-ABI  --  B200 (sm_100a) kernels for the Mix-StAGE generator hot path.
 *
 * Drop-in boundary: the reference has no native layer at all -- its hot
 * path is `JointLateClusterSoftStyle4_G.forward` / `Speech2Gesture_D.forward` /
 * `GAN.forward` calling torch.nn ops (cuDNN/ATen in fp64).  Each entry point below
 * replaces one ATen op family at the reference call site quoted next to it
 * (paths relative to /root/reference/src/model/).  The Python host mirror
 * (mixstage_b200/*.py) binds them with ctypes; INTEGRATION.md shows the binding and
 * how the classes are installed under the reference's own names.
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a cudaError_t value (or a negative
 *     MS_E* code for argument errors); nothing is printed, nothing throws.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises.  Functions are re-entrant per device (one process per GPU).
 *   - activations are channels-last: 1-D (B, L, C) is passed as H=1, W=L.
 *     2-D is (B, H, W, C).  fp32 everywhere in the `_f32` family; bf16 operands with fp32
 *     accumulation in the `_bf16` (tcgen05) family.
 *   - `pdt` arguments give the dtype of caller-owned parameter tensors
 *     (MS_F32 / MS_F64): the reference keeps fp64 master parameters
 *     (trainer.py:138 `.double()`), so parameters/grad buffers are read and written in
 *     the caller's dtype while all arithmetic runs in fp32/bf16.
 */
#ifndef MIXSTAGE_B200_H
#define MIXSTAGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_F32 0
#define MS_F64 1
#define MS_BF16 2
#define MS_BF16X2 3      /* split-bf16: a hi plane and a lo plane, v ~= hi + lo (~16 mantissa bits) */

#define MS_EINVAL (-1)   /* bad argument (shape/stride/alignment not supported)   */
#define MS_ENOTSUP (-2)  /* valid request this build cannot serve (e.g. no sm_100) */

/* Geometry of one convolution, NHWC activations.  Replaces the nn.Conv1d/nn.Conv2d
 * construction in ConvNormRelu (layers.py:58-70): Cin/Cout are TOTAL channel counts
 * (already multiplied by groups as at layers.py:58-59). */
typedef struct ms_conv_desc {
  int32_t B, H, W, Cin, Cout;
  int32_t kh, kw, sh, sw, ph, pw, groups;
  int32_t Ho, Wo;
} ms_conv_desc;

int ms_version(void);
/* 1 when the current device is compute capability 10.x (tcgen05 path usable). */
int ms_device_is_sm100(void);

/* ---- weight packing (replaces nothing in the reference: ATen consumes (Cout,Cin/g,kh,kw)
 * directly; we re-tile once per parameter version) --------------------------------- */
/* w (Cout, Cin/g, kh, kw) in dtype pdt -> wf[g][tap][c][n] (forward / wgrad layout) and
 * wt[g][tap][n][c] (dgrad layout), both fp32.  Either output may be NULL. */
int ms_pack_conv_weight_f32(const void* w, int pdt, const ms_conv_desc* d, float* wf, float* wt, void* stream);
/* dwf[g][tap][c][n] fp32 -> dw (Cout, Cin/g, kh, kw) in dtype pdt. */
int ms_unpack_conv_wgrad(const float* dwf, const ms_conv_desc* d, void* dw, int pdt, int accumulate, void* stream);
/* accumulate != 0 (here and below): dw += ... instead of dw = ..., so a train step can write parameter gradients
 * straight into its flat, pre-zeroed gradient buffer (no per-tensor accumulation kernels afterwards). */
/* dst[i] = (T_dst) src[i]; dtypes MS_F32/MS_F64/MS_BF16. */
int ms_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream);
/* dst[i] = (T_dst)(src[i] * scale), MS_F32/MS_F64, src == dst allowed for equal dtypes.  The staging casts of the
 * data-parallel gradient exchange (fp64 flat gradients <-> fp32 all-reduce buffer, scaled by 1/world): replaces the
 * `grad / world_size` + NCCL bucket copies a DistributedDataParallel wrapper around trainer.py:1138-1146 would do. */
int ms_scale_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, double scale, void* stream);

/* ---- convolution as implicit GEMM, fp32 SIMT (conv1d/conv2d call sites:
 * layers.py:60-70 via :78; speech2gesture.py:76,90; jlcss.py:83; layers.py:459) ------ */
/* y (B,Ho,Wo,Cout) = conv(x (B,H,W,Cin), wf) + bias (bias nullable; bias in fp32).
 * act: 0 none, 1 LeakyReLU(slope) applied in the epilogue (speech2gesture.py:76-77). */
int ms_conv_fwd_f32(const float* x, const float* wf, const float* bias, float* y,
                    const ms_conv_desc* d, int act, float slope, void* stream);
/* dx (B,H,W,Cin) = conv_transpose(dy (B,Ho,Wo,Cout), wt)  (aten::convolution_backward, input grad) */
int ms_conv_dgrad_f32(const float* dy, const float* wt, float* dx, const ms_conv_desc* d, void* stream);
/* dwf[g][tap][c][n] = sum_pos x[pos+tap] * dy[pos]   (convolution_backward, weight grad).
 * Overwrites dwf (zeroed internally; split-K partials are combined with fp32 atomics). */
int ms_conv_wgrad_f32(const float* x, const float* dy, float* dwf, const ms_conv_desc* d, void* stream);

/* Inference form of the C_in = 1 convolution (audio_encoder.conv.0, layers.py:167: 3x3 over the (T, F) log-mel image,
 * K = 9 -- HBM-bound, not tensor-core work) with the eval-mode BatchNorm folded to scale/shift (conv bias inside shift)
 * and LeakyReLU(slope) applied in the same pass: v = act(conv(x) * scale[n] + shift[n]) written as fp32 (y, nullable)
 * and/or as the next layer's bf16 operand planes (nullable; pfmt MS_BF16 | MS_BF16X2, lo plane at + pstride).
 * wf is the [tap][0][n] fp32 tiling of ms_pack_conv_weight_f32.  Cout % 8 == 0. */
int ms_conv_cin1_bnact(const float* x, const float* wf, const float* scale, const float* shift, float slope,
                       const ms_conv_desc* d, float* y, void* planes, int pfmt, int64_t pstride, void* stream);

/* ---- implicit-GEMM convolution on tcgen05 tensor cores (bf16 operands, fp32 accumulate in TMEM,
 * operands staged by TMA).  Same call sites as the fp32 family above; this is the fast path
 * ("precision=bf16").  One descriptor type serves the forward pass and the input gradient:
 *
 *   out[b,h,w, off[q] + n] = sum_t sum_c A5[base[q] + taps[q][t].chan + c, w + taps.dw, taps.par, h + taps.dh, b]
 *                                        * Wp[q*class_n + n][t][c]          (+ bias | * scale + shift, LeakyReLU)
 *
 * A5 is the bf16 activation seen as a 5-D tensor (channel, w-like, h-parity, h-like, batch) with
 * element strides a_strides; reads outside it are zero (the convolution's padding).  Wp is the
 * re-tiled weight [classes*class_n][ntaps][cchunks*64] bf16 (ms_pack_igemm_weight_bf16).
 * box = {64, bw, 1, bh, bb} with bw*bh*bb = 128 rows per tile.  Requires sm_100. */
#define MS_IGEMM_MAX_CLASSES 16
#define MS_IGEMM_MAX_TAPS 48
typedef struct ms_igemm_desc {
  int32_t a_dims[5];
  int64_t a_strides[5];
  int32_t box[5];
  int32_t out_dims[3];      /* w-like, h-like, batch extents of the output */
  int64_t out_strides[3];   /* element strides of the output for (w, h, b)  */
  int32_t num_classes, class_n, block_n;
  int32_t ntaps, cchunks, shared_taps;
  int32_t a_chan_base[MS_IGEMM_MAX_CLASSES];
  int64_t out_off[MS_IGEMM_MAX_CLASSES];
  int16_t taps[MS_IGEMM_MAX_TAPS][4];   /* chan offset, w shift, h-parity coordinate, h shift */
  int32_t out_dtype;        /* MS_F32, MS_BF16 or MS_BF16X2 (hi plane at out, lo plane at out + out_plane_stride) */
  int32_t epilogue;         /* 0: + bias (nullable); 1: * scale[n] + shift[n] then LeakyReLU(slope);
                               2: + bias (nullable) then LeakyReLU(slope) */
  float slope;
  /* Split-bf16 operands ("bf16x3", ~fp32 accuracy on the bf16 tensor pipe): planes == 2 means A and Wp each
   * come as a hi plane and a lo plane (lo at + *_plane_stride elements) and every k-step is issued three
   * times: hi*hi + hi*lo + lo*hi.  planes == 1: plain bf16 operands.  For ms_wgrad_bf16 the dz operand's
   * lo plane sits at dz + out_plane_stride. */
  int32_t planes;
  int64_t a_plane_stride, w_plane_stride, out_plane_stride;
  /* Split-K for launches whose tile count cannot fill the 148 SMs (batch-16 layers, K up to 6144).
   * ms_igemm_bf16: split_k > 1 slices the k-steps of every tile over gridDim.z; partial tiles are combined with fp32
   *   vector reductions into `out`, which is zero-filled first (out_numel fp32 elements; MS_F32 output, epilogue 0 only).
   * ms_wgrad_bf16: split_k slices the pixel rows; slice s writes its partial dWp at dwp + s * wp_numel (the caller
   *   allocates split_k partials; ms_unpack_igemm_wgrad sums them).  0/1 = no split. */
  int32_t split_k;
  int64_t out_numel;
  int32_t wgrad_c_tile;     /* ms_wgrad_bf16: input-channel columns per CTA tile (64/128/192/256; 0 = 256) */
} ms_igemm_desc;
int ms_igemm_bf16(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                  const float* shift, void* out, void* stream);
/* Inference form of the same launch (eval-mode ConvNormRelu, layers.py:78 with the BatchNorm folded to scale/shift):
 * the epilogue result goes out as the NEXT layer's operand planes (`out`, d->out_dtype MS_BF16 / MS_BF16X2) and, when
 * out_f32 is not NULL, also as fp32 at the same element offsets -- no fp32 round trip through HBM between layers.
 * up2 != 0 fuses UNet1D's `upconv(x) + residual` (layers.py:150-151, 1-D only): GEMM row (b, w) produces output rows
 * (b, 2w) and (b, 2w+1) of a tensor with 2*out_dims[0] rows per sequence (out_strides[2] is the stride of the
 * UN-doubled tensor), each plus the residual row read from `res` (bf16 planes laid out like that output; res_planes 1|2,
 * lo plane at + res_pstride elements).  split_k must be <= 1. */
int ms_igemm_bf16_fused(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                        const float* shift, void* out, float* out_f32, const void* res, int res_planes,
                        int64_t res_pstride, int up2, void* stream);
/* The soft cluster mixture (index_select_outputs, joint_late_cluster_soft_style.py:106-115, called at :194) folded into
 * the sub-decoder GEMMs so that the per-cluster outputs (B,T,K*P) never reach HBM:
 *   row_w_mode 1 (last grouped decoder block, classes = clusters): result[row, class q columns] *= row_w[row*row_w_stride + q]
 *                 applied after scale/shift + LeakyReLU, before the bf16 conversion;
 *   row_w_mode 2 (the grouped 1x1 `logits` conv run as ONE dense GEMM over the K*256 weighted channels, whose accumulator
 *                 is then sum_k w_k * logits_k): result[row, n] += sum_{k<mix_k} row_w[row*row_w_stride + k] * bias[k*N + n],
 *                 N = num_classes*class_n <= 128, epilogue 0, mix_k <= 16.
 * row = (b*out_dims[1] + h)*out_dims[0] + w.  Same outputs as ms_igemm_bf16_fused (planes and/or fp32); split_k <= 1. */
int ms_igemm_bf16_mix(const ms_igemm_desc* d, const void* a, const void* w, const float* bias, const float* scale,
                      const float* shift, void* out, float* out_f32, const float* row_w, int row_w_stride, int row_w_mode,
                      int mix_k, void* stream);
/* Re-tile a conv weight (Cout, Cin/g, kh*kw) of dtype pdt into Wp (bf16).
 * mode 0 (forward):  Wp[q*class_n + r][t][c] = w[q*class_n + r][c][srctap[t]]            (c < Cin/g, else 0)
 * mode 1 (dgrad):    Wp[q*class_n + r][t][n] = w[g*Cout/g + n][r][srctap[q*ntaps + t]]   (n < Cout/g, r < Cin/g, else 0)
 *                    with g = q when groups > 1 (classes are groups) and g = 0 otherwise (classes are parities).
 * srctap_host: HOST array of ntaps (mode 0) or num_classes*ntaps (mode 1) source tap indices.
 * wp_lo (nullable): lo plane for the split-bf16 mode, bf16(w - float(wp)). */
int ms_pack_igemm_weight_bf16(const void* w, int pdt, int Cout, int Cin_g, int taps_total, int groups, int mode,
                              int num_classes, int class_n, int ntaps, int kpad, const int16_t* srctap_host,
                              void* wp, void* wp_lo, void* stream);

/* The same re-tiling for MANY weights in one launch: table_dev is a DEVICE array of n_entries records (arguments of
 * ms_pack_igemm_weight_bf16, srctap inline).  Used by the train step to refresh every packed weight of the sub-network
 * that the optimiser just updated. */
typedef struct ms_pack_entry {
  const void* w;
  void* wp;
  void* wp_lo;              /* nullable */
  int32_t pdt, Cout, Cin_g, taps_total, groups, mode, num_classes, class_n, ntaps, kpad;
  int16_t srctap[MS_IGEMM_MAX_TAPS];
} ms_pack_entry;
/* blocks: CTAs of the launch (<= 0: 8 per SM); work units are spread over the entries in proportion to their size.
 * wp / wp_lo are 16-byte aligned (a thread converts eight adjacent k positions and stores them at once). */
int ms_pack_igemm_weight_multi(const ms_pack_entry* table_dev, int n_entries, int blocks, void* stream);

/* Weight gradient on tcgen05 (aten::convolution_backward, weight grad): with d the FORWARD descriptor,
 *   dwp[q*class_n + n][t][c] = sum_{b,h,w} dz[b,h,w, off[q] + n] * A5[base[q] + taps[t].chan + c, w + dw, par, h + dh, b]
 * x and dz are bf16 (dz has the forward output's shape), dwp is fp32 [classes*class_n][ntaps][cchunks*64]
 * per split-K slice (see split_k above; every partial is fully overwritten).  Both operands are MN-major UMMA operands. */
int ms_wgrad_bf16(const ms_igemm_desc* d, const void* x, const void* dz, float* dwp, void* stream);
/* sum of the nsplit partial dwp (forward tiling, one source tap per tap) -> dw (Cout, Cin/g, taps) in dtype pdt. */
int ms_unpack_igemm_wgrad(const float* dwp, int Cout, int Cin_g, int taps_total, int ntaps, int kpad, void* dw, int pdt,
                          int nsplit, int accumulate, void* stream);

/* ---- Fused TRAINING blocks (csrc/conv_train.cu): ConvNormRelu.forward in train mode (layers.py:78: conv -> BatchNorm with
 * batch statistics -> LeakyReLU, plus UNet1D's `upconv(x) + residual`, layers.py:151) and its backward, ONE persistent
 * cooperative launch each (grid <= #SMs, device-wide barriers between the phases) instead of three dependent kernels. */
typedef struct ms_block_bn {
  int32_t C;                /* channels = num_classes * class_n of the GEMM */
  int32_t pdt;              /* dtype of gamma / beta / conv_bias / running statistics: MS_F32 or MS_F64 */
  int32_t training;         /* forward: must be 1; backward: 1 = batch-statistics formula, 0 = dz = scale * dy * act' */
  float momentum, eps, slope;
  const void* gamma;
  const void* beta;
  const void* conv_bias;    /* nullable; the GEMM output excludes it (enters the running-mean update only) */
  void* running_mean;       /* updated in place (forward) */
  void* running_var;
  int64_t* num_batches_tracked;   /* nullable, += 1 (forward) */
  double* sums;             /* [2][C] ZERO-FILLED by the caller: forward sum / sum of squares; backward dgamma / dbeta */
  float* ss;                /* [4][C] scale, shift, mean, rstd: written by the forward, read by the backward */
} ms_block_bn;
/* Forward: z = igemm(d, a, w) (fp32, dense (rows, C); ZERO-FILLED by the caller when d->split_k > 1: k-slices combine with
 * vector reductions), batch statistics of z, finalize (as ms_bn_finalize, training) and y = act(z*scale + shift)
 * [up2: y[b,2l+r,:] = act(..)[b,l,:] + res[b,2l+r,:]] as fp32 (y, nullable) and/or operand planes (nullable).
 * d: forward descriptor with out_dtype MS_F32, epilogue 0.  sync: 4 zero-filled bytes (device barrier counter). */
int ms_conv_block_train_fwd(const ms_igemm_desc* d, const void* a, const void* w, float* z, const ms_block_bn* bn,
                            float* y, void* planes, int pfmt, int64_t pstride, const float* res, const void* res_planes,
                            int res_pfmt, int64_t res_pstride, int up2, void* sync, void* stream);
/* bn->training == 0 turns the same launch into the small-batch form of the INFERENCE block (eval-mode ConvNormRelu with
 * the BatchNorm folded into bn->ss = [scale, shift] by ms_bn_finalize; no statistics phase, nothing updated): at batch
 * 16 a layer has fewer tiles than SMs, and k-slices across the whole machine + one barrier beat ms_igemm_bf16_fused's one
 * CTA per tile.  The skip tensor of up2 comes as fp32 (res) or as operand planes (res_planes, res_pfmt, res_pstride). */
/* Backward: dz = d(act(bn(z)))/dz . dy -> operand planes dz_planes (row stride C; the A operand of the input-gradient
 * GEMM and of ms_wgrad_bf16[_acc]); grad_gamma / grad_beta (nullable, dtype gdt) += dgamma / dbeta; then, when dg is not
 * NULL, dx = igemm(dg, dz_planes, wt) (fp32; ZERO-FILLED by the caller when dg->split_k > 1).  dy has 2*rows_per_seq rows
 * per sequence when up2 (both upsampled rows feed the same z row). */
int ms_conv_block_train_bwd(const ms_igemm_desc* dg, const float* dy, const float* z, const ms_block_bn* bn,
                            int64_t rows, int up2, int rows_per_seq, void* dz_planes, int pfmt, int64_t pstride,
                            void* grad_gamma, void* grad_beta, int gdt, const void* wt, float* dx, void* sync,
                            void* stream);
/* CHAINS: up to MS_CHAIN_MAX consecutive blocks in ONE cooperative launch (block i+1 reads the operand planes block i
 * wrote; a device-wide barrier separates them).  At batch 16 a train step is bound by the ~280 dependent launches it is
 * made of, not by their arithmetic; a chain turns the 6-12 launches of a conv stack (UNet1D, ClusterClassify, the grouped
 * sub-decoders, AudioEncoder, PoseStyleEncoder -- layers.py:80-157,159-199,246-289,446-467) into one.
 * Forward layer = the arguments of ms_conv_block_train_fwd; `sync` (4 zero-filled bytes) is shared by the chain. */
#define MS_CHAIN_MAX 16
typedef struct ms_chain_fwd_layer {
  const ms_igemm_desc* d;
  const void* a;
  const void* w;
  float* z;
  const ms_block_bn* bn;
  float* y;
  void* planes;
  int32_t pfmt;
  int64_t pstride;
  const float* res;
  const void* res_planes;
  int32_t res_pfmt;
  int64_t res_pstride;
  int32_t up2;
} ms_chain_fwd_layer;
int ms_conv_chain_fwd(const ms_chain_fwd_layer* layers, int n, void* sync, void* stream);
/* Backward layers in EXECUTION order (last block of the forward chain first) = the arguments of ms_conv_block_train_bwd
 * plus dy2 (nullable): a second addend of the incoming gradient, laid out like dy -- the gradient a later block of the
 * chain received for an output that used this block's output as its skip tensor (UNet1D, layers.py:150-152). */
typedef struct ms_chain_bwd_layer {
  const ms_igemm_desc* dg;
  const float* dy;
  const float* dy2;
  const float* z;
  const ms_block_bn* bn;
  int64_t rows;
  int32_t up2, rows_per_seq;
  void* dz_planes;
  int32_t pfmt;
  int64_t pstride;
  void* grad_gamma;
  void* grad_beta;
  int32_t gdt;
  const void* wt;
  float* dx;
} ms_chain_bwd_layer;
int ms_conv_chain_bwd(const ms_chain_bwd_layer* layers, int n, void* sync, void* stream);
/* SM budget of the chain launches (0 = whole device).  Data-parallel training sets it a little below the SM count so that the
 * NCCL kernels of the overlapped gradient exchange find free SMs beside a chain launch (which is persistent and would
 * otherwise hold all of them until it ends). */
int ms_set_chain_sm_budget(int sms);
/* The weight gradients of up to MS_CHAIN_MAX blocks (arguments of ms_wgrad_bf16_acc each) in ONE launch. */
typedef struct ms_wgrad_item {
  const ms_igemm_desc* d;
  const void* x;
  const void* dz;
  float* acc;
} ms_wgrad_item;
int ms_wgrad_bf16_acc_multi(const ms_wgrad_item* items, int n, void* stream);
/* Timing experiments only (MS_PHASE_TS=1 in the environment when the library is first used): %globaltimer stamps of
 * CTA 0 at the phase boundaries of the LAST fused launch, 8 per block of the chain (8 * MS_CHAIN_MAX values) copied to host
 * memory. */
int ms_debug_phase_ts(unsigned long long* out16);
/* Debugging aid (MS_PHASE_TS=1): where a bounded spin inside a chain launch gave up, from mapped host memory (readable after
 * the context died): [0] site (1 producer/empty, 2 MMA/accumulator free, 3 MMA/operands, 4 epilogue/accumulator full,
 * 5 producer drain, 6 device barrier), [1] CTA, [2] thread, [3] block of the chain (+100: backward), [4..5] site data. */
int ms_debug_trap_info(int* out8);
/* ms_wgrad_bf16 with every pixel slice ADDING its tile into one fp32 accumulator acc[classes*class_n][ntaps][cchunks*64]
 * (zero-filled by the caller once per step) -- no per-slice partials, no summing kernel. */
int ms_wgrad_bf16_acc(const ms_igemm_desc* d, const void* x, const void* dz, float* acc, void* stream);
/* Every weight-gradient accumulator of a sub-network -> its parameter-gradient tensor (Cout, Cin/g, taps) in ONE launch:
 * dw[o][c][t] (+)= acc[o][t][c].  table_dev: DEVICE array. */
typedef struct ms_wgrad_entry {
  const void* acc;
  void* dw;
  int32_t pdt, Cout, Cin_g, taps, kpad, accumulate;
} ms_wgrad_entry;
int ms_unpack_wgrad_multi(const ms_wgrad_entry* table_dev, int n_entries, int blocks, void* stream);

/* ---- BatchNorm (+LeakyReLU) pieces, nn.BatchNorm1d/2d at layers.py:64,70 and
 * nn.LeakyReLU(0.2) at layers.py:72-73, applied as in ConvNormRelu.forward (:78) ------ */
/* column sums over rows: sum[c] += x[r,c], sumsq[c] += x[r,c]^2 (double accumulators,
 * caller zeroes them).  sumsq nullable.  Also used for conv-bias gradients. */
int ms_col_stats_f32(const float* x, int64_t rows, int C, double* sum, double* sumsq, void* stream);
/* training: batch mean / biased var from (sum,sumsq,rows) -> scale = gamma*rstd,
 * shift = beta - mean*scale, save mean/rstd; update running_mean/var in place
 * (momentum, unbiased var) in dtype pdt.  eval (training=0): scale/shift from the running
 * statistics, sum/sumsq ignored.  conv_bias (nullable, dtype pdt): bias of the producing convolution when
 * the GEMM output x was stored WITHOUT it (tensor-core path): batch-stat BN cancels a per-channel constant,
 * so it only enters the running-mean update (training) or the shift (eval). */
int ms_bn_finalize(const double* sum, const double* sumsq, int64_t rows, int C,
                   const void* gamma, const void* beta, const void* conv_bias, void* running_mean, void* running_var, int pdt,
                   int training, float momentum, float eps,
                   float* scale, float* shift, float* mean, float* rstd, void* stream);
/* Training-mode statistics and finalize in ONE launch: column sums as ms_col_stats_f32 (sum/sumsq/ticket zeroed by
 * the caller; ticket = 4 bytes), then the block that finishes last runs the training branch of ms_bn_finalize for every
 * channel and adds 1 to num_batches_tracked (nullable, device int64). */
int ms_bn_stats_finalize(const float* x, int64_t rows, int C, double* sum, double* sumsq, void* ticket,
                         const void* gamma, const void* beta, const void* conv_bias, void* running_mean, void* running_var,
                         int64_t* num_batches_tracked, int pdt, float momentum, float eps,
                         float* scale, float* shift, float* mean, float* rstd, void* stream);
/* y = act(x*scale[c] + shift[c]) over rows x C; act LeakyReLU(slope) when slope != 1.
 * up2 != 0 fuses UNet1D's `upconv(x) + residual` (layers.py:151): y has 2*L rows per
 * sequence, y[b,2l+r,:] = act(..)[b,l,:] + res[b,2l+r,:]  (L = rows_per_seq). */
int ms_bn_act_fwd_f32(const float* x, const float* scale, const float* shift, float slope,
                      int64_t rows, int C, float* y, const float* res, int up2, int rows_per_seq,
                      void* planes, int pfmt, int64_t pstride, void* stream);
/* "planes": optional second output of the activation as bf16 tensor-core operand planes, same element
 * order as y: pfmt MS_BF16 (one plane) or MS_BF16X2 (hi plane, then lo = bf16(v - hi) at + pstride elements).
 * y may be NULL when only the planes are wanted. */
/* x (rows, C) fp32 -> planes with row stride row_stride >= C (pad columns zero). */
int ms_to_planes(const float* x, int64_t rows, int C, int row_stride, void* planes, int pfmt, int64_t pstride, void* stream);
/* inverse: planes (row stride row_stride) -> x (rows, C) fp32 = hi (+ lo). */
int ms_planes_to_f32(const void* planes, int pfmt, int64_t pstride, int64_t rows, int C, int row_stride, float* x, void* stream);
/* backward reductions of act(bn(x)):  dz = dy * (z>0 ? 1 : slope) with z = x*scale+shift,
 * dbeta[c] += sum dz, dgamma_hat[c] += sum dz * xhat  (xhat = (x-mean)*rstd), doubles,
 * caller zeroes.  up2: dy has 2*L rows per sequence and the two rows of a pair are summed. */
int ms_bn_act_bwd_reduce_f32(const float* dy, const float* x, const float* scale, const float* shift,
                             const float* mean, const float* rstd, float slope, int64_t rows, int C,
                             int up2, int rows_per_seq, double* dgamma, double* dbeta, void* stream);
/* dx = scale * (dz - dbeta/rows - xhat*dgamma/rows)  (training) or scale*dz (training=0). */
int ms_bn_act_bwd_apply_f32(const float* dy, const float* x, const float* scale, const float* shift,
                            const float* mean, const float* rstd, float slope, int64_t rows, int C,
                            int up2, int rows_per_seq, const double* dgamma, const double* dbeta,
                            int training, float* dx, void* planes, int pfmt, int64_t pstride,
                            void* grad_gamma, void* grad_beta, int gdt, void* stream);
/* grad_gamma / grad_beta (nullable, dtype gdt): the affine parameters' gradient buffers, += dgamma / dbeta. */
/* dz = dy * (y > 0 ? 1 : slope) for a plain conv + LeakyReLU (speech2gesture.py:76-77); y is the
 * activation output. */
int ms_lrelu_bwd_f32(const float* dy, const float* y, float slope, int64_t n, float* dz, void* planes, int pfmt,
                     int64_t pstride, void* stream);
/* out[i] (dtype pdt) = (T) in[i] for small per-channel vectors (dgamma/dbeta/dbias). */
int ms_store_param_grad(const double* src, int n, void* dst, int pdt, int accumulate, void* stream);

/* ---- resize / glue ---------------------------------------------------------------- */
/* torch.nn.functional.interpolate(x, size=(T,1), mode='bilinear') + squeeze (layers.py:197-198):
 * x (B,Hi,Wi,C) -> y (B,T,C). */
int ms_bilinear_to_T_fwd_f32(const float* x, int B, int Hi, int Wi, int C, int T, float* y, void* stream);
int ms_bilinear_to_T_bwd_f32(const float* dy, int B, int Hi, int Wi, int C, int T, float* dx, void* stream);
/* style embedding + concat (layers.py:659-663, jlcss.py:175-180):
 * out (rows, C+sd): out[r,:C] = x[r,:], out[r,C:] = emb[idx[r/rep]] ('emb' mode, idx int64)
 * or soft[r/rep,:] @ emb ('lin' mode, soft (rows/rep, S) fp32).  emb (S, sd) in dtype pdt.
 * rep = T when the style is constant over the sequence and given per sequence, else 1. */
int ms_style_concat_fwd_f32(const float* x, int64_t rows, int C, const int64_t* idx, const float* soft,
                            int rep, const void* emb, int pdt, int S, int sd, float* out, void* stream);
/* Warp-per-row form of the same concat (C % 128 == 0, sd <= 32): 16-byte loads, 8-byte stores, and the row is ALSO emitted as
 * bf16 operand planes (nullable; row stride rs >= C + sd elements, padding zero-filled, hi [+ lo at pstride]) for the
 * tensor-core consumers (ClusterClassify.conv.0 / decoder.0), so no separate fp32 -> planes pass runs over the features. */
int ms_style_concat_planes_fwd_f32(const float* x, int64_t rows, int C, const int64_t* idx, const float* soft, int rep,
                                   const void* emb, int pdt, int S, int sd, float* out, void* planes, int pfmt,
                                   int64_t pstride, int rs, void* stream);
/* backward: dx[r,:] = dout[r,:C]; demb[s,:] += sum_{r: idx=s} dout[r,C:] (fp32 table, caller
 * zeroes; 'lin': demb += soft^T dout_style, dsoft[q,:] = sum_{r in q} dout_style[r,:] @ emb^T). */
int ms_style_concat_bwd_f32(const float* dout, int64_t rows, int C, const int64_t* idx, const float* soft,
                            int rep, const void* emb, int pdt, int S, int sd,
                            float* dx, float* demb, float* dsoft, void* stream);
/* softmax over K + cross-entropy (mean over rows) + argmax (jlcss.py:183-187, :159-165, :203):
 * score (rows,K) -> soft (rows,K) (nullable), amax (rows) int64 (nullable),
 * loss_sum += sum_r -log soft[r,target[r/trep]] (double, caller zeroes; nullable with target). */
int ms_softmax_ce_fwd_f32(const float* score, int64_t rows, int K, const int64_t* target, int trep,
                          float* soft, int64_t* amax, double* loss_sum, void* stream);
/* dscore = g_ce/rows * (soft - onehot) + soft * (dsoft - sum_k soft*dsoft); g_ce device scalar
 * (nullable), dsoft nullable. */
int ms_softmax_ce_bwd_f32(const float* soft, int64_t rows, int K, const int64_t* target, int trep,
                          const float* g_ce, const float* dsoft, float* dscore, void* stream);
/* index_select_outputs (jlcss.py:106-115): out[r,p] = sum_k w[r,k] * z[r,k*P+p]. */
int ms_mixture_fwd_f32(const float* z, const float* w, int64_t rows, int K, int P, float* out, void* stream);
int ms_mixture_bwd_f32(const float* dout, const float* z, const float* w, int64_t rows, int K, int P,
                       float* dz, float* dw, void* stream);
/* x.mean(-1) of PoseStyleEncoder (layers.py:287): (B,L,C) -> (B,C) and its adjoint. */
int ms_mean_rows_fwd_f32(const float* x, int B, int L, int C, float* y, void* stream);
int ms_mean_rows_bwd_f32(const float* dy, int B, int L, int C, float* dx, void* stream);

/* ---- GAN losses (gan.py:47-52, 64-75) ---------------------------------------------- */
/* velocity: v[b,0,:]=0, v[b,t,:]=x[b,t,:]-x[b,t-1,:]; adjoint in _bwd. */
int ms_velocity_fwd_f32(const float* x, int B, int T, int P, float* v, void* stream);
int ms_velocity_bwd_f32(const float* dv, int B, int T, int P, float* dx, void* stream);
/* L1Loss(reduction='none') then mean: loss_sum += sum |a - b| (b nullable -> constant c);
 * sgn (nullable) receives sign(a-b) for the backward. */
int ms_l1_fwd_f32(const float* a, const float* b, float c, int64_t n, double* loss_sum, float* sgn, void* stream);
/* da = g[0] * sgn / n  (g device scalar) */
int ms_l1_bwd_f32(const float* sgn, const float* g, int64_t n, float* da, void* stream);
/* The same backward WITHOUT a stored sign tensor (ms_l1_fwd_f32 with sgn = NULL): da = sign(a - b) * g[0] / n recomputed from
 * the operands (b NULL: constant c) -- 12 B of traffic per element instead of 16 B and no (n,) fp32 temporary. */
int ms_l1_bwd_ab_f32(const float* a, const float* b, float c, const float* g, int64_t n, float* da, void* stream);
/* out[0] = (float)(scale * in[0]) : turns a double accumulator into a loss scalar. */
int ms_scalar_finish(const double* in, double scale, float* out, void* stream);
/* The loss bookkeeping of one train step (trainer.py:1268-1285: `loss = sum(internal_losses)` after gan.py:113-135 scaled the
 * terms by lambda_D / lambda_gan and jlcss.py:197-205 by lambda_id) in ONE launch instead of ~20 scalar casts, multiplies and
 * adds: losses = HOST array of n (<= MS_LOSS_MAX_TERMS) device pointers to fp32 scalars, weights / lam_idx = host arrays;
 *   w_i = weights[i] * (lam_idx[i] >= 0 ? lam_dev[lam_idx[i]] : 1);  report[i] = w_i * l_i (fp64, nullable);  total = sum_i.
 * The host arrays are read at call time (by-value kernel parameters: safe under CUDA-graph capture); lam_dev is read by the
 * kernel, so a replayed graph follows a lambda schedule.  ms_loss_combine_bwd: g[i] = w_i * gtotal[0]. */
#define MS_LOSS_MAX_TERMS 8
int ms_loss_combine(const float* const* losses, const double* weights, const int* lam_idx, int n, const double* lam_dev,
                    float* total, double* report, void* stream);
int ms_loss_combine_bwd(const float* gtotal, const double* weights, const int* lam_idx, int n, const double* lam_dev, float* g,
                        void* stream);

/* ---- fused clip_grad_norm_(params, max_norm) + Adam.step() over flat buffers (trainer.py:1138-1146, :262-287).
 * All parameters of one sub-network live in ONE contiguous buffer of dtype dt (MS_F32 / MS_F64), likewise
 * gradients and the two Adam moments. */
/* acc[0] = sum g[i]^2 (zeroed internally); step (nullable, device int64) is incremented by one. */
int ms_grad_sqnorm(const void* g, int dt, int64_t n, double* acc, int64_t* step, void* stream);
/* coef = min(1, max_norm / (sqrt(sqnorm[0]) + 1e-6)) (max_norm <= 0: no clipping); t = step[0] (already advanced);
 * m = b1 m + (1-b1) g c; v = b2 v + (1-b2) (g c)^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps). */
int ms_clip_adam(void* p, const void* g, void* m, void* v, int dt, int64_t n, const double* sqnorm, const int64_t* step,
                 double lr, double beta1, double beta2, double eps, double max_norm, const double* lr_dev, void* stream);
/* lr_dev (nullable, device double): overrides lr, so a captured CUDA graph follows a learning-rate schedule. */
/* The same update, 4 elements per thread, with the moments m / v stored as state_dt (MS_F32 or dt) while p / g are dt:
 * fp64 master parameters with fp32 moments move 40 instead of 56 bytes per element.  Arithmetic in fp64 as above.
 * All four buffers 32-byte aligned.  (torch.optim.Adam keeps its state in the parameter dtype, trainer.py:262-287: with
 * state_dt == dt this is that update.) */
int ms_clip_adam_mixed(void* p, const void* g, void* m, void* v, int dt, int state_dt, int64_t n, const double* sqnorm,
                       const int64_t* step, double lr, double beta1, double beta2, double eps, double max_norm,
                       const double* lr_dev, void* stream);

/* ---- the step before the hot path: pose preprocessing on the device, fp64 like the reference.
 * Replaces, per batch, trainer.py:1290-1308 = KMeans.predict(RemoveJoints(pose)) (src/data/transform.py:352-410, 463-510)
 * and RemoveJoints(ZNorm(pose)) (src/data/transform.py:221-226) with one pass over the raw pose batch.
 *   x (B,T,Pr) raw pose; cols[P] kept columns (RemoveJoints as a gather); mean/var (Pr) ZNorm statistics;
 *   centers (K,D) k-means centres; feats_host[nfeats] in {1 pose, 2 velocity, 3 speed, 4 acceleration}, D = sum of widths;
 *   y (B,T,P) normalised pose (nullable); labels (B,T) int64 argmin (nullable); soft (B,T,K) soft labels (nullable). */
int ms_pose_prepare(const double* x, const double* mean, const double* var, const int32_t* cols, const double* centers,
                    int B, int T, int Pr, int P, int K, const int32_t* feats_host, int nfeats, double eps, double* y,
                    int64_t* labels, double* soft, void* stream);
/* ZNorm.inv_znorm (src/data/transform.py:228-229): out = x * sqrt(var) + mean over the last dimension C. */
int ms_inv_znorm(const double* x, const double* mean, const double* var, int64_t rows, int C, double* out, void* stream);

/* Evaluation metrics of TrainerBase.calculate_metrics (src/model/trainer.py:865-907) in one pass on the device:
 * L1 / VelL1 (src/evaluation/metrics.py:94-131) on the normalised full-width poses y, gt (B,T,2J) and PCK
 * (metrics.py:247-303) on the un-normalised (mean/var, 2J), root-centred frames.  keep[J]: 1 for joints outside the mask.
 * acc[2] = {sum |y-gt|, sum |vel(y)-vel(gt)|} over kept joints; cnt[nalpha*J] = frames with dist_j < alpha * max(h, w). */
int ms_pose_metrics(const double* y, const double* gt, const double* mean, const double* var, const uint8_t* keep,
                    int B, int T, int J, const double* alphas_host, int nalpha, double* acc, uint64_t* cnt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXSTAGE_B200_H */
