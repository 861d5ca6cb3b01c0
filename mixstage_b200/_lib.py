"""ctypes binding of libmixstage_b200.so (include/mixstage_b200.h).

There is NO CPU fallback: if the shared library is missing or an entry point fails,
a RuntimeError is raised."""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmixstage_b200.so")

MS_F32, MS_F64, MS_BF16, MS_BF16X2 = 0, 1, 2, 3
LOSS_MAX_TERMS = 8          # MS_LOSS_MAX_TERMS


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "H", "W", "Cin", "Cout", "kh", "kw", "sh", "sw", "ph", "pw", "groups", "Ho", "Wo")]


MAX_CLASSES, MAX_TAPS = 16, 48


class IgemmDesc(ctypes.Structure):
    """ms_igemm_desc (include/mixstage_b200.h)."""
    _fields_ = [
        ("a_dims", ctypes.c_int32 * 5), ("a_strides", ctypes.c_int64 * 5), ("box", ctypes.c_int32 * 5),
        ("out_dims", ctypes.c_int32 * 3), ("out_strides", ctypes.c_int64 * 3),
        ("num_classes", ctypes.c_int32), ("class_n", ctypes.c_int32), ("block_n", ctypes.c_int32),
        ("ntaps", ctypes.c_int32), ("cchunks", ctypes.c_int32), ("shared_taps", ctypes.c_int32),
        ("a_chan_base", ctypes.c_int32 * MAX_CLASSES), ("out_off", ctypes.c_int64 * MAX_CLASSES),
        ("taps", (ctypes.c_int16 * 4) * MAX_TAPS),
        ("out_dtype", ctypes.c_int32), ("epilogue", ctypes.c_int32), ("slope", ctypes.c_float),
        ("planes", ctypes.c_int32),
        ("a_plane_stride", ctypes.c_int64), ("w_plane_stride", ctypes.c_int64), ("out_plane_stride", ctypes.c_int64),
        ("split_k", ctypes.c_int32), ("out_numel", ctypes.c_int64), ("wgrad_c_tile", ctypes.c_int32),
    ]


class PackEntry(ctypes.Structure):
    """ms_pack_entry (include/mixstage_b200.h)."""
    _fields_ = [("w", ctypes.c_void_p), ("wp", ctypes.c_void_p), ("wp_lo", ctypes.c_void_p)] + [
        (n, ctypes.c_int32) for n in ("pdt", "Cout", "Cin_g", "taps_total", "groups", "mode", "num_classes", "class_n",
                                      "ntaps", "kpad")] + [("srctap", ctypes.c_int16 * MAX_TAPS)]


class BlockBn(ctypes.Structure):
    """ms_block_bn (include/mixstage_b200.h): the BatchNorm + activation side of a fused block."""
    _fields_ = [("C", ctypes.c_int32), ("pdt", ctypes.c_int32), ("training", ctypes.c_int32),
                ("momentum", ctypes.c_float), ("eps", ctypes.c_float), ("slope", ctypes.c_float),
                ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("conv_bias", ctypes.c_void_p),
                ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p), ("num_batches_tracked", ctypes.c_void_p),
                ("sums", ctypes.c_void_p), ("ss", ctypes.c_void_p)]


CHAIN_MAX = 16


class ChainFwdLayer(ctypes.Structure):
    """ms_chain_fwd_layer (include/mixstage_b200.h)."""
    _fields_ = [("d", ctypes.POINTER(IgemmDesc)), ("a", ctypes.c_void_p), ("w", ctypes.c_void_p), ("z", ctypes.c_void_p),
                ("bn", ctypes.POINTER(BlockBn)), ("y", ctypes.c_void_p), ("planes", ctypes.c_void_p), ("pfmt", ctypes.c_int32),
                ("pstride", ctypes.c_int64), ("res", ctypes.c_void_p), ("res_planes", ctypes.c_void_p),
                ("res_pfmt", ctypes.c_int32), ("res_pstride", ctypes.c_int64), ("up2", ctypes.c_int32)]


class ChainBwdLayer(ctypes.Structure):
    """ms_chain_bwd_layer (include/mixstage_b200.h)."""
    _fields_ = [("dg", ctypes.POINTER(IgemmDesc)), ("dy", ctypes.c_void_p), ("dy2", ctypes.c_void_p), ("z", ctypes.c_void_p),
                ("bn", ctypes.POINTER(BlockBn)), ("rows", ctypes.c_int64), ("up2", ctypes.c_int32),
                ("rows_per_seq", ctypes.c_int32), ("dz_planes", ctypes.c_void_p), ("pfmt", ctypes.c_int32),
                ("pstride", ctypes.c_int64), ("grad_gamma", ctypes.c_void_p), ("grad_beta", ctypes.c_void_p),
                ("gdt", ctypes.c_int32), ("wt", ctypes.c_void_p), ("dx", ctypes.c_void_p)]


class WgradItem(ctypes.Structure):
    """ms_wgrad_item (include/mixstage_b200.h)."""
    _fields_ = [("d", ctypes.POINTER(IgemmDesc)), ("x", ctypes.c_void_p), ("dz", ctypes.c_void_p), ("acc", ctypes.c_void_p)]


class WgradEntry(ctypes.Structure):
    """ms_wgrad_entry (include/mixstage_b200.h)."""
    _fields_ = [("acc", ctypes.c_void_p), ("dw", ctypes.c_void_p)] + [
        (n, ctypes.c_int32) for n in ("pdt", "Cout", "Cin_g", "taps", "kpad", "accumulate")]


_P, _I, _L, _F, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double
_BN = ctypes.POINTER(BlockBn)
_CD = ctypes.POINTER(ConvDesc)
_GD = ctypes.POINTER(IgemmDesc)
_S16 = ctypes.POINTER(ctypes.c_int16)
_S32 = ctypes.POINTER(ctypes.c_int32)

# name -> argtypes (mirrors include/mixstage_b200.h; tests/test_capi_symbols.py checks both ways)
PROTOTYPES = {
    "ms_version": [],
    "ms_device_is_sm100": [],
    "ms_pack_conv_weight_f32": [_P, _I, _CD, _P, _P, _P],
    "ms_unpack_conv_wgrad": [_P, _CD, _P, _I, _I, _P],
    "ms_cast": [_P, _I, _P, _I, _L, _P],
    "ms_scale_cast": [_P, _I, _P, _I, _L, _D, _P],
    "ms_conv_fwd_f32": [_P, _P, _P, _P, _CD, _I, _F, _P],
    "ms_conv_dgrad_f32": [_P, _P, _P, _CD, _P],
    "ms_conv_wgrad_f32": [_P, _P, _P, _CD, _P],
    "ms_conv_cin1_bnact": [_P, _P, _P, _P, _F, _CD, _P, _P, _I, _L, _P],
    "ms_igemm_bf16": [_GD, _P, _P, _P, _P, _P, _P, _P],
    "ms_igemm_bf16_fused": [_GD, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _P],
    "ms_igemm_bf16_mix": [_GD, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ms_pack_igemm_weight_bf16": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _S16, _P, _P, _P],
    "ms_pack_igemm_weight_multi": [_P, _I, _I, _P],
    "ms_wgrad_bf16": [_GD, _P, _P, _P, _P],
    "ms_conv_block_train_fwd": [_GD, _P, _P, _P, _BN, _P, _P, _I, _L, _P, _P, _I, _L, _I, _P, _P],
    "ms_conv_block_train_bwd": [_GD, _P, _P, _BN, _L, _I, _I, _P, _I, _L, _P, _P, _I, _P, _P, _P, _P],
    "ms_set_chain_sm_budget": [_I],
    "ms_debug_phase_ts": [_P],
    "ms_debug_trap_info": [_P],
    "ms_conv_chain_fwd": [_P, _I, _P, _P],
    "ms_conv_chain_bwd": [_P, _I, _P, _P],
    "ms_wgrad_bf16_acc_multi": [_P, _I, _P],
    "ms_wgrad_bf16_acc": [_GD, _P, _P, _P, _P],
    "ms_unpack_wgrad_multi": [_P, _I, _I, _P],
    "ms_unpack_igemm_wgrad": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    "ms_col_stats_f32": [_P, _L, _I, _P, _P, _P],
    "ms_bn_finalize": [_P, _P, _L, _I, _P, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P],
    "ms_bn_act_fwd_f32": [_P, _P, _P, _F, _L, _I, _P, _P, _I, _I, _P, _I, _L, _P],
    "ms_to_planes": [_P, _L, _I, _I, _P, _I, _L, _P],
    "ms_planes_to_f32": [_P, _I, _L, _L, _I, _I, _P, _P],
    "ms_bn_act_bwd_reduce_f32": [_P, _P, _P, _P, _P, _P, _F, _L, _I, _I, _I, _P, _P, _P],
    "ms_bn_act_bwd_apply_f32": [_P, _P, _P, _P, _P, _P, _F, _L, _I, _I, _I, _P, _P, _I, _P, _P, _I, _L, _P, _P, _I, _P],
    "ms_bn_stats_finalize": [_P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _P, _P, _P, _P, _P],
    "ms_lrelu_bwd_f32": [_P, _P, _F, _L, _P, _P, _I, _L, _P],
    "ms_store_param_grad": [_P, _I, _P, _I, _I, _P],
    "ms_bilinear_to_T_fwd_f32": [_P, _I, _I, _I, _I, _I, _P, _P],
    "ms_bilinear_to_T_bwd_f32": [_P, _I, _I, _I, _I, _I, _P, _P],
    "ms_style_concat_fwd_f32": [_P, _L, _I, _P, _P, _I, _P, _I, _I, _I, _P, _P],
    "ms_style_concat_planes_fwd_f32": [_P, _L, _I, _P, _P, _I, _P, _I, _I, _I, _P, _P, _I, _L, _I, _P],
    "ms_style_concat_bwd_f32": [_P, _L, _I, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P],
    "ms_softmax_ce_fwd_f32": [_P, _L, _I, _P, _I, _P, _P, _P, _P],
    "ms_softmax_ce_bwd_f32": [_P, _L, _I, _P, _I, _P, _P, _P, _P],
    "ms_mixture_fwd_f32": [_P, _P, _L, _I, _I, _P, _P],
    "ms_mixture_bwd_f32": [_P, _P, _P, _L, _I, _I, _P, _P, _P],
    "ms_mean_rows_fwd_f32": [_P, _I, _I, _I, _P, _P],
    "ms_mean_rows_bwd_f32": [_P, _I, _I, _I, _P, _P],
    "ms_velocity_fwd_f32": [_P, _I, _I, _I, _P, _P],
    "ms_velocity_bwd_f32": [_P, _I, _I, _I, _P, _P],
    "ms_l1_fwd_f32": [_P, _P, _F, _L, _P, _P, _P],
    "ms_l1_bwd_f32": [_P, _P, _L, _P, _P],
    "ms_l1_bwd_ab_f32": [_P, _P, _F, _P, _L, _P, _P],
    "ms_scalar_finish": [_P, _D, _P, _P],
    "ms_loss_combine": [_P, _P, _P, _I, _P, _P, _P, _P],
    "ms_loss_combine_bwd": [_P, _P, _P, _I, _P, _P, _P],
    "ms_grad_sqnorm": [_P, _I, _L, _P, _P, _P],
    "ms_clip_adam": [_P, _P, _P, _P, _I, _L, _P, _P, _D, _D, _D, _D, _D, _P, _P],
    "ms_clip_adam_mixed": [_P, _P, _P, _P, _I, _I, _L, _P, _P, _D, _D, _D, _D, _D, _P, _P],
    "ms_pose_prepare": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _S32, _I, _D, _P, _P, _P, _P],
    "ms_inv_znorm": [_P, _P, _P, _L, _I, _P, _P],
    "ms_pose_metrics": [_P, _P, _P, _P, _P, _I, _I, _I, ctypes.POINTER(ctypes.c_double), _I, _P, _P, _P],
}

_LIB = None
LAUNCHES = 0          # number of C-ABI kernel entry points called (bench.py's gpu_launches claim)


class MixStageError(RuntimeError):
    pass


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MixStageError(
                "libmixstage_b200.so is not built (%s). Run `python -m mixstage_b200.build`; "
                "mixstage_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def call(name, *args):
    """Invoke an entry point; raises on a non-zero status (cudaError_t or MS_E*)."""
    global LAUNCHES
    rc = getattr(load(), name)(*args)
    LAUNCHES += 1
    if rc != 0:
        raise MixStageError("%s failed with status %d" % (name, rc))


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dt_code(dtype):
    if dtype == torch.float32:
        return MS_F32
    if dtype == torch.float64:
        return MS_F64
    if dtype == torch.bfloat16:
        return MS_BF16
    raise MixStageError("unsupported dtype %s" % dtype)
