"""ctypes binding of libsedt_b200.so (include/sedt_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing or does not export the ABI this module raises, and every product
entry point above it fails with it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsedt_b200.so")
ABI_VERSION = 1

F32, BF16 = 0, 1
KERNEL_CLASSES = ("gemm_tcgen05", "gemm_cuda_core", "stem", "attention", "norm", "matcher", "other")


class SedtConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "enc_layers", "dec_layers", "num_queries", "num_classes", "hidden_dim", "nheads", "dim_feedforward",
        "dec_at", "pre_norm", "dilation", "self_sup", "feature_recon", "num_patches", "aux_loss",
        "precision", "use_tensor_cores")]


class SedtOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("hs", "logits", "boxes", "at", "memory", "pred_feature", "gt_feature", "feat")]


class SedtConvDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("in_", "w", "scale", "bias", "residual", "out")] + \
               [(n, C.c_int32) for n in ("in_dtype", "out_dtype", "B", "H", "W", "Cin", "lda", "Ho", "Wo", "Cout", "ldc",
                                         "ld_res", "R", "S", "stride", "dil", "pad", "relu")]


class SedtOptimTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("group", C.c_int32), ("reserved", C.c_int32)]


class SedtAugmentParams(C.Structure):
    _fields_ = [("tm_t0", C.c_int32), ("tm_t", C.c_int32), ("fm_f0", C.c_int32), ("fm_f", C.c_int32), ("fm_mode", C.c_int32),
                ("fm_const", C.c_float), ("fs_shift", C.c_int32), ("reserved", C.c_int32)]


class SedtMixRow(C.Structure):
    _fields_ = [("i1", C.c_int32), ("i2", C.c_int32), ("a", C.c_float), ("b", C.c_float)]


class SedtAdamWGroup(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("decay", "w1", "beta2", "w2", "bc2_sqrt", "eps", "neg_step", "reserved")]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/sedt_b200.h declares
SIGNATURES = {
    "sedt_last_error": (C.c_char_p, []),
    "sedt_abi_version": (_i, []),
    "sedt_launch_count": (C.c_ulonglong, []),
    "sedt_kernel_kinds": (_i, []),
    "sedt_kernel_kind_name": (C.c_char_p, [_i]),
    "sedt_kernel_kind_count": (C.c_ulonglong, [_i]),
    "sedt_profile_enable": (_i, [_i]),
    "sedt_profile_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "sedt_model_create": (_i, [C.POINTER(SedtConfig), C.POINTER(_vp)]),
    "sedt_model_destroy": (None, [_vp]),
    "sedt_model_num_weights": (_i, [_vp]),
    "sedt_model_weight_name": (C.c_char_p, [_vp, _i]),
    "sedt_model_weight_numel": (_i64, [_vp, _i]),
    "sedt_model_packed_bytes": (_i64, [_vp]),
    "sedt_model_pack": (_i, [_vp, C.POINTER(_vp), _vp, _i64, _vp]),
    "sedt_feature_shape": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "sedt_workspace_bytes": (_i64, [_vp, _i, _i, _i, _i, _i]),
    "sedt_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i64, C.POINTER(SedtOutputs), _vp]),
    "sedt_model_set_bucket_event": (_i, [_vp, _vp]),
    "sedt_train_tape_bytes": (_i64, [_vp, _i, _i, _i, _i]),
    "sedt_backward_workspace_bytes": (_i64, [_vp, _i, _i, _i]),
    "sedt_grad_numel": (_i64, [_vp]),
    "sedt_grad_offset": (_i64, [_vp, _i]),
    "sedt_forward_train": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i64, C.POINTER(SedtOutputs), _f, C.c_uint64, _vp]),
    "sedt_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i, _f, _vp]),
    "sedt_train_tape_bytes_sp": (_i64, [_vp, _i, _i, _i, _i, _i, _i]),
    "sedt_backward_workspace_bytes_sp": (_i64, [_vp, _i, _i, _i, _i, _i]),
    "sedt_forward_train_sp": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i64, C.POINTER(SedtOutputs), _f, C.c_uint64, _vp]),
    "sedt_backward_sp": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _f, _vp]),
    "sedt_matcher": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sedt_matcher_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _f, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sedt_lsap": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sedt_set_criterion": (_i, [_vp] * 9 + [_i] * 7 + [_f] * 5 + [_vp] * 10),
    "sedt_decode_events": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _f, _f] + [_vp] * 9),
    "sedt_pseudo_labels": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "sedt_prepare_clips": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sedt_augment_clips": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "sedt_mix_rows": (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    "sedt_query_patches": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sedt_optim_chunk_elems": (_i, []),
    "sedt_grad_norm": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sedt_clip_grads": (_i, [_vp, _vp, _i, _vp, _f, _vp]),
    "sedt_adamw_step": (_i, [_vp, _vp, _i, C.POINTER(SedtAdamWGroup), _i, _vp, _f, _vp]),
    "sedt_op_conv": (_i, [C.POINTER(SedtConvDesc), _i, _vp]),
    "sedt_op_conv_tc_supported": (_i, [C.POINTER(SedtConvDesc)]),
    "sedt_op_ffn": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "sedt_op_bneck_tail": (_i, [_vp] * 7 + [_i, _i, _i, _vp]),
    "sedt_op_enc_attn": (_i, [_vp] * 8 + [_i, _i, _vp, _vp, _vp, _vp]),
    "sedt_op_repack_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sedt_op_upsample2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sedt_op_relu_mask": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "sedt_op_colsum": (_i, [_vp, _i, _i64, _vp, _i64, _i, _vp]),
    "sedt_op_layernorm_bwd": (_i, [_vp] * 9 + [_i64, _vp]),
    "sedt_op_attention_bwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "sedt_op_dropout_mask": (_i, [_vp, _i64, C.c_uint64, C.c_uint64, C.c_uint32, _f, _vp]),
    "sedt_op_conv_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "sedt_op_repack_conv": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sedt_op_cast": (_i, [_vp, _vp, _i, _i64, _vp]),
    "sedt_op_stem": (_i, [_vp] * 10 + [_i, _i, _i, _i, _vp]),
    "sedt_op_stem_tc": (_i, [_vp] * 10 + [_i, _i, _i, _vp]),
    "sedt_op_layernorm": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _i64, _vp]),
    "sedt_op_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "sedt_op_pos_table": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


class SedtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sedt_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Load the library (once).  Raises if it is missing: build it with
    `python -m sound_event_detection_transformer_b200.build` or __graft_entry__.build()."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension has not been built "
                "(run `python -m sound_event_detection_transformer_b200.build`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.sedt_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libsedt_b200.so ABI {lib.sedt_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def kernel_kind_counts() -> dict:
    """Eager launches so far per size-dependent kernel kind (graph replays are not counted)."""
    lib = load()
    return {lib.sedt_kernel_kind_name(i).decode(): int(lib.sedt_kernel_kind_count(i)) for i in range(lib.sedt_kernel_kinds())}


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sedt_last_error()
        raise SedtError(rc, msg.decode() if msg else "")


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
