"""ctypes binding of libpcd_b200.so (the C ABI declared in include/pcd_b200.h).

There is no CPU fallback: importing the package works anywhere, but every compute
entry point raises if the shared library or a CUDA device is missing.
"""
import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# PCD_B200_LIB: A/B timing of two builds of the library in one process tree (tools/ab.sh); never a fallback
LIB_PATH = os.environ.get("PCD_B200_LIB") or os.path.join(HERE, "libpcd_b200.so")

ABI_VERSION = 4
PCD_F32, PCD_BF16 = 0, 1
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2
EPI_RESIDUAL_STATS, EPI_LN_BIAS, EPI_LN_BIAS_GELU = 3, 4, 5
# pcd_attn_variant (tensor-core attention kernels, per call) and pcd_model_desc.flags
ATTN_DEFAULT, ATTN_GROUPED, ATTN_GROUPED_TOKEN, ATTN_GROUPED_STEPTAIL, ATTN_GROUPED_WIDE = 0, 8, 9, 10, 11
ATTN_PAIRED, ATTN_PAIRED_POLY4, ATTN_PAIRED_POLY2 = 5, 6, 7
MODEL_SEPARATE_LAYERNORM, MODEL_ATTN_VARIANT_SHIFT = 1, 8

c_float_p = C.POINTER(C.c_float)
vp = C.c_void_p


class AttnOperand(C.Structure):
    _fields_ = [("ptr", vp), ("batch_stride", C.c_int64), ("row_stride", C.c_int64),
                ("head_stride", C.c_int64)]


class StepScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("c_in", "coef_x", "coef_eps", "sigma", "dt", "guidance",
                                          "clip", "next_c_in", "next_noise", "dt2", "mode")]


class BlockWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_g", "ln1_b", "ln2_g", "ln2_b", "w_qkv", "w_proj", "w_fc",
                                  "w_fc2", "b_qkv", "b_proj", "b_fc", "b_fc2",
                                  "w_qkv_ln", "w_fc_ln", "qkv_colsum", "qkv_const", "fc_colsum", "fc_const")]


class GemmArgs(C.Structure):
    _fields_ = [("A", vp), ("lda", C.c_int), ("W", vp), ("ldw", C.c_int), ("bias", vp), ("residual", vp),
                ("ldr", C.c_int), ("C", vp), ("ldc", C.c_int), ("out_precision", C.c_int), ("C2", vp),
                ("ldc2", C.c_int), ("stats_out", vp), ("stats_in", vp), ("colsum", vp), ("ln_eps", C.c_float),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("epilogue", C.c_int), ("debug", C.c_int)]


class DdpmArgs(C.Structure):  # include/pcd_b200.h: pcd_ddpm_args
    _fields_ = ([(n, vp) for n in ("x", "model_out", "noise", "t", "table", "ch_scale", "ch_bias", "x_next",
                                   "pred_xstart", "sample_unscaled", "mean", "log_variance")]
                + [(n, C.c_int) for n in ("batch", "channels", "n_points", "out_channels", "var_mode",
                                          "clip_denoised", "unscale", "num_timesteps")])


DDPM_COLS = 8
DDPM_MEAN_X0, DDPM_MEAN_XT = 2, 3   # posterior_mean_coef1 / coef2 columns (PCD_DDPM_MEAN_X0 / _XT)
VAR_FIXED, VAR_LEARNED_RANGE, VAR_LEARNED = 0, 1, 2


class ModelDesc(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("precision", "width", "heads", "layers", "c_in", "c_out",
                                         "n_points", "n_prefix", "time_slot")]
                + [("ln_eps", C.c_float), ("flags", C.c_int)]
                + [(n, vp) for n in ("time_fc_w", "time_fc_b", "time_proj_w", "time_proj_b", "freqs",
                                     "ln_pre_g", "ln_pre_b", "ln_post_g", "ln_post_b", "in_w", "in_b",
                                     "out_w", "out_b")]
                + [("blocks", C.POINTER(BlockWeights))])


_SIGS = {
    "pcd_abi_version": (C.c_int, []),
    "pcd_last_error": (C.c_char_p, []),
    "pcd_launch_count": (C.c_ulonglong, []),
    "pcd_check_device": (C.c_int, []),
    "pcd_timestep_embed": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]),
    "pcd_layernorm": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp]),
    "pcd_embed_tokens": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp,
                                   C.c_float, vp, C.c_int, C.c_int, vp, vp, vp]),
    "pcd_add_layernorm": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_float, vp]),
    "pcd_output_proj": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_float, vp, vp,
                                  C.c_int, vp, vp]),
    "pcd_gemm_f32": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, vp]),
    "pcd_gemm_bf16": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, vp]),
    "pcd_gemm_bf16_ex": (C.c_int, [C.POINTER(GemmArgs), vp]),
    "pcd_add_f32": (C.c_int, [vp, vp, vp, C.c_int64, vp]),
    "pcd_cast_rowstats": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp]),
    "pcd_attention": (C.c_int, [C.POINTER(AttnOperand), C.POINTER(AttnOperand), C.POINTER(AttnOperand), vp,
                                C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_float, vp, C.c_int, C.c_int, vp]),
    "pcd_attention_hd32": (C.c_int, [C.POINTER(AttnOperand), C.POINTER(AttnOperand), C.POINTER(AttnOperand), vp,
                                     C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp]),
    "pcd_attention_hd32_bf16": (C.c_int, [C.POINTER(AttnOperand), C.POINTER(AttnOperand), C.POINTER(AttnOperand), vp,
                                          C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                          C.c_int, vp]),
    "pcd_rope_bf16": (C.c_int, [C.POINTER(AttnOperand), vp, C.c_int, C.c_int, C.c_int, vp]),
    "pcd_farthest_point_sample": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "pcd_nearest_points": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "pcd_fscore": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp, vp]),
    "pcd_sampler_begin": (C.c_int, [vp, vp, vp, C.POINTER(StepScalars), C.c_int64, vp]),
    "pcd_sampler_predictor": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                        C.POINTER(StepScalars), C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "pcd_sampler_corrector": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, C.POINTER(StepScalars),
                                        C.c_int, C.c_int, C.c_int, vp]),
    "pcd_ddpm_step": (C.c_int, [C.POINTER(DdpmArgs), vp]),
    "pcd_chamfer": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "pcd_model_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(vp)]),
    "pcd_model_destroy": (C.c_int, [vp]),
    "pcd_model_workspace_bytes": (C.c_size_t, [vp, C.c_int]),
    "pcd_model_forward": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_size_t, C.c_int, vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None
_lock = threading.Lock()


class PcdError(RuntimeError):
    pass


def load():
    """Load (building in-tree if needed) the shared library. Raises loudly on failure."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.environ.get("PCD_B200_LIB"):
            # rebuilds only when the sources' content hash differs from the one the .so was built from: a stale
            # library would mis-read the ctypes structs below
            from . import build as _build
            _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
            fn.restype, fn.argtypes = res, args
        if lib.pcd_abi_version() != ABI_VERSION:
            raise PcdError(f"{LIB_PATH}: ABI version {lib.pcd_abi_version()}, this binding expects {ABI_VERSION}")
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().pcd_last_error().decode(errors="replace")
        raise PcdError(f"{what or 'pcd call'} failed (status {rc}): {msg}")


def require_cuda(*tensors):
    """The product path is CUDA-only: fail loudly instead of falling back."""
    if not torch.cuda.is_available():
        raise PcdError("pcd_b200 requires a CUDA device (no CPU fallback exists)")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PcdError("pcd_b200 kernels take CUDA tensors only")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
