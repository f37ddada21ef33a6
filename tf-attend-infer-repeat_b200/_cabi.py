"""ctypes binding of libair_b200.so (the C ABI declared in include/air_b200.h).

There is deliberately no CPU fallback: if the shared library is missing, or a compute
entry point is called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libair_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "air_b200.h")

_c_f = ctypes.c_void_p  # device pointers travel as void*
_SIGS = {
    "air_abi_version": (ctypes.c_int, []),
    "air_last_error": (ctypes.c_char_p, []),
    "air_launch_count": (ctypes.c_int64, []),
    "air_crc32c": (ctypes.c_uint32, [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32]),
    "air_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3),
    "air_st_forward": (ctypes.c_int, [_c_f, _c_f, _c_f, ctypes.c_int64] + [ctypes.c_int] * 5 + [_c_f]),
    "air_st_backward": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_int64] + [ctypes.c_int] * 5 + [_c_f]),
    "air_st_writeback_canvas_fwd": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_float] + [_c_f] * 2 + [ctypes.c_int64] +
                                    [ctypes.c_int] * 4 + [_c_f]),
    "air_st_forward_steps": (ctypes.c_int, [_c_f, _c_f, _c_f, ctypes.c_int64] + [ctypes.c_int] * 6 + [_c_f]),
    "air_st_backward_steps": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_int64] + [ctypes.c_int] * 6 + [_c_f]),
    "air_st_writeback_canvas_bwd_steps": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_int64, ctypes.c_float] + [_c_f] * 4 +
                                          [ctypes.c_int, ctypes.c_int64] + [ctypes.c_int] * 5 + [_c_f]),
    "air_st_writeback_canvas_fwd_steps": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_int64, ctypes.c_float, _c_f, _c_f, ctypes.c_int64] +
                                          [ctypes.c_int] * 5 + [_c_f]),
    "air_st_writeback_canvas_bwd": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_float] + [_c_f] * 4 + [ctypes.c_int, ctypes.c_int64] +
                                    [ctypes.c_int] * 4 + [_c_f]),
    "air_concrete_step_fwd": (ctypes.c_int, [_c_f] * 6 + [ctypes.c_float, ctypes.c_float, ctypes.c_int] + [_c_f] * 7 +
                              [ctypes.c_int64, _c_f]),
    "air_concrete_step_bwd": (ctypes.c_int, [_c_f] * 6 + [ctypes.c_float, ctypes.c_int, _c_f, ctypes.c_int64, _c_f]),
    "air_gemm": (ctypes.c_int, [_c_f] * 6 + [ctypes.c_int64] + [ctypes.c_int] * 9 + [_c_f]),
    "air_gemm_ex": (ctypes.c_int, [_c_f] * 6 + [ctypes.c_int64] + [ctypes.c_int] * 8 + [ctypes.c_float, ctypes.c_int, _c_f]),
    "air_gemm_ws": (ctypes.c_int, [_c_f] * 6 + [ctypes.c_int64] + [ctypes.c_int] * 8 +
                    [ctypes.c_float, ctypes.c_int, _c_f, ctypes.c_int64, _c_f]),
    "air_lstm_fwd": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_lstm_bwd": (ctypes.c_int, [_c_f] * 8 + [ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_heads_fwd": (ctypes.c_int, [_c_f] * 14 + [ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_heads_bwd_workspace": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int]),
    "air_heads_bwd": (ctypes.c_int, [_c_f] * 10 + [ctypes.c_float] + [_c_f] * 3 + [ctypes.c_int, _c_f, ctypes.c_int64,
                                                                                 ctypes.c_int, _c_f]),
    "air_vae_latent_fwd": (ctypes.c_int, [_c_f] * 4 + [ctypes.c_int] + [_c_f] * 2 + [ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_vae_latent_bwd": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_float, _c_f, ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_sigmoid_noise_fwd": (ctypes.c_int, [_c_f, _c_f, ctypes.c_float, _c_f, ctypes.c_int64, _c_f]),
    "air_sigmoid_bwd": (ctypes.c_int, [_c_f, _c_f, _c_f, ctypes.c_int64, _c_f]),
    "air_bce_loss": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_float, ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_finalize_loss": (ctypes.c_int, [_c_f] * 6 + [ctypes.c_int64, _c_f]),
    "air_colsum_workspace": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int]),
    "air_colsum": (ctypes.c_int, [_c_f, ctypes.c_int, _c_f, ctypes.c_int, _c_f, ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_adam_workspace": (ctypes.c_int64, [ctypes.c_int64]),
    "air_adam_step": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_float] * 5 + [_c_f, ctypes.c_int64, _c_f]),
    "air_adam_step_ex": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_float] * 5 + [_c_f, ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_anneal": (ctypes.c_int, [_c_f] + [ctypes.c_float] * 3 + [ctypes.c_int] + [ctypes.c_float] * 2 +
                   [ctypes.c_int, _c_f, _c_f]),
    "air_colsum_multi_workspace": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    "air_colsum_multi": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_f, _c_f]),
    "air_reduce_rows": (ctypes.c_int, [_c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_f, ctypes.c_int, _c_f]),
    "air_synth_canvases": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, _c_f, _c_f, ctypes.c_int64, ctypes.c_int,
                                          ctypes.c_int, _c_f]),
    "air_synth_canvases_ex": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, _c_f, _c_f, _c_f, _c_f, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_int, _c_f]),
    "air_noise_fill": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_int64, ctypes.c_int, _c_f]),
    "air_rng_normals": (ctypes.c_int, [_c_f, ctypes.c_int, _c_f, ctypes.c_int64, _c_f]),
    "air_rng_uniforms": (ctypes.c_int, [_c_f, ctypes.c_int, _c_f, ctypes.c_int64, _c_f]),
    "air_zero_buffers": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _c_f]),
    "air_expand_u8": (ctypes.c_int, [_c_f, _c_f, ctypes.c_int64, _c_f]),
    "air_tfrecord_index": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "air_shuffle_order": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]),
    "air_gather_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p,
                                       ctypes.c_int]),
    "air_conv5x5_fwd": (ctypes.c_int, [_c_f] * 5 + [ctypes.c_int64] + [ctypes.c_int] * 5 + [_c_f]),
    "air_conv5x5_bwd_workspace": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "air_conv5x5_bwd": (ctypes.c_int, [_c_f] * 8 + [ctypes.c_int, _c_f, ctypes.c_int64] + [ctypes.c_int] * 5 + [_c_f]),
}


class Hyper(ctypes.Structure):
    """air_hyper_t (include/air_b200.h)."""
    _fields_ = [(n, ctypes.c_float) for n in (
        "scale_prior_mean", "scale_prior_variance", "shift_prior_mean", "shift_prior_variance", "vae_prior_mean",
        "vae_prior_variance", "vae_likelihood_std", "z_pres_temperature", "stopping_threshold")] + [("train", ctypes.c_int32)]


# enum air_field
(F_SCALE_MEAN, F_SCALE_LV, F_SHIFT_MEAN_X, F_SHIFT_MEAN_Y, F_SHIFT_LV_X, F_SHIFT_LV_Y, F_LOG_ODDS, F_S, F_X, F_Y,
 F_YPRE, F_Z, F_ZPROB, F_KL_Z, F_KL_SCALE, F_KL_SHIFT, F_KL_VAE, F_STOP_PREV, F_STOP_NEW) = range(19)
NF = 20
EPI_NONE, EPI_RELU, EPI_SOFTPLUS, EPI_MUL_DRELU, EPI_MUL_DSOFTPLUS, EPI_SIGMOID_NOISE, EPI_SIGMOID_RNG = range(7)
RNG_SCALE, RNG_SHIFT, RNG_LATENT, RNG_CONCRETE, RNG_LIKE = 1, 2, 3, 4, 5
GEMM_MODES = {"fp32": 0, "tf32": 1, "tf32x3": 2}

_lib = None


class ColsumItem(ctypes.Structure):
    """air_colsum_item_t of include/air_b200.h."""
    _fields_ = [("X", ctypes.c_void_p), ("out", ctypes.c_void_p), ("rows", ctypes.c_int64), ("ld", ctypes.c_int),
                ("N", ctypes.c_int), ("accumulate", ctypes.c_int)]


class AirError(RuntimeError):
    """Raised when a C-ABI entry point returns a negative AIR_ERR_* code."""


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libair_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libair_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AirError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(this package has no CPU / PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def declared_symbols():
    return list(_SIGS.keys())


_DEBUG_CAPTURE = os.environ.get("AIR_DEBUG_CAPTURE", "0") != "0"


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().air_last_error().decode("utf-8", "replace")
        raise AirError(f"{what} failed with code {code}: {msg}")
    if _DEBUG_CAPTURE:   # diagnostics: name the first call after which the stream's capture has been invalidated
        import torch
        try:
            torch.cuda.is_current_stream_capturing()
        except Exception as e:
            raise AirError(f"stream capture is invalid after {what}: {e}") from None


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AirError("air_b200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise AirError("air_b200 ops need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def launch_count() -> int:
    return int(lib().air_launch_count())
