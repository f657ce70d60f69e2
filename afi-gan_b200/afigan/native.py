"""ctypes binding of libafigan_b200.so (include/afigan_b200.h).  PyTorch only supplies device memory and streams.

There is NO CPU or eager-PyTorch fallback: if the shared library is missing, or the device is not an sm_100
GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AFIGAN_LIB_PATH") or os.path.join(_HERE, "libafigan_b200.so")      # (override: A/B runs of kernel build variants)

PREC_FP32, PREC_BF16, PREC_BF16_SIMT, PREC_SPLIT = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16_simt": PREC_BF16_SIMT, "split": PREC_SPLIT}
MAX_RDB = 4
ABI_VERSION = 4


def default_precision() -> str:
    """AFIGAN_PRECISION = split (default: fp32 storage, every GEMM as a six-term bf16x3 split product on the tcgen05 tensor cores --
    meets the reference's fp32 results within 1e-3 on gradients / 1e-5 on losses) | bf16 (tcgen05 throughput mode: bf16 storage and
    operands, ~10 % gradient noise, EXPLICIT opt-in) | fp32 (CUDA-core FFMA cross-check) | bf16_simt (CUDA-core cross-check of bf16)."""
    p = os.environ.get("AFIGAN_PRECISION", "split").lower()
    if p not in PRECISIONS:
        raise ValueError(f"AFIGAN_PRECISION={p!r}; expected one of {sorted(PRECISIONS)}")
    return p


class View4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_longlong), ("sc", C.c_longlong), ("sh", C.c_longlong), ("sw", C.c_longlong),
                ("dtype", C.c_int), ("reserved", C.c_int)]


class GParams(C.Structure):
    _fields_ = [("n_rdb", C.c_int), ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("rdb_w", (C.c_void_p * 5) * MAX_RDB),
                ("post_w", C.c_void_p), ("post_b", C.c_void_p), ("up_w", C.c_void_p), ("up_b", C.c_void_p),
                ("out_w", C.c_void_p), ("out_b", C.c_void_p)]


GGrads = GParams  # same layout (non-const pointers)


class Lateral(C.Structure):
    _fields_ = [("lat_x", View4), ("lat_c", C.c_int), ("lat_w", C.c_void_p), ("lat_b", C.c_void_p), ("scale", C.c_float)]


class GCall(C.Structure):
    _fields_ = [("x", View4), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("y", C.c_void_p), ("oh", C.c_int), ("ow", C.c_int),
                ("dy", View4), ("dx", C.c_void_p), ("lateral", C.POINTER(Lateral)), ("lat_dx", C.c_void_p), ("lat_gw", C.c_void_p),
                ("lat_gb", C.c_void_p), ("ws", C.c_void_p), ("ws_bytes", C.c_size_t), ("fuse_cur", View4), ("fuse_w", C.c_void_p)]


class DCall(C.Structure):
    _fields_ = [("x", View4), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("logits", C.c_void_p), ("dlogits", C.c_void_p),
                ("dx", C.c_void_p), ("ws", C.c_void_p), ("ws_bytes", C.c_size_t), ("input_staged", C.c_int)]


MAX_CALLS = 10


class DParams(C.Structure):
    _fields_ = [("w", C.c_void_p * 4), ("b", C.c_void_p * 4), ("gamma", C.c_void_p * 3), ("beta", C.c_void_p * 3),
                ("running_mean", C.c_void_p * 3), ("running_var", C.c_void_p * 3), ("num_batches_tracked", C.c_void_p * 3)]


class DGrads(C.Structure):
    _fields_ = [("w", C.c_void_p * 4), ("b", C.c_void_p * 4), ("gamma", C.c_void_p * 3), ("beta", C.c_void_p * 3)]


_SIGNATURES = {
    "afi_abi_version": (C.c_int, []),
    "afi_last_error": (C.c_char_p, []),
    "afi_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "afi_destroy": (None, [C.c_void_p]),
    "afi_g_packed_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "afi_g_gradacc_bytes": (C.c_size_t, [C.c_int]),
    "afi_g_workspace_bytes": (C.c_size_t, [C.c_int] * 7),
    "afi_g_pack": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GParams), C.c_void_p, C.c_void_p]),
    "afi_g_forward": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GParams), C.c_void_p, C.POINTER(GCall), C.c_int, C.c_int, C.c_void_p]),
    "afi_g_backward": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GParams), C.c_void_p, C.POINTER(GCall), C.c_int, C.c_void_p, C.c_void_p]),
    "afi_g_unpack_grads": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(GGrads), C.c_float, C.c_int, C.c_void_p]),
    "afi_d_packed_bytes": (C.c_size_t, [C.c_int]),
    "afi_d_gradacc_bytes": (C.c_size_t, []),
    "afi_d_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "afi_d_pack": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DParams), C.c_void_p, C.c_void_p]),
    "afi_d_forward": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DParams), C.c_void_p, C.POINTER(DCall), C.c_int, C.c_int, C.c_float,
                                C.c_float, C.c_int, C.c_void_p]),
    "afi_d_update_running": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DParams), C.POINTER(DCall), C.c_int, C.c_float, C.c_void_p]),
    "afi_d_backward": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DParams), C.c_void_p, C.POINTER(DCall), C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p]),
    "afi_d_unpack_grads": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(DGrads), C.c_float, C.c_int, C.c_void_p]),
    "afi_bce_with_logits": (C.c_int, [C.c_void_p, C.c_longlong, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float,
                                      C.c_void_p]),
    "afi_l1_loss": (C.c_int, [View4, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float,
                              C.c_void_p]),
    "afi_sgd_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                               C.c_void_p]),
    "afi_sgd_step_multi": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                     C.c_int, C.c_void_p]),
    "afi_sizeof": (C.c_size_t, [C.c_int]),
    "afi_zero": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv3x3": (C.c_int, [C.c_void_p, C.c_int, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv3x3_backward": (C.c_int, [C.c_void_p, C.c_int, View4, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv3x3_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "afi_conv1x1": (C.c_int, [C.c_void_p, C.c_int, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv1x1_backward": (C.c_int, [C.c_void_p, C.c_int, View4, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv1x1_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "afi_conv3x3s2": (C.c_int, [C.c_void_p, C.c_int, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv3x3s2_backward": (C.c_int, [C.c_void_p, C.c_int, View4, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_conv3x3s2_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "afi_sepconv": (C.c_int, [C.c_void_p, C.c_int, View4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "afi_sepconv_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "afi_bifpn_fuse_down": (C.c_int, [View4, View4, View4, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "afi_bifpn_fuse_act": (C.c_int, [View4, View4, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "afi_bifpn_fuse_act_backward": (C.c_int, [View4, C.c_void_p, View4, View4, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "afi_launch_count": (C.c_longlong, [C.c_int]),
    "afi_profile_begin": (C.c_int, [C.c_int]),
    "afi_profile_end": (C.c_int, [C.POINTER(C.c_int)]),
    "afi_profile_get": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None
_ctx = {}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python afi-gan_b200/build.py` "
                               "(there is no CPU / PyTorch fallback for the AFI-GAN hot path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        # a library built from another revision of include/afigan_b200.h must not be driven with these struct layouts
        if l.afi_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {l.afi_abi_version()}, this binding expects {ABI_VERSION}: rebuild it")
        for which, st in enumerate((View4, GParams, Lateral, GCall, DParams, DCall, GGrads, DGrads)):
            if l.afi_sizeof(which) != C.sizeof(st):
                raise RuntimeError(f"struct {st.__name__}: the library says {l.afi_sizeof(which)} bytes, the binding {C.sizeof(st)}")
        _lib = l
    return _lib


class AfiError(RuntimeError):
    pass


def check(code: int) -> None:
    if code != 0:
        raise AfiError(f"libafigan_b200 error {code}: {lib().afi_last_error().decode()}")


def context(device: torch.device) -> C.c_void_p:
    """One afi_ctx per CUDA device of this process."""
    if device.type != "cuda":
        raise RuntimeError("the AFI-GAN hot path runs on sm_100a CUDA devices only (no CPU fallback); got device " + str(device))
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx:
        with torch.cuda.device(idx):
            h = C.c_void_p()
            check(lib().afi_create(C.byref(h)))
            _ctx[idx] = h
    return _ctx[idx]


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


DT_F32, DT_BF16 = 0, 1


def view4(t: torch.Tensor) -> View4:
    """Strided view of a 4-D fp32 or bf16 tensor (any strides: contiguous, channels_last, crops) -- no copy, no up-cast."""
    if t.dtype not in (torch.float32, torch.bfloat16) or t.dim() != 4:
        raise TypeError(f"expected a 4-D float32 / bfloat16 tensor, got {tuple(t.shape)} {t.dtype}")
    s = t.stride()
    return View4(t.data_ptr(), s[0], s[1], s[2], s[3], DT_BF16 if t.dtype == torch.bfloat16 else DT_F32, 0)


def boundary(t: torch.Tensor) -> torch.Tensor:
    """What crosses the boundary as it is: fp32 and bf16 tensors.  Anything else (fp16 under autocast, fp64) is cast to fp32 once."""
    return t if t.dtype in (torch.float32, torch.bfloat16) else t.float()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()
