"""ctypes binding of librd_b200.so (the C ABI declared in include/rd_b200.h).

The library holds every hand-written sm_100a kernel of this package.  There is no CPU or
PyTorch fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librd_b200.so")

RD_F32, RD_F16, RD_BF16 = 0, 1, 2
RD_MAX_GROUPS, RD_MAX_WINDOW = 8, 8


def XATTN_TILES(n: int) -> int:
    return (n & 3) << 4


def XATTN_POLY(n: int) -> int:
    return (n & 0xF) << 8


def XATTN_F32(n: int) -> int:
    return (n & 3) << 14

_DTYPES = {torch.float32: RD_F32, torch.float16: RD_F16, torch.bfloat16: RD_BF16}


class RdError(RuntimeError):
    pass


class CoalignProblem(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("w", C.c_int32), ("N", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32),
        ("border", C.c_int32), ("factor", C.c_int32),
        ("Hs", C.c_int32), ("Ws", C.c_int32),
        ("P", C.c_int32), ("P_pad", C.c_int32),
        ("d", C.c_int32 * RD_MAX_GROUPS),
        ("n", C.c_int32 * RD_MAX_GROUPS),
        ("row0", C.c_int32 * RD_MAX_GROUPS),
        ("n_total", C.c_int32),
        ("dtype", C.c_int32),
    ]


class CoalignHParams(C.Structure):
    _fields_ = [
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("lmda2", C.c_float), ("lmda3", C.c_float),
        ("depth_loss_weight", C.c_float), ("loss_scale", C.c_float),
    ]


_lib: Optional[C.CDLL] = None

# number of kernels of librd_b200.so launched through this binding (bench.py reports it as gpu_launches)
LAUNCHES = 0


def count_launches(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n

# name -> (restype, argtypes); every symbol include/rd_b200.h declares
_I64, _I, _F, _P = C.c_int64, C.c_int, C.c_float, C.c_void_p
_PP = C.POINTER(C.c_void_p)
_PRB, _HP = C.POINTER(CoalignProblem), C.POINTER(CoalignHParams)
SIGNATURES = {
    "rd_version": (_I, []),
    "rd_last_error": (C.c_char_p, []),
    "rd_device_arch": (_I, []),
    "rd_xattn_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I] + [_I64] * 12 + [_I, _F, _I, _P]),
    "rd_ddim_step": (_I, [_P, _P, _P, _I64, _F, _F, _F, _F, _I, _F, _I, _P]),
    "rd_ddim_add_noise": (_I, [_P, _P, _P, _I64, _F, _F, _I, _P]),
    "rd_coalign_problem_init": (_I, [_PRB]),
    "rd_reduce_scratch_bytes": (_I64, []),
    "rd_global_min": (_I, [_PP, _PRB, _P, _P, _P]),
    "rd_coalign_prep": (_I, [_PP, _PRB, _P, _P, _P]),
    "rd_coalign_set_config": (_I, [_I]),
    "rd_coalign_workspace_bytes": (_I64, [_PRB, _I]),
    "rd_coalign_optimize": (_I, [_P, _PRB, _HP, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I64, _P]),
    "rd_coalign_begin": (_I, [_PRB, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I64, _P]),
    "rd_coalign_pass": (_I, [_P, _PRB, _HP, _I, _I, _I, _I, _P, _P]),
    "rd_coalign_exchange": (_I, [_PRB, _P, _I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(_I64)]),
    "rd_coalign_end": (_I, [_PRB, _HP, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rd_merge": (_I, [_PP, _PRB, _P, _P, _P, _P, _P, _P, _P]),
    "rd_renorm": (_I, [_P, _I64, _P, _I, _P]),
    "rd_groupnorm_workspace_bytes": (_I64, [_I, _I]),
    "rd_groupnorm_nhwc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _I, _P, _I64, _P]),
    "rd_bias_residual_nhwc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "rd_upsample_nearest_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rd_phase_shuffle_bias_nhwc": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rd_geglu": (_I, [_P, _P, _I64, _I, _I, _P]),
    "rd_softmax_rows": (_I, [_P, _P, _I64, _I, _F, _I, _P]),
    "rd_layernorm": (_I, [_P, _P, _P, _P, _I64, _I, _F, _I, _P]),
}


def lib() -> C.CDLL:
    """Load librd_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RdError(
                f"{LIB_PATH} is missing: build it with `python -m rollingdepth_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)"
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().rd_last_error().decode(errors="replace")
        raise RdError(f"{what} failed with status {code}: {msg}")


def device_guard(device: torch.device):
    """Context manager making `device` the CURRENT CUDA device: the C ABI launches on the calling thread's current
    device (kernel attributes and the SM count are cached per device ordinal inside the library), while tensors
    and the stream come from `device` -- without the guard a pipeline living on cuda:1 in a process whose current
    device is cuda:0 would launch on the wrong GPU."""
    return torch.cuda.device(device)


def call(name: str, device: torch.device, *args) -> None:
    """lib().<name>(*args) with `device` current; raises RdError on a non-zero status."""
    with torch.cuda.device(device):
        code = getattr(lib(), name)(*args)
    check(code, name)


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise RdError(f"unsupported dtype {dt}; the kernels take float32, float16, bfloat16") from None


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RdError("rollingdepth_b200 kernels need CUDA tensors (no CPU fallback)")


def stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def ptr_array(ts: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr
