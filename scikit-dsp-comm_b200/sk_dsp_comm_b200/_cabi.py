"""ctypes binding of libb200dsp.so (C ABI declared in include/b200dsp.h).

There is NO fallback: if the CUDA library has not been built, importing this module raises
ImportError; if no CUDA device is present, the first call raises RuntimeError.
PyTorch is used only as plumbing (device memory, streams, torch.distributed).
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200DSP_LIB", os.path.join(_HERE, "_lib", "libb200dsp.so"))

F32, F64, C64, C128 = 0, 1, 2, 3
DTYPE_CODE = {torch.float32: F32, torch.float64: F64, torch.complex64: C64, torch.complex128: C128}
REAL_OF = {torch.float32: torch.float32, torch.float64: torch.float64,
           torch.complex64: torch.float32, torch.complex128: torch.float64}

# every symbol include/b200dsp.h declares (tests check the library exports all of them)
SYMBOLS = (
    "b200dsp_version", "b200dsp_last_error", "b200dsp_device_info",
    "b200dsp_fir_plan_create", "b200dsp_fir_plan_destroy", "b200dsp_fir_plan_ntaps",
    "b200dsp_fir_filter", "b200dsp_fir_filter_batch", "b200dsp_fir_up", "b200dsp_fir_up_hist_len", "b200dsp_fir_dn",
    "b200dsp_sos_plan_create", "b200dsp_sos_plan_destroy", "b200dsp_sos_plan_nsec",
    "b200dsp_sos_workspace_bytes", "b200dsp_sos_filter", "b200dsp_sos_filter_batch",
    "b200dsp_upsample", "b200dsp_downsample", "b200dsp_combine_complex",
    "b200dsp_launch_count", "b200dsp_launch_count_reset", "b200dsp_set_fir_variant",
    "b200dsp_set_sos_variant",
)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libb200dsp.so not found at %s -- build it with "
            "`python scikit-dsp-comm_b200/build.py` (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    P, I64, I32, I, SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, ctypes.c_size_t
    PP = ctypes.POINTER(ctypes.c_void_p)
    lib.b200dsp_version.restype = I
    lib.b200dsp_last_error.restype = ctypes.c_char_p
    lib.b200dsp_device_info.argtypes = [ctypes.POINTER(I)] * 3
    lib.b200dsp_fir_plan_create.argtypes = [P, I32, PP]
    lib.b200dsp_fir_plan_destroy.argtypes = [P]
    lib.b200dsp_fir_plan_destroy.restype = None
    lib.b200dsp_fir_plan_ntaps.argtypes = [P]
    lib.b200dsp_fir_plan_ntaps.restype = I32
    lib.b200dsp_fir_filter.argtypes = [P, I, P, P, P, I64, P]
    lib.b200dsp_fir_filter_batch.argtypes = [P, I, P, P, I64, I64, I64, I64, P]
    lib.b200dsp_sos_filter_batch.argtypes = [P, I, P, P, I64, I64, I64, I64, P, SZ, P]
    lib.b200dsp_combine_complex.argtypes = [I, P, P, P, I64, P]
    lib.b200dsp_fir_up.argtypes = [P, I, P, P, P, I64, I32, P]
    lib.b200dsp_fir_up_hist_len.argtypes = [P, I32]
    lib.b200dsp_fir_up_hist_len.restype = I32
    lib.b200dsp_fir_dn.argtypes = [P, I, P, P, P, I64, I32, P]
    lib.b200dsp_sos_plan_create.argtypes = [P, I32, PP]
    lib.b200dsp_sos_plan_destroy.argtypes = [P]
    lib.b200dsp_sos_plan_destroy.restype = None
    lib.b200dsp_sos_plan_nsec.argtypes = [P]
    lib.b200dsp_sos_plan_nsec.restype = I32
    lib.b200dsp_sos_workspace_bytes.argtypes = [P, I, I64, I32, I32]
    lib.b200dsp_sos_workspace_bytes.restype = SZ
    lib.b200dsp_sos_filter.argtypes = [P, I, P, P, I64, I32, I32, P, P, P, SZ, P]
    lib.b200dsp_upsample.argtypes = [I, P, P, I64, I32, P]
    lib.b200dsp_downsample.argtypes = [I, P, P, I64, I32, I32, P]
    lib.b200dsp_launch_count.restype = I64
    lib.b200dsp_launch_count_reset.restype = None
    lib.b200dsp_set_fir_variant.argtypes = [I]
    lib.b200dsp_set_fir_variant.restype = None
    lib.b200dsp_set_sos_variant.argtypes = [I]
    lib.b200dsp_set_sos_variant.restype = None
    return lib


lib = _load()


class B200DspError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.b200dsp_last_error()
        raise B200DspError("%s failed (code %d): %s" % (what or "b200dsp call", rc,
                                                       msg.decode() if msg else "?"))


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("sk_dsp_comm_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib.b200dsp_launch_count())


def launch_count_reset():
    lib.b200dsp_launch_count_reset()
