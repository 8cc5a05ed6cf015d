"""Device-tensor level wrappers over the C ABI: plans + one function per entry point.

Everything here takes 1-D contiguous ``torch.cuda`` tensors (float32/float64/complex64/
complex128) and returns freshly allocated tensors on the same device.  Dtype policy, numpy /
host staging and error contracts of the reference live one level up (multirate_helper.py,
sigsys.py); multi-GPU sharding lives in sharded.py.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _cabi
from ._cabi import lib, check, DTYPE_CODE, REAL_OF


def _dev_index(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _check_x(x: torch.Tensor):
    if not x.is_cuda:
        raise ValueError("device-level call needs a CUDA tensor")
    if x.dtype not in DTYPE_CODE:
        raise NotImplementedError("input type '%s' not supported" % x.dtype)
    if x.dim() != 1 or not x.is_contiguous():
        raise ValueError("device-level call needs a 1-D contiguous tensor")


class FirPlan:
    """Device-resident taps of one FIR filter (b200dsp_fir_plan), one handle per device."""

    def __init__(self, b):
        self.taps = np.ascontiguousarray(np.asarray(b, dtype=np.float64).ravel())
        if self.taps.size < 1:
            raise ValueError("FIR needs at least one tap")
        self.ntaps = int(self.taps.size)
        self._handles = {}

    def handle(self, dev: int):
        h = self._handles.get(dev)
        if h is None:
            _cabi.require_cuda()
            out = ctypes.c_void_p()
            with torch.cuda.device(dev):
                check(lib.b200dsp_fir_plan_create(self.taps.ctypes.data_as(ctypes.c_void_p),
                                                  self.ntaps, ctypes.byref(out)), "fir_plan_create")
            h = out.value
            self._handles[dev] = h
        return h

    def up_hist_len(self, L: int) -> int:
        return (self.ntaps - 1 + L - 1) // L

    def __del__(self):
        try:
            for h in self._handles.values():
                lib.b200dsp_fir_plan_destroy(h)
        except Exception:
            pass
        self._handles = {}


class SosPlan:
    """Device-resident biquad cascade + scan matrices (b200dsp_sos_plan), one handle per device."""

    def __init__(self, sos):
        s = np.atleast_2d(np.asarray(sos, dtype=np.float64))
        self.sos = np.ascontiguousarray(s)
        self.nsec = int(self.sos.shape[0])
        self._handles = {}

    def handle(self, dev: int):
        h = self._handles.get(dev)
        if h is None:
            _cabi.require_cuda()
            out = ctypes.c_void_p()
            with torch.cuda.device(dev):
                check(lib.b200dsp_sos_plan_create(self.sos.ctypes.data_as(ctypes.c_void_p),
                                                  self.nsec, ctypes.byref(out)), "sos_plan_create")
            h = out.value
            self._handles[dev] = h
        return h

    def __del__(self):
        try:
            for h in self._handles.values():
                lib.b200dsp_sos_plan_destroy(h)
        except Exception:
            pass
        self._handles = {}


def _hist_ptr(hist, x, need):
    if hist is None:
        return None
    if hist.dtype != x.dtype or hist.device != x.device or not hist.is_contiguous() or hist.numel() != need:
        raise ValueError("hist must be a contiguous %s tensor of %d samples on %s"
                         % (x.dtype, need, x.device))
    return hist.data_ptr() if need > 0 else None


def fir_filter(plan: FirPlan, x: torch.Tensor, hist=None, out=None) -> torch.Tensor:
    """y[i] = sum_k b[k] xe[i-k]; hist = the ntaps-1 samples preceding x (or None = zeros)."""
    _check_x(x)
    dev = _dev_index(x)
    y = torch.empty_like(x) if out is None else out
    with torch.cuda.device(dev):
        check(lib.b200dsp_fir_filter(plan.handle(dev), DTYPE_CODE[x.dtype], x.data_ptr(),
                                     _hist_ptr(hist, x, plan.ntaps - 1), y.data_ptr(), x.numel(),
                                     _cabi.stream_ptr(dev)), "fir_filter")
    return y


def _check_rows(x: torch.Tensor):
    if not x.is_cuda:
        raise ValueError("device-level call needs a CUDA tensor")
    if x.dtype not in DTYPE_CODE:
        raise NotImplementedError("input type '%s' not supported" % x.dtype)
    if x.dim() != 2 or not x.is_contiguous():
        raise ValueError("batched call needs a contiguous (rows, n) tensor")


def fir_filter_batch(plan: FirPlan, x: torch.Tensor) -> torch.Tensor:
    """Every row of the (rows, n) tensor filtered from zero state: one C call (b200dsp_fir_filter_batch)."""
    _check_rows(x)
    dev = _dev_index(x)
    y = torch.empty_like(x)
    rows, n = x.shape
    with torch.cuda.device(dev):
        check(lib.b200dsp_fir_filter_batch(plan.handle(dev), DTYPE_CODE[x.dtype], x.data_ptr(), y.data_ptr(),
                                           rows, n, n, n, _cabi.stream_ptr(dev)), "fir_filter_batch")
    return y


def sos_filter_batch(plan: SosPlan, x: torch.Tensor) -> torch.Tensor:
    """Every row of the (rows, n) tensor through the cascade from zero state: one C call."""
    _check_rows(x)
    dev = _dev_index(x)
    y = torch.empty_like(x)
    rows, n = x.shape
    if rows and n:
        code = DTYPE_CODE[x.dtype]
        with torch.cuda.device(dev):
            h = plan.handle(dev)
            nbytes = int(lib.b200dsp_sos_workspace_bytes(h, code, n, 1, 1))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            check(lib.b200dsp_sos_filter_batch(h, code, x.data_ptr(), y.data_ptr(), rows, n, n, n, ws.data_ptr(),
                                               nbytes, _cabi.stream_ptr(dev)), "sos_filter_batch")
    return y


def combine_complex(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + 1j*b on the device (b200dsp_combine_complex): real a, b -> complex; complex a, b -> complex."""
    _check_x(a)
    if b.dtype != a.dtype or b.shape != a.shape or not b.is_cuda or not b.is_contiguous():
        raise ValueError("combine_complex needs two matching contiguous CUDA tensors")
    dev = _dev_index(a)
    odt = {torch.float32: torch.complex64, torch.float64: torch.complex128}.get(a.dtype, a.dtype)
    y = torch.empty(a.numel(), dtype=odt, device=a.device)
    with torch.cuda.device(dev):
        check(lib.b200dsp_combine_complex(DTYPE_CODE[a.dtype], a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(),
                                          _cabi.stream_ptr(dev)), "combine_complex")
    return y


def fir_up(plan: FirPlan, x: torch.Tensor, L: int, hist=None, out=None) -> torch.Tensor:
    _check_x(x)
    dev = _dev_index(x)
    y = torch.empty(x.numel() * L, dtype=x.dtype, device=x.device) if out is None else out
    with torch.cuda.device(dev):
        check(lib.b200dsp_fir_up(plan.handle(dev), DTYPE_CODE[x.dtype], x.data_ptr(),
                                 _hist_ptr(hist, x, plan.up_hist_len(L)), y.data_ptr(), x.numel(),
                                 L, _cabi.stream_ptr(dev)), "fir_up")
    return y


def fir_dn(plan: FirPlan, x: torch.Tensor, M: int, hist=None, out=None) -> torch.Tensor:
    _check_x(x)
    dev = _dev_index(x)
    y = torch.empty(x.numel() // M, dtype=x.dtype, device=x.device) if out is None else out
    with torch.cuda.device(dev):
        check(lib.b200dsp_fir_dn(plan.handle(dev), DTYPE_CODE[x.dtype], x.data_ptr(),
                                 _hist_ptr(hist, x, plan.ntaps - 1), y.data_ptr(), x.numel(),
                                 M, _cabi.stream_ptr(dev)), "fir_dn")
    return y


def sos_filter(plan: SosPlan, x: torch.Tensor, L: int = 1, M: int = 1, zi=None, return_zf=False):
    """sosfilt (L=M=1), sosfilt(L*upsample(x,L)) (L>1) or downsample(sosfilt(x),M) (M>1).

    zi / zf: (nsec, 2) tensors in the real scalar type (complex x: (nsec, 2, 2) with the last
    axis = re/im channel), scipy's layout.
    """
    _check_x(x)
    dev = _dev_index(x)
    n = x.numel()
    code = DTYPE_CODE[x.dtype]
    y = torch.empty((n * L) // M, dtype=x.dtype, device=x.device)
    rdt = REAL_OF[x.dtype]
    nch = 2 if x.is_complex() else 1
    zshape = (plan.nsec, 2) if nch == 1 else (plan.nsec, 2, 2)
    zf = torch.empty(zshape, dtype=rdt, device=x.device) if return_zf else None
    if zi is not None:
        zi = zi.to(device=x.device, dtype=rdt).contiguous()
        if tuple(zi.shape) != zshape:
            raise ValueError("zi must have shape %s" % (zshape,))
    if n > 0:
        with torch.cuda.device(dev):
            h = plan.handle(dev)
            nbytes = int(lib.b200dsp_sos_workspace_bytes(h, code, n, L, M))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            check(lib.b200dsp_sos_filter(h, code, x.data_ptr(), y.data_ptr(), n, L, M,
                                         zi.data_ptr() if zi is not None else None,
                                         zf.data_ptr() if zf is not None else None,
                                         ws.data_ptr(), nbytes, _cabi.stream_ptr(dev)), "sos_filter")
    elif return_zf:
        zf.copy_(zi) if zi is not None else zf.zero_()
    return (y, zf) if return_zf else y


def upsample(x: torch.Tensor, L: int) -> torch.Tensor:
    _check_x(x)
    dev = _dev_index(x)
    y = torch.empty(x.numel() * L, dtype=x.dtype, device=x.device)
    with torch.cuda.device(dev):
        check(lib.b200dsp_upsample(DTYPE_CODE[x.dtype], x.data_ptr(), y.data_ptr(), x.numel(), L,
                                   _cabi.stream_ptr(dev)), "upsample")
    return y


def downsample(x: torch.Tensor, M: int, p: int = 0) -> torch.Tensor:
    _check_x(x)
    dev = _dev_index(x)
    y = torch.empty(x.numel() // M, dtype=x.dtype, device=x.device)
    with torch.cuda.device(dev):
        check(lib.b200dsp_downsample(DTYPE_CODE[x.dtype], x.data_ptr(), y.data_ptr(), x.numel(), M, p,
                                     _cabi.stream_ptr(dev)), "downsample")
    return y
