"""Container / dtype policy of the drop-in surface.

* numpy in  -> numpy out, with the REFERENCE's dtypes: lfilter/sosfilt promote through their
  float64 coefficients, so float32/ints -> float64 and complex64 -> complex128
  (``np.result_type(b, a, x)``).  The data crosses PCIe in its narrow input dtype and is widened
  on the device.
* torch.cuda tensor in -> torch.cuda tensor out on the same device, dtype-PRESERVING for
  float32 / complex64 / float64 / complex128 (the one documented deviation from the reference,
  which would widen float32/complex64: the throughput configs of BASELINE.json are quoted on
  dtype-preserving streams).  Other torch dtypes are widened to float64 like the reference.
* torch CPU tensor in -> staged through pinned memory, torch CPU tensor out (same dtype rule as
  torch.cuda).
Everything is computed by the CUDA kernels; there is no host arithmetic path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi

_SUPPORTED = (torch.float32, torch.float64, torch.complex64, torch.complex128)


def _ref_numpy_dtype(x: np.ndarray) -> np.dtype:
    """dtype scipy would compute in: result_type(float64 coefficients, x) restricted to d/D."""
    if x.dtype.kind not in "biufc":
        raise NotImplementedError("input type '%s' not supported" % x.dtype)
    return np.dtype(np.complex128) if x.dtype.kind == "c" else np.dtype(np.float64)


class Staged:
    """A host/device input normalised to a CUDA tensor plus the way back."""

    def __init__(self, x, device=None, widen=True):
        _cabi.require_cuda()
        self.kind = "cuda"
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                t = x
            else:
                self.kind = "cpu"
                dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
                src = x if x.is_pinned() else x.contiguous()
                t = src.to(dev, non_blocking=True)
            if t.dtype not in _SUPPORTED:
                if t.dtype == torch.complex32:
                    t = t.to(torch.complex128)
                elif t.dtype.is_floating_point or t.dtype in (torch.int8, torch.int16, torch.int32,
                                                              torch.int64, torch.uint8, torch.bool):
                    t = t.to(torch.float64)
                else:
                    raise NotImplementedError("input type '%s' not supported" % t.dtype)
        else:
            self.kind = "numpy"
            a = np.asarray(x)
            tgt = _ref_numpy_dtype(a) if widen else a.dtype
            if a.dtype == np.float16 or a.dtype.kind == "b" or a.dtype == np.longdouble \
                    or a.dtype == np.clongdouble or a.dtype.itemsize > 16:
                a = a.astype(tgt)
            if a.dtype.kind == "u" and a.dtype.itemsize > 1:      # torch lacks most unsigned types
                a = a.astype(np.float64)
            dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
            t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            if widen:
                t = t.to(torch.complex128 if tgt.kind == "c" else torch.float64)
        self.tensor = t

    def finish(self, y: torch.Tensor):
        if self.kind == "numpy":
            return y.cpu().numpy()
        if self.kind == "cpu":
            return y.cpu()
        return y
