"""B200 drop-in for ``sk_dsp_comm.sigsys.upsample`` / ``downsample``
(reference: src/sk_dsp_comm/sigsys.py:3031-3083).

Same names, argument meaning, defaults and error behaviour; the index maps run as CUDA kernels
(``b200dsp_upsample`` / ``b200dsp_downsample``) and are bit exact.

dtype / container rules (see ``_io.py``): numpy in -> numpy out with the reference's dtypes
(``upsample`` widens to >= float64 because the reference builds the zeros with ``np.zeros``;
``downsample`` preserves dtype); torch tensors keep their dtype.  ``downsample`` returns a fresh
contiguous array where the reference returns a strided view of its argument.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _engine
from ._io import Staged


_butter24 = {}


def _butter10(M):
    """10th-order Butterworth low-pass with cutoff 1/M as a GPU cascade (cached per M)."""
    r = _butter24.get(M)
    if r is None:
        from . import _design
        from .multirate_helper import _SosRunner
        r = _SosRunner(_design.butter(10, 1.0 / M)[2])
        _butter24[M] = r
    return r


def interp24(x):
    """
    Interpolate by L = 24 using Butterworth filters.

    The interpolation is done using three stages. Upsample by
    L = 2 and lowpass filter, upsample by 3 and lowpass filter, then
    upsample by L = 4 and lowpass filter. In all cases the lowpass
    filter is a 10th-order Butterworth lowpass (reference: sigsys.py:2945-2985);
    every stage runs as one fused zero-stuff + biquad-cascade pass on the GPU.
    """
    y = _butter10(2).up(x, 2)
    y = _butter10(3).up(y, 3)
    return _butter10(4).up(y, 4)


def deci24(x):
    """
    Decimate by L = 24 using Butterworth filters: lowpass + downsample by 2, 3 and 4
    (reference: sigsys.py:2988-3028), each stage one fused cascade + decimation pass on the GPU.
    """
    y = _butter10(2).dn(x, 2)
    y = _butter10(3).dn(y, 3)
    return _butter10(4).dn(y, 4)


def _is_arraylike_1d(x):
    if isinstance(x, torch.Tensor):
        if x.dim() != 1:
            # reference: x.reshape(N_input, 1) on an N-D array -> ValueError (sigsys.py:3051)
            raise ValueError("cannot reshape array of size %d into shape (%d,1)"
                             % (x.numel(), x.shape[0] if x.dim() else 0))
        return
    if not hasattr(x, "reshape"):
        # reference calls x.reshape(...) directly: a list raises AttributeError
        raise AttributeError("'%s' object has no attribute 'reshape'" % type(x).__name__)
    if np.ndim(x) != 1:
        raise ValueError("cannot reshape array of size %d into shape (%d,1)" % (np.size(x), len(x)))


def upsample(x, L):
    """
    Upsample by factor L

    Insert L - 1 zero samples in between each input sample.

    Parameters
    ----------
    x : ndarray (or torch tensor) of input signal values
    L : upsample factor

    Returns
    -------
    y : the output signal values, ``len(y) == L*len(x)``

    Examples
    --------
    >>> y = upsample(x,3)
    """
    _is_arraylike_1d(x)
    Li = int(L - 1) + 1          # the reference zero-stuffs int(L-1) samples (sigsys.py:3051)
    if Li < 1:
        raise ValueError("negative dimensions are not allowed")
    st = Staged(x, widen=True)
    t = st.tensor.contiguous()
    if t.numel() == 0:
        return st.finish(torch.empty(0, dtype=t.dtype, device=t.device))
    return st.finish(_engine.upsample(t, Li))


def downsample(x, M, p=0):
    """
    Downsample by factor M

    Keep every Mth sample of the input. The phase of the input samples
    kept can be selected.

    Parameters
    ----------
    x : ndarray (or torch tensor) of input signal values
    M : downsample factor
    p : phase of decimated value, 0 (default), 1, ..., M-1

    Returns
    -------
    y : the output signal values, ``len(y) == floor(len(x)/M)``

    Examples
    --------
    >>> y = downsample(x,3)
    >>> y = downsample(x,3,1)
    """
    if not isinstance(M, int):
        raise TypeError("M must be an int")          # sigsys.py:3078-3079
    if M == 0:
        raise ZeroDivisionError("float division by zero")
    n = len(x)
    if isinstance(x, torch.Tensor):
        if x.dim() != 1:
            raise ValueError("downsample expects a 1-D signal")
    elif np.ndim(x) != 1:
        raise ValueError("downsample expects a 1-D signal")
    p = int(p)
    if p >= M or p < -M:
        # reference: x[:, p] on an (n/M, M) array (sigsys.py:3082)
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (p, M))
    if p < 0:
        p += M
    is_np = not isinstance(x, torch.Tensor)
    if is_np:
        a = np.asarray(x)
        orig_dtype = a.dtype
        if a.dtype.kind not in "biufc":
            raise NotImplementedError("input type '%s' not supported" % a.dtype)
        if a.dtype in (np.float32, np.float64, np.complex64, np.complex128):
            moved = a
        elif a.dtype.kind in "iu" and a.dtype.itemsize == 8:
            moved = a.view(np.float64)          # moved as raw 8-byte words: bit exact
        elif a.dtype.kind in "iu" and a.dtype.itemsize == 4:
            moved = a.view(np.float32)          # raw 4-byte words
        else:
            moved = a.astype(np.float64)        # int8/16, bool, float16: exact in float64
        st = Staged(moved, widen=False)
    else:
        narrow = x.dtype not in (torch.float32, torch.float64, torch.complex64, torch.complex128)
        st = Staged(x)
    t = st.tensor.contiguous()
    if n // M == 0:
        y = torch.empty(0, dtype=t.dtype, device=t.device)
    else:
        y = _engine.downsample(t, M, p)
    out = st.finish(y)
    if is_np:
        if out.dtype != orig_dtype:
            if orig_dtype.kind in "iu" and orig_dtype.itemsize in (4, 8):
                out = out.view(orig_dtype)
            else:
                out = out.astype(orig_dtype)
    elif narrow:
        out = out.to(x.dtype)
    return out
