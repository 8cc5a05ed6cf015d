"""B200 drop-in for ``sk_dsp_comm.multirate_helper`` (reference:
src/sk_dsp_comm/multirate_helper.py:85-208).

``multirate_FIR`` / ``multirate_IIR`` keep the reference's constructor, attributes
(``b``, ``sos``, ``N_forder``), method names, keyword names and defaults
(``L_change=12``, ``M_change=12``) and its two ``log.info`` lines, so notebooks written against
the reference run unchanged:

    import sk_dsp_comm_b200.multirate_helper as mrh
    y = mrh.multirate_FIR(b).filter(x)

The arithmetic (``scipy.signal.lfilter`` / ``sosfilt`` in the reference) runs in hand-written
sm_100a CUDA kernels through the ctypes C ABI (``include/b200dsp.h``); see ``_io.py`` for the
container / dtype rules.  Plot helpers (``freq_resp``, ``zplane``) are out of scope
 and raise ``NotImplementedError``.
"""
from __future__ import annotations

import warnings
from logging import getLogger

import numpy as np
import torch

from . import _design, _engine
from ._io import Staged

log = getLogger(__name__)

# inputs at least this long that live on the HOST go through the chunked, copy/compute
# overlapped pipeline in hostpipe.py instead of one monolithic H2D -> kernel -> D2H
_HOST_PIPE_MIN = 1 << 22


def _rows(t: torch.Tensor):
    """lfilter/sosfilt filter along the last axis of an N-D input."""
    if t.dim() == 1:
        return [t.contiguous()], None
    shape = t.shape
    flat = t.reshape(-1, shape[-1]).contiguous()
    return [flat[i] for i in range(flat.shape[0])], shape


def _require_1d(x, what):
    nd = x.dim() if isinstance(x, torch.Tensor) else np.ndim(x)
    if nd != 1:
        # the reference reaches sigsys.upsample/downsample, whose reshape rejects N-D input
        raise ValueError("%s expects a 1-D signal" % what)


class _SosRunner(object):
    """sosfilt / sosfilt(L*upsample) / downsample(sosfilt) on the GPU cascade for one sos array."""

    def __init__(self, sos):
        self._plan = _engine.SosPlan(np.asarray(sos, dtype=np.float64))

    def filter(self, x):
        st = Staged(x)
        rows, shape = _rows(st.tensor)
        outs = [_engine.sos_filter(self._plan, r) for r in rows]
        y = outs[0] if shape is None else torch.stack(outs).reshape(shape)
        return st.finish(y)

    def up(self, x, L_change):
        _require_1d(x, "up")
        L = int(L_change - 1) + 1
        if L < 1:
            raise ValueError("negative dimensions are not allowed")
        st = Staged(x)
        y = _engine.sos_filter(self._plan, st.tensor.contiguous(), L=L)
        if L != L_change:
            y = y * (float(L_change) / L)
        return st.finish(y)

    def dn(self, x, M_change):
        if not isinstance(M_change, int):
            raise TypeError("M must be an int")
        _require_1d(x, "dn")
        st = Staged(x)
        return st.finish(_engine.sos_filter(self._plan, st.tensor.contiguous(), M=M_change))


class rate_change(object):
    """
    A simple class for encapsulating the upsample/filter and
    filter/downsample operations used in modeling a comm
    system. Objects of this class will hold the required filter
    coefficients once an object is instantiated.

    B200-native re-implementation of the reference class of the same name
    (multirate_helper.py:45-83): same constructor, ``M``/``fc``/``N_forder``/``b``/``a`` attributes and
    warning for an unknown ``ftype``.  The reference filters with the transfer-function form
    ``lfilter(b, a, .)``; here the same Butterworth / Chebyshev-I design (``_design.py``) runs as a
    second-order-section cascade on the GPU with the rate change fused in.
    """

    def __init__(self, M_change=12, fcutoff=0.9, N_filt_order=8, ftype='butter'):
        """
        Object constructor method
        """
        self.M = M_change            # Rate change factor M or L
        self.fc = fcutoff * .5       # must be fs/(2*M), but scale by fcutoff
        self.N_forder = N_filt_order
        self._iir = None
        if ftype.lower() == 'butter':
            self.b, self.a, sos = _design.butter(self.N_forder, 2 / self.M * self.fc)
        elif ftype.lower() == 'cheby1':
            # Set the ripple to 0.05 dB
            self.b, self.a, sos = _design.cheby1(self.N_forder, 0.05, 2 / self.M * self.fc)
        else:
            warnings.warn('ftype must be "butter" or "cheby1"')
            return
        self._iir = _SosRunner(sos)

    def up(self, x):
        """
        Upsample and filter the signal
        """
        return self._iir.up(x, self.M)       # AttributeError for a bad ftype, like the reference (no self.b)

    def dn(self, x):
        """
        Downsample and filter the signal
        """
        return self._iir.dn(x, self.M)


class multirate_FIR(object):
    """
    A simple class for encapsulating FIR filtering, or FIR upsample/
    filter, or FIR filter/downsample operations used in modeling a comm
    system. Objects of this class will hold the required filter
    coefficients once an object is instantiated.

    B200-native re-implementation of the reference class of the same name
    (multirate_helper.py:85-143).
    """

    def __init__(self, b):
        """
        Object constructor method
        """
        self.N_forder = len(b)
        self.b = b
        log.info('FIR filter taps = %d' % self.N_forder)
        barr = np.asarray(b)
        if barr.ndim != 1 or barr.size == 0:
            # scipy.signal.lfilter: "object of too small depth" / empty numerator
            raise ValueError("numerator b must be a non-empty 1-D sequence of taps")
        # complex taps (scipy.signal.lfilter accepts them, multirate_helper.py:108): by linearity two real-tap
        # passes, y = FIR(b.real, x) + 1j * FIR(b.imag, x), combined on the device (b200dsp_combine_complex)
        self._plan_im = None
        if barr.dtype.kind == "c":
            self._plan = _engine.FirPlan(np.ascontiguousarray(barr.real, dtype=np.float64))
            self._plan_im = _engine.FirPlan(np.ascontiguousarray(barr.imag, dtype=np.float64))
        else:
            self._plan = _engine.FirPlan(barr.astype(np.float64))

    def _both(self, run, t):
        """run(plan, tensor) for the real-tap plan, and the imaginary-tap plan when the taps are complex."""
        yr = run(self._plan, t)
        if self._plan_im is None:
            return yr
        yi = run(self._plan_im, t)
        shape = yr.shape
        return _engine.combine_complex(yr.reshape(-1).contiguous(), yi.reshape(-1).contiguous()).reshape(shape)

    # -- reference: y = signal.lfilter(self.b,[1],x)  (multirate_helper.py:104-109)
    def filter(self, x, zi=None):
        """
        Filter the signal

        ``zi`` (extension): the ``len(b)-1`` filter delay values of
        ``scipy.signal.lfilter(b, [1], x, zi=zi)``; when given, ``(y, zf)`` is returned and ``zf``
        fed back as the next call's ``zi`` continues the stream exactly where this block ended.
        """
        if zi is not None:
            return self._filter_stateful(x, zi)
        if isinstance(x, torch.Tensor) and not x.is_cuda and x.dim() == 1 and self._plan_im is None \
                and x.numel() >= _HOST_PIPE_MIN and x.dtype in _engine.DTYPE_CODE:
            from . import hostpipe
            return hostpipe.fir_filter_host(self._plan, x)
        st = Staged(x)
        t = st.tensor
        if t.numel() == 0:
            y = t.clone() if self._plan_im is None or t.is_complex() else \
                t.to(torch.complex64 if t.dtype == torch.float32 else torch.complex128)
        elif t.dim() == 1:
            y = self._both(_engine.fir_filter, t.contiguous())
        else:
            # lfilter filters the last axis of an N-D array: all rows in one C call
            flat = t.reshape(-1, t.shape[-1]).contiguous()
            y = self._both(_engine.fir_filter_batch, flat).reshape(t.shape)
        return st.finish(y)

    def _filter_stateful(self, x, zi):
        """lfilter's transposed-direct-form state for an FIR.  With zero input the state simply drains,
        so ``y[n] += zi[n]`` for ``n < K-1``; the final state is the filter's zero-input response after the
        block, ``zf[i] = sum_j b[i+1+j] x[N-1-j] (+ zi[i+N])`` (scipy _signaltools.lfilter / _linear_filter)
        -- i.e. what the FIR kernel outputs for K-1 zero samples whose history is the block's tail.  Both
        the block and that (K-1)-sample launch run on the device; nothing is read back, so a stream of
        blocks chained through ``zf`` never synchronises with the host."""
        _require_1d(x, "filter(zi=...)")
        if self._plan_im is not None:
            raise NotImplementedError("filter(zi=...) with complex taps is not supported by the B200 engine")
        K = self._plan.ntaps
        st = Staged(x)
        t = st.tensor.contiguous()
        N = t.numel()
        if isinstance(zi, torch.Tensor):
            z = zi.detach().reshape(-1)
        else:
            z = torch.from_numpy(np.ascontiguousarray(np.asarray(zi)).reshape(-1))
        if z.numel() != K - 1:
            raise ValueError("Unexpected shape for zi: expected (%d,), found %s." % (K - 1, tuple(np.shape(zi))))
        if z.is_complex() and not t.is_complex():
            t = t.to(torch.complex64 if t.dtype == torch.float32 else torch.complex128)
        z = z.to(device=t.device, dtype=t.dtype)
        y = _engine.fir_filter(self._plan, t) if N else t.clone()
        m = min(K - 1, N)
        if m:
            y[:m] += z[:m]
        if K == 1:
            zf = z.clone()
        else:
            if N >= K - 1:
                tail = t[N - (K - 1):]
            else:                       # block shorter than the filter memory: zeros stand in for x[<0]
                tail = torch.cat([torch.zeros(K - 1 - N, dtype=t.dtype, device=t.device), t])
            zf = _engine.fir_filter(self._plan, torch.zeros(K - 1, dtype=t.dtype, device=t.device),
                                    hist=tail.contiguous())
            if N < K - 1:
                zf[:K - 1 - N] += z[N:]
        if st.kind == "numpy":
            # numpy streams are computed in float64 / complex128 like the reference
            return st.finish(y), zf.cpu().numpy()
        return st.finish(y), (zf if st.kind == "cuda" else zf.cpu())

    # -- reference: y = L*upsample(x,L); y = lfilter(b,[1],y)  (multirate_helper.py:112-118)
    def up(self, x, L_change=12):
        """
        Upsample and filter the signal
        """
        _require_1d(x, "up")
        L = int(L_change - 1) + 1        # sigsys.upsample truncates float factors: int(L-1) zeros
        if L < 1:
            raise ValueError("negative dimensions are not allowed")
        st = Staged(x)
        t = st.tensor.contiguous()
        if t.numel() == 0:
            return st.finish(t.clone())
        y = self._both(lambda pl, tt: _engine.fir_up(pl, tt, L), t)
        if L != L_change:
            y = y * (float(L_change) / L)     # reference gain is the un-truncated L_change
        return st.finish(y)

    # -- reference: y = lfilter(b,[1],x); y = downsample(y,M)  (multirate_helper.py:121-127)
    def dn(self, x, M_change=12):
        """
        Downsample and filter the signal
        """
        if not isinstance(M_change, int):
            raise TypeError("M must be an int")
        _require_1d(x, "dn")
        st = Staged(x)
        t = st.tensor.contiguous()
        if t.numel() // M_change == 0:
            return st.finish(torch.empty(0, dtype=t.dtype, device=t.device))
        return st.finish(self._both(lambda pl, tt: _engine.fir_dn(pl, tt, M_change), t))

    def freq_resp(self, mode='dB', fs=8000, ylim=[-100, 2]):
        raise NotImplementedError("plot helpers are out of scope of the B200 engine (SURVEY.md section 2)")

    def zplane(self, auto_scale=True, size=2, detect_mult=True, tol=0.001):
        raise NotImplementedError("plot helpers are out of scope of the B200 engine (SURVEY.md section 2)")


def _validate_sos(sos):
    """scipy.signal._validate_sos: (n_sections, 6) with sos[:, 3] == 1, else ValueError."""
    sos = np.atleast_2d(np.asarray(sos))
    if sos.ndim != 2:
        raise ValueError('sos array must be 2D')
    n_sections, m = sos.shape
    if m != 6:
        raise ValueError('sos array must be shape (n_sections, 6)')
    if not (sos[:, 3] == 1).all():
        raise ValueError('sos[:, 3] should be all ones')
    return sos


class multirate_IIR(object):
    """
    A simple class for encapsulating IIR filtering, or IIR upsample/
    filter, or IIR filter/downsample operations used in modeling a comm
    system. All filtering is done on a cascade of second-order sections,
    y = sosfilt(sos,x).

    B200-native re-implementation of the reference class of the same name
    (multirate_helper.py:146-208): the cascade runs as a parallel-prefix recurrence
    (csrc/sos_scan.cu).
    """

    def __init__(self, sos):
        """
        Object constructor method
        """
        self.N_forder = np.sum(np.sign(np.abs(sos[:, 2]))) \
                      + np.sum(np.sign(np.abs(sos[:, 1])))
        self.sos = sos
        log.info('IIR filter order = %d' % self.N_forder)
        self._plan = None

    def _get_plan(self):
        if self._plan is None:
            sos = _validate_sos(self.sos)          # sosfilt validates at call time in the reference
            if sos.dtype.kind == "c":
                raise NotImplementedError("complex sos coefficients are not supported by the B200 engine")
            self._plan = _engine.SosPlan(sos.astype(np.float64))
        return self._plan

    # -- reference: y = signal.sosfilt(self.sos,x)  (multirate_helper.py:169-174)
    def filter(self, x, zi=None):
        """
        Filter the signal using second-order sections

        ``zi`` (extension): ``(n_sections, 2)`` section delay values in
        ``scipy.signal.sosfilt``'s layout; when given, ``(y, zf)`` is returned.  The parallel-prefix
        kernel carries the same transposed-direct-form-II state, so chunked calls chained through
        ``zf`` reproduce the monolithic result.
        """
        plan = self._get_plan()
        if zi is not None:
            _require_1d(x, "filter(zi=...)")
            st = Staged(x)
            t = st.tensor.contiguous()
            z = zi if isinstance(zi, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(zi)))
            if tuple(z.shape) != (plan.nsec, 2):
                raise ValueError("Invalid zi shape. With axis=-1, an input with shape %s, and an sos array "
                                 "with %d sections, zi must have shape (%d, 2), got %s."
                                 % (tuple(t.shape), plan.nsec, plan.nsec, tuple(z.shape)))
            if z.is_complex():
                if not t.is_complex():
                    t = t.to(torch.complex64 if t.dtype == torch.float32 else torch.complex128)
                z = torch.view_as_real(z.to(torch.complex128).contiguous())          # (nsec, 2, re/im)
            elif t.is_complex():
                z = torch.stack([z.to(torch.float64), torch.zeros_like(z, dtype=torch.float64)], dim=-1)
            y, zf = _engine.sos_filter(plan, t, zi=z, return_zf=True)
            if t.is_complex():
                zf = torch.view_as_complex(zf.to(torch.float64).contiguous())
            else:
                zf = zf.to(torch.float64)
            if st.kind == "numpy":
                return st.finish(y), zf.cpu().numpy()
            return st.finish(y), (zf if st.kind == "cuda" else zf.cpu())
        st = Staged(x)
        t = st.tensor
        if t.dim() == 1:
            y = _engine.sos_filter(plan, t.contiguous())
        else:
            # sosfilt filters the last axis of an N-D array: all rows in one C call
            y = _engine.sos_filter_batch(plan, t.reshape(-1, t.shape[-1]).contiguous()).reshape(t.shape)
        return st.finish(y)

    # -- reference: y = L*upsample(x,L); y = sosfilt(sos,y)  (multirate_helper.py:177-183)
    def up(self, x, L_change=12):
        """
        Upsample and filter the signal
        """
        plan = self._get_plan()
        _require_1d(x, "up")
        L = int(L_change - 1) + 1
        if L < 1:
            raise ValueError("negative dimensions are not allowed")
        st = Staged(x)
        t = st.tensor.contiguous()
        y = _engine.sos_filter(plan, t, L=L)
        if L != L_change:
            y = y * (float(L_change) / L)
        return st.finish(y)

    # -- reference: y = sosfilt(sos,x); y = downsample(y,M)  (multirate_helper.py:186-192)
    def dn(self, x, M_change=12):
        """
        Downsample and filter the signal
        """
        if not isinstance(M_change, int):
            raise TypeError("M must be an int")
        plan = self._get_plan()
        _require_1d(x, "dn")
        st = Staged(x)
        t = st.tensor.contiguous()
        return st.finish(_engine.sos_filter(plan, t, M=M_change))

    def freq_resp(self, mode='dB', fs=8000, ylim=[-100, 2]):
        raise NotImplementedError("plot helpers are out of scope of the B200 engine (SURVEY.md section 2)")

    def zplane(self, auto_scale=True, size=2, detect_mult=True, tol=0.001):
        raise NotImplementedError("plot helpers are out of scope of the B200 engine (SURVEY.md section 2)")
