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


# ---------------------------------------------------------------------------------------------
# Pulse-shaped line codes (reference: sigsys.py:1847-2202).  The waveform is
# one launch of the fused zero-stuff + FIR kernel (``_pulse.shape_symbols``); bits come from the
# legacy ``np.random`` generator in the reference's order so seeded runs reproduce its output.

from . import _pulse                                             # noqa: E402
from ._pulse import rc_imp, sqrt_rc_imp, m_seq, pn_gen           # noqa: E402,F401


def _nrz_pulse(pulse, ns, alpha, m):
    b = _pulse.pulse_taps(pulse, ns, alpha, m)
    if b is None:
        raise ValueError('pulse type must be rec, rc, or src')
    return b


def nrz_bits(n_bits, ns, pulse='rect', alpha=0.25, m=6):
    """
    Generate non-return-to-zero (NRZ) data bits with pulse shaping.

    A baseband digital data signal using +/-1 amplitude signal values and including pulse
    shaping ('rect', 'rc' or 'src'; ``2*m*ns + 1`` taps for the latter two).  Returns the
    waveform, ``b / ns`` and the 0/1 bits (reference: sigsys.py:2099-2150).
    """
    data = np.random.randint(0, 2, n_bits)
    b = _nrz_pulse(pulse, ns, alpha, m)
    x = _pulse.shape_symbols(2.0 * data - 1, b, ns)
    return x, b / float(ns), data


def nrz_bits2(data, Ns, pulse='rect', alpha=0.25, M=6):
    """
    NRZ with user supplied 0/1 ``data`` (reference: sigsys.py:2153-2202); returns the waveform
    and ``b / Ns``.
    """
    b = _nrz_pulse(pulse, Ns, alpha, M)
    d = np.asarray(data)
    x = _pulse.shape_symbols(2 * d.reshape(len(d)) - 1, b, Ns)
    return x, b / float(Ns)


def bpsk_tx(N_bits, Ns, ach_fc=2.0, ach_lvl_dB=-100, pulse='rect', alpha=0.25, M=6):
    """
    BPSK transmitter with adjacent channel interference: the wanted NRZ stream plus two more at
    ``+/- ach_fc/Ns`` cycles/sample and ``ach_lvl_dB`` (reference: sigsys.py:2053-2096; only
    'rect' and 'src' pulses are accepted there).
    """
    if pulse not in ('rect', 'src'):
        raise ValueError('Pulse shape must be \'rect\' or \'src\'')
    x0, b, data0 = nrz_bits(N_bits, Ns, pulse, alpha, M)
    x1p, b, _ = nrz_bits(N_bits, Ns, pulse, alpha, M)
    x1m, b, _ = nrz_bits(N_bits, Ns, pulse, alpha, M)
    n = np.arange(len(x0))
    x1p = x1p * np.exp(1j * 2 * np.pi * ach_fc / float(Ns) * n)
    x1m = x1m * np.exp(-1j * 2 * np.pi * ach_fc / float(Ns) * n)
    ach_lvl = 10 ** (ach_lvl_dB / 20.)
    return x0 + ach_lvl * (x1p + x1m), b, data0


# ---------------------------------------------------------------------------------------------
# Block FIR filtering (reference: sigsys.py:482-598).  The reference walks
# the signal in FFT frames of N samples (overlap-save / overlap-add) to obtain exactly the linear
# convolution truncated to len(x); the FIR kernel already is an overlap-save over shared-memory
# tiles, so the filtered output is one launch.  ``mode=1`` rebuilds the per-frame diagnostic rows
# (circular convolution of each frame for overlap-save, the zero-padded frame's linear
# convolution for overlap-add) with one launch per frame.

def _block_filter_args(x, h, N):
    h = np.asarray(h)
    if h.ndim != 1 or h.size == 0:
        raise ValueError("h must be a non-empty 1-D sequence of taps")
    if h.dtype.kind == "c":
        raise NotImplementedError("complex FIR taps are not supported by the B200 engine")
    P = len(h)
    L = int(N) - P + 1
    if L < 1:
        raise ValueError("FFT size N must be at least len(h)")
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("x must be 1-D")
    # the reference keeps np.real(ifft(.)): with real taps that is the filtered real part of x
    x = np.ascontiguousarray(x.real, dtype=np.float64)
    return x, h.astype(np.float64), P, L


def _run_fir(plan, seg, hist=None):
    st = Staged(seg)
    hd = None if hist is None else torch.from_numpy(np.ascontiguousarray(hist)).to(st.tensor.device)
    return st.finish(_engine.fir_filter(plan, st.tensor, hist=hd))


def os_filter(x, h, N, mode=0):
    """
    Overlap and save transform domain FIR filtering.

    Returns ``y`` (``len(x)`` samples of ``h * x``) and, for ``mode == 1``, the diagnostic matrix
    whose row ``k`` holds frame ``k``'s circular-convolution output placed at its position in the
    stream (reference: sigsys.py:482-540).
    """
    x, h, P, L = _block_filter_args(x, h, N)
    N = int(N)
    plan = _engine.FirPlan(h)
    n_x = len(x)
    y = _run_fir(plan, x) if n_x else x.copy()
    if mode != 1:
        return y
    Nx = n_x + P - 1
    Nframe = int(np.ceil(Nx / float(L)))
    xp = np.hstack((np.zeros(P - 1), x, np.zeros(Nframe * L - Nx + (N - L))))
    y_mat = np.zeros((Nframe, int(Nframe * N)))
    for k in range(Nframe):
        xk = xp[k * L:k * L + N]
        if len(xk) < N:
            xk = np.hstack((xk, np.zeros(N - len(xk))))
        # circular convolution of the frame == linear filter whose history is the frame's own tail
        hist = xk[N - (P - 1):] if P > 1 else None
        y_mat[k, k * L:k * L + N] = _run_fir(plan, xk, hist)
    return y, y_mat[:, P - 1:Nx]


def oa_filter(x, h, N, mode=0):
    """
    Overlap and add transform domain FIR filtering.

    Returns ``y`` (``len(x)`` samples of ``h * x``) and, for ``mode == 1``, the diagnostic matrix
    whose row ``k`` holds the ``N``-sample response of frame ``k`` (``L = N - len(h) + 1`` inputs)
    at its position in the stream (reference: sigsys.py:543-598).
    """
    x, h, P, L = _block_filter_args(x, h, N)
    N = int(N)
    plan = _engine.FirPlan(h)
    Nx = len(x)
    y = _run_fir(plan, x) if Nx else x.copy()
    if mode != 1:
        return y
    Nframe = int(np.ceil(Nx / float(L)))
    xp = np.hstack((x, np.zeros(Nframe * L - Nx)))
    y_mat = np.zeros((Nframe, Nframe * N))
    for k in range(Nframe):
        xk = np.hstack((xp[k * L:(k + 1) * L], np.zeros(N - L)))
        y_mat[k, k * L:k * L + N] = _run_fir(plan, xk)
    return y, y_mat[:, 0:Nx]


# ---------------------------------------------------------------------------------------------
# Ten-band graphic equaliser: the reference's other in-package caller of the biquad-cascade
# arithmetic (sigsys.py:96-141, 202-252; literal known-answer vector tests/test_sigsys.py:28-34).
# The reference chains ten ``lfilter(B_k, A_k, .)`` calls; the same ten transposed-direct-form-II
# sections run as ONE parallel-prefix cascade launch here.

def peaking(GdB, fc, Q=3.5, fs=44100.):
    """
    A second-order peaking filter having GdB gain at fc and approximately 0 dB otherwise
    (reference: sigsys.py:202-252).  Returns ``b, a`` (three coefficients each, ``a[0] == 1``).
    """
    mu = 10 ** (GdB / 20.)
    kq = 4 / (1 + mu) * np.tan(2 * np.pi * fc / fs / (2 * Q))
    Cpk = (1 + kq * mu) / (1 + kq)
    c0 = -2 * np.cos(2 * np.pi * fc / fs)
    b = Cpk * np.array([1, c0 / (1 + kq * mu), (1 - kq * mu) / (1 + kq * mu)])
    a = np.array([1, c0 / (1 + kq), (1 - kq) / (1 + kq)])
    return b, a


def ten_band_eq_filt(x, GdB, Q=3.5):
    """
    Filter the input signal x with a ten-band equalizer having octave gain values in ndarray GdB
    (band centres 31.25 Hz ... 16 kHz at fs = 44.1 kHz; reference: sigsys.py:96-141).
    """
    NB = len(GdB)
    if not NB == 10:
        raise ValueError("GdB length not equal to ten")
    Fc = 31.25 * 2 ** np.arange(NB)
    sos = np.array([np.hstack(peaking(GdB[k], Fc[k], Q)) for k in range(NB)])
    from .multirate_helper import _SosRunner
    return _SosRunner(sos).filter(x)


def cic(m, k):
    """
    A functional form implementation of a cascade of integrator comb (CIC) filters: the taps of
    ``k`` cascaded length-``m`` boxcars, unity gain at DC (reference: sigsys.py:62-93).  Host-side
    design; feed the result to ``multirate_FIR`` to run it.
    """
    b = np.ones(m)
    for _ in range(1, k):
        b = np.convolve(b, np.ones(m))
    return b / np.sum(b)
