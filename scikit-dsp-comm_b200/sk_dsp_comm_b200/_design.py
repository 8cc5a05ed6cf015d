"""Host-side digital low-pass designs needed by ``rate_change`` and ``interp24`` / ``deci24``
(reference: ``scipy.signal.butter`` / ``cheby1`` called at src/sk_dsp_comm/multirate_helper.py:61-65 and
src/sk_dsp_comm/sigsys.py:2971-3025).  O(order) numpy work: analog prototype poles -> frequency
pre-warp -> bilinear transform -> second-order sections for the CUDA cascade, plus the (b, a)
polynomials the reference objects expose.  The product package deliberately does not import scipy.
"""
from __future__ import annotations

import numpy as np


def _butter_proto(N):
    k = np.arange(-N + 1, N, 2)
    return -np.exp(1j * np.pi * k / (2 * N)), 1.0            # poles, gain


def _cheby1_proto(N, rp):
    eps = np.sqrt(10.0 ** (0.1 * rp) - 1.0)
    mu = np.arcsinh(1.0 / eps) / N
    m = np.arange(-N + 1, N, 2)
    theta = np.pi * m / (2 * N)
    p = -np.sinh(mu + 1j * theta)
    k = np.prod(-p).real
    if N % 2 == 0:
        k = k / np.sqrt(1.0 + eps * eps)
    return p, k


def _lowpass_zpk(N, Wn, proto):
    """Digital low-pass with normalised cutoff Wn (1 = Nyquist): zeros, poles, gain."""
    if not 0.0 < Wn < 1.0:
        raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
    p, k = proto
    fs = 2.0
    warped = 2.0 * fs * np.tan(np.pi * Wn / fs)
    p = warped * p                                           # lp2lp
    k = k * warped ** N
    fs2 = 2.0 * fs
    pd = (fs2 + p) / (fs2 - p)                               # bilinear
    kd = k * np.real(1.0 / np.prod(fs2 - p))
    zd = -np.ones(N)
    return zd, pd, kd


def _zpk2sos_lowpass(N, pd, kd):
    """All zeros at z = -1: pair complex-conjugate poles; a real pole (odd N) gets a first-order section.
    Sections are ordered by increasing pole radius (least resonant first), the gain goes to section 0."""
    pos = sorted([p for p in pd if p.imag > 1e-12], key=lambda z: abs(z))
    real = sorted([p.real for p in pd if abs(p.imag) <= 1e-12], key=abs)
    sos = []
    while len(real) >= 2:                                   # two real poles share a section
        r1, r2 = real.pop(0), real.pop(0)
        sos.append([1.0, 2.0, 1.0, 1.0, -(r1 + r2), r1 * r2])
    if real:
        r = real.pop(0)
        sos.append([1.0, 1.0, 0.0, 1.0, -r, 0.0])
    for p in pos:
        sos.append([1.0, 2.0, 1.0, 1.0, -2.0 * p.real, abs(p) ** 2])
    sos = np.array(sos, dtype=np.float64)
    sos[0, :3] *= kd
    assert sos.shape[0] == (N + 1) // 2
    return sos


def _ba(zd, pd, kd):
    b = kd * np.real(np.poly(zd))
    a = np.real(np.poly(pd))
    return b, a


def butter(N, Wn):
    """(b, a, sos) of the N-th order Butterworth low-pass ``scipy.signal.butter(N, Wn)``."""
    zd, pd, kd = _lowpass_zpk(N, Wn, _butter_proto(N))
    b, a = _ba(zd, pd, kd)
    return b, a, _zpk2sos_lowpass(N, pd, kd)


def cheby1(N, rp, Wn):
    """(b, a, sos) of the Chebyshev-I low-pass ``scipy.signal.cheby1(N, rp, Wn)``."""
    zd, pd, kd = _lowpass_zpk(N, Wn, _cheby1_proto(N, rp))
    b, a = _ba(zd, pd, kd)
    return b, a, _zpk2sos_lowpass(N, pd, kd)
