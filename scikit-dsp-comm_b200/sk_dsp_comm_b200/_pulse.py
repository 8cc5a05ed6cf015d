"""Shared host logic of the pulse-shaping transmitters.

The reference builds every digital-modulation waveform the same way
(digitalcom.py:488,666,1047,1676,1821; sigsys.py:2149,2201): one symbol per ``ns`` samples,
zero-stuffed, then ``signal.lfilter(b, 1, .)`` with a rect / RC / SRC pulse.  That is the
``.up`` polyphase path without the ``L`` gain, so all of them funnel into ``shape_symbols``
below, which runs the fused zero-stuff + FIR kernel (``b200dsp_fir_up``) -- the zero-stuffed
intermediate the reference materialises never exists.

What stays on the host is what the reference also does in numpy before the filter: drawing
the symbols from ``np.random`` (the seeded golden vectors of tests/test_digitalcom.py:64-310
depend on the legacy global generator being consumed in the same order), the Gray maps, and the
pulse designs.
"""
from __future__ import annotations

import numpy as np

# test seam, same idea as sharded.ShardedFIR(compute=...): the gloo/CPU tests replace the kernel
# call with the checker so the host logic can be exercised without a GPU.  None = CUDA kernels.
_compute_up = None
_compute_filter = None

_plan_cache = {}


def _plan_for(b, scale):
    from . import _engine
    key = (b.tobytes(), scale)
    p = _plan_cache.get(key)
    if p is None:
        if len(_plan_cache) > 64:
            _plan_cache.clear()
        p = _engine.FirPlan(b / scale if scale != 1 else b)
        _plan_cache[key] = p
    return p


def shape_symbols(symbols, b, ns):
    """``lfilter(b, 1, upsample(symbols, int(ns)))`` -- one pulse ``b`` per symbol, no rate gain.

    The kernel computes ``L * sum_q b'[L q + r] x[m - q]`` (the ``.up`` contract,
    multirate_helper.py:116-117); the plan holds ``b' = b / L`` so the product is the plain
    pulse train (exact for power-of-two ``L``, one rounding of the tap otherwise).
    """
    L = int(ns)
    if L < 1:
        raise ValueError("negative dimensions are not allowed")
    sym = np.ascontiguousarray(symbols)
    sym = sym.astype(np.complex128 if sym.dtype.kind == "c" else np.float64, copy=False)
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).ravel())
    if _compute_up is not None:
        return _compute_up(b, sym, L)
    if sym.size == 0:
        return sym.copy()
    from . import _engine
    from ._io import Staged
    st = Staged(sym)
    if L == 1:
        return st.finish(_engine.fir_filter(_plan_for(b, 1), st.tensor))
    return st.finish(_engine.fir_up(_plan_for(b, L), st.tensor, L))


def fir_apply(b, x):
    """``lfilter(b, 1, x)`` on a host array (gmsk_bb's Gaussian pre-modulation filter)."""
    x = np.ascontiguousarray(x)
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).ravel())
    if _compute_filter is not None:
        return _compute_filter(b, x)
    from . import _engine
    from ._io import Staged
    st = Staged(x)
    return st.finish(_engine.fir_filter(_plan_for(b, 1), st.tensor))


# ------------------------------------------------------------------ pulse designs (host, numpy)

def rc_imp(ns, alpha, m=6):
    """
    A truncated raised cosine pulse used in digital communications.

    ``ns`` samples per symbol, excess bandwidth ``alpha`` on (0, 1), one-sided truncation of
    ``m`` symbols: ``2*m*ns + 1`` taps (reference: digitalcom.py:893-941, sigsys.py:1847-1891).
    Evaluated vectorised, with the reference's association of the products so that the taps
    (and the position of the removable singularity at ``|n| = ns/(2 alpha)``) are identical.
    """
    n = np.arange(-m * ns, m * ns + 1)
    ns = ns * 1.0
    den = 1 - 4 * (alpha * n / ns) ** 2
    sing = den == 0
    safe = np.where(sing, 1.0, den)
    b = np.sinc(n / ns) * np.cos(np.pi * alpha * n / ns) / safe
    return np.where(sing, np.pi / 4 * np.sinc(1 / (2. * alpha)), b)


def sqrt_rc_imp(ns, alpha, m=6):
    """
    A truncated square root raised cosine pulse (reference: digitalcom.py:944-995,
    sigsys.py:1894-1944): ``2*m*ns + 1`` taps, the matched-filter partner of ``rc_imp``.
    """
    n = np.arange(-m * ns, m * ns + 1)
    ns = ns * 1.0
    a = alpha
    den = 1 - 16 * a ** 2 * (n / ns) ** 2
    sing = np.abs(den) <= np.finfo(np.float32).eps / 2
    safe = np.where(sing, 1.0, den)
    b = 4 * a / (np.pi * safe)
    b = b * (np.cos((1 + a) * np.pi * n / ns) + np.sinc((1 - a) * n / ns) * (1 - a) * np.pi / (4. * a))
    edge = 1 / 2. * ((1 + a) * np.sin((1 + a) * np.pi / (4. * a)) - (1 - a) * np.cos((1 - a) * np.pi / (4. * a))
                     + (4 * a) / np.pi * np.sin((1 - a) * np.pi / (4. * a)))
    return np.where(sing, edge, b)


def pulse_taps(pulse, ns, alpha, m):
    """'rect' | 'rc' | 'src' -> taps, or None for an unknown name (callers own the error)."""
    kind = pulse.lower()
    if kind == 'rect':
        return np.ones(int(ns))
    if kind == 'rc':
        return rc_imp(ns, alpha, m)
    if kind == 'src':
        return sqrt_rc_imp(ns, alpha, m)
    return None


# ------------------------------------------------------------------ PN sequences (host)

# feedback positions k (1 <= k < m) of the reference's generator polynomials (sigsys.py:2004-2034)
_LFSR_FEEDBACK = {2: (1,), 3: (2,), 4: (3,), 5: (3,), 6: (5,), 7: (4,), 8: (4, 5, 6), 9: (5,),
                  10: (7,), 11: (9,), 12: (6, 8, 11), 13: (9, 10, 12), 14: (4, 8, 13), 15: (11,),
                  16: (2, 11, 15)}


def m_seq(m):
    """
    Generate an m-sequence ndarray using an all-ones initialization (reference:
    sigsys.py:1983-2050); one period, ``2**m - 1`` values of 0./1.
    """
    fb = _LFSR_FEEDBACK.get(m)
    if fb is None:
        raise ValueError('Invalid length specified')
    q = 2 ** m - 1
    c = np.zeros(q)
    reg = (1 << m) - 1                       # bit i = shift-register stage i, all ones
    top = m - 1
    for i in range(q):
        out = (reg >> top) & 1
        c[i] = out
        x = 0
        for k in fb:                          # each active tap contributes (last ^ stage[m-1-k])
            x ^= out ^ ((reg >> (top - k)) & 1)
        reg = ((reg << 1) & ((1 << m) - 1)) | x
    return c


def pn_gen(n_bits, m=5):
    """
    Maximal length sequence signal generator: ``n_bits`` 0./1. values from the period
    ``2**m - 1`` m-sequence (reference: sigsys.py:1947-1980).
    """
    c = m_seq(m)
    reps = int(np.ceil(n_bits / float(len(c))))
    return np.tile(c, max(reps, 0))[:n_bits].copy()
