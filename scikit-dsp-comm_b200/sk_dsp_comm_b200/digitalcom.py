"""B200 drop-in for the pulse-shaping transmitters of ``sk_dsp_comm.digitalcom``
(reference: src/sk_dsp_comm/digitalcom.py:418-492, 585-721, 849-1048,
1584-1826).

Same names, arguments, defaults, return tuples and error behaviour as the reference.  Symbols
are drawn on the host from the legacy ``np.random`` global generator exactly as the reference
draws them (so ``np.random.seed(k)`` reproduces the reference's waveforms, which is what
tests/test_digitalcom.py pins); the waveform itself -- ``lfilter(b, 1, upsample(symbols, ns))`` in
the reference -- is one launch of the fused zero-stuff + FIR CUDA kernel (``_pulse.shape_symbols``).
Arrays come back as numpy, like the reference's.

Receivers, error counters, OFDM and the plot helpers of the reference module are out of scope
(SURVEY.md section 2).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _pulse
from ._pulse import rc_imp, sqrt_rc_imp                       # noqa: F401  (public, like the reference)
from .sigsys import upsample, downsample, nrz_bits, nrz_bits2, pn_gen, m_seq   # noqa: F401

def _prefix_xor(v):
    r = 0
    while v:
        r ^= v
        v >>= 1
    return r


# the reference's ``bin2gray`` LUTs (digitalcom.py:1611-1614, 1766-1771) are the Gray -> binary
# map, i.e. the running XOR of the word's bits from the msb down; one table per word width
_REF_GRAY = {n: np.array([_prefix_xor(v) for v in range(1 << n)]) for n in range(1, 6)}


def _words(bits, width):
    """(n, width) 0/1 rows, msb first -> integer per row."""
    w = 2 ** np.arange(width - 1, -1, -1)
    return np.asarray(bits).reshape(-1, width).astype(np.int64) @ w


def _shaping_pulse(pulse, ns, alpha, m, msg):
    b = _pulse.pulse_taps(pulse, ns, alpha, m)
    if b is None:
        raise ValueError(msg)
    return b


def qam_bb(n_symb, ns, mod='16qam', pulse='rect', alpha=0.35):
    """
    A complex baseband transmitter (reference: digitalcom.py:418-492).

    Returns ``x`` (complex baseband waveform), ``b`` (shaping filter scaled to unity DC gain) and
    ``tx_data = xI + 1j*xQ`` (the odd-integer symbol coordinates before normalisation).
    """
    b = _shaping_pulse(pulse, ns, alpha, 6, 'pulse shape must be src, rc, or rect')
    levels = {'qpsk': 2, '16qam': 4, '64qam': 8, '256qam': 16}.get(mod.lower())
    if levels is None:
        raise ValueError('Unknown mod_type')
    xI = 2 * np.random.randint(0, levels, n_symb) - (levels - 1)
    xQ = 2 * np.random.randint(0, levels, n_symb) - (levels - 1)
    symb = xI + 1j * xQ
    if levels > 2:
        symb = symb / (levels - 1)
    x = _pulse.shape_symbols(symb, b, ns)
    return x, b / sum(b), xI + 1j * xQ


def mpsk_bb(n_symb, ns, mod, pulse='rect', alpha=0.25, m=6):
    """
    Generate a complex baseband MPSK signal with pulse shaping (reference:
    digitalcom.py:613-669).  Returns ``x``, ``b / ns`` and the symbol indices.
    """
    data = np.random.randint(0, mod, n_symb)
    xs = np.exp(1j * 2 * np.pi / mod * data)
    b = _shaping_pulse(pulse, ns, alpha, m, 'pulse type must be rec, rc, or src')
    x = _pulse.shape_symbols(xs, b, ns)
    if mod == 4:
        x = x * np.exp(1j * np.pi / 4)          # QPSK points in the quadrants
    return x, b / float(ns), data


def qpsk_bb(n_symb, ns, lfsr_len=5, pulse='src', alpha=0.25, m=6):
    """
    QPSK from two NRZ streams: PN data when ``lfsr_len > 0``, random bits otherwise
    (reference: digitalcom.py:704-720).
    """
    if lfsr_len > 0:
        data = pn_gen(2 * n_symb, lfsr_len)
        xI, b = nrz_bits2(data[0::2], ns, pulse, alpha, m)
        xQ, b = nrz_bits2(data[1::2], ns, pulse, alpha, m)
    else:
        data = np.zeros(2 * n_symb)
        xI, b, data[0::2] = nrz_bits(n_symb, ns, pulse, alpha, m)
        xQ, b, data[1::2] = nrz_bits(n_symb, ns, pulse, alpha, m)
    return (xI + 1j * xQ) / np.sqrt(2.), b, data


def bpsk_tx(n_bits, ns, ach_fc=2.0, ach_lvl_dB=-100, pulse='rect', alpha=0.25, m=6):
    """
    BPSK transmitter with two adjacent-channel interferers at ``+/- ach_fc/ns`` and level
    ``ach_lvl_dB`` (reference: digitalcom.py:849-890).
    """
    x0, b, data0 = nrz_bits(n_bits, ns, pulse, alpha, m)
    x1p, b, _ = nrz_bits(n_bits, ns, pulse, alpha, m)
    x1m, b, _ = nrz_bits(n_bits, ns, pulse, alpha, m)
    rot = np.exp(1j * (2 * np.pi * ach_fc / float(ns) * np.arange(len(x0))))
    ach_lvl = 10 ** (ach_lvl_dB / 20.)
    return x0 + ach_lvl * (x1p * rot + x1m * np.conj(rot)), b, data0


def rz_bits(n_bits, ns, pulse='rect', alpha=0.25, m=6):
    """
    Generate return-to-zero (RZ) data bits with pulse shaping: 0/1 amplitudes (reference:
    digitalcom.py:998-1048).  An unknown pulse name only warns in the reference and then fails on
    the unbound pulse; here it warns and raises ``UnboundLocalError`` the same way.
    """
    data = np.random.randint(0, 2, n_bits)
    b = _pulse.pulse_taps(pulse, ns, alpha, m)
    if b is None:
        warnings.warn('pulse type must be rec, rc, or src')
        raise UnboundLocalError("cannot access local variable 'b' where it is not associated with a value")
    x = _pulse.shape_symbols(data.astype(np.float64), b, ns)
    return x, b / float(ns), data


def gmsk_bb(n_bits, ns, msk=0, bt=0.35):
    """
    MSK/GMSK complex baseband modulation (reference: digitalcom.py:585-610): rectangular NRZ,
    optional Gaussian pre-modulation filter (``msk != 0``), phase accumulation.  The two FIR
    passes run on the GPU; the cumulative phase and ``exp`` stay in numpy as in the reference.
    """
    x, b, data = nrz_bits(n_bits, ns)
    span = 4
    n = np.arange(-span * ns, span * ns + 1)
    p = np.exp(-2 * np.pi ** 2 * bt ** 2 / np.log(2) * (n / float(ns)) ** 2)
    p = p / np.sum(p)
    if msk != 0:
        x = _pulse.fir_apply(p, x)
    return np.exp(1j * np.pi / 2 * np.cumsum(x) / ns), data


def _gray_bits(n_symb, ext_data, bits_per_symbol):
    if n_symb is None:
        n_symb = int(np.floor(len(ext_data) / float(bits_per_symbol)))
        return n_symb, ext_data[:n_symb * bits_per_symbol]
    return n_symb, np.random.randint(0, 2, size=bits_per_symbol * n_symb)


def _finish_gray(x_IQ, ns, pulse, alpha, m_span, scale, data):
    if ns > 1:
        b = _shaping_pulse(pulse, ns, alpha, m_span, 'pulse shape must be src, rc, or rect')
        x = _pulse.shape_symbols(x_IQ, b, ns)
        return x / scale, b / sum(b), data
    return x_IQ / scale, 1, data


def qam_gray_encode_bb(n_symb, ns, mod=4, pulse='rect', alpha=0.35, m_span=6, ext_data=None):
    """
    A Gray code mapped QAM complex baseband transmitter, ``mod`` in {2, 4, 16, 64, 256}
    (reference: digitalcom.py:1584-1681).  The serial bit stream is
    ``[Ibits, Qbits, Ibits, Qbits, ...]`` msb first; returns ``x / (sqrt(mod)-1)``, the
    unity-DC-gain pulse and the bits.
    """
    bps = int(np.log2(mod))
    n_symb, data = _gray_bits(n_symb, ext_data, bps)
    x_m = np.sqrt(mod) - 1
    if mod == 2:
        x_IQ = 2 * data - 1
        x_m = 1
    elif mod in (4, 16, 64, 256):
        half = bps // 2
        idx = _words(data, half).reshape(n_symb, 2)             # columns: I word, Q word
        lut = _REF_GRAY[half]
        x_IQ = (2 * lut[idx[:, 0]] - x_m) + 1j * (2 * lut[idx[:, 1]] - x_m)
    else:
        raise ValueError('M must be 2, 4, 16, 64, 256')
    return _finish_gray(x_IQ, ns, pulse, alpha, m_span, x_m, data)


def mpsk_gray_encode_bb(n_symb, ns, mod=4, pulse='rect', alpha=0.35, m_span=6, ext_data=None):
    """
    A Gray code mapped MPSK complex baseband transmitter, ``mod`` in {2, 4, 8, 16, 32}
    (reference: digitalcom.py:1742-1826); QPSK is rotated by ``pi/4`` into the quadrants.
    """
    bps = int(np.log2(mod))
    n_symb, data = _gray_bits(n_symb, ext_data, bps)
    if mod == 2:
        x_IQ = 2 * data - 1
    elif mod in (4, 8, 16, 32):
        phase = 2 * np.pi * _REF_GRAY[bps][_words(data, bps)] / mod
        if mod == 4:
            phase = phase + np.pi / mod
        x_IQ = np.exp(1j * phase)
    else:
        raise ValueError('M must be 2, 4, 8, 16, or 32')
    return _finish_gray(x_IQ, ns, pulse, alpha, m_span, 1, data)
