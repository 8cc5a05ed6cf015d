"""Coefficient interchange with the reference's C-header formats (SURVEY.md 8f rank 4;
reference: src/sk_dsp_comm/coeff2header.py:42-155).

Writers produce byte-for-byte the files the reference writes (``fir_header``,
``fir_fix_header``, ``iir_sos_header`` -- the CMSIS-DSP ``b0,b1,b2,-a1,-a2`` stage layout), so a
filter that runs on the GPU engine can be dropped into the same embedded projects.  The
readers are the way back: taps / an ``(n_sections, 6)`` scipy-layout ``sos`` array ready for
``multirate_FIR`` / ``multirate_IIR``.  Pure host code; nothing here touches the device.

``freqz_resp_list`` (plot) and ``ca_code_header`` (GPS code tables) are out of scope.
"""
from __future__ import annotations

import re

import numpy as np

_BAR_FIR = '/************************************************************************/\n'
_BAR_SOS = '/*********************************************************/\n'


def _array_body(cells, per_line, indent):
    """cells joined with ',', a line break + indent after every ``per_line`` cells."""
    rows = [','.join(cells[i:i + per_line]) for i in range(0, len(cells), per_line)]
    return (',\n' + ' ' * indent).join(rows)


def _write_fir(fname_out, ctype, cells, per_line, indent):
    text = ('//define a FIR coefficient Array\n\n'
            '#include <stdint.h>\n\n'
            '#ifndef M_FIR\n'
            '#define M_FIR %d\n'
            '#endif\n' % len(cells))
    text += _BAR_FIR
    text += '/*                         FIR Filter Coefficients                      */\n'
    text += '%s h_FIR[M_FIR] = {' % ctype + _array_body(cells, per_line, indent) + '};\n'
    text += _BAR_FIR
    with open(fname_out, 'wt') as f:
        f.write(text)


def fir_header(fname_out, h):
    """
    Write FIR Filter Header Files: ``float32_t h_FIR[M_FIR]``, three ``%15.12f`` values per
    line (reference: coeff2header.py:42-74).
    """
    _write_fir(fname_out, 'float32_t', ['%15.12f' % v for v in h], 3, 26)


def fir_fix_header(fname_out, h):
    """
    Write FIR Fixed-Point Filter Header Files: Q15 ``int16_t h_FIR[M_FIR]``, eight values per
    line (reference: coeff2header.py:77-110).
    """
    hq = np.int16(np.rint(np.asarray(h) * 2 ** 15))
    _write_fir(fname_out, 'int16_t', ['%5d' % v for v in hq], 8, 24)


def iir_sos_header(fname_out, SOS_mat):
    """
    Write IIR SOS Header Files, compatible with the CMSIS-DSP IIR direct form II functions:
    ``ba_coeff[5*STAGES]`` holding ``b0,b1,b2,-a1,-a2`` per stage (reference:
    coeff2header.py:113-155).
    """
    SOS_mat = np.asarray(SOS_mat)
    Ns, Mcol = SOS_mat.shape
    stages = ['    %+-13e, %+-13e, %+-13e,\n    %+-13e, %+-13e'
              % (r[0], r[1], r[2], -r[4], -r[5]) for r in SOS_mat]
    text = ('//define a IIR SOS CMSIS-DSP coefficient array\n\n'
            '#include <stdint.h>\n\n'
            '#ifndef STAGES\n'
            '#define STAGES %d\n'
            '#endif\n' % Ns)
    text += _BAR_SOS
    text += '/*                     IIR SOS Filter Coefficients       */\n'
    text += 'float32_t ba_coeff[%d] = { //b0,b1,b2,a1,a2,... by stage\n' % (5 * Ns)
    text += ',\n'.join(stages) + '\n};\n'
    text += _BAR_SOS
    with open(fname_out, 'wt') as f:
        f.write(text)


# ------------------------------------------------------------------------------------ readers

_NUM = r'[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?'


def _braces(text, what):
    m = re.search(r'\{(.*?)\}', text, re.S)
    if m is None:
        raise ValueError('no %s initialiser found' % what)
    body = re.sub(r'//[^\n]*', '', m.group(1))              # the SOS header has a trailing comment
    return np.array([float(v) for v in re.findall(_NUM, body)])


def _declared(text, macro):
    m = re.search(r'#define\s+%s\s+(\d+)' % macro, text)
    return int(m.group(1)) if m else None


def read_fir_header(fname):
    """Taps from a ``fir_header`` / ``fir_fix_header`` file (Q15 integers are scaled by 2**-15).

    Returns a float64 array; raises ``ValueError`` if the count disagrees with ``M_FIR``.
    """
    with open(fname, 'rt') as f:
        text = f.read()
    h = _braces(text, 'h_FIR')
    n = _declared(text, 'M_FIR')
    if n is not None and n != len(h):
        raise ValueError('M_FIR says %d taps, found %d' % (n, len(h)))
    if re.search(r'int16_t\s+h_FIR', text):
        h = h / 2.0 ** 15
    return h


def read_iir_sos_header(fname):
    """``(STAGES, 6)`` scipy-layout ``sos`` (``[b0 b1 b2 1 a1 a2]``) from an ``iir_sos_header``
    file (stored CMSIS style as ``b0,b1,b2,-a1,-a2``)."""
    with open(fname, 'rt') as f:
        text = f.read()
    v = _braces(text, 'ba_coeff')
    n = _declared(text, 'STAGES')
    if len(v) % 5 or (n is not None and 5 * n != len(v)):
        raise ValueError('expected 5 coefficients per stage, found %d values for STAGES=%s' % (len(v), n))
    v = v.reshape(-1, 5)
    sos = np.ones((v.shape[0], 6))
    sos[:, :3] = v[:, :3]
    sos[:, 4:] = -v[:, 3:]
    return sos
