#!/usr/bin/env python
"""Header-file fixtures for the coefficient interchange row: the text the
UNMODIFIED reference's coeff2header writers produce for the fixture filters.

    python tests/golden/make_header_golden.py     # writes tests/golden/headers.npz

Also checks that the reference still reproduces its own committed fixture
(/root/reference/tests/sig_mean_var.h, tests/test_coeff2header.py:28-45) for the same input, so the
text stored here is pinned to that known-answer file without copying it.
"""
import os
import sys
import tempfile
import types

import numpy as np

for m in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "matplotlib.mlab"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")
import sk_dsp_comm.coeff2header as c2h      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def text_of(fn, arg):
    with tempfile.NamedTemporaryFile(suffix=".h") as t:
        fn(t.name, arg)
        return open(t.name, "rt").read()


def main():
    filt = np.load(os.path.join(OUT, "filters.npz"))
    out = {}
    n = np.arange(0, 501)
    x = 3 * np.cos(2 * np.pi * 1000 / 48000 * n) + 2 * np.sin(2 * np.pi * 400 / 48000 * n)
    t = text_of(c2h.fir_header, x)
    assert t == open("/root/reference/tests/sig_mean_var.h").read()
    out["sig_mean_var"] = np.array(t)
    for k in ("b101", "b256", "b7", "b1", "b33_remez_bpf"):
        out["fir_" + k] = np.array(text_of(c2h.fir_header, filt[k]))
        out["fix_" + k] = np.array(text_of(c2h.fir_fix_header, filt[k]))
    for k in ("sos6", "sos_butter5", "sos_sharp_lpf", "sos_tenband"):
        out["sos_" + k] = np.array(text_of(c2h.iir_sos_header, filt[k]))
    out["sos_single"] = np.array(text_of(c2h.iir_sos_header, filt["sos6"][:1]))
    for m in (2, 8, 9, 16, 17):          # line-wrap boundaries of the 3- and 8-per-line layouts
        out["fir_len%d" % m] = np.array(text_of(c2h.fir_header, filt["b101"][:m]))
        out["fix_len%d" % m] = np.array(text_of(c2h.fir_fix_header, filt["b101"][40:40 + m]))
    np.savez_compressed(os.path.join(OUT, "headers.npz"), **out)
    print("wrote", len(out), "header texts")


if __name__ == "__main__":
    main()
