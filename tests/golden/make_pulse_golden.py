#!/usr/bin/env python
"""Golden fixtures for the pulse-shaping transmitters, produced by running
the UNMODIFIED reference in the dev container (same matplotlib stub as make_golden.py).

    python tests/golden/make_pulse_golden.py      # writes tests/golden/pulse_cases.npz

Every case seeds the legacy ``np.random`` generator (the reference draws its symbols from it,
tests/test_digitalcom.py:64-65), calls one reference function and stores every returned array.
``table`` inside the .npz is the JSON list of cases: ``module`` ('dc' = digitalcom, 'ss' = sigsys),
function, args, kwargs, seed, number of outputs.  ``mseq_<m>`` hold ``sigsys.m_seq(m)``.
"""
import json
import os
import sys
import types

import numpy as np

for m in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "matplotlib.mlab"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")

import sk_dsp_comm.digitalcom as dc     # noqa: E402
import sk_dsp_comm.sigsys as ss         # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = []


def case(module, func, *args, seed=100, **kwargs):
    CASES.append(dict(module=module, func=func, args=list(args), kwargs=kwargs, seed=seed))


# the reference's own seeded tests (tests/test_digitalcom.py:64-310): same calls, same seed
for pulse in ("src", "rc", "rect"):
    case("dc", "qam_bb", 10, 2, mod="qpsk", pulse=pulse)
for mod in ("16qam", "64qam", "256qam"):
    case("dc", "qam_bb", 10, 2, mod=mod, pulse="rect")
case("dc", "mpsk_bb", 500, 10, 8, "rect", 0.35)
case("dc", "mpsk_bb", 500, 10, 8, "rc", 0.35)
case("dc", "mpsk_bb", 500, 10, 8, "src", 0.35)
case("dc", "gmsk_bb", 10, 2)
# wider coverage: odd samples/symbol (tap pre-scale is inexact for non powers of two), long pulses
case("dc", "qam_bb", 300, 7, mod="64qam", pulse="src", alpha=0.2, seed=7)
case("dc", "qam_bb", 4096, 4, mod="256qam", pulse="rc", alpha=0.5, seed=8)
case("dc", "mpsk_bb", 257, 3, 4, "src", 0.25, 8, seed=9)
case("dc", "mpsk_bb", 64, 16, 16, "rect", seed=10)
case("dc", "qpsk_bb", 200, 8, seed=11)
case("dc", "qpsk_bb", 200, 8, 0, "rc", 0.35, 4, seed=12)
case("dc", "qpsk_bb", 100, 5, 7, "rect", seed=13)
case("dc", "bpsk_tx", 10, 10, pulse="src")
case("dc", "bpsk_tx", 300, 6, 1.5, -20, "rect", seed=14)
case("dc", "rz_bits", 100, 10, seed=15)
case("dc", "rz_bits", 150, 4, "src", 0.3, 5, seed=16)
case("dc", "rz_bits", 90, 9, "rc", seed=17)
case("dc", "gmsk_bb", 400, 8, 1, 0.3, seed=18)
case("dc", "gmsk_bb", 400, 8, 0, seed=19)
for mod in (2, 4, 16, 64, 256):
    case("dc", "qam_gray_encode_bb", 120, 4, mod, "src", seed=20 + mod)
case("dc", "qam_gray_encode_bb", 100, 1, 16, seed=30)
case("dc", "qam_gray_encode_bb", 333, 5, 64, "rc", 0.25, 4, seed=31)
case("dc", "qam_gray_encode_bb", 64, 8, 16, "rect", seed=32)
for mod in (2, 4, 8, 16, 32):
    case("dc", "mpsk_gray_encode_bb", 120, 4, mod, "src", seed=40 + mod)
case("dc", "mpsk_gray_encode_bb", 100, 1, 8, seed=50)
case("dc", "mpsk_gray_encode_bb", 200, 6, 8, "rect", seed=51)
case("ss", "nrz_bits", 100, 10, seed=60)
case("ss", "nrz_bits", 1000, 4, "src", 0.35, seed=61)
case("ss", "nrz_bits", 77, 3, "rc", 0.5, 3, seed=62)
case("ss", "bpsk_tx", 200, 8, 2.0, -30, "src", seed=63)
case("ss", "bpsk_tx", 50, 10, seed=64)
case("ss", "pn_gen", 100, 5)
case("ss", "pn_gen", 1000, 7)
case("ss", "pn_gen", 3, 4)


def main():
    out = {}
    for i, c in enumerate(CASES):
        np.random.seed(c["seed"])
        fn = getattr(dc if c["module"] == "dc" else ss, c["func"])
        res = fn(*c["args"], **c["kwargs"])
        if not isinstance(res, tuple):
            res = (res,)
        c["nout"] = len(res)
        for j, r in enumerate(res):
            out["c%03d_o%d" % (i, j)] = np.asarray(r)
    # externally supplied bits (ext_data) and user data (nrz_bits2) use the reference's PN source
    bits = ss.pn_gen(600, 9).astype(int)
    out["ext_bits"] = bits
    for j, r in enumerate(dc.qam_gray_encode_bb(None, 4, 64, "src", ext_data=bits)):
        out["ext_qam_o%d" % j] = np.asarray(r)
    for j, r in enumerate(dc.mpsk_gray_encode_bb(None, 3, 8, "rc", ext_data=bits)):
        out["ext_mpsk_o%d" % j] = np.asarray(r)
    for j, r in enumerate(ss.nrz_bits2(ss.m_seq(5), 10)):
        out["nrz2_mseq5_o%d" % j] = np.asarray(r)
    for j, r in enumerate(ss.nrz_bits2(bits, 6, "src", 0.4, 5)):
        out["nrz2_src_o%d" % j] = np.asarray(r)
    for m in range(2, 17):
        out["mseq_%d" % m] = ss.m_seq(m).astype(np.uint8)      # 0/1 values; the function returns float64
    # pulse designs on their own (several ns / alpha / span; includes the singular points)
    for k, (ns, a, m) in enumerate([(10, 0.35, 6), (2, 0.25, 6), (4, 0.5, 4), (7, 0.2, 8), (8, 0.125, 6), (3, 1.0, 2)]):
        out["rc_%d" % k] = dc.rc_imp(ns, a, m)
        out["src_%d" % k] = dc.sqrt_rc_imp(ns, a, m)
        out["pulse_args_%d" % k] = np.array([ns, a, m])
    out["table"] = np.array(json.dumps(CASES))
    np.savez_compressed(os.path.join(OUT, "pulse_cases.npz"), **out)
    print("wrote", len(CASES), "cases,", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
