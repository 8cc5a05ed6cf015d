#!/usr/bin/env python
"""Fixtures for SURVEY.md 8f rank 3 (block filtering and streaming state), produced in the dev
container by the UNMODIFIED reference (``sigsys.os_filter`` / ``oa_filter``, sigsys.py:482-598)
and by the scipy calls the reference's filter classes sit on (``signal.lfilter(b,[1],x,zi=)`` /
``signal.sosfilt(sos,x,zi=)``, multirate_helper.py:108,173) with their ``zi`` argument.

    python tests/golden/make_stream_golden.py     # writes tests/golden/stream_cases.npz
"""
import json
import os
import sys
import types

import numpy as np
import scipy
import scipy.signal as signal

for m in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "matplotlib.mlab"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")
import sk_dsp_comm.sigsys as ss         # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    F = np.load(os.path.join(OUT, "filters.npz"))
    rng = np.random.default_rng(815)
    out, table = {}, []

    def cx(n):
        return rng.standard_normal(n) + 1j * rng.standard_normal(n)

    # ---- os_filter / oa_filter
    blk = [("cos20", np.cos(2 * np.pi * 0.05 * np.arange(20)), np.ones(10), 2 ** 10),     # tests/test_sigsys.py:688-706
           ("r1000_b101", rng.standard_normal(1000), F["b101"], 256),
           ("r300_b7", rng.standard_normal(300), F["b7"], 16),
           ("r77_b7_tight", rng.standard_normal(77), F["b7"], 7),                         # L = 1
           ("c200_b33", cx(200), F["b33_remez_bpf"], 64),                                 # np.real() of a complex input
           ("r5000_b256", rng.standard_normal(5000), F["b256"], 1024)]
    for name, x, h, N in blk:
        out["blk_%s_x" % name], out["blk_%s_h" % name] = x, h
        for fn in ("os_filter", "oa_filter"):
            y, ym = getattr(ss, fn)(x, h, N, 1)
            assert np.array_equal(y, getattr(ss, fn)(x, h, N))
            out["blk_%s_%s_y" % (name, fn)] = y
            if ym.size <= 400000:
                out["blk_%s_%s_ymat" % (name, fn)] = ym
        table.append(dict(kind="blk", name=name, N=N))

    # ---- FIR with lfilter's zi / zf
    for name, K in (("b256", None), ("b101", None), ("b7", None), ("b1", None)):
        b = F[name]
        for tag, x in (("f64_1000", rng.standard_normal(1000)), ("c128_777", cx(777)),
                       ("f64_100", rng.standard_normal(100)), ("f64_1", rng.standard_normal(1)),
                       ("f32_4096", rng.standard_normal(4096).astype(np.float32))):
            zi = rng.standard_normal(len(b) - 1)
            if tag.startswith("c128"):
                zi = zi + 1j * rng.standard_normal(len(b) - 1)
            y, zf = signal.lfilter(b, [1], x, zi=zi)
            key = "firz_%s_%s" % (name, tag)
            out[key + "_x"], out[key + "_zi"], out[key + "_y"], out[key + "_zf"] = x, zi, y, zf
            table.append(dict(kind="firz", key=key, filt=name))
    # real x with complex zi promotes
    b = F["b7"]
    x = rng.standard_normal(50)
    zi = cx(6)
    y, zf = signal.lfilter(b, [1], x, zi=zi)
    out["firz_promote_x"], out["firz_promote_zi"], out["firz_promote_y"], out["firz_promote_zf"] = x, zi, y, zf

    # ---- SOS with sosfilt's zi / zf
    for name in ("sos6", "sos_butter5", "sos_sharp_lpf"):
        sos = F[name]
        for tag, x in (("f64_3000", rng.standard_normal(3000)), ("c128_1500", cx(1500)),
                       ("f64_1", rng.standard_normal(1)), ("f32_70000", rng.standard_normal(70000).astype(np.float32))):
            if tag == "f32_70000" and name != "sos6":      # one case longer than a scan tile is enough
                continue
            zi = rng.standard_normal((sos.shape[0], 2)) * 0.1
            if tag.startswith("c128"):
                zi = zi + 1j * rng.standard_normal((sos.shape[0], 2)) * 0.1
            y, zf = signal.sosfilt(sos, x, zi=zi)
            key = "sosz_%s_%s" % (name, tag)
            out[key + "_x"], out[key + "_zi"], out[key + "_y"], out[key + "_zf"] = x, zi, y, zf
            table.append(dict(kind="sosz", key=key, filt=name))
    # ---- ten-band equaliser (sigsys.py:96-141): literal test input (tests/test_helper.py seed 100) + longer runs
    np.random.seed(100)
    out["eq_w10"] = np.random.randn(10)
    out["eq_y10"] = ss.ten_band_eq_filt(out["eq_w10"], [g for g in range(1, 11)])
    out["eq_x"] = rng.standard_normal(50000)
    out["eq_gdb"] = np.array([3., -2., 0., 6., -6., 1.5, -1.5, 4., -4., 2.])
    out["eq_y"] = ss.ten_band_eq_filt(out["eq_x"], out["eq_gdb"])
    out["eq_y_q2"] = ss.ten_band_eq_filt(out["eq_x"][:5000], out["eq_gdb"], 2.0)
    out["peak_b"], out["peak_a"] = ss.peaking(2.0, 500, 3.5, 44100)
    out["cic_4_7"], out["cic_10_2"], out["cic_5_1"] = ss.cic(4, 7), ss.cic(10, 2), ss.cic(5, 1)
    out["table"] = np.array(json.dumps(table))
    out["scipy_version"] = np.array(scipy.__version__)
    np.savez_compressed(os.path.join(OUT, "stream_cases.npz"), **out)
    print("wrote", len(table), "cases", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
