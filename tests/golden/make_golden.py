#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference in the dev container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (mwickert/scikit-dsp-comm, /root/reference, read-only) is pure Python and
cannot travel to the GPU box, so its outputs are committed as small fixtures.  The only
shim is a 4-entry ``sys.modules`` stub for matplotlib (absent from this image; the
reference imports pylab/pyplot at module top, multirate_helper.py:32-33) -- no reference
source is modified or copied.

Fixtures written:
  filters.npz   -- coefficient sets produced by the reference's own design helpers
                   (fir_design_helper.py:62-88, iir_design_helper.py:143-191)
  ref_cases.npz -- inputs + outputs of multirate_FIR/multirate_IIR .filter/.up/.dn and
                   sigsys.upsample/downsample for several dtypes / rate factors
  cfg1_windows.npz -- BASELINE.json configs[0] (101-tap Kaiser LPF on 2^20 float64):
                   head/middle/tail output windows + a float64 checksum
"""
import os
import sys
import types

import numpy as np

for m in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "matplotlib.mlab"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")

import sk_dsp_comm.multirate_helper as mrh      # noqa: E402
import sk_dsp_comm.fir_design_helper as fir_d   # noqa: E402
import sk_dsp_comm.iir_design_helper as iir_d   # noqa: E402
import sk_dsp_comm.sigsys as ss                 # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make_filters():
    f = {}
    f["b101"] = fir_d.firwin_kaiser_lpf(0.2, 0.25, 80, fs=1.0, n_bump=-1)       # cfg1
    f["b256"] = fir_d.firwin_kaiser_lpf(0.10, 0.12, 80, fs=1.0, n_bump=4)       # cfg2/3/5
    f["sos6"] = iir_d.IIR_bpf(0.07, 0.1, 0.2, 0.23, 0.5, 60, 1.0, 'ellip')[2]   # cfg4
    # extra shapes: odd length / non-multiple-of-anything taps, a short filter, a 1-tap gain,
    # a Remez band-pass (asymmetric use), and the notebook's sharp elliptic low-pass
    f["b33_remez_bpf"] = fir_d.fir_remez_bpf(0.1, 0.15, 0.3, 0.35, 0.5, 50, fs=1.0, n_bump=0)
    f["b7"] = np.array([0.1, -0.2, 0.3, 0.9, 0.3, -0.2, 0.1])
    f["b1"] = np.array([0.5])
    f["sos_sharp_lpf"] = iir_d.IIR_lpf(1950, 2050, 0.5, 80, 8000, 'ellip')[2]
    f["sos_butter6"] = iir_d.IIR_lpf(0.1, 0.2, 0.5, 30, 1.0, 'butter')[2]
    import scipy.signal as _sig
    f["sos_butter5"] = _sig.butter(5, 0.2, output='sos')    # odd order -> a first-order section (b2 = a2 = 0)
    # ten-band equaliser cascade of tests/test_sigsys.py:28-34 (sigsys.py:96-141): rows [B_k, A_k]
    # from the reference's own ``peaking`` designer, so the literal 10-value known-answer
    # vector of that test pins the oracle's biquad cascade.
    f["sos_tenband"] = np.array([np.hstack(ss.peaking(g, fc, 3.5))
                                 for g, fc in zip(range(1, 11), 31.25 * 2 ** np.arange(10))])
    np.random.seed(100)                       # tests/test_helper.py:4-7
    f["tenband_w"] = np.random.randn(10)
    assert len(f["b101"]) == 101 and len(f["b256"]) == 256 and f["sos6"].shape == (6, 6)
    return f


def rand(rng, n, dt):
    dt = np.dtype(dt)
    if dt.kind == "c":
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dt)
    if dt.kind == "i":
        return rng.integers(-2000, 2000, n).astype(dt)
    return rng.standard_normal(n).astype(dt)


def main():
    filt = make_filters()
    # "next" row: rate_change + interp24/deci24.  Coefficients come from the
    # reference's own scipy calls; inputs of the two literal goldens are the reference's m-sequences.
    import scipy.signal as _sig2
    for M in (2, 3, 4):
        bb, aa = _sig2.butter(10, 1.0 / M)
        filt["butter10_%d_b" % M], filt["butter10_%d_a" % M] = bb, aa
    filt["mseq2"], filt["mseq3"] = ss.m_seq(2), ss.m_seq(3)
    filt["iir_orders"] = np.array([mrh.multirate_IIR(filt[k]).N_forder
                                   for k in ("sos6", "sos_sharp_lpf", "sos_butter5", "sos_butter6")])
    np.savez_compressed(os.path.join(OUT, "filters.npz"), **filt)

    cases = {}
    rng = np.random.default_rng(20260924)
    idx = 0

    def add(kind, fname, x, y, **kw):
        nonlocal idx
        key = "case%03d" % idx
        idx += 1
        cases[key + "_x"] = x
        cases[key + "_y"] = np.ascontiguousarray(y)
        meta = dict(kind=kind, filt=fname, **kw)
        cases[key + "_meta"] = np.array(repr(meta))

    # FIR
    for fname in ("b101", "b256", "b33_remez_bpf", "b7", "b1"):
        fir = mrh.multirate_FIR(filt[fname])
        for dt, n in (("float64", 1501), ("float32", 1024), ("complex64", 769),
                      ("complex128", 500), ("int16", 333)):
            x = rand(rng, n, dt)
            add("fir_filter", fname, x, fir.filter(x))
        for L in (4, 3, 12):
            x = rand(rng, 301, "float32")
            add("fir_up", fname, x, fir.up(x, L), L=L)
            x = rand(rng, 167, "complex128")
            add("fir_up", fname, x, fir.up(x, L), L=L)
        for M in (4, 3, 12):
            x = rand(rng, 2051, "float64")
            add("fir_dn", fname, x, fir.dn(x, M), M=M)
            x = rand(rng, 1250, "complex64")
            add("fir_dn", fname, x, fir.dn(x, M), M=M)
    # short inputs (shorter than the filter) and a single sample
    for n in (1, 2, 5, 100, 255, 256, 257):
        x = rand(rng, n, "float64")
        add("fir_filter", "b256", x, mrh.multirate_FIR(filt["b256"]).filter(x))
    # default rate factors (L_change = M_change = 12, multirate_helper.py:112,121)
    x = rand(rng, 240, "float64")
    add("fir_up", "b256", x, mrh.multirate_FIR(filt["b256"]).up(x), L=12)
    add("fir_dn", "b256", x, mrh.multirate_FIR(filt["b256"]).dn(x), M=12)

    # IIR
    for fname in ("sos6", "sos_sharp_lpf", "sos_butter5", "sos_butter6"):
        iir = mrh.multirate_IIR(filt[fname])
        for dt, n in (("float64", 2500), ("float32", 2048), ("complex64", 750),
                      ("complex128", 499), ("int16", 256), ("float64", 1)):
            x = rand(rng, n, dt)
            add("sos_filter", fname, x, iir.filter(x))
        if fname == "sos6":      # longer than one scan tile of the GPU kernel
            x = rand(rng, 40000, "float32")
            add("sos_filter", fname, x, iir.filter(x))
        for L in (4, 12):
            x = rand(rng, 350, "float32")
            add("sos_up", fname, x, iir.up(x, L), L=L)
        for M in (4, 12):
            x = rand(rng, 3000, "float64")
            add("sos_dn", fname, x, iir.dn(x, M), M=M)

    # rate_change (multirate_helper.py:45-83)
    for M, fcut, N, ftype in ((12, 0.9, 8, "butter"), (4, 0.8, 6, "cheby1"), (3, 0.9, 5, "butter")):
        rc = mrh.rate_change(M, fcut, N, ftype)
        x = rand(rng, 400, "float64")
        add("rc_up", "", x, rc.up(x), M=M, fcut=fcut, N=N, ftype=ftype, b=rc.b.tolist(), a=rc.a.tolist())
        x = rand(rng, 5000, "float32")
        add("rc_dn", "", x, rc.dn(x), M=M, fcut=fcut, N=N, ftype=ftype, b=rc.b.tolist(), a=rc.a.tolist())
    x = rand(rng, 100, "float64")
    add("interp24", "", x, ss.interp24(x))
    x = rand(rng, 6000, "float64")
    add("deci24", "", x, ss.deci24(x))

    # upsample / downsample (sigsys.py:3031-3083)
    for dt in ("float32", "float64", "complex64", "complex128", "int16"):
        x = rand(rng, 101, dt)
        for L in (1, 2, 3, 4, 12):
            add("upsample", "", x, ss.upsample(x, L), L=L)
        for M, p in ((1, 0), (2, 1), (3, 0), (3, 2), (4, 1), (12, 11), (200, 0)):
            add("downsample", "", x, ss.downsample(x, M, p), M=M, p=p)
    np.savez_compressed(os.path.join(OUT, "ref_cases.npz"), **cases)

    # cfg1: 101-tap Kaiser LPF on 2^20 float64 (SURVEY.md 8d synthetic input)
    x = np.random.default_rng(100).standard_normal(2 ** 20)
    y = mrh.multirate_FIR(filt["b101"]).filter(x)
    W = 4096
    mid = 2 ** 19
    np.savez_compressed(os.path.join(OUT, "cfg1_windows.npz"),
                        head=y[:W], mid=y[mid:mid + W], tail=y[-W:],
                        x_head=x[:8], checksum=np.array([y.sum(), np.abs(y).sum(), (y * y).sum()]))
    print("wrote", idx, "cases;", {k: v.shape for k, v in filt.items()})


if __name__ == "__main__":
    main()
