"""Streaming state (zi/zf) and block filtering against fixtures from
scipy's lfilter/sosfilt ``zi`` argument and the reference's os_filter/oa_filter
(tests/golden/make_stream_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

S = np.load(os.path.join(GOLDEN, "stream_cases.npz"))
F = np.load(os.path.join(GOLDEN, "filters.npz"))
TABLE = json.loads(str(S["table"]))
FIRZ = [c for c in TABLE if c["kind"] == "firz"]
SOSZ = [c for c in TABLE if c["kind"] == "sosz"]
BLK = [c for c in TABLE if c["kind"] == "blk"]

FIR_TOL, IIR_TOL = 1e-11, 1e-9


def _rel(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.abs(got - want).max()) / max(float(np.abs(want).max()), 1e-300) if want.size else 0.0


# ------------------------------------------------------------------------------ CPU: host logic

@pytest.mark.parametrize("c", FIRZ, ids=[c["key"] for c in FIRZ])
def test_fir_final_state_is_zero_input_response(c):
    """The identity the stateful FIR call relies on: lfilter's zf == the filter's output for K-1 zero
    samples whose history is the block's tail (+ what is left of zi when the block is shorter than K-1),
    and y == FIR(x) with zi added to the first K-1 outputs.  Checked with the oracle on scipy's fixtures."""
    import oracle
    k = c["key"]
    b, x, zi = F[c["filt"]], S[k + "_x"], S[k + "_zi"]
    K, N = len(b), len(x)
    dt = np.complex128 if (np.iscomplexobj(x) or np.iscomplexobj(zi)) else np.float64
    x64 = x.astype(dt)
    y = oracle.fir_filter(b, x64, backend="numpy")
    m = min(K - 1, N)
    y[:m] += zi[:m]
    assert _rel(y, S[k + "_y"]) <= 1e-13
    if K == 1:
        return
    tail = x64[N - (K - 1):] if N >= K - 1 else np.concatenate([np.zeros(K - 1 - N, dt), x64])
    zf = oracle.fir_filter(b, np.zeros(K - 1, dt), hist=tail, backend="numpy")
    if N < K - 1:
        zf[:K - 1 - N] += zi[N:]
    assert _rel(zf, S[k + "_zf"]) <= 1e-13


# ------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("c", FIRZ, ids=[c["key"] for c in FIRZ])
def test_gpu_fir_zi_zf(c):
    import sk_dsp_comm_b200.multirate_helper as mrh
    k = c["key"]
    fir = mrh.multirate_FIR(F[c["filt"]])
    y, zf = fir.filter(S[k + "_x"], zi=S[k + "_zi"])
    assert y.dtype == S[k + "_y"].dtype and zf.dtype == S[k + "_zf"].dtype
    assert _rel(y, S[k + "_y"]) <= FIR_TOL
    if zf.size:
        assert _rel(zf, S[k + "_zf"]) <= FIR_TOL


@pytest.mark.gpu
def test_gpu_fir_zi_promotes_and_validates():
    import sk_dsp_comm_b200.multirate_helper as mrh
    fir = mrh.multirate_FIR(F["b7"])
    y, zf = fir.filter(S["firz_promote_x"], zi=S["firz_promote_zi"])
    assert y.dtype == np.complex128
    assert _rel(y, S["firz_promote_y"]) <= FIR_TOL and _rel(zf, S["firz_promote_zf"]) <= FIR_TOL
    with pytest.raises(ValueError, match="Unexpected shape for zi"):
        fir.filter(np.zeros(10), zi=np.zeros(5))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "complex64"])
def test_gpu_fir_chunked_stream_equals_monolithic(dtype):
    """Chaining zf -> zi over ragged chunks (one shorter than the filter) reproduces one call."""
    import torch
    import sk_dsp_comm_b200.multirate_helper as mrh
    rng = np.random.default_rng(4)
    n = 100000
    x = rng.standard_normal(n)
    if dtype == "complex64":
        x = (x + 1j * rng.standard_normal(n)).astype(np.complex64)
    fir = mrh.multirate_FIR(F["b256"])
    xt = torch.from_numpy(x).cuda()
    whole = fir.filter(xt)
    z = np.zeros(255)
    parts = []
    for lo, hi in ((0, 40000), (40000, 40100), (40100, 40101), (40101, n)):
        y, z = fir.filter(xt[lo:hi], zi=z)
        assert y.is_cuda and z.is_cuda and y.dtype == xt.dtype
        parts.append(y)
    got = torch.cat(parts)
    tol = FIR_TOL if dtype == "float64" else 1e-6
    assert float((got - whole).abs().max()) <= tol * float(whole.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("c", SOSZ, ids=[c["key"] for c in SOSZ])
def test_gpu_sos_zi_zf(c):
    import sk_dsp_comm_b200.multirate_helper as mrh
    k = c["key"]
    iir = mrh.multirate_IIR(F[c["filt"]])
    y, zf = iir.filter(S[k + "_x"], zi=S[k + "_zi"])
    assert y.dtype == S[k + "_y"].dtype and zf.shape == S[k + "_zf"].shape and zf.dtype == S[k + "_zf"].dtype
    assert _rel(y, S[k + "_y"]) <= IIR_TOL
    assert _rel(zf, S[k + "_zf"]) <= IIR_TOL


@pytest.mark.gpu
def test_gpu_sos_chunked_stream_equals_monolithic():
    import torch
    import sk_dsp_comm_b200.multirate_helper as mrh
    x = torch.from_numpy(np.random.default_rng(6).standard_normal(300000)).cuda()
    iir = mrh.multirate_IIR(F["sos6"])
    whole = iir.filter(x)
    z = np.zeros((6, 2))
    parts = []
    for lo, hi in ((0, 70001), (70001, 70002), (70002, 200000), (200000, 300000)):
        y, z = iir.filter(x[lo:hi], zi=z)
        parts.append(y)
    got = torch.cat(parts)
    assert float((got - whole).abs().max()) <= IIR_TOL * float(whole.abs().max())
    with pytest.raises(ValueError, match="Invalid zi shape"):
        iir.filter(x[:10], zi=np.zeros((5, 2)))


@pytest.mark.gpu
@pytest.mark.parametrize("c", BLK, ids=[c["name"] for c in BLK])
def test_gpu_block_filters(c):
    import sk_dsp_comm_b200.sigsys as ss
    name, N = c["name"], c["N"]
    x, h = S["blk_%s_x" % name], S["blk_%s_h" % name]
    for fn in ("os_filter", "oa_filter"):
        want = S["blk_%s_%s_y" % (name, fn)]
        y = getattr(ss, fn)(x, h, N)
        assert y.dtype == np.float64 and _rel(y, want) <= 1e-10          # the reference's FFT path carries ~1e-13
        key = "blk_%s_%s_ymat" % (name, fn)
        if key in S.files and S[key].shape[0] <= 64:
            y1, ymat = getattr(ss, fn)(x, h, N, 1)
            assert np.array_equal(y1, y)
            assert ymat.shape == S[key].shape
            assert np.abs(ymat - S[key]).max() <= 1e-10 * max(np.abs(S[key]).max(), 1.0)


@pytest.mark.gpu
def test_gpu_block_filter_literal_golden():
    """tests/test_sigsys.py:688-706."""
    import sk_dsp_comm_b200.sigsys as ss
    y_test = [1., 1.95105652, 2.76007351, 3.34785876, 3.65687576, 3.65687576, 3.34785876, 2.76007351,
              1.95105652, 1., -1., -2.90211303, -4.52014702, -5.69571753, -6.31375151,
              -6.31375151, -5.69571753, -4.52014702, -2.90211303, -1.]
    x = np.cos(2 * np.pi * 0.05 * np.arange(0, 20))
    np.testing.assert_almost_equal(ss.os_filter(x, np.ones(10), 2 ** 10), y_test)
    np.testing.assert_almost_equal(ss.oa_filter(x, np.ones(10), 2 ** 10), y_test)
    with pytest.raises(ValueError):
        ss.os_filter(x, np.ones(10), 8)


# ------------------------------------------------------------------------------ ten-band equaliser

def test_peaking_and_cic_designs_match_reference():
    import sk_dsp_comm_b200.sigsys as ss
    b, a = ss.peaking(2.0, 500, 3.5, 44100)
    np.testing.assert_almost_equal(b, [1.00458357, -1.95961252, 0.96001185])     # tests/test_sigsys.py:42-47
    np.testing.assert_almost_equal(a, [1., -1.95961252, 0.96459542])
    assert np.abs(b - S["peak_b"]).max() <= 1e-15 and np.abs(a - S["peak_a"]).max() <= 1e-15
    for (m, k) in ((4, 7), (10, 2), (5, 1)):
        assert np.abs(ss.cic(m, k) - S["cic_%d_%d" % (m, k)]).max() <= 1e-16
    assert np.sum(ss.cic(10, 1) - np.ones(10) / 10) == 0                          # tests/test_sigsys.py:13-19


@pytest.mark.gpu
def test_gpu_ten_band_equaliser():
    import sk_dsp_comm_b200.sigsys as ss
    y_test = [-4.23769156, 0.097137, 4.18516645, -0.54460053, 2.2257584, 1.60147407, -0.76767407, -1.95402381,
              -1.0580526, 0.9111369]                                              # tests/test_sigsys.py:28-34
    y = ss.ten_band_eq_filt(S["eq_w10"], [g for g in range(1, 11)])
    np.testing.assert_almost_equal(y, y_test)
    assert _rel(y, S["eq_y10"]) <= IIR_TOL
    assert _rel(ss.ten_band_eq_filt(S["eq_x"], S["eq_gdb"]), S["eq_y"]) <= IIR_TOL
    assert _rel(ss.ten_band_eq_filt(S["eq_x"][:5000], S["eq_gdb"], 2.0), S["eq_y_q2"]) <= IIR_TOL
    with pytest.raises(ValueError, match="GdB length not equal to ten"):
        ss.ten_band_eq_filt(S["eq_w10"], [g for g in range(1, 9)])
