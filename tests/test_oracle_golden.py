"""Pin the CPU oracle: (1) the reference's own literal known-answer vectors, (2) fixtures
produced by the unmodified reference classes (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import numpy.testing as npt
import pytest

import oracle
from conftest import ref_cases


# ---- (1) literal vectors copied from the reference's test-suite --------------------------
def test_boxcar_fir_known_answer():
    # /root/reference/tests/test_sigsys.py:688-706 (test_os_filter_0 / test_oa_filter_0):
    # 10-tap boxcar on a cosine == lfilter(ones(10), 1, x)
    y_test = [1., 1.95105652, 2.76007351, 3.34785876, 3.65687576, 3.65687576, 3.34785876,
              2.76007351, 1.95105652, 1., -1., -2.90211303, -4.52014702, -5.69571753,
              -6.31375151, -6.31375151, -5.69571753, -4.52014702, -2.90211303, -1.]
    n = np.arange(0, 20)
    x = np.cos(2 * np.pi * 0.05 * n)
    b = np.ones(10)
    npt.assert_almost_equal(oracle.fir_filter(b, x), y_test)
    npt.assert_almost_equal(oracle.fir_filter(b, x, backend="c"), y_test)


def test_ten_band_biquad_cascade_known_answer(filters):
    # /root/reference/tests/test_sigsys.py:28-34 (test_ten_band_equalizer): ten cascaded
    # biquads (sigsys.py:96-141) on randn(10), seed 100.
    y_test = [-4.23769156, 0.097137, 4.18516645, -0.54460053, 2.2257584, 1.60147407,
              -0.76767407, -1.95402381, -1.0580526, 0.9111369]
    y = oracle.sos_filter(filters["sos_tenband"], filters["tenband_w"])
    npt.assert_almost_equal(y, y_test)
    y_py, _ = oracle._sosfilt_py(filters["sos_tenband"], filters["tenband_w"])
    npt.assert_almost_equal(y_py, y_test)


def test_upsample_downsample_reference_tests():
    # /root/reference/tests/test_sigsys.py:655-668
    npt.assert_equal(oracle.upsample(np.zeros(1), 3), np.zeros(3))
    npt.assert_equal(oracle.downsample(np.zeros(3), 3), np.zeros(1))
    with pytest.raises(TypeError, match="M must be an int"):
        oracle.downsample(np.zeros(0), 3.0)


# ---- (2) fixtures from the unmodified reference -----------------------------------------
def _run_oracle(c, filters, backend):
    k = c["kind"]
    x = c["x"]
    if k == "fir_filter":
        return oracle.fir_filter(filters[c["filt"]], x, backend=backend)
    if k == "fir_up":
        return oracle.fir_up(filters[c["filt"]], x, c["L"], backend=backend)
    if k == "fir_dn":
        return oracle.fir_dn(filters[c["filt"]], x, c["M"], backend=backend)
    if k == "sos_filter":
        return oracle.sos_filter(filters[c["filt"]], x)
    if k == "sos_up":
        return oracle.sos_up(filters[c["filt"]], x, c["L"])
    if k == "sos_dn":
        return oracle.sos_dn(filters[c["filt"]], x, c["M"])
    if k == "upsample":
        return oracle.upsample(x, c["L"])
    if k == "downsample":
        return oracle.downsample(x, c["M"], c["p"])
    raise AssertionError(k)


@pytest.mark.parametrize("backend", ["numpy", "c"])
def test_oracle_matches_reference_fixtures(filters, backend):
    n = 0
    for c in ref_cases():
        y = _run_oracle(c, filters, backend)
        ref = c["y"]
        assert y.dtype == ref.dtype, (c["name"], c["kind"], y.dtype, ref.dtype)
        assert y.shape == ref.shape, (c["name"], c["kind"], y.shape, ref.shape)
        if c["kind"] in ("upsample", "downsample"):
            assert np.array_equal(y, ref), c["name"]          # index maps: bit exact
        elif c["kind"].startswith("sos") or backend == "numpy":
            # same arithmetic in the same order -> identical bits
            assert np.array_equal(y, ref), (c["name"], c["kind"], np.abs(y - ref).max())
        else:
            # C FIR sums in tap order, numpy's convolve uses its own dot kernel: ~1 ulp apart
            scale = max(np.abs(ref).max(), 1e-300)
            assert np.abs(y - ref).max() <= 1e-13 * scale, (c["name"], c["kind"])
        n += 1
    assert n > 150


def test_cfg1_windows(filters):
    """BASELINE.json configs[0]: 101-tap Kaiser LPF on 2^20 float64."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "cfg1_windows.npz"))
    x = np.random.default_rng(100).standard_normal(2 ** 20)
    npt.assert_array_equal(x[:8], g["x_head"])
    y = oracle.fir_filter(filters["b101"], x, backend="c")
    W = 4096
    mid = 2 ** 19
    for got, ref in ((y[:W], g["head"]), (y[mid:mid + W], g["mid"]), (y[-W:], g["tail"])):
        npt.assert_allclose(got, ref, rtol=1e-9, atol=1e-13)
    npt.assert_allclose([y.sum(), np.abs(y).sum(), (y * y).sum()], g["checksum"], rtol=1e-9)


def test_halo_and_state_carry_are_exact(filters):
    """Overlap-save halo (FIR) and zi/zf carry (SOS) reproduce the monolithic result --
    the property the sharded multi-GPU path relies on."""
    rng = np.random.default_rng(7)
    b = filters["b256"]
    x = rng.standard_normal(5000) + 1j * rng.standard_normal(5000)
    y = oracle.fir_filter(b, x, backend="c")
    cut = 1777
    y2 = oracle.fir_filter(b, x[cut:], hist=x[cut - 255:cut], backend="c")
    assert np.array_equal(y[cut:], y2)
    y2n = oracle.fir_filter(b, x[cut:], hist=x[cut - 255:cut], backend="numpy")
    npt.assert_allclose(y2n, y[cut:], rtol=0, atol=1e-13)
    sos = filters["sos6"]
    xr = rng.standard_normal(9000)
    yr = oracle.sos_filter(sos, xr)
    ya, zf = oracle.sos_filter(sos, xr[:4000], return_zf=True)
    yb = oracle.sos_filter(sos, xr[4000:], zi=zf)
    assert np.array_equal(np.concatenate([ya, yb]), yr)


def test_iir_order(filters):
    names = ("sos6", "sos_sharp_lpf", "sos_butter5", "sos_butter6")
    got = [oracle.iir_order(filters[k]) for k in names]
    assert got == list(filters["iir_orders"]) == [12, 11, 5, 6]
