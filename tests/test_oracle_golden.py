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


def oracle_interp24(x, filters):
    """sigsys.interp24 (sigsys.py:2971-2985) restated on the oracle's lfilter; (b, a) = the reference's own
    scipy.signal.butter(10, 1/M) coefficients from the fixture file."""
    y = np.asarray(x, dtype=np.float64)
    for M in (2, 3, 4):
        y = oracle.lfilter_ba(filters["butter10_%d_b" % M], filters["butter10_%d_a" % M], M * oracle.upsample(y, M))
    return y


def oracle_deci24(x, filters):
    """sigsys.deci24 (sigsys.py:3014-3028)."""
    y = np.asarray(x, dtype=np.float64)
    for M in (2, 3, 4):
        y = np.ascontiguousarray(oracle.downsample(oracle.lfilter_ba(filters["butter10_%d_b" % M], filters["butter10_%d_a" % M], y), M))
    return y


INTERP24_GOLDEN = np.array([   # /root/reference/tests/test_sigsys.py:617-644 (test_interp24, x = m_seq(2))
    8.95202944e-11, 1.34163933e-09, 9.65059549e-09, 4.44437220e-08, 1.48633219e-07, 3.93481543e-07,
    8.92384996e-07, 1.86143255e-06, 3.71505394e-06, 7.08928675e-06, 1.27731258e-05, 2.18065111e-05,
    3.58949144e-05, 5.76136754e-05, 8.98878361e-05, 1.35638003e-04, 1.98845903e-04, 2.85711249e-04,
    4.03316692e-04, 5.57555026e-04, 7.55072659e-04, 1.00722733e-03, 1.32812089e-03, 1.72879988e-03,
    2.22023892e-03, 2.82320495e-03, 3.56559842e-03, 4.46845856e-03, 5.54968392e-03, 6.84570710e-03,
    8.40812804e-03, 1.02717590e-02, 1.24562018e-02, 1.50108479e-02, 1.80160711e-02, 2.15202204e-02,
    2.55317912e-02, 3.01035822e-02, 3.53501673e-02, 4.13390064e-02, 4.80615549e-02, 5.55741246e-02,
    6.40436130e-02, 7.35682641e-02, 8.41085273e-02, 9.57097399e-02, 1.08597610e-01, 1.22897001e-01,
    1.38490810e-01, 1.55356681e-01, 1.73756981e-01, 1.93835690e-01, 2.15364192e-01, 2.38212644e-01,
    2.62680245e-01, 2.88955795e-01, 3.16705554e-01, 3.45687757e-01, 3.76257702e-01, 4.08669655e-01,
    4.42456034e-01, 4.77197735e-01, 5.13261795e-01, 5.50948788e-01, 5.89612250e-01, 6.28576388e-01,
    6.68168565e-01, 7.08757382e-01, 7.49579596e-01, 7.89751182e-01, 8.29598667e-01, 8.69628354e-01])
DECI24_GOLDEN = np.array([     # /root/reference/tests/test_sigsys.py:646-653 (test_deci24, x = interp24(m_seq(3)))
    3.33911797e-22, 3.71880014e-10, 4.33029514e-06, 1.16169513e-03, 4.34891180e-02, 4.08255952e-01,
    1.16839852e+00])


def test_interp24_deci24_known_answers(filters):
    npt.assert_almost_equal(oracle_interp24(filters["mseq2"], filters), INTERP24_GOLDEN)
    npt.assert_almost_equal(oracle_deci24(oracle_interp24(filters["mseq3"], filters), filters), DECI24_GOLDEN)


def test_product_filter_designs_match_reference_coefficients(filters):
    """rate_change / interp24 design their Butterworth / Chebyshev-I filters with the product's own numpy
    designer (_design.py, no scipy): it must reproduce the (b, a) the reference obtained from scipy."""
    from sk_dsp_comm_b200 import _design
    for M in (2, 3, 4):
        b, a, sos = _design.butter(10, 1.0 / M)
        npt.assert_allclose(b, filters["butter10_%d_b" % M], rtol=1e-12, atol=0)
        npt.assert_allclose(a, filters["butter10_%d_a" % M], rtol=1e-12, atol=0)
        assert sos.shape == (5, 6) and np.all(sos[:, 3] == 1)
    n = 0
    for c in ref_cases():
        if c["kind"] in ("rc_up", "rc_dn"):
            fc = c["fcut"] * 0.5
            if c["ftype"] == "butter":
                b, a, sos = _design.butter(c["N"], 2 / c["M"] * fc)
            else:
                b, a, sos = _design.cheby1(c["N"], 0.05, 2 / c["M"] * fc)
            npt.assert_allclose(b, c["b"], rtol=1e-11, atol=0)
            npt.assert_allclose(a, c["a"], rtol=1e-11, atol=0)
            # the cascade realises the same transfer function: impulse responses agree
            imp = np.zeros(400)
            imp[0] = 1.0
            npt.assert_allclose(oracle.sos_filter(sos, imp), oracle.lfilter_ba(c["b"], c["a"], imp), rtol=0, atol=1e-9)
            n += 1
    assert n == 6


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
    if k == "rc_up":
        return oracle.rate_change_up(c["b"], c["a"], x, c["M"])
    if k == "rc_dn":
        return oracle.rate_change_dn(c["b"], c["a"], x, c["M"])
    if k == "interp24":
        return oracle_interp24(x, filters)
    if k == "deci24":
        return oracle_deci24(x, filters)
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
        elif c["kind"] in ("rc_up", "rc_dn", "interp24", "deci24"):
            scale = max(np.abs(ref).max(), 1e-300)
            assert np.abs(y - ref).max() <= 1e-13 * scale, (c["name"], c["kind"], np.abs(y - ref).max())
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
