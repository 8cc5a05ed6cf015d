"""GPU parity tests: the CUDA path (through the public drop-in API -> ctypes -> C ABI) against
(1) fixtures produced by the unmodified reference, (2) the reference's literal known-answer
vectors, (3) the CPU oracle on seeded inputs, (4) size-independent properties at full size.

Tolerances (BASELINE.json north_star):
  * float64 / complex128 arithmetic (every numpy input, like the reference): FIR rtol 1e-6,
    IIR rtol 1e-4 are the bars; the kernels compute in fp64, so the tests assert a much tighter
    |d| <= 1e-11*max|y| (FIR) / 1e-9*max|y| (IIR).
  * dtype-preserving float32 / complex64 streams (torch tensors): FIR |d| <= 1e-6*max|y_ref|,
    IIR |d| <= 1e-4*max|y_ref|, y_ref = oracle on the same values widened to float64.
  * upsample / downsample: bit exact.
"""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

import oracle
from conftest import ref_cases, GOLDEN

pytestmark = pytest.mark.gpu

FIR_TOL64, IIR_TOL64 = 1e-11, 1e-9
FIR_TOL32, IIR_TOL32 = 1e-6, 1e-4


@pytest.fixture(scope="module")
def mods():
    import sk_dsp_comm_b200.multirate_helper as mrh
    import sk_dsp_comm_b200.sigsys as ss
    import sk_dsp_comm_b200._cabi as cabi
    return mrh, ss, cabi


def _maxerr(y, ref):
    y = np.asarray(y)
    ref = np.asarray(ref)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    if ref.size == 0:
        return 0.0, 1.0
    return float(np.abs(y - ref).max()), max(float(np.abs(ref).max()), 1e-300)


def _run(mods, c, filters, x):
    mrh, ss, _ = mods
    k = c["kind"]
    if k == "fir_filter":
        return mrh.multirate_FIR(filters[c["filt"]]).filter(x)
    if k == "fir_up":
        return mrh.multirate_FIR(filters[c["filt"]]).up(x, c["L"])
    if k == "fir_dn":
        return mrh.multirate_FIR(filters[c["filt"]]).dn(x, c["M"])
    if k == "sos_filter":
        return mrh.multirate_IIR(filters[c["filt"]]).filter(x)
    if k == "sos_up":
        return mrh.multirate_IIR(filters[c["filt"]]).up(x, c["L"])
    if k == "sos_dn":
        return mrh.multirate_IIR(filters[c["filt"]]).dn(x, c["M"])
    if k == "rc_up":
        return mrh.rate_change(c["M"], c["fcut"], c["N"], c["ftype"]).up(x)
    if k == "rc_dn":
        return mrh.rate_change(c["M"], c["fcut"], c["N"], c["ftype"]).dn(x)
    if k == "interp24":
        return ss.interp24(x)
    if k == "deci24":
        return ss.deci24(x)
    if k == "upsample":
        return ss.upsample(x, c["L"])
    if k == "downsample":
        return ss.downsample(x, c["M"], c["p"])
    raise AssertionError(k)


# ---------------------------------------------------------------- (1) reference fixtures
def test_reference_fixtures_numpy_in_numpy_out(mods, filters):
    """numpy in -> numpy out reproduces the reference's dtype, shape and values."""
    launches0 = mods[2].launch_count()
    n = 0
    for c in ref_cases():
        y = _run(mods, c, filters, c["x"])
        ref = c["y"]
        assert isinstance(y, np.ndarray)
        assert y.dtype == ref.dtype, (c["name"], c["kind"], y.dtype, ref.dtype)
        if c["kind"] in ("upsample", "downsample"):
            assert y.shape == ref.shape and np.array_equal(y, ref), (c["name"], c["kind"])
        else:
            err, scale = _maxerr(y, ref)
            tol = FIR_TOL64 if c["kind"].startswith("fir") else IIR_TOL64
            if c["kind"] in ("rc_up", "rc_dn", "interp24", "deci24"):
                tol = 1e-7       # cascade vs the reference's transfer-function form (bar: rtol 1e-4)
            assert err <= tol * scale, (c["name"], c["kind"], c.get("filt"), err, scale)
        n += 1
    assert n > 150
    assert mods[2].launch_count() - launches0 >= n      # the CUDA kernels did the work


def test_reference_fixtures_torch_dtype_preserving(mods, filters):
    """torch.cuda in -> torch.cuda out, same dtype; f32/c64 within the fp32 tolerance."""
    for c in ref_cases():
        x = c["x"]
        if x.dtype not in (np.float32, np.complex64, np.float64, np.complex128):
            continue
        xt = torch.from_numpy(x).cuda()
        y = _run(mods, c, filters, xt)
        assert isinstance(y, torch.Tensor) and y.is_cuda and y.dtype == xt.dtype, c["name"]
        yh = y.cpu().numpy()
        ref = c["y"]
        if c["kind"] in ("upsample", "downsample"):
            assert np.array_equal(yh, ref.astype(x.dtype)), c["name"]
            continue
        err, scale = _maxerr(yh, ref)
        single = x.dtype in (np.float32, np.complex64)
        if c["kind"].startswith("fir"):
            tol = FIR_TOL32 if single else FIR_TOL64
        else:
            tol = IIR_TOL32 if single else IIR_TOL64
        assert err <= tol * scale, (c["name"], c["kind"], c.get("filt"), str(x.dtype), err, scale)


def test_cpu_tensor_in_cpu_tensor_out(mods, filters):
    mrh, ss, _ = mods
    x = torch.randn(5000, dtype=torch.float64)
    y = mrh.multirate_FIR(filters["b101"]).filter(x)
    assert isinstance(y, torch.Tensor) and not y.is_cuda and y.dtype == x.dtype
    err, scale = _maxerr(y.numpy(), oracle.fir_filter(filters["b101"], x.numpy()))
    assert err <= FIR_TOL64 * scale


# ---------------------------------------------------------------- (2) literal known answers
def test_boxcar_fir_known_answer(mods):
    # /root/reference/tests/test_sigsys.py:688-706
    y_test = [1., 1.95105652, 2.76007351, 3.34785876, 3.65687576, 3.65687576, 3.34785876,
              2.76007351, 1.95105652, 1., -1., -2.90211303, -4.52014702, -5.69571753,
              -6.31375151, -6.31375151, -5.69571753, -4.52014702, -2.90211303, -1.]
    x = np.cos(2 * np.pi * 0.05 * np.arange(0, 20))
    npt.assert_almost_equal(mods[0].multirate_FIR(np.ones(10)).filter(x), y_test)


def test_ten_band_biquad_cascade_known_answer(mods, filters):
    # /root/reference/tests/test_sigsys.py:28-34 -- a 10-section cascade (two launch groups)
    y_test = [-4.23769156, 0.097137, 4.18516645, -0.54460053, 2.2257584, 1.60147407,
              -0.76767407, -1.95402381, -1.0580526, 0.9111369]
    y = mods[0].multirate_IIR(filters["sos_tenband"]).filter(filters["tenband_w"])
    npt.assert_almost_equal(y, y_test)


def test_interp24_deci24_known_answers(mods, filters):
    # /root/reference/tests/test_sigsys.py:617-653: literal goldens through upsample + Butterworth chains
    from test_oracle_golden import INTERP24_GOLDEN, DECI24_GOLDEN
    ss = mods[1]
    npt.assert_almost_equal(ss.interp24(filters["mseq2"]), INTERP24_GOLDEN)
    npt.assert_almost_equal(ss.deci24(ss.interp24(filters["mseq3"])), DECI24_GOLDEN)


def test_rate_change_surface(mods):
    """rate_change keeps the reference's constructor / attributes / warning (multirate_helper.py:45-83)."""
    import inspect
    import warnings as _w
    mrh = mods[0]
    assert str(inspect.signature(mrh.rate_change.__init__)) == "(self, M_change=12, fcutoff=0.9, N_filt_order=8, ftype='butter')"
    rc = mrh.rate_change()
    assert rc.M == 12 and rc.fc == 0.45 and rc.N_forder == 8 and len(rc.b) == 9 and len(rc.a) == 9
    with _w.catch_warnings(record=True) as rec:
        _w.simplefilter("always")
        mrh.rate_change(ftype="bessel")
    assert any("butter" in str(r.message) for r in rec)
    x = torch.randn(100000, dtype=torch.float32, device="cuda")
    y = rc.up(x)
    assert y.is_cuda and y.dtype == torch.float32 and y.numel() == 12 * x.numel()
    assert rc.dn(x).numel() == x.numel() // 12


def test_upsample_downsample_reference_tests(mods):
    # /root/reference/tests/test_sigsys.py:655-668
    ss = mods[1]
    npt.assert_equal(ss.upsample(np.zeros(1), 3), np.zeros(3))
    npt.assert_equal(ss.downsample(np.zeros(3), 3), np.zeros(1))
    with pytest.raises(TypeError, match="M must be an int"):
        ss.downsample(np.zeros(0), 3.0)


def test_cfg1_101tap_2e20_float64(mods, filters):
    """BASELINE.json configs[0] against the reference's own output windows."""
    g = np.load(os.path.join(GOLDEN, "cfg1_windows.npz"))
    x = np.random.default_rng(100).standard_normal(2 ** 20)
    y = mods[0].multirate_FIR(filters["b101"]).filter(x)
    W, mid = 4096, 2 ** 19
    for got, ref in ((y[:W], g["head"]), (y[mid:mid + W], g["mid"]), (y[-W:], g["tail"])):
        npt.assert_allclose(got, ref, rtol=1e-6, atol=1e-12)     # north_star bar: rtol 1e-6
        assert np.abs(got - ref).max() <= 1e-12
    npt.assert_allclose([y.sum(), np.abs(y).sum(), (y * y).sum()], g["checksum"], rtol=1e-10)


# ---------------------------------------------------------------- (3) seeded inputs vs oracle
@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 4095, 4096, 4097, 12289, 70001])
@pytest.mark.parametrize("dt", ["float32", "complex64", "float64", "complex128"])
def test_fir_sizes_across_tile_boundaries(mods, filters, n, dt):
    rng = np.random.default_rng(n + 17)
    x = rng.standard_normal(n)
    if "complex" in dt:
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    b = filters["b256"]
    y = mods[0].multirate_FIR(b).filter(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = oracle.fir_filter(b, x, backend="c")
    err, scale = _maxerr(y, ref)
    tol = FIR_TOL32 if dt in ("float32", "complex64") else FIR_TOL64
    assert err <= tol * scale, (n, dt, err, scale)


@pytest.mark.parametrize("ntaps", [1, 2, 3, 31, 32, 33, 64, 100, 511, 1024, 3000])
def test_fir_tap_counts(mods, ntaps):
    rng = np.random.default_rng(ntaps)
    b = rng.standard_normal(ntaps) / np.sqrt(ntaps)
    x = rng.standard_normal(9000)
    fir = mods[0].multirate_FIR(b)
    for L in (1, 3):
        y = fir.filter(x) if L == 1 else fir.up(x[:1500], L)
        ref = oracle.fir_filter(b, x, backend="c") if L == 1 else oracle.fir_up(b, x[:1500], L, backend="c")
        err, scale = _maxerr(y, ref)
        assert err <= FIR_TOL64 * scale, (ntaps, L, err)
    y = fir.dn(x, 5)
    err, scale = _maxerr(y, oracle.fir_dn(b, x, 5, backend="c"))
    assert err <= FIR_TOL64 * scale, (ntaps, err)


@pytest.mark.parametrize("factor", [1, 2, 3, 4, 7, 12, 25])
def test_fir_up_dn_factors(mods, filters, factor):
    rng = np.random.default_rng(factor)
    b = filters["b256"]
    fir = mods[0].multirate_FIR(b)
    for dt in ("float32", "complex128"):
        x = rng.standard_normal(5003)
        if "complex" in dt:
            x = x + 1j * rng.standard_normal(5003)
        x = x.astype(dt)
        xt = torch.from_numpy(x).cuda()
        tol = FIR_TOL32 if dt == "float32" else FIR_TOL64
        yu = fir.up(xt, factor).cpu().numpy()
        err, scale = _maxerr(yu, oracle.fir_up(b, x, factor, backend="c"))
        assert err <= tol * scale, ("up", factor, dt, err, scale)
        yd = fir.dn(xt, factor).cpu().numpy()
        err, scale = _maxerr(yd, oracle.fir_dn(b, x, factor, backend="c"))
        assert err <= tol * scale, ("dn", factor, dt, err, scale)


def test_fir_halo_equals_monolithic(mods, filters):
    """hist = overlap-save halo: chunked == monolithic.  Bit for bit on the CUDA-core kernels
    (fp64, and complex64 below the tensor-core length threshold); on the tensor-core complex64 path
    the tile grid moves with the cut, so the fp32 accumulation order changes: equal to 2e-6*max|y| (each within 1e-6 of the oracle)."""
    from sk_dsp_comm_b200 import _engine
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    for dt, n in ((torch.complex64, 20000), (torch.float64, 50000), (torch.complex64, 300000)):
        x = torch.randn(n, dtype=dt, device="cuda")
        y = _engine.fir_filter(plan, x)
        exact = not (dt == torch.complex64 and n >= 32768)
        scale = y.abs().max().item()

        def same(a, b_):
            if exact:
                return torch.equal(a, b_)
            return (a - b_).abs().max().item() <= 2e-6 * scale      # two fp32 results, each within 1e-6 of the truth

        for cut in (255, 4096, 17777):
            y2 = _engine.fir_filter(plan, x[cut:].contiguous(), hist=x[cut - 255:cut].contiguous())
            assert same(y[cut:], y2), (dt, n, cut)
        # up / dn halos (CUDA-core kernels: exact)
        yu = _engine.fir_up(plan, x, 4)
        cut = 8000
        hl = plan.up_hist_len(4)
        yu2 = _engine.fir_up(plan, x[cut:].contiguous(), 4, hist=x[cut - hl:cut].contiguous())
        assert torch.equal(yu[4 * cut:], yu2)
        yd = _engine.fir_dn(plan, x, 4)
        yd2 = _engine.fir_dn(plan, x[cut:].contiguous(), 4, hist=x[cut - 255:cut].contiguous())
        assert torch.equal(yd[cut // 4:], yd2)


@pytest.mark.parametrize("ntaps", [1, 2, 63, 64, 65, 101, 128, 129, 192, 193, 255, 256])
def test_tensor_core_fir_tap_counts(mods, ntaps):
    """tcgen05 path for every filter length it accepts (<= 256 taps): leading all-zero Toeplitz
    k-blocks are skipped for short filters.  Also non-symmetric taps and the hist (halo) argument."""
    from sk_dsp_comm_b200 import _engine
    rng = np.random.default_rng(ntaps)
    b = rng.standard_normal(ntaps) / np.sqrt(ntaps)
    plan = _engine.FirPlan(b)
    n = 70001
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    xt = torch.from_numpy(x).cuda()
    ref = oracle.fir_filter(b, x.astype(np.complex128), backend="c")
    try:
        for variant in (12, 15, 14, 13):                 # 64 / 80 / 96 / 128-row tiles
            mods[2].lib.b200dsp_set_fir_variant(variant)
            y = _engine.fir_filter(plan, xt).cpu().numpy()
            err, scale = _maxerr(y, ref)
            assert err <= FIR_TOL32 * scale, (ntaps, variant, err, scale)
        if ntaps > 1:
            cut = 4096 * 3
            hist = xt[cut - (ntaps - 1):cut].contiguous()
            y2 = _engine.fir_filter(plan, xt[cut:].contiguous(), hist=hist).cpu().numpy()
            err, scale = _maxerr(y2, ref[cut:])
            assert err <= FIR_TOL32 * scale, (ntaps, "hist", err, scale)
    finally:
        mods[2].lib.b200dsp_set_fir_variant(0)


@pytest.mark.parametrize("fname", ["b256", "b101", "b33_remez_bpf", "b7", "b1"])
def test_tensor_core_f32_filter_up_dn(mods, filters, fname):
    """float32 streams long enough for the tcgen05 kernels (fir_tc_real.cu): filter, up(2..4), dn(2..4);
    combinations the tensor-core path cannot take (too many taps per phase) silently use the CUDA cores."""
    from sk_dsp_comm_b200 import _engine
    b = filters[fname]
    plan = _engine.FirPlan(b)
    rng = np.random.default_rng(len(b))
    for n in (70001, 262144 + 5):
        x = rng.standard_normal(n).astype(np.float32)
        x64 = x.astype(np.float64)
        xt = torch.from_numpy(x).cuda()
        err, scale = _maxerr(_engine.fir_filter(plan, xt).cpu().numpy(), oracle.fir_filter(b, x64, backend="c"))
        assert err <= FIR_TOL32 * scale, (fname, n, "filter", err, scale)
        for F in (2, 3, 4):
            err, scale = _maxerr(_engine.fir_up(plan, xt, F).cpu().numpy(), oracle.fir_up(b, x64, F, backend="c"))
            assert err <= FIR_TOL32 * scale, (fname, n, "up", F, err, scale)
            err, scale = _maxerr(_engine.fir_dn(plan, xt, F).cpu().numpy(), oracle.fir_dn(b, x64, F, backend="c"))
            assert err <= FIR_TOL32 * scale, (fname, n, "dn", F, err, scale)
    # halo argument on the tensor-core paths (cuts on tile boundaries keep the tile grid: bit identical)
    x = torch.randn(300000, dtype=torch.float32, device="cuda")
    cut = 65536
    if len(b) > 1:
        k1 = len(b) - 1
        y = _engine.fir_filter(plan, x)
        assert torch.equal(y[cut:], _engine.fir_filter(plan, x[cut:].contiguous(), hist=x[cut - k1:cut].contiguous()))
        hl = plan.up_hist_len(4)
        yu = _engine.fir_up(plan, x, 4)
        assert torch.equal(yu[4 * cut:], _engine.fir_up(plan, x[cut:].contiguous(), 4, hist=x[cut - hl:cut].contiguous()))
        yd = _engine.fir_dn(plan, x, 4)
        assert torch.equal(yd[cut // 4:], _engine.fir_dn(plan, x[cut:].contiguous(), 4, hist=x[cut - k1:cut].contiguous()))


def test_tensor_core_fir_block_scaling(mods, filters):
    """The tcgen05 path splits fp32 into fp16 hi/lo around a per-tile power-of-two scale: results
    must be scale invariant (1e-30 .. 1e30), survive tiles that are all zero, and handle a stream
    whose amplitude jumps by 2^40 between tiles."""
    from sk_dsp_comm_b200 import _engine
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    rng = np.random.default_rng(11)
    n = 200000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    ref = oracle.fir_filter(b, x.astype(np.complex64).astype(np.complex128), backend="c")
    y1 = _engine.fir_filter(plan, torch.from_numpy(x.astype(np.complex64)).cuda())
    for sc in (2.0 ** -100, 2.0 ** -20, 2.0 ** 30, 2.0 ** 100):
        ys = _engine.fir_filter(plan, torch.from_numpy((x * sc).astype(np.complex64)).cuda())
        assert torch.equal(ys, y1 * sc), sc                  # power-of-two scaling is exact
    err, scale = _maxerr(y1.cpu().numpy(), ref)
    assert err <= FIR_TOL32 * scale
    xz = x.astype(np.complex64).copy()
    xz[50000:120000] = 0
    xz[120000:] *= np.float32(2.0 ** 40)
    yz = _engine.fir_filter(plan, torch.from_numpy(xz).cuda()).cpu().numpy()
    refz = oracle.fir_filter(b, xz.astype(np.complex128), backend="c")
    assert np.all(yz[50300:119900] == 0)
    # per-region tolerance: each region is judged against its own amplitude
    for lo, hi in ((0, 50000), (121000, n)):
        e, s = _maxerr(yz[lo:hi], refz[lo:hi])
        assert e <= FIR_TOL32 * s, (lo, hi, e, s)
    # CUDA-core and tensor-core kernels agree
    mods[2].lib.b200dsp_set_fir_variant(9)
    y_cc = _engine.fir_filter(plan, torch.from_numpy(x.astype(np.complex64)).cuda())
    mods[2].lib.b200dsp_set_fir_variant(0)
    assert (y_cc - y1).abs().max().item() <= 1.5e-6 * scale


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 4097, 65535, 65536, 65537, 131073, 200003])
@pytest.mark.parametrize("dt", ["float32", "complex64", "float64", "complex128"])
def test_sos_sizes_across_tile_boundaries(mods, filters, n, dt):
    rng = np.random.default_rng(n + 3)
    x = rng.standard_normal(n)
    if "complex" in dt:
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    for fname in ("sos6", "sos_butter5"):
        sos = filters[fname]
        y = mods[0].multirate_IIR(sos).filter(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = oracle.sos_filter(sos, x)
        err, scale = _maxerr(y, ref)
        tol = IIR_TOL32 if dt in ("float32", "complex64") else IIR_TOL64
        assert err <= tol * scale, (fname, n, dt, err, scale)


@pytest.mark.parametrize("nsec", [1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 17])
def test_sos_section_counts(mods, nsec):
    """1..17 sections: identity padding of odd counts and chaining of >8-section cascades."""
    rng = np.random.default_rng(nsec)
    sos = np.zeros((nsec, 6))
    for s in range(nsec):
        r, th = rng.uniform(0.5, 0.97), rng.uniform(0.2, 2.8)
        sos[s] = [rng.uniform(0.2, 1.0), rng.uniform(-1, 1), rng.uniform(-1, 1), 1.0,
                  -2 * r * np.cos(th), r * r]
    sos[:, :3] /= np.abs(sos[:, :3]).sum(axis=1, keepdims=True)
    x = rng.standard_normal(40011)
    iir = mods[0].multirate_IIR(sos)
    ref = oracle.sos_filter(sos, x)
    err, scale = _maxerr(iir.filter(x), ref)
    assert err <= IIR_TOL64 * scale, (nsec, err, scale)
    err, scale = _maxerr(iir.dn(x, 3), oracle.sos_dn(sos, x, 3))
    assert err <= IIR_TOL64 * scale, (nsec, "dn", err, scale)
    err, scale = _maxerr(iir.up(x[:5000], 3), oracle.sos_up(sos, x[:5000], 3))
    assert err <= IIR_TOL64 * scale, (nsec, "up", err, scale)


@pytest.mark.parametrize("dt", ["float32", "float64", "complex64"])
def test_sos_up_dn_long_streams(mods, filters, dt):
    """multirate_IIR.up/.dn on streams long enough for the staged full-rate path (>= 2^16 filter-rate
    samples), incl. a 10-section cascade (two launch groups) and the final state zf."""
    from sk_dsp_comm_b200 import _engine
    rng = np.random.default_rng(5)
    n = 100003
    x = rng.standard_normal(n)
    if "complex" in dt:
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    tol = IIR_TOL32 if dt in ("float32", "complex64") else IIR_TOL64
    # the ten-band equaliser (10 sections = two launch groups, poles at radius 0.9988) is checked for the real
    # dtypes: float64 on the scan kernels, float32 on the tensor-core kernel (float64 in-tile scan); a complex64
    # stream still runs the float32 recurrence of the scan kernels, which is only good to ~1e-4 there
    for fname in (("sos6", "sos_tenband") if dt != "complex64" else ("sos6", "sos_butter5")):
        sos = filters[fname]
        iir = mods[0].multirate_IIR(sos)
        xt = torch.from_numpy(x).cuda()
        for F in (4, 12):
            err, scale = _maxerr(iir.dn(xt, F).cpu().numpy(), oracle.sos_dn(sos, x, F))
            assert err <= tol * scale, (fname, dt, "dn", F, err, scale)
            xs = x[:30011]
            err, scale = _maxerr(iir.up(torch.from_numpy(xs).cuda(), F).cpu().numpy(), oracle.sos_up(sos, xs, F))
            assert err <= tol * scale, (fname, dt, "up", F, err, scale)
    if dt == "float64":
        plan = _engine.SosPlan(filters["sos6"])
        _, zf = _engine.sos_filter(plan, torch.from_numpy(x).cuda(), M=4, return_zf=True)
        _, zf_ref = oracle.sos_filter(filters["sos6"], x, return_zf=True)
        npt.assert_allclose(zf.cpu().numpy(), zf_ref, rtol=0, atol=1e-9 * max(1.0, np.abs(zf_ref).max()))


def test_sos_state_carry(mods, filters):
    """zi/zf chaining reproduces the monolithic run (the multi-GPU / streaming hook)."""
    from sk_dsp_comm_b200 import _engine
    for fname, dt in (("sos6", torch.float64), ("sos_tenband", torch.float64), ("sos6", torch.complex128)):
        plan = _engine.SosPlan(filters[fname])
        x = torch.randn(90001, dtype=dt, device="cuda")
        y = _engine.sos_filter(plan, x)
        cut = 33333
        ya, zf = _engine.sos_filter(plan, x[:cut].contiguous(), return_zf=True)
        yb = _engine.sos_filter(plan, x[cut:].contiguous(), zi=zf)
        y2 = torch.cat([ya, yb])
        scale = y.abs().max().item()
        assert (y - y2).abs().max().item() <= 1e-10 * scale, fname
        # zf itself equals the oracle's final state
        xr = x.cpu().numpy()
        if dt == torch.float64:
            _, zf_ref = oracle.sos_filter(filters[fname], xr[:cut], return_zf=True)
            npt.assert_allclose(zf.cpu().numpy(), zf_ref, rtol=0, atol=1e-9 * max(1.0, np.abs(zf_ref).max()))


def test_nd_input_filters_last_axis(mods, filters):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 2, 700))
    y = mods[0].multirate_FIR(filters["b33_remez_bpf"]).filter(x)
    ref = oracle.fir_filter(filters["b33_remez_bpf"], x)
    assert y.shape == x.shape
    npt.assert_allclose(y, ref, rtol=0, atol=1e-12)
    y = mods[0].multirate_IIR(filters["sos6"]).filter(x)
    npt.assert_allclose(y, oracle.sos_filter(filters["sos6"], x), rtol=0, atol=1e-10)


def test_updown_bit_exact_all_dtypes(mods):
    ss = mods[1]
    rng = np.random.default_rng(9)
    for dt in ("int8", "int16", "int32", "int64", "uint8", "float16", "float32", "float64",
               "complex64", "complex128", "bool"):
        if dt == "bool":
            x = rng.integers(0, 2, 1001).astype(bool)
        elif "int" in dt:
            x = rng.integers(0, 100, 1001).astype(dt)
        elif "complex" in dt:
            x = (rng.standard_normal(1001) + 1j * rng.standard_normal(1001)).astype(dt)
        else:
            x = rng.standard_normal(1001).astype(dt)
        for M, p in ((1, 0), (2, 1), (7, 3), (7, -1), (1001, 1000), (2000, 0)):
            y = ss.downsample(x, M, p)
            ref = oracle.downsample(x, M, p)
            assert y.dtype == ref.dtype and np.array_equal(y, ref), (dt, M, p)
        for L in (1, 2, 5, 3.0):
            y = ss.upsample(x, L)
            ref = oracle.upsample(x, L)
            assert y.dtype == ref.dtype and np.array_equal(y, ref), (dt, L)


# ---------------------------------------------------------------- (4) full-size properties
def test_cfg2_full_size_windows_and_properties(mods, filters):
    """BASELINE.json configs[1]: 256 taps over 2^28 complex64.  Windows against the oracle on
    the same values widened to complex128, plus size-independent properties."""
    from sk_dsp_comm_b200 import _engine
    n = 2 ** 28
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    torch.manual_seed(100)
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    y = _engine.fir_filter(plan, x)
    W = 1 << 16
    worst = 0.0
    for start in (0, n // 2 - W // 2, n - W, 4096 * 1000 - 300, 12345678):
        lo = max(start - 255, 0)
        xs = x[lo:start + W].cpu().numpy().astype(np.complex128)
        ref = oracle.fir_filter(b, xs, backend="c")[start - lo:]
        got = y[start:start + W].cpu().numpy()
        err, scale = _maxerr(got, ref)
        worst = max(worst, err / scale)
        assert err <= FIR_TOL32 * scale, (start, err, scale)
    # linearity / homogeneity: FIR(2x) == 2 FIR(x) exactly in binary floating point
    y2 = _engine.fir_filter(plan, x * 2)
    assert torch.equal(y2, y * 2)
    del y2
    # DC gain: a constant stream settles to sum(b) * c after the transient
    c = torch.full((1 << 20,), 1.5 - 0.5j, dtype=torch.complex64, device="cuda")
    yc = _engine.fir_filter(plan, c)[255:].cpu().numpy()
    assert np.abs(yc - (1.5 - 0.5j) * b.sum()).max() <= 1e-6 * 1.6
    # shift invariance across tile boundaries: filtering a delayed copy delays the output
    d = 4099
    xs = torch.zeros(1 << 22, dtype=torch.complex64, device="cuda")
    xs[d:] = x[:(1 << 22) - d]
    ys = _engine.fir_filter(plan, xs)
    # (the tensor-core tile grid is anchored at sample 0, so a shift changes the fp32 summation order)
    assert (ys[d:] - y[:(1 << 22) - d]).abs().max().item() <= 2e-6 * y[:(1 << 22)].abs().max().item()
    print("cfg2 worst window error / max|y| = %.3g" % worst)


def test_cfg3_up4_dn4_2e26_float32(mods, filters):
    from sk_dsp_comm_b200 import _engine
    n = 2 ** 26
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    torch.manual_seed(101)
    x = torch.randn(n, dtype=torch.float32, device="cuda")
    yu = _engine.fir_up(plan, x, 4)
    yd = _engine.fir_dn(plan, x, 4)
    assert yu.numel() == 4 * n and yd.numel() == n // 4
    W = 1 << 15
    for start in (0, n // 2, n - W):
        lo = max(start - 255, 0)
        xs = x[lo:start + W].cpu().numpy().astype(np.float64)
        ref_u = oracle.fir_up(b, xs, 4, backend="c")[4 * (start - lo):]
        err, scale = _maxerr(yu[4 * start:4 * (start + W)].cpu().numpy(), ref_u)
        assert err <= FIR_TOL32 * scale, ("up", start, err, scale)
    # decimation windows: starts that are multiples of M keep phase 0 aligned
    for start in (0, n // 2, n - 4 * W):
        lo = max(start - 256, 0)
        xs = x[lo:start + 4 * W].cpu().numpy().astype(np.float64)
        ref_d = oracle.fir_dn(b, xs, 4, backend="c")[(start - lo) // 4:]
        err, scale = _maxerr(yd[start // 4:start // 4 + W].cpu().numpy(), ref_d)
        assert err <= FIR_TOL32 * scale, ("dn", start, err, scale)
    # identity: decimating the interpolated stream's phase-0 polyphase branch
    assert torch.equal(_engine.downsample(_engine.upsample(x, 4), 4, 0), x)


def test_cfg4_sos6_2e28_float32(mods, filters):
    from sk_dsp_comm_b200 import _engine
    n = 2 ** 28
    sos = filters["sos6"]
    plan = _engine.SosPlan(sos)
    torch.manual_seed(102)
    x = torch.randn(n, dtype=torch.float32, device="cuda")
    y = _engine.sos_filter(plan, x)
    # head window exactly; later windows via warm-up (impulse response < 1e-8 after 1400 samples,
    # SURVEY.md 8e) -- 16384 samples of history make the truncation error negligible vs 1e-4
    W, H = 1 << 16, 1 << 14
    for start in (0, n // 2 + 777, n - W):
        lo = max(start - H, 0)
        xs = x[lo:start + W].cpu().numpy().astype(np.float64)
        ref = oracle.sos_filter(sos, xs)[start - lo:]
        err, scale = _maxerr(y[start:start + W].cpu().numpy(), ref)
        assert err <= IIR_TOL32 * scale, (start, err, scale)
    # homogeneity is exact for power-of-two gains
    y2 = _engine.sos_filter(plan, x * 4)
    assert torch.equal(y2, y * 4)


def test_host_pipeline_equals_device_path(mods, filters):
    """Chunked H2D|kernel|D2H pipeline (hostpipe.py) vs the monolithic device call.  The halo makes
    the chunked FILTER exact; on the tensor-core complex64 path the tile grid restarts at every chunk,
    so fp32 summation order differs (both within 1e-6*max|y| of the oracle); bit for bit on float64."""
    from sk_dsp_comm_b200 import _engine, hostpipe
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    torch.manual_seed(7)
    x = torch.randn((1 << 22) + 12345, dtype=torch.complex64).pin_memory()
    y_host = hostpipe.fir_filter_host(plan, x, chunk=1 << 20)
    y_dev = _engine.fir_filter(plan, x.cuda()).cpu()
    scale = y_dev.abs().max().item()
    # two fp32 evaluations with different tile grids: each is within 1e-6 of the float64 truth
    assert (y_host - y_dev).abs().max().item() <= 2e-6 * scale
    W = 1 << 15
    for start in (0, (1 << 20) - 100, (3 << 20) + 5):          # windows straddling chunk boundaries
        lo = max(start - 255, 0)
        ref = oracle.fir_filter(b, x[lo:start + W].numpy().astype(np.complex128), backend="c")[start - lo:]
        err, sc = _maxerr(y_host[start:start + W].numpy(), ref)
        assert err <= FIR_TOL32 * sc, (start, err, sc)
    # public API routes long host tensors through the same pipeline; a brand-new filter object whose very
    # first launch happens on a side stream must see its freshly uploaded tap matrices (regression: the
    # plan upload is now synchronised before the plan is handed out)
    for _ in range(3):
        y_api = mods[0].multirate_FIR(b).filter(x)
        assert (y_api - y_dev).abs().max().item() <= 2e-6 * scale
    x64 = torch.randn((1 << 20) + 77, dtype=torch.float64).pin_memory()
    y_host = hostpipe.fir_filter_host(plan, x64, chunk=1 << 18)
    assert torch.equal(y_host, _engine.fir_filter(plan, x64.cuda()).cpu())


# ---------------------------------------------------------------- short-phase up kernel (pulse shaping shapes)

@pytest.mark.parametrize("dt", ["float32", "complex64", "float64", "complex128"])
@pytest.mark.parametrize("fname,L", [("b33_remez_bpf", 4), ("b101", 8), ("b101", 12), ("b256", 8),
                                     ("b256", 16), ("b7", 2), ("b7", 8), ("b1", 4), ("b101", 6)])
def test_fir_up_short_phase_kernel(filters, fname, L, dt):
    """up(L) with few taps per phase (fir_up_short_kernel where L*sizeof(sample) % 16 == 0, else the
    polyphase kernel): oracle parity, overlap-save halo == monolithic bit for bit, and agreement with the
    polyphase kernel forced through variant 8."""
    from sk_dsp_comm_b200 import _engine, _cabi
    b = filters[fname]
    plan = _engine.FirPlan(b)
    rng = np.random.default_rng(L * 131 + len(b))
    n = 10007
    x = rng.standard_normal(n)
    if "complex" in dt:
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    xt = torch.from_numpy(x).cuda()
    tol = FIR_TOL32 if dt in ("float32", "complex64") else FIR_TOL64
    y = _engine.fir_up(plan, xt, L)
    assert y.shape == (n * L,)
    err, scale = _maxerr(y.cpu().numpy(), oracle.fir_up(b, x, L, backend="c"))
    assert err <= tol * scale, (fname, L, dt, err, scale)
    hl = plan.up_hist_len(L)
    for cut in (hl, 256, 4001):
        if cut < hl:
            continue
        y2 = _engine.fir_up(plan, xt[cut:].contiguous(), L, hist=xt[cut - hl:cut].contiguous())
        assert torch.equal(y[L * cut:], y2), (fname, L, dt, cut)
    _cabi.lib.b200dsp_set_fir_variant(8)
    try:
        yp = _engine.fir_up(plan, xt, L)
    finally:
        _cabi.lib.b200dsp_set_fir_variant(0)
    assert float((yp - y).abs().max()) <= 2 * tol * scale


def test_fir_up_short_phase_long_stream(filters):
    """2^22 complex64 symbols, 8 samples/symbol, 97-tap pulse: whole output against the oracle."""
    from sk_dsp_comm_b200 import _engine
    rng = np.random.default_rng(77)
    n, L = (1 << 22) + 3, 8
    k = np.arange(-48, 49) / 8.0
    b = np.sinc(k) * np.hamming(97)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    y = _engine.fir_up(_engine.FirPlan(b), torch.from_numpy(x).cuda(), L).cpu().numpy()
    err, scale = _maxerr(y, oracle.fir_up(b, x, L, backend="c"))
    assert err <= FIR_TOL32 * scale, (err, scale)
