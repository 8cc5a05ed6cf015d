"""GPU tests of the round-2 additions to the FIR side of the boundary:
  * decimation factors / filter lengths whose staging exceeds shared memory (unstaged kernel),
  * complex taps (scipy.signal.lfilter accepts them; two real-tap passes + device-side combination),
  * N-D input through the batched C entries,
  * plans are complete after *_plan_create: a first call can be captured in a CUDA graph,
  * element-wise precision report for the float32 / complex64 kernels (a strict allclose(rtol=1e-6) census)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mrh():
    import sk_dsp_comm_b200.multirate_helper as m
    return m


def _rel(y, ref):
    return float(np.abs(np.asarray(y) - ref).max() / max(np.abs(ref).max(), 1e-300))


@pytest.mark.parametrize("M", [32, 64, 100, 257])
@pytest.mark.parametrize("dt", ["float32", "complex64", "float64", "complex128"])
def test_dn_any_factor(mrh, filters, M, dt):
    """the reference accepts any M (multirate_helper.py:121-127); large factors overflow the polyphase kernel's
    shared-memory tile and take the unstaged kernel"""
    rng = np.random.default_rng(M)
    n = 40000 + M
    x = rng.standard_normal(n)
    if "complex" in dt:
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    fir = mrh.multirate_FIR(filters["b256"])
    y = fir.dn(torch.from_numpy(x).cuda(), M).cpu().numpy()
    ref = oracle.fir_dn(filters["b256"], x, M)
    assert y.shape == ref.shape
    assert _rel(y, ref) <= (1e-6 if dt in ("float32", "complex64") else 1e-11), (M, dt)


def test_very_long_filter(mrh):
    """20001 taps in float64: beyond the shared-memory tile of the polyphase kernel"""
    rng = np.random.default_rng(0)
    b = rng.standard_normal(20001) / 100.0
    x = rng.standard_normal(50000)
    y = mrh.multirate_FIR(b).filter(x)
    assert _rel(y, oracle.fir_filter(b, x)) <= 1e-11
    yu = mrh.multirate_FIR(b).up(x[:3000], 3)
    assert _rel(yu, oracle.fir_up(b, x[:3000], 3)) <= 1e-11


@pytest.mark.parametrize("dt", ["float32", "complex64", "float64", "complex128"])
def test_complex_taps(mrh, filters, dt):
    rng = np.random.default_rng(3)
    b = filters["b101"] * np.exp(2j * np.pi * 0.11 * np.arange(101))          # a frequency-shifted (analytic) lowpass
    n = 70001
    x = rng.standard_normal(n)
    if "complex" in dt:
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    fir = mrh.multirate_FIR(b)
    tol = 2e-6 if dt in ("float32", "complex64") else 1e-11
    xt = torch.from_numpy(x).cuda()
    y = fir.filter(xt)
    assert y.is_complex()
    assert _rel(y.cpu().numpy(), oracle.fir_filter(b, x)) <= tol
    assert _rel(fir.up(xt[:9001].contiguous(), 4).cpu().numpy(), oracle.fir_up(b, x[:9001], 4)) <= tol
    assert _rel(fir.dn(xt, 5).cpu().numpy(), oracle.fir_dn(b, x, 5)) <= tol
    # numpy in -> numpy out in the reference's dtype
    yn = fir.filter(x[:5000])
    assert yn.dtype == np.complex128
    assert _rel(yn, oracle.fir_filter(b, x[:5000])) <= 1e-11


def test_nd_batched(mrh, filters):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5, 3, 40000)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    y = mrh.multirate_FIR(filters["b256"]).filter(xt)
    assert y.shape == xt.shape
    assert _rel(y.cpu().numpy(), oracle.fir_filter(filters["b256"], x.astype(np.float64))) <= 1e-6
    y = mrh.multirate_IIR(filters["sos6"]).filter(xt)
    assert _rel(y.cpu().numpy(), oracle.sos_filter(filters["sos6"], x.astype(np.float64))) <= 1e-4


def test_first_call_inside_graph_capture(filters):
    """nothing is allocated or synchronised inside a filter call, not even the first one of a plan"""
    from sk_dsp_comm_b200 import _engine
    plan = _engine.FirPlan(filters["b256"])
    splan = _engine.SosPlan(filters["sos6"])
    x = torch.randn(1 << 20, dtype=torch.float32, device="cuda")
    plan.handle(0 if x.device.index is None else x.device.index)
    splan.handle(0 if x.device.index is None else x.device.index)
    yu = torch.empty(4 << 20, dtype=torch.float32, device="cuda")
    yd = torch.empty((1 << 20) // 4, dtype=torch.float32, device="cuda")
    yf = torch.empty_like(x)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            _engine.fir_up(plan, x, 4, out=yu)                 # tensor-core up(4): first use of that mode
            _engine.fir_dn(plan, x, 4, out=yd)
            _engine.fir_filter(plan, x, out=yf)
    g.replay()
    torch.cuda.synchronize()
    xs = x.cpu().numpy().astype(np.float64)
    assert _rel(yf.cpu().numpy(), oracle.fir_filter(filters["b256"], xs, backend="c")) <= 1e-6
    assert _rel(yu.cpu().numpy(), oracle.fir_up(filters["b256"], xs, 4, backend="c")) <= 1e-6
    assert _rel(yd.cpu().numpy(), oracle.fir_dn(filters["b256"], xs, 4, backend="c")) <= 1e-6


def test_elementwise_precision_census(filters):
    """How many samples of the float32-class kernels miss a STRICT element-wise allclose(rtol=1e-6, atol=1e-6 rms)
    against the float64 oracle?  (The parity bar is max-norm: |d| <= 1e-6 max|y|.)  The census is printed and, on
    the GPU box, written to gpurun_out/ so it can be kept under profiles/."""
    from sk_dsp_comm_b200 import _engine, _cabi
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    n = 1 << 22
    torch.manual_seed(100)
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    ref = oracle.fir_filter(b, x.cpu().numpy().astype(np.complex128), backend="c")
    rms = float(np.sqrt(np.mean(np.abs(ref) ** 2)))
    out = {"n": n, "rtol": 1e-6, "atol": "1e-6 * rms(y_ref)", "rms": rms, "max": float(np.abs(ref).max())}
    for name, variant in (("tcgen05 (fir_tc2_kernel)", 0), ("CUDA cores (fir_poly_kernel)", 9)):
        _cabi.lib.b200dsp_set_fir_variant(variant)
        try:
            y = _engine.fir_filter(plan, x).cpu().numpy()
        finally:
            _cabi.lib.b200dsp_set_fir_variant(0)
        d = np.abs(y - ref)
        fail = d > (1e-6 * rms + 1e-6 * np.abs(ref))
        out[name] = {"frac_failing_strict_allclose": float(fail.mean()), "max_err_over_max": float(d.max() / np.abs(ref).max()),
                     "rms_err_over_rms": float(np.sqrt(np.mean(d ** 2)) / rms),
                     "frac_failing_rtol_only": float((d > 1e-6 * np.abs(ref)).mean())}
        assert out[name]["max_err_over_max"] <= 1e-6
        assert out[name]["frac_failing_strict_allclose"] <= 0.02
    # float32 SOS cascade (tensor-core kernel) against the 1e-4 IIR bar, element-wise
    sos = filters["sos6"]
    xr = torch.randn(n, dtype=torch.float32, device="cuda")
    refs = oracle.sos_filter(sos, xr.cpu().numpy().astype(np.float64))
    ys = _engine.sos_filter(_engine.SosPlan(sos), xr).cpu().numpy()
    ds = np.abs(ys - refs)
    rmss = float(np.sqrt(np.mean(refs ** 2)))
    out["sos_tc_kernel (float32, 6 sections)"] = {
        "frac_failing_allclose_rtol1e-4_atol1e-4rms": float((ds > 1e-4 * rmss + 1e-4 * np.abs(refs)).mean()),
        "frac_failing_allclose_rtol1e-6_atol1e-6rms": float((ds > 1e-6 * rmss + 1e-6 * np.abs(refs)).mean()),
        "max_err_over_max": float(ds.max() / np.abs(refs).max())}
    assert out["sos_tc_kernel (float32, 6 sections)"]["frac_failing_allclose_rtol1e-4_atol1e-4rms"] == 0.0
    print(json.dumps(out, indent=1))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/precision_census.json", "w"), indent=1)
    except OSError:
        pass


def _lowpass(K, fc=0.1):
    n = np.arange(K) - (K - 1) / 2
    return np.sinc(2 * fc * n) * np.kaiser(K, 8.0) * 2 * fc


@pytest.mark.parametrize("dt", ["complex64", "float32"])
@pytest.mark.parametrize("K", [2, 33, 257, 1024, 2049])
def test_fft_overlap_save_kernel(K, dt):
    """csrc/fir_fft.cu (4096-point FFT in shared memory, overlap-save) against the oracle: forced for any filter
    it can take (variant 16), with and without a history, on lengths around the frame boundaries"""
    from sk_dsp_comm_b200 import _engine, _cabi
    rng = np.random.default_rng(K)
    b = _lowpass(K) if K > 2 else np.array([0.7, -0.2])
    plan = _engine.FirPlan(b)
    valid = 4096 - (K - 1)
    for n in (1, valid - 1, valid, 2 * valid + 1, 40000, 123457):
        x = rng.standard_normal(n)
        hist = rng.standard_normal(K - 1)
        if dt == "complex64":
            x = x + 1j * rng.standard_normal(n)
            hist = hist + 1j * rng.standard_normal(K - 1)
        x, hist = x.astype(dt), hist.astype(dt)
        wide = np.complex128 if dt == "complex64" else np.float64
        _cabi.lib.b200dsp_set_fir_variant(16)
        try:
            y = _engine.fir_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy()
            yh = _engine.fir_filter(plan, torch.from_numpy(x).cuda(), hist=torch.from_numpy(hist).cuda()).cpu().numpy()
        finally:
            _cabi.lib.b200dsp_set_fir_variant(0)
        assert y.dtype == np.dtype(dt) and y.shape == (n,)
        # a transform's error scales with the frame, not with the local output: a stream shorter than the filter
        # and without history is only the filter's leading tail (1e-5 of the passband gain), so the bar there is
        # taken against the filter's full-overlap output level instead
        ref = oracle.fir_filter(b, x.astype(wide))
        floor = np.abs(b).sum() * np.abs(x).max() if n < K else 0.0
        assert float(np.abs(y - ref).max()) <= 1e-6 * max(np.abs(ref).max(), floor), (K, n)
        assert _rel(yh, oracle.fir_filter(b, x.astype(wide), hist=hist.astype(wide))) <= 1e-6, (K, n)


@pytest.mark.parametrize("dt", ["complex64", "float32"])
def test_long_filter_takes_the_fft_kernel(mrh, dt):
    """multirate_FIR.filter with more than 256 taps on a long float32 / complex64 stream: the automatic choice is
    the overlap-save kernel (bit-identical to forcing it) and the result holds the 1e-6 bar"""
    from sk_dsp_comm_b200 import _engine, _cabi
    rng = np.random.default_rng(7)
    b = _lowpass(1024)
    n = 1 << 20
    x = rng.standard_normal(n)
    if dt == "complex64":
        x = x + 1j * rng.standard_normal(n)
    x = x.astype(dt)
    y = mrh.multirate_FIR(b).filter(torch.from_numpy(x).cuda()).cpu().numpy()     # torch input keeps its dtype
    assert y.dtype == np.dtype(dt)
    plan = _engine.FirPlan(b)
    _cabi.lib.b200dsp_set_fir_variant(16)
    try:
        yf = _engine.fir_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy()
    finally:
        _cabi.lib.b200dsp_set_fir_variant(0)
    assert np.array_equal(np.asarray(y), yf)
    ref = oracle.fir_filter(b, x[:200000].astype(np.complex128 if dt == "complex64" else np.float64))
    assert _rel(np.asarray(y)[:200000], ref) <= 1e-6


@pytest.mark.parametrize("dt", ["float32", "complex64"])
@pytest.mark.parametrize("M", [2, 3, 5, 12, 13])
def test_dn_decimating_tensor_core_filter(filters, dt, M):
    """long float32 / complex64 dn(M) the phase-stream kernel does not take runs the tensor-core filter kernel with
    decimating stores (csrc/fir_tc2.cu DEC, csrc/fir_tc_real.cu <MODE_FILTER, 0>): oracle parity on lengths around
    the tile and factor boundaries, and agreement with the CUDA-core polyphase kernel (variant 9)"""
    from sk_dsp_comm_b200 import _engine, _cabi
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    rng = np.random.default_rng(M)
    for n in (32768, 6144 * 7 + 1, 200003, 12 * 4096 * 5):
        x = rng.standard_normal(n)
        if dt == "complex64":
            x = x + 1j * rng.standard_normal(n)
        x = x.astype(dt)
        xt = torch.from_numpy(x).cuda()
        y = _engine.fir_dn(plan, xt, M).cpu().numpy()
        _cabi.lib.b200dsp_set_fir_variant(9)
        try:
            y9 = _engine.fir_dn(plan, xt, M).cpu().numpy()
        finally:
            _cabi.lib.b200dsp_set_fir_variant(0)
        ref = oracle.fir_dn(b, x.astype(np.complex128 if dt == "complex64" else np.float64), M)
        assert y.shape == ref.shape == y9.shape and y.dtype == np.dtype(dt)
        assert _rel(y, ref) <= 1e-6, (M, n)
        assert _rel(y, y9) <= 2e-6, (M, n)


@pytest.mark.parametrize("L", [2, 3, 4, 5, 7])
def test_up_complex64_tensor_core_filter(filters, L):
    """long complex64 up(L) with more than 32 taps per phase runs the tensor-core filter kernel on the zero-stuffed
    stream, the converter warps doing the stuffing (csrc/fir_tc2.cu MODE 2): oracle parity with and without history
    on lengths around the tile boundaries, agreement with the CUDA-core polyphase kernel (variant 9)"""
    from sk_dsp_comm_b200 import _engine, _cabi
    b = filters["b256"]
    plan = _engine.FirPlan(b)
    rng = np.random.default_rng(L)
    nh = plan.up_hist_len(L)
    for n in (32768 // L + 1, 6144 * 3 // L + 5, 50001, 6144 * 4):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        hist = (rng.standard_normal(nh) + 1j * rng.standard_normal(nh)).astype(np.complex64)
        xt = torch.from_numpy(x).cuda()
        y = _engine.fir_up(plan, xt, L).cpu().numpy()
        yh = _engine.fir_up(plan, xt, L, hist=torch.from_numpy(hist).cuda()).cpu().numpy()
        _cabi.lib.b200dsp_set_fir_variant(9)
        try:
            y9 = _engine.fir_up(plan, xt, L).cpu().numpy()
        finally:
            _cabi.lib.b200dsp_set_fir_variant(0)
        ref = oracle.fir_up(b, x.astype(np.complex128), L)
        xe = np.concatenate([hist, x]).astype(np.complex128)
        refh = oracle.fir_up(b, xe, L)[nh * L:]
        assert y.shape == ref.shape and y.dtype == np.complex64
        assert _rel(y, ref) <= 1e-6, (L, n)
        assert _rel(yh, refh) <= 1e-6, (L, n)
        assert _rel(y, y9) <= 2e-6, (L, n)
