"""GPU parity of the single-pass tensor-core SOS cascade kernel (csrc/sos_tc.cu; float32 streams of
multirate_IIR.filter/.up/.dn, reference src/sk_dsp_comm/multirate_helper.py:169-192) against the CPU oracle
(scipy.signal.sosfilt's arithmetic in float64).  Bar: |d| <= 1e-4 * max|y_ref| (north_star, IIR); the kernel is
asserted at 5e-6 (measured 2e-7 ... 4e-6).  The kernel is forced with b200dsp_set_sos_variant(2) so that sizes
below the automatic threshold exercise it too."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
TOL = 5e-6


@pytest.fixture(scope="module")
def eng():
    from sk_dsp_comm_b200 import _engine, _cabi
    return _engine, _cabi.lib


def _rel(y, ref):
    return float(np.abs(np.asarray(y, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-300))


def _tc(eng, sos, x, **kw):
    _engine, lib = eng
    lib.b200dsp_set_sos_variant(2)
    try:
        out = _engine.sos_filter(_engine.SosPlan(sos), torch.from_numpy(x).cuda(), **kw)
        torch.cuda.synchronize()
    finally:
        lib.b200dsp_set_sos_variant(0)
    return out


@pytest.mark.parametrize("fname", ["sos6", "sos_sharp_lpf", "sos_butter6", "sos_butter5", "sos_tenband"])
@pytest.mark.parametrize("n", [1, 127, 8192, 8193, 3 * 8192 + 17, 200003, (1 << 21) + 12345])
def test_filter_sizes_and_cascades(eng, filters, fname, n):
    """one tile, partial tiles, several blocks with state warm-up; 3 / 6 / 10 sections (10 = two launch groups,
    the ten-band equaliser with poles at radius 0.9988 runs its in-tile scan in float64)"""
    sos = filters[fname]
    x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
    y, zf = _tc(eng, sos, x, return_zf=True)
    ref, zref = oracle.sos_filter(sos, x.astype(np.float64), return_zf=True)
    assert _rel(y.cpu().numpy(), ref) <= TOL, (fname, n)
    assert np.abs(zf.cpu().numpy() - zref).max() <= 2e-5 * max(np.abs(zref).max(), np.abs(ref).max()), (fname, n)


@pytest.mark.parametrize("fname", ["sos6", "sos_tenband"])
def test_state_carry(eng, filters, fname):
    """zi / zf chaining reproduces the monolithic run (streaming and multi-GPU hook)"""
    sos = filters[fname]
    rng = np.random.default_rng(7)
    x = rng.standard_normal(300007).astype(np.float32)
    ref = oracle.sos_filter(sos, x.astype(np.float64))
    cut = 123457
    ya, zf = _tc(eng, sos, x[:cut], return_zf=True)
    yb = _tc(eng, sos, x[cut:].copy(), zi=zf)
    y = np.concatenate([ya.cpu().numpy(), yb.cpu().numpy()])
    # the carried state itself is float32 (scipy's zi/zf dtype for a float32 stream): on the ten-band cascade a
    # 6e-8 rounding of the state is worth ~1e-5 of the output, so the chained run is held to the 1e-4 IIR bar
    assert _rel(y, ref) <= (TOL if fname == "sos6" else 1e-4)
    # an arbitrary but natural initial state: where another signal left the cascade
    _, zi = oracle.sos_filter(sos, rng.standard_normal(50000), return_zf=True)
    zi = (zi * 3.0).astype(np.float32)
    y2 = _tc(eng, sos, x[:70001], zi=torch.from_numpy(zi))
    ref2 = oracle.sos_filter(sos, x[:70001].astype(np.float64), zi=zi.astype(np.float64))
    assert _rel(y2.cpu().numpy(), ref2) <= (TOL if fname == "sos6" else 1e-4), _rel(y2.cpu().numpy(), ref2)


@pytest.mark.parametrize("F", [2, 3, 4, 12])
def test_up_dn_fused(eng, filters, F):
    """zero stuffing fused into the tile load, decimation into the stores (multirate_IIR.up / .dn)"""
    sos = filters["sos6"]
    rng = np.random.default_rng(F)
    x = rng.standard_normal(100003).astype(np.float32)
    y = _tc(eng, sos, x, L=F)
    assert _rel(y.cpu().numpy(), oracle.sos_up(sos, x.astype(np.float64), F)) <= TOL
    x = rng.standard_normal((1 << 20) + 5).astype(np.float32)
    y = _tc(eng, sos, x, M=F)
    assert _rel(y.cpu().numpy(), oracle.sos_dn(sos, x.astype(np.float64), F)) <= TOL
    # two launch groups + decimation
    x = rng.standard_normal(150001).astype(np.float32)
    y = _tc(eng, filters["sos_tenband"], x, M=F)
    assert _rel(y.cpu().numpy(), oracle.sos_dn(filters["sos_tenband"], x.astype(np.float64), F)) <= TOL


def test_block_scale_window(eng, filters):
    """loud burst, then near silence: the chunk start states are far larger than the current samples; the block
    scale looks at the cascade's memory, so the fp16 operands neither overflow nor lose the ringing.  Also scale
    invariance over 2^+-40."""
    sos = filters["sos6"]
    rng = np.random.default_rng(11)
    n = 20 * 8192
    x = (rng.standard_normal(n) * 1e-6).astype(np.float32)
    x[3 * 8192 - 500:3 * 8192] += (rng.standard_normal(500) * 50.0).astype(np.float32)
    x[9 * 8192:10 * 8192] = 0.0
    ref = oracle.sos_filter(sos, x.astype(np.float64))
    y = _tc(eng, sos, x).cpu().numpy().astype(np.float64)
    assert np.isfinite(y).all()
    assert _rel(y, ref) <= TOL
    # the quiet tail (after the burst has rung out) is resolved relative to ITS level, not the burst's
    tail = slice(12 * 8192, n)
    assert np.abs(y[tail] - ref[tail]).max() <= 1e-4 * np.abs(ref[tail]).max()
    for e in (-40, 40):
        xs = (x.astype(np.float64) * 2.0 ** e).astype(np.float32)
        ys = _tc(eng, sos, xs).cpu().numpy().astype(np.float64) * 2.0 ** (-e)
        assert _rel(ys, ref) <= TOL, e


@pytest.mark.parametrize("nsec", [1, 2, 3, 5, 8, 9, 13, 17])
def test_section_counts(eng, nsec):
    """1..17 random sections: state padding to multiples of 4, chaining of 8-section groups through the workspace"""
    rng = np.random.default_rng(nsec)
    sos = np.zeros((nsec, 6))
    for s in range(nsec):
        r, th = rng.uniform(0.5, 0.97), rng.uniform(0.2, 2.8)
        sos[s] = [rng.uniform(0.2, 1.0), rng.uniform(-1, 1), rng.uniform(-1, 1), 1.0, -2 * r * np.cos(th), r * r]
    sos[:, :3] /= np.abs(sos[:, :3]).sum(axis=1, keepdims=True)
    x = rng.standard_normal(70001).astype(np.float32)
    y = _tc(eng, sos, x)
    assert _rel(y.cpu().numpy(), oracle.sos_filter(sos, x.astype(np.float64))) <= TOL, nsec


def test_unaligned_and_auto_path(eng, filters):
    """a stream that does not start on a 16-byte boundary (no bulk TMA: the converters fetch the tiles), and the
    automatic kernel choice on a long stream (tensor-core kernel) against the scan kernels"""
    _engine, lib = eng
    sos = filters["sos6"]
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1 << 20) + 9).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    lib.b200dsp_set_sos_variant(2)
    try:
        y = _engine.sos_filter(_engine.SosPlan(sos), xt[1:])
    finally:
        lib.b200dsp_set_sos_variant(0)
    assert _rel(y.cpu().numpy(), oracle.sos_filter(sos, x[1:].astype(np.float64))) <= TOL
    n = 1 << 24
    xl = torch.randn(n, dtype=torch.float32, device="cuda")
    plan = _engine.SosPlan(sos)
    y_auto = _engine.sos_filter(plan, xl)
    lib.b200dsp_set_sos_variant(1)
    try:
        y_scan = _engine.sos_filter(plan, xl)
    finally:
        lib.b200dsp_set_sos_variant(0)
    scale = y_scan.abs().max().item()
    assert (y_auto - y_scan).abs().max().item() <= 2e-6 * scale
    w = slice(n - (1 << 20), n)
    ref = oracle.sos_filter(sos, xl[w].cpu().numpy().astype(np.float64))
    assert _rel(y_auto[w].cpu().numpy()[1 << 16:], ref[1 << 16:]) <= TOL


def test_tenband_float32_meets_the_iir_bar(filters):
    """the cascade the round-1 float32 scan kernels missed (1.1e-4): through the public API, default kernel choice"""
    import sk_dsp_comm_b200.multirate_helper as mrh
    sos = filters["sos_tenband"]
    x = np.random.default_rng(5).standard_normal(100003).astype(np.float32)
    y = mrh.multirate_IIR(sos).filter(torch.from_numpy(x).cuda()).cpu().numpy()
    assert _rel(y, oracle.sos_filter(sos, x.astype(np.float64))) <= 1e-5
