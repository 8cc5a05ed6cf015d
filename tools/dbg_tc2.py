"""TC kernel: error structure (gain vs noise), k-block order effect, bottleneck flags."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
b = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"]
plan = _engine.FirPlan(b)
_cabi.lib.b200dsp_set_fir_variant(int(os.environ.get('B200DSP_VARIANT', '10')))
what = sys.argv[1]
if what == "err":
    n = 1 << 21
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = oracle.fir_filter(b, x.astype(np.complex128), backend="c")
    y = _engine.fir_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy().astype(np.complex128)
    e = y - ref
    alpha = np.vdot(ref, e) / np.vdot(ref, ref)
    res = e - alpha * ref
    print("order", os.environ.get("B200DSP_TC_ORDER"), "max %.3g rms %.3g | gain alpha %.3g%+.3gj | residual max %.3g rms %.3g (all / max|y|, rms / rms y)" % (
        np.abs(e).max() / np.abs(ref).max(), np.sqrt((np.abs(e) ** 2).mean() / (np.abs(ref) ** 2).mean()), alpha.real, alpha.imag,
        np.abs(res).max() / np.abs(ref).max(), np.sqrt((np.abs(res) ** 2).mean() / (np.abs(ref) ** 2).mean())), flush=True)
    ec = np.abs(e[:(n // 64) * 64]).reshape(-1, 64).max(axis=0) / np.abs(ref).max()
    print("  max error by output position mod 64 (x1e-7):", " ".join("%.0f" % (v * 1e7) for v in ec), flush=True)
    # fp32 CUDA-core kernel for comparison
    _cabi.lib.b200dsp_set_fir_variant(0)
    y0 = _engine.fir_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy().astype(np.complex128)
    e0 = y0 - ref
    print("  cuda-core fp32: max %.3g rms %.3g" % (np.abs(e0).max() / np.abs(ref).max(), np.sqrt((np.abs(e0) ** 2).mean() / (np.abs(ref) ** 2).mean())))
else:
    n = 1 << 28
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    for _ in range(3): _engine.fir_filter(plan, x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): _engine.fir_filter(plan, x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("dbg", os.environ.get("B200DSP_TC_DBG"), "2^28: %.3f ms  %.1f GS/s" % (ms, n / ms / 1e6), flush=True)
