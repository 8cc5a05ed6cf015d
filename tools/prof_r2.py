"""Minimal drivers for the round-2 ncu captures: python tools/prof_r2.py <which>  (a few launches of one kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine, _cabi
F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
which = sys.argv[1]
dev = "cuda"
if which == "fir_c64":
    plan = _engine.FirPlan(F["b256"]); x = torch.randn(1 << 28, dtype=torch.complex64, device=dev)
    fn = lambda: _engine.fir_filter(plan, x)
elif which == "sos_f32":
    plan = _engine.SosPlan(F["sos6"]); x = torch.randn(1 << 28, dtype=torch.float32, device=dev)
    fn = lambda: _engine.sos_filter(plan, x)
elif which == "sos_f64":
    plan = _engine.SosPlan(F["sos6"]); x = torch.randn(1 << 26, dtype=torch.float64, device=dev)
    fn = lambda: _engine.sos_filter(plan, x)
elif which == "fir_f64":
    plan = _engine.FirPlan(F["b256"]); x = torch.randn(1 << 26, dtype=torch.float64, device=dev)
    fn = lambda: _engine.fir_filter(plan, x)
elif which == "pulse_c64":
    from sk_dsp_comm_b200 import _pulse
    b = np.hanning(97) / 8.0
    plan = _engine.FirPlan(b); x = torch.randn(1 << 24, dtype=torch.complex64, device=dev)
    fn = lambda: _engine.fir_up(plan, x, 8)
elif which == "upsample":
    x = torch.randn(1 << 26, dtype=torch.float32, device=dev)
    fn = lambda: _engine.upsample(x, 4)
elif which == "downsample":
    x = torch.randn(1 << 26, dtype=torch.float32, device=dev)
    fn = lambda: _engine.downsample(x, 4)
elif which == "fir_up12_c64":
    plan = _engine.FirPlan(F["b256"]); x = torch.randn(1 << 23, dtype=torch.complex64, device=dev)
    fn = lambda: _engine.fir_up(plan, x, 12)
elif which == "fir_dn12_f32":
    plan = _engine.FirPlan(F["b256"]); x = torch.randn(1 << 26, dtype=torch.float32, device=dev)
    fn = lambda: _engine.fir_dn(plan, x, 12)
elif which == "fir_fft_c64":
    rng = np.random.default_rng(0)
    plan = _engine.FirPlan(rng.standard_normal(1024) / 32); x = torch.randn(1 << 26, dtype=torch.complex64, device=dev)
    fn = lambda: _engine.fir_filter(plan, x)
else:
    raise SystemExit("unknown " + which)
for _ in range(4):
    y = fn()
torch.cuda.synchronize()
