"""up(12) / dn(12) -- the reference's default rate factors (multirate_helper.py:112,121) -- on 256 taps: CUDA-event timing
against the measured HBM peak (algorithmic bytes: up s(1+L), dn s(1+1/M) per input sample)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine
F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
PEAK = 6542.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
plan = _engine.FirPlan(F["b256"])
out = []
def timed(name, fn, nbytes):
    for _ in range(3): y = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): y = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out.append({"config": name, "ms": ms, "frac": nbytes / ms / 1e6 / PEAK})
    print("%-44s %.4f ms  %.3f of HBM" % (name, ms, nbytes / ms / 1e6 / PEAK), flush=True)
for dt, s in ((torch.float32, 4), (torch.complex64, 8)):
    nu = (1 << 27) // 12 // 4 * 4 if dt == torch.float32 else (1 << 26) // 12 // 4 * 4
    xu = torch.randn(nu, dtype=dt, device="cuda")
    timed("up(12) 256 taps %s %d in" % (str(dt)[6:], nu), lambda: _engine.fir_up(plan, xu, 12), s * 13 * nu)
    del xu
    nd = 1 << 27 if dt == torch.float32 else 1 << 26
    xd = torch.randn(nd, dtype=dt, device="cuda")
    timed("dn(12) 256 taps %s %d in" % (str(dt)[6:], nd), lambda: _engine.fir_dn(plan, xd, 12), s * nd * 13 // 12)
    timed("dn(4)  256 taps %s %d in" % (str(dt)[6:], nd), lambda: _engine.fir_dn(plan, xd, 4), s * nd * 5 // 4)
    timed("up(4)  256 taps %s %d in" % (str(dt)[6:], nd // 8), lambda: _engine.fir_up(plan, xd[: nd // 8], 4), s * 5 * (nd // 8))
    del xd
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rate12.json"), "w"), indent=1)
