"""Role-by-role timing of the tensor-core SOS kernel through its B200DSP_STC_DBG bring-up bits."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine, _cabi
F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
name = sys.argv[1] if len(sys.argv) > 1 else "sos6"
plan = _engine.SosPlan(F[name][:8])
n = 1 << 28
x = torch.randn(n, dtype=torch.float32, device="cuda")
_cabi.lib.b200dsp_set_sos_variant(2)
for dbg in [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,4,16,3,6,18,22,23,8".split(","))]:
    os.environ["B200DSP_STC_DBG"] = str(dbg)
    for _ in range(2 if dbg != 8 else 0):
        y = _engine.sos_filter(plan, x)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    reps = 5 if dbg != 8 else 1
    for _ in range(reps):
        y = _engine.sos_filter(plan, x)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    print("dbg %2d: %.4f ms  %.1f GS/s  frac %.3f" % (dbg, ms, n / ms / 1e6, 8 * n / ms / 1e6 / 6542.1), flush=True)
