import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine, hostpipe
import sk_dsp_comm_b200.multirate_helper as mrh
b = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"]
plan = _engine.FirPlan(b)
torch.manual_seed(7)
n = (1 << 22) + 12345
x = torch.randn(n, dtype=torch.complex64).pin_memory()
y_dev = _engine.fir_filter(plan, x.cuda()).cpu()
scale = y_dev.abs().max().item()
def report(name, y):
    d = (y - y_dev).abs()
    bad = torch.nonzero(d > 2e-6 * scale).flatten()
    print(name, "max diff/scale %.3g" % (d.max().item() / scale), "nbad", bad.numel(),
          "first/last bad", (bad[0].item(), bad[-1].item()) if bad.numel() else None, flush=True)
report("hostpipe chunk=2^20 same plan", hostpipe.fir_filter_host(plan, x, chunk=1 << 20))
report("hostpipe default chunk same plan", hostpipe.fir_filter_host(plan, x))
report("hostpipe default chunk same plan (2nd)", hostpipe.fir_filter_host(plan, x))
f = mrh.multirate_FIR(b)
report("API kept object", f.filter(x))
report("API temp object", mrh.multirate_FIR(b).filter(x))
report("hostpipe chunk=n-1", hostpipe.fir_filter_host(plan, x, chunk=n - 1))
report("hostpipe chunk=3000001", hostpipe.fir_filter_host(plan, x, chunk=3000001))
