import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
name, lo, hi, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sos = F[name][lo:hi]
plan = _engine.SosPlan(sos)
x = np.random.default_rng(0).standard_normal(n).astype(np.float32)
_cabi.lib.b200dsp_set_sos_variant(2)
y = _engine.sos_filter(plan, torch.from_numpy(x).cuda())
torch.cuda.synchronize()
y = y.cpu().numpy().astype(np.float64)
ref = oracle.sos_filter(sos, x.astype(np.float64))
print(name, lo, hi, n, "err/max %.3e" % (np.abs(y - ref).max() / np.abs(ref).max()), flush=True)
if len(sys.argv) > 5:
    y2, zf = _engine.sos_filter(plan, torch.from_numpy(x).cuda(), return_zf=True)
    torch.cuda.synchronize()
    _, zref = oracle.sos_filter(sos, x.astype(np.float64), return_zf=True)
    print("zf err %.3e" % (np.abs(zf.cpu().numpy() - zref).max() / np.abs(zref).max()), flush=True)
