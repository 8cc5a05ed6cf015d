import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
name = sys.argv[1] if len(sys.argv) > 1 else "sos6"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192 * 4
sos = F[name][:8]
plan = _engine.SosPlan(sos)
x = np.random.default_rng(0).standard_normal(n).astype(np.float32)
_cabi.lib.b200dsp_set_sos_variant(2)
y = _engine.sos_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
ref = oracle.sos_filter(sos, x.astype(np.float64))
err = np.abs(y - ref) / np.abs(ref).max()
nt = n // 8192
E = err[:nt * 8192].reshape(nt, 64, 128)
print("per tile max:", E.max(axis=(1, 2)))
print("per chunk max (tile 1):", np.array2string(E[min(1, nt - 1)].max(axis=1), precision=1, max_line_width=200))
print("per position max (tile 1):", np.array2string(E[min(1, nt - 1)].max(axis=0), precision=1, max_line_width=200))
