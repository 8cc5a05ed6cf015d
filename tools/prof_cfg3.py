import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine
plan = _engine.FirPlan(np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"])
x = torch.randn(2 ** 26, dtype=torch.float32, device="cuda")
for _ in range(4):
    yu = _engine.fir_up(plan, x, 4); yd = _engine.fir_dn(plan, x, 4)
torch.cuda.synchronize()
