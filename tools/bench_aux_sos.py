import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine
f = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
sos6 = _engine.SosPlan(f["sos6"])
x = torch.randn(2 ** 28, dtype=torch.float32, device="cuda")
for _ in range(2):
    y = _engine.sos_filter(sos6, x)
torch.cuda.synchronize()
