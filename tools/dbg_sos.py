import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine
f = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
for fname in ("sos6", "b1sec"):
    sos = f["sos6"] if fname == "sos6" else f["sos6"][:1]
    plan = _engine.SosPlan(sos)
    for dt, LC in ((np.float64, 32), (np.float32, 64)):
        for n in (1, LC // 2, LC, LC + 1, 2 * LC, 3 * LC, 32 * LC, 33 * LC, 64 * LC, 512 * LC, 512 * LC + 5, 3 * 512 * LC):
            x = np.random.default_rng(n).standard_normal(n).astype(dt)
            y = _engine.sos_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy()
            ref = oracle.sos_filter(sos, x.astype(np.float64))
            d = np.abs(y - ref)
            bad = np.nonzero(d > 1e-4 * np.abs(ref).max())[0]
            print(fname, dt.__name__, "n=%d" % n, "maxerr %.3g" % d.max(), "first bad", bad[:3] if bad.size else None, flush=True)
