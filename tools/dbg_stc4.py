import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
sos = F["sos_tenband"]
rng = np.random.default_rng(7)
x = rng.standard_normal(300007).astype(np.float32)
ref, zfull = oracle.sos_filter(sos, x.astype(np.float64), return_zf=True)
cut = 123457
_cabi.lib.b200dsp_set_sos_variant(2)
plan = _engine.SosPlan(sos)
ya, zf = _engine.sos_filter(plan, torch.from_numpy(x[:cut]).cuda(), return_zf=True)
_, zref = oracle.sos_filter(sos, x[:cut].astype(np.float64), return_zf=True)
print("zf err rel to max", np.abs(zf.cpu().numpy() - zref).max() / np.abs(zref).max())
print("zf", zf.cpu().numpy().ravel()[:6], zref.ravel()[:6])
yb = _engine.sos_filter(plan, torch.from_numpy(x[cut:].copy()).cuda(), zi=zf)
yb64 = _engine.sos_filter(plan, torch.from_numpy(x[cut:].copy()).cuda(), zi=torch.from_numpy(zref.astype(np.float32)).cuda())
sc = np.abs(ref).max()
print("first part err", np.abs(ya.cpu().numpy() - ref[:cut]).max() / sc)
print("second part err (kernel zf)", np.abs(yb.cpu().numpy() - ref[cut:]).max() / sc)
print("second part err (oracle zf rounded to f32)", np.abs(yb64.cpu().numpy() - ref[cut:]).max() / sc)
refb = oracle.sos_filter(sos, x[cut:].astype(np.float64), zi=zref.astype(np.float32).astype(np.float64))
print("oracle itself with f32-rounded zi vs full:", np.abs(refb - ref[cut:]).max() / sc)
e = np.abs(yb.cpu().numpy() - ref[cut:]) / sc
print("where:", e.argmax(), e[:5], e[8192-3:8192+3])
