"""Bring-up of the float32 tensor-core FIR kernels (filter / up / dn) vs the oracle + timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
f = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
def rel(y, ref): return np.abs(y - ref).max() / np.abs(ref).max()
rng = np.random.default_rng(3)
for fname in ("b256", "b101", "b7"):
    b = f[fname]; plan = _engine.FirPlan(b)
    for n in (70001, 200000):
        x = rng.standard_normal(n).astype(np.float32); xt = torch.from_numpy(x).cuda(); x64 = x.astype(np.float64)
        print(fname, n, "filter %.3g" % rel(_engine.fir_filter(plan, xt).cpu().numpy(), oracle.fir_filter(b, x64, backend="c")), flush=True)
        for F in (2, 3, 4):
            try:
                print(fname, n, "up%d %.3g" % (F, rel(_engine.fir_up(plan, xt, F).cpu().numpy(), oracle.fir_up(b, x64, F, backend="c"))), flush=True)
                print(fname, n, "dn%d %.3g" % (F, rel(_engine.fir_dn(plan, xt, F).cpu().numpy(), oracle.fir_dn(b, x64, F, backend="c"))), flush=True)
            except Exception as e:
                print(fname, n, F, "ERR", repr(e)[:200], flush=True)
b = f["b256"]; plan = _engine.FirPlan(b)
# hist
x = rng.standard_normal(300000).astype(np.float32); xt = torch.from_numpy(x).cuda()
cut = 65536
for name, fn, hl, oc in (("filter", lambda t, h: _engine.fir_filter(plan, t, hist=h), 255, cut),
                         ("up4", lambda t, h: _engine.fir_up(plan, t, 4, hist=h), 64, 4 * cut),
                         ("dn4", lambda t, h: _engine.fir_dn(plan, t, 4, hist=h), 255, cut // 4)):
    yfull = fn(xt, None); y2 = fn(xt[cut:].contiguous(), xt[cut - hl:cut].contiguous())
    print("hist", name, "%.3g" % ((yfull[oc:] - y2).abs().max().item() / yfull.abs().max().item()), flush=True)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
x = torch.randn(2 ** 26, dtype=torch.float32, device="cuda")
for v in (0, 9):
    _cabi.lib.b200dsp_set_fir_variant(v)
    tu = timeit(lambda: _engine.fir_up(plan, x, 4)); td = timeit(lambda: _engine.fir_dn(plan, x, 4)); tf = timeit(lambda: _engine.fir_filter(plan, x))
    print("variant", v, "2^26 f32: up4 %.3f ms (%.1f%%)  dn4 %.3f ms (%.1f%%)  filter %.3f ms (%.1f%%)" % (
        tu, 20 * 2**26 / tu / 1e6 / 6481 * 100, td, 5 * 2**26 / td / 1e6 / 6481 * 100, tf, 8 * 2**26 / tf / 1e6 / 6481 * 100), flush=True)
