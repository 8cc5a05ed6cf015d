"""Bring-up of the tcgen05 FIR kernel: correctness of both descriptor base-offset modes vs the oracle,
then timing at 2^28."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
b = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"]
plan = _engine.FirPlan(b)
modes = [int(a) for a in sys.argv[1:]] or [10, 11]
for mode in modes:
    _cabi.lib.b200dsp_set_fir_variant(mode)
    for n in (64, 8192, 8193, 20000, 1 << 20):
        rng = np.random.default_rng(n)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        y = _engine.fir_filter(plan, torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        y = y.cpu().numpy()
        ref = oracle.fir_filter(b, x.astype(np.complex128), backend="c")
        d = np.abs(y - ref)
        bad = np.nonzero(d > 1e-6 * np.abs(ref).max())[0]
        print("variant", mode, "n", n, "maxerr/max %.3g" % (d.max() / np.abs(ref).max()),
              "rms err/rms %.3g" % (np.sqrt((d ** 2).mean()) / np.sqrt((np.abs(ref) ** 2).mean())),
              "nbad", bad.size, "first bad", bad[:6], flush=True)
    # scaled inputs: block floating point must make the result scale-invariant
    x = (np.random.default_rng(5).standard_normal(50000) + 1j * np.random.default_rng(6).standard_normal(50000))
    for sc in (1e-20, 1e-3, 1e4, 1e20):
        xs = (x * sc).astype(np.complex64)
        y = _engine.fir_filter(plan, torch.from_numpy(xs).cuda()).cpu().numpy()
        ref = oracle.fir_filter(b, xs.astype(np.complex128), backend="c")
        print("variant", mode, "scale", sc, "maxerr/max %.3g" % (np.abs(y - ref).max() / np.abs(ref).max()), flush=True)
    n = 1 << 28
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    for _ in range(3): _engine.fir_filter(plan, x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): _engine.fir_filter(plan, x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("variant", mode, "2^28: %.3f ms  %.1f GS/s  %.1f%% of 6481 GB/s" % (ms, n / ms / 1e6, 16 * n / ms / 1e6 / 6481 * 100), flush=True)
    W = 1 << 16
    st = n // 2 + 12345
    ref = oracle.fir_filter(b, x[st - 255:st + W].cpu().numpy().astype(np.complex128), backend="c")[255:]
    got = y[st:st + W].cpu().numpy()
    print("variant", mode, "2^28 window err/max %.3g" % (np.abs(got - ref).max() / np.abs(ref).max()), flush=True)
    del x, y
