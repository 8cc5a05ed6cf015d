"""Bring-up of the overlap-save FFT FIR kernel (csrc/fir_fft.cu): parity vs the oracle, timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi
lib = _cabi.lib
rng = np.random.default_rng(0)
fails = 0
def lowpass(K, fc=0.1):
    n = np.arange(K) - (K - 1) / 2
    return np.sinc(2 * fc * n) * np.kaiser(K, 8.0) * 2 * fc
for K in (2, 33, 256, 257, 1024, 2049):
    b = lowpass(K) if K > 2 else np.array([0.7, -0.2])
    plan = _engine.FirPlan(b)
    for n in (4096, 40000, 123457):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        hist = (rng.standard_normal(K - 1) + 1j * rng.standard_normal(K - 1)).astype(np.complex64)
        lib.b200dsp_set_fir_variant(16)
        y = _engine.fir_filter(plan, torch.from_numpy(x).cuda()).cpu().numpy()
        yh = _engine.fir_filter(plan, torch.from_numpy(x).cuda(), hist=torch.from_numpy(hist).cuda()).cpu().numpy()
        lib.b200dsp_set_fir_variant(0)
        ref = oracle.fir_filter(b, x.astype(np.complex128), backend="c")
        refh = oracle.fir_filter(b, x.astype(np.complex128), hist=hist.astype(np.complex128), backend="c")
        e1 = np.abs(y - ref).max() / np.abs(ref).max(); e2 = np.abs(yh - refh).max() / np.abs(refh).max()
        ok = e1 <= 1e-6 and e2 <= 1e-6
        fails += (not ok)
        print("K=%4d n=%6d err %.2e  with hist %.2e %s" % (K, n, e1, e2, "ok" if ok else "FAIL"), flush=True)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
x = torch.randn(1 << 26, dtype=torch.complex64, device="cuda")
for K in (257, 512, 1024, 2049):
    plan = _engine.FirPlan(lowpass(K))
    ms = timeit(lambda: _engine.fir_filter(plan, x))
    lib.b200dsp_set_fir_variant(9)
    ms9 = timeit(lambda: _engine.fir_filter(plan, x), reps=2)
    lib.b200dsp_set_fir_variant(0)
    print("K=%4d 2^26 c64: fft %.3f ms (%.3f of HBM), CUDA-core direct %.3f ms" % (K, ms, 16 * (1 << 26) / ms / 1e6 / 6542.1, ms9), flush=True)
x = torch.randn(1 << 27, dtype=torch.float32, device="cuda")
for K in (257, 1024, 2049):
    plan = _engine.FirPlan(lowpass(K))
    ms = timeit(lambda: _engine.fir_filter(plan, x))
    lib.b200dsp_set_fir_variant(9)
    ms9 = timeit(lambda: _engine.fir_filter(plan, x), reps=2)
    lib.b200dsp_set_fir_variant(0)
    print("K=%4d 2^27 f32: fft %.3f ms (%.3f of HBM), CUDA-core direct %.3f ms" % (K, ms, 8 * (1 << 27) / ms / 1e6 / 6542.1, ms9), flush=True)
x = torch.randn(1 << 28, dtype=torch.complex64, device="cuda")
plan = _engine.FirPlan(np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"])
ms0 = timeit(lambda: _engine.fir_filter(plan, x))
lib.b200dsp_set_fir_variant(16)
ms16 = timeit(lambda: _engine.fir_filter(plan, x))
lib.b200dsp_set_fir_variant(0)
print("cfg2 256 taps 2^28: tcgen05 %.4f ms, fft overlap-save %.4f ms (%.3f)" % (ms0, ms16, 16 * (1 << 28) / ms16 / 1e6 / 6542.1))
print("FAILS", fails)
