"""SURVEY.md 8f rank 2/3 measurement: pulse shaping (fused zero-stuff + FIR) and the stateful
filter call, device-resident, CUDA events, algorithmic bytes vs the measured HBM peak; plus the
wall time of the public host-array calls (numpy in -> numpy out, PCIe inside)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine, _pulse, _cabi
import sk_dsp_comm_b200.digitalcom as dc
import sk_dsp_comm_b200.multirate_helper as mrh
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]
out = []
def report(name, n_in, algo_bytes, ms):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    r = dict(config=name, ms=ms, Msymbols_per_s=n_in / (ms * 1e-3) / 1e6, algo_GBps=gbs, frac_of_measured_hbm=gbs / peak)
    out.append(r); print(json.dumps(r), flush=True)
ns = 8
b = dc.sqrt_rc_imp(ns, 0.35, 6)                      # 97 taps
plan = _engine.FirPlan(b / ns)
for dt, esz in ((torch.complex128, 16), (torch.complex64, 8), (torch.float32, 4)):
    n = 1 << 24
    s = torch.randn(n, dtype=dt, device="cuda")
    y = torch.empty(n * ns, dtype=dt, device="cuda")
    report("pulse shaping SRC 97 taps ns=8 %s 2^24 symbols" % str(dt).split(".")[1], n, esz * n * (1 + ns),
           timeit(lambda: _engine.fir_up(plan, s, ns, out=y)))
    _cabi.lib.b200dsp_set_fir_variant(8)
    report("  same, polyphase kernel (variant 8)", n, esz * n * (1 + ns), timeit(lambda: _engine.fir_up(plan, s, ns, out=y)))
    _cabi.lib.b200dsp_set_fir_variant(0)
    del s, y
b256 = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"]
p256 = _engine.FirPlan(b256)
for L in (8, 12, 16):
    n = 1 << 23
    s = torch.randn(n, dtype=torch.complex64, device="cuda")
    y = torch.empty(n * L, dtype=torch.complex64, device="cuda")
    report("fir256 up%d complex64 2^23 (default)" % L, n, 8 * n * (1 + L), timeit(lambda: _engine.fir_up(p256, s, L, out=y)))
    _cabi.lib.b200dsp_set_fir_variant(8)
    report("  same, polyphase kernel (variant 8)", n, 8 * n * (1 + L), timeit(lambda: _engine.fir_up(p256, s, L, out=y)))
    _cabi.lib.b200dsp_set_fir_variant(0)
    del s, y
# stateful FIR block call (zi/zf): the kernel + the K-1 state bookkeeping, device tensors
fir = mrh.multirate_FIR(np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"])
x = torch.randn(1 << 26, dtype=torch.complex64, device="cuda")
z = np.zeros(255)
t0 = timeit(lambda: fir.filter(x), reps=5)
t1 = timeit(lambda: fir.filter(x, zi=z), reps=5)
report("fir256 c64 2^26 stateless", x.numel(), 16 * x.numel(), t0)
report("fir256 c64 2^26 with zi/zf (lfilter state)", x.numel(), 16 * x.numel(), t1)
del x
# public host calls (numpy in/out): symbols drawn + shaped + copied back
for n in (1 << 16, 1 << 20, 1 << 22):
    np.random.seed(1)
    dc.qam_bb(1024, ns, '16qam', 'src')
    t = time.perf_counter(); xw, bb, d = dc.qam_bb(n, ns, '16qam', 'src'); dt_ = time.perf_counter() - t
    r = dict(config="dc.qam_bb(%d, 8, 16qam, src) wall" % n, ms=dt_ * 1e3, Msymbols_per_s=n / dt_ / 1e6)
    out.append(r); print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_pulse.json"), "w"), indent=1)
