"""Secondary configs (BASELINE.json configs[0,2,3] + index maps): device-resident timing with
CUDA events, algorithmic bytes per SURVEY.md 8d, fraction of the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
from sk_dsp_comm_b200 import _engine
f = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
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
    r = dict(config=name, ms=ms, Msamples_per_s_input=n_in / (ms * 1e-3) / 1e6, algo_GBps=gbs, frac_of_measured_hbm=gbs / peak)
    out.append(r); print(json.dumps(r), flush=True)
fir256 = _engine.FirPlan(f["b256"]); fir101 = _engine.FirPlan(f["b101"]); sos6 = _engine.SosPlan(f["sos6"])
x = torch.randn(2 ** 20, dtype=torch.float64, device="cuda")
report("cfg1 fir101 f64 2^20", x.numel(), 16 * x.numel(), timeit(lambda: _engine.fir_filter(fir101, x)))
x = torch.randn(2 ** 26, dtype=torch.float32, device="cuda")
report("cfg3 fir256 up4 f32 2^26", x.numel(), 20 * x.numel(), timeit(lambda: _engine.fir_up(fir256, x, 4)))
report("cfg3 fir256 dn4 f32 2^26", x.numel(), 5 * x.numel(), timeit(lambda: _engine.fir_dn(fir256, x, 4)))
report("upsample4 f32 2^26", x.numel(), 20 * x.numel(), timeit(lambda: _engine.upsample(x, 4)))
report("downsample4 f32 2^26 (s + s/M)", x.numel(), 5 * x.numel(), timeit(lambda: _engine.downsample(x, 4, 0)))
x = torch.randn(2 ** 28, dtype=torch.float32, device="cuda")
report("fir256 f32 2^28", x.numel(), 8 * x.numel(), timeit(lambda: _engine.fir_filter(fir256, x), reps=5))
report("cfg4 sos6 f32 2^28", x.numel(), 8 * x.numel(), timeit(lambda: _engine.sos_filter(sos6, x), reps=5))
x26 = x[:2 ** 26]
report("sos6 dn4 f32 2^26 (multirate_IIR.dn)", x26.numel(), 5 * x26.numel(), timeit(lambda: _engine.sos_filter(sos6, x26, M=4), reps=5))
x24 = x[:2 ** 24]
report("sos6 up4 f32 2^24 (multirate_IIR.up)", x24.numel(), 20 * x24.numel(), timeit(lambda: _engine.sos_filter(sos6, x24, L=4), reps=5))
del x
x = torch.randn(2 ** 26, dtype=torch.float64, device="cuda")
report("sos6 f64 2^26", x.numel(), 16 * x.numel(), timeit(lambda: _engine.sos_filter(sos6, x), reps=5))
report("fir256 f64 2^26", x.numel(), 16 * x.numel(), timeit(lambda: _engine.fir_filter(fir256, x), reps=5))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_aux.json"), "w"), indent=1)
