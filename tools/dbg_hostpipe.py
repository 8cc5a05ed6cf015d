import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, hostpipe
b = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"]
plan = _engine.FirPlan(b)
torch.manual_seed(100)
x = torch.randn((1 << 22) + 12345, dtype=torch.complex64).pin_memory()
y_host = hostpipe.fir_filter_host(plan, x, chunk=1 << 20)
y_dev = _engine.fir_filter(plan, x.cuda()).cpu()
d = (y_host - y_dev).abs()
scale = y_dev.abs().max().item()
idx = torch.argsort(d, descending=True)[:12]
print("scale", scale, "max diff/scale", d.max().item() / scale)
print("worst idx", idx.tolist())
for i in idx[:4].tolist():
    lo = max(i - 255, 0)
    ref = oracle.fir_filter(b, x[lo:i + 1].numpy().astype(np.complex128), backend="c")[-1]
    print(i, i % (1 << 20), "host err %.3g dev err %.3g" % (abs(y_host[i].item() - ref) / scale, abs(y_dev[i].item() - ref) / scale))
# error distribution vs oracle on a big window for the device path
W = 1 << 21
ref = oracle.fir_filter(b, x[:W].numpy().astype(np.complex128), backend="c")
e = np.abs(y_dev[:W].numpy() - ref) / scale
print("dev path: max err %.3g  99.99%% %.3g  rms %.3g" % (e.max(), np.quantile(e, 0.9999), np.sqrt((e ** 2).mean())))
e = np.abs(y_host[:W].numpy() - ref) / scale
print("host path: max err %.3g at %d" % (e.max(), int(e.argmax())))
