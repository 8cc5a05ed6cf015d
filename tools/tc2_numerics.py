"""CPU numerics study for the tcgen05 FIR: how much accuracy does the
x_lo * b_lo product buy?  numpy emulation of the fp16 hi/lo split with per-tile block scaling and
fp32 accumulation (round-to-nearest here; the tensor core truncates, so absolute levels are
slightly optimistic) against the float64 oracle.  Not used by the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np
import oracle

f32 = np.float32
b = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))["b256"]
rng = np.random.default_rng(1)
n = 1 << 19
x = rng.standard_normal(n).astype(f32)
ref = oracle.fir_filter(b, x.astype(np.float64), backend="c")


def split(a, e):
    s = a.astype(np.float64) * 2.0 ** e
    hi = s.astype(np.float16)
    lo = (s - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(f32), lo.astype(f32)


sb = -int(np.floor(np.log2(np.abs(b).max()))) + 0
bh, bl = split(b, sb)
TILE = 6144
y = {k: np.zeros(n) for k in ("4 products", "3 products (no lo*lo)", "2 products (hi*x only)", "hi*hi only")}
K = len(b)
for t0 in range(0, n, TILE):
    t1 = min(t0 + TILE, n)
    seg = x[max(t0 - (K - 1), 0):t1]
    if t0 < K - 1:
        seg = np.concatenate([np.zeros(K - 1 - t0, f32), seg])
    e = -int(np.floor(np.log2(max(np.abs(seg).max(), 1e-30))))
    xh, xl = split(seg, e)
    sc = 2.0 ** (-(e + sb))
    def conv(bt, xt):
        return np.convolve(bt.astype(f32), xt.astype(f32)).astype(f32)[K - 1:K - 1 + (t1 - t0)]
    hh, hl, lh, ll = conv(bh, xh), conv(bh, xl), conv(bl, xh), conv(bl, xl)
    y["4 products"][t0:t1] = ((ll + hl) + (lh + hh)).astype(np.float64) * sc
    y["3 products (no lo*lo)"][t0:t1] = (hl + (lh + hh)).astype(np.float64) * sc
    y["2 products (hi*x only)"][t0:t1] = (hl + hh).astype(np.float64) * sc
    y["hi*hi only"][t0:t1] = hh.astype(np.float64) * sc
sc_ = np.abs(ref).max()
rms = np.sqrt((ref ** 2).mean())
for k, v in y.items():
    e = v - ref
    print("%-26s max %.3g  rms %.3g   (of max|y| / rms y)" % (k, np.abs(e).max() / sc_, np.sqrt((e ** 2).mean()) / rms))
y32 = np.convolve(b.astype(f32), x)[:n]
e = y32.astype(np.float64) - ref
print("%-26s max %.3g  rms %.3g" % ("plain fp32 convolution", np.abs(e).max() / sc_, np.sqrt((e ** 2).mean()) / rms))
