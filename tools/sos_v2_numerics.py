"""CPU numerics study for the next SOS-cascade kernel: does the planned
arithmetic hold the 1e-4 bar?  Pure numpy emulation of the fp32 / fp16-split data paths; nothing
here runs on the GPU or is used by the product.

  (a) second pass replaced by the linear correction  y[n] = y0[n] + (C A^n) s0   (fp32 table)
  (b) chunk carries from a dense contraction  carry = sum_n A^(LC-1-n) B x[n]  with fp16 hi/lo
      split operands and fp32 accumulation (what a tcgen05 kind::f16 GEMM would compute)
  (c) the correction itself as an fp16 hi/lo GEMM (3 products, fp32 accumulate)
against the float64 oracle, for the cascades of the fixture set, chunk lengths 64 / 128 / 256.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np
import oracle

F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
f32 = np.float32


def state_space(sos):
    """(A, B, C, D) of the DF-II-T cascade in scipy's state order, by probing one step in float64."""
    sos = np.asarray(sos, dtype=np.float64)
    ns = sos.shape[0]
    nd = 2 * ns

    def step(z, x):
        z = z.copy()
        v = x
        for s in range(ns):
            b0, b1, b2, _, a1, a2 = sos[s]
            y = b0 * v + z[2 * s]
            z[2 * s], z[2 * s + 1] = b1 * v - a1 * y + z[2 * s + 1], b2 * v - a2 * y
            v = y
        return z, v
    A = np.zeros((nd, nd)); C = np.zeros(nd)
    for d in range(nd):
        e = np.zeros(nd); e[d] = 1
        A[:, d], C[d] = step(e, 0.0)
    B, D = step(np.zeros(nd), 1.0)
    return A, B, C, D


def cascade_f32(sos, x, z0):
    """fp32 DF-II-T cascade over chunks in parallel: x (nchunk, LC), z0 (nchunk, nd) -> y, z_end."""
    c = sos.astype(f32)
    z = z0.astype(f32).copy()
    y = np.empty_like(x, dtype=f32)
    for n in range(x.shape[1]):
        v = x[:, n].astype(f32)
        for s in range(c.shape[0]):
            b0, b1, b2, _, a1, a2 = c[s]
            o = b0 * v + z[:, 2 * s]
            z[:, 2 * s] = (b1 * v - a1 * o) + z[:, 2 * s + 1]
            z[:, 2 * s + 1] = b2 * v - a2 * o
            v = o
        y[:, n] = v
    return y, z


def split16(a):
    """power-of-two block scale + fp16 hi / fp16 residual lo (as fir_tc2 does)."""
    m = np.abs(a).max()
    e = 0 if m == 0 else int(np.floor(np.log2(m)))
    sc = 2.0 ** (-e)
    hi = (a * sc).astype(np.float16)
    lo = ((a * sc) - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(f32), lo.astype(f32), 1.0 / sc


def gemm16(Wm, Xm):
    """sum over the shared axis with fp16 hi/lo operands, 3 products, fp32 accumulation."""
    wh, wl, ws = split16(Wm)
    xh, xl, xs = split16(Xm)
    acc = (wh @ xl + wl @ xh) + wh @ xh           # small terms first, all fp32 matmuls
    return acc.astype(np.float64) * (ws * xs)


def study(name, sos, LC, n=1 << 18, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype(f32)
    ref, _ = oracle.sos_filter(sos, x.astype(np.float64), return_zf=True)
    A, B, C, D = state_space(sos)
    nd = len(B)
    nch = n // LC
    X = x[:nch * LC].reshape(nch, LC)
    # true chunk start states in float64 (stand-in for the scan, which stays as it is today)
    Apow = [np.eye(nd)]
    for _ in range(LC):
        Apow.append(A @ Apow[-1])
    W = np.stack([Apow[LC - 1 - k] @ B for k in range(LC)], axis=1)           # nd x LC
    carry64 = X.astype(np.float64) @ W.T                                       # zero-state chunk carries
    s0 = np.zeros((nch, nd))
    for c in range(1, nch):
        s0[c] = Apow[LC] @ s0[c - 1] + carry64[c - 1]
    scale = np.abs(ref).max()
    out = {}
    # baseline: what the kernel does today (fp32 cascade from the true start state)
    y_b, _ = cascade_f32(sos, X, s0)
    out["today"] = np.abs(y_b.reshape(-1) - ref[:nch * LC]).max() / scale
    # (a) zero-state fp32 cascade + fp32 table correction
    y0, zend = cascade_f32(sos, X, np.zeros((nch, nd)))
    T = np.stack([C @ Apow[k] for k in range(LC)], axis=0)                     # LC x nd
    y_a = y0 + (s0.astype(f32) @ T.astype(f32).T)
    out["table_f32"] = np.abs(y_a.reshape(-1) - ref[:nch * LC]).max() / scale
    # (c) the correction as an fp16-split GEMM
    y_c = y0.astype(np.float64) + gemm16(s0, T.T)
    out["table_tc"] = np.abs(y_c.reshape(-1) - ref[:nch * LC]).max() / scale
    # (b) carries from an fp16-split GEMM vs the fp32 cascade's end state
    c16 = gemm16(X.astype(np.float64), W.T)
    cs = np.abs(carry64).max()
    out["carry_tc"] = np.abs(c16 - carry64).max() / cs
    out["carry_f32"] = np.abs(zend.astype(np.float64) - carry64).max() / cs
    out["rho"] = float(np.abs(np.linalg.eigvals(A)).max())
    print("%-14s LC=%3d  pole %.4f | y err/max: today %.2e  table-f32 %.2e  table-tc %.2e | carry err: cascade-f32 %.2e  gemm-fp16x3 %.2e"
          % (name, LC, out["rho"], out["today"], out["table_f32"], out["table_tc"], out["carry_f32"], out["carry_tc"]), flush=True)
    return out


if __name__ == "__main__":
    for name in ("sos6", "sos_sharp_lpf", "sos_butter6", "sos_tenband"):
        sos = F[name] if name != "sos_tenband" else F[name][:8]              # one group of <= 8 sections
        for LC in (64, 128, 256):
            study(name, sos, LC)
