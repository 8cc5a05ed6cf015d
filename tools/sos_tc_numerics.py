"""CPU numerics study for the tensor-core SOS cascade kernel (csrc/sos_tc.cu): numpy emulation of
its data path against the float64 oracle.  Nothing here runs on the GPU or is used by the product.

Per chunk of LC samples (X = chunk matrix, one row per chunk):
  zero-state outputs   Y0 = X T^T          T[n,k] = h[n-k] (lower-triangular Toeplitz of the cascade's
                                           impulse response), fp16 hi/lo operands, fp32 accumulate
  chunk carries        E  = X W^T          W[:,k] = R A^(LC-1-k) B   (state basis R: see below)
  scan                 s_{c+1} = A' s_c + E_c        float64
  correction           Y  = Y0 + S Q^T     Q = orthonormal basis of the zero-input responses O = C A^n,
                                           O = Q R (thin QR), so the state is carried as s' = R s:
                                           |Q s'| = |s'| -- no cancellation left for the fp32 product.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np
import oracle
from sos_v2_numerics import state_space, split16, F

f32 = np.float32


def gemm16(Wm, Xm, products=4):
    wh, wl, ws = split16(Wm)
    xh, xl, xs = split16(Xm)
    acc = wh @ xl + wl @ xh
    if products == 4:
        acc = acc + wl @ xl
    acc = acc + wh @ xh
    return acc.astype(np.float64) * (ws * xs)


def study(name, sos, LC, n=1 << 18, seed=0, basis="qr", products=4, scan=np.float64):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype(f32)
    ref = oracle.sos_filter(sos, x.astype(np.float64))
    A, B, C, D = state_space(sos)
    nd = len(B)
    nch = n // LC
    X = x[:nch * LC].reshape(nch, LC).astype(np.float64)
    Apow = [np.eye(nd)]
    for _ in range(LC):
        Apow.append(A @ Apow[-1])
    h = np.array([D] + [C @ Apow[k] @ B for k in range(LC - 1)])
    T = np.zeros((LC, LC))
    for i in range(LC):
        T[i, :i + 1] = h[:i + 1][::-1]
    O = np.stack([C @ Apow[k] for k in range(LC)], axis=0)                    # LC x nd
    if basis == "qr":
        Q, R = np.linalg.qr(O)
        cond = np.linalg.cond(R)
    else:
        Q, R, cond = O, np.eye(nd), 1.0
    Rinv = np.linalg.inv(R)
    W = np.stack([R @ Apow[LC - 1 - k] @ B for k in range(LC)], axis=1)       # nd x LC
    Ap = R @ Apow[LC] @ Rinv
    # tile-wise (96 chunks) block scale like the kernel: emulate per 96-chunk groups
    Y0 = np.empty((nch, LC)); E = np.empty((nch, nd))
    for c0 in range(0, nch, 96):
        Xt = X[c0:c0 + 96]
        Y0[c0:c0 + 96] = gemm16(Xt, T.T, products)
        E[c0:c0 + 96] = gemm16(Xt, W.T, products)
    E = E.astype(f32).astype(np.float64)
    s = np.zeros((nch, nd), dtype=scan)
    Aps = Ap.astype(scan)
    for c in range(1, nch):
        s[c] = Aps @ s[c - 1] + E[c - 1].astype(scan)
    corr = (s.astype(f32) @ Q.astype(f32).T)                                   # fp32 products/accumulate
    Y = (Y0.astype(f32) + corr).astype(np.float64)
    scale = np.abs(ref).max()
    err = np.abs(Y.reshape(-1) - ref[:nch * LC]).max() / scale
    print("%-14s LC=%3d nd=%2d basis=%-4s cond(R)=%.1e prod=%d scan=%s: err/max = %.2e   max|s'|/max|y| = %.1e"
          % (name, LC, nd, basis, cond, products, np.dtype(scan).name, err, np.abs(s).max() / scale), flush=True)
    return err


if __name__ == "__main__":
    for name in ("sos6", "sos_sharp_lpf", "sos_butter6", "sos_tenband"):
        sos = F[name]
        for LC in (64, 128):
            for basis in ("raw", "qr"):
                study(name, sos, LC, basis=basis)
        study(name, sos, 64, basis="qr", products=3)
        study(name, sos, 64, basis="qr", scan=np.float32)
