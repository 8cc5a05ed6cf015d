"""Bring-up of the tensor-core SOS cascade kernel (csrc/sos_tc.cu): parity against the oracle and the scan
kernels, zi/zf, up/dn, timing.  Run by hand on the GPU box; not part of the product."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")]
import numpy as np, torch
import oracle
from sk_dsp_comm_b200 import _engine, _cabi

F = np.load(os.path.join(ROOT, "tests/golden/filters.npz"))
lib = _cabi.lib
fails = 0


def rel(y, ref):
    return float(np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-300))


def check(name, err, tol):
    global fails
    ok = err <= tol
    fails += (not ok)
    print("%-58s err/max %.2e  %s" % (name, err, "ok" if ok else "FAIL (tol %.0e)" % tol), flush=True)


def run(sosname, n, L=1, M=1, zi=None, variant=2, seed=0, nsec=None):
    sos = F[sosname] if nsec is None else F[sosname][:nsec]
    plan = _engine.SosPlan(sos)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    lib.b200dsp_set_sos_variant(variant)
    zit = None if zi is None else torch.from_numpy(zi.astype(np.float32)).cuda()
    y, zf = _engine.sos_filter(plan, xt, L=L, M=M, zi=zit, return_zf=True)
    torch.cuda.synchronize()
    lib.b200dsp_set_sos_variant(0)
    x64 = x.astype(np.float64)
    if L > 1:
        x64 = oracle.upsample(x64, L) * L
    ref, zref = oracle.sos_filter(sos, x64, zi=None if zi is None else zi.astype(np.float64), return_zf=True)
    if M > 1:
        ref = oracle.downsample(ref, M)
    return y.cpu().numpy(), zf.cpu().numpy(), ref, zref


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    for name, tol in (("sos6", 2e-6), ("sos_sharp_lpf", 2e-6), ("sos_butter6", 2e-6)):
        for n in (8192, 8192 * 3 + 17, 1 << 20, (1 << 22) + 12345):
            y, zf, ref, zref = run(name, n)
            check("%s n=%d tc" % (name, n), rel(y, ref), tol)
            check("%s n=%d tc zf" % (name, n), rel(zf, zref), 1e-5)
    # ten-band: 10 sections = two groups (8 + 2), poles at 0.9988
    y, zf, ref, zref = run("sos_tenband", 1 << 22)
    check("sos_tenband n=2^22 tc (2 groups)", rel(y, ref), 1e-4)
    y, zf, ref, zref = run("sos_tenband", 1 << 22, nsec=8)
    check("sos_tenband[:8] n=2^22 tc", rel(y, ref), 1e-4)
    # zi
    rng = np.random.default_rng(5)
    zi = rng.standard_normal((6, 2)) * 0.3
    y, zf, ref, zref = run("sos6", 1 << 20, zi=zi)
    check("sos6 zi n=2^20", rel(y, ref), 2e-6)
    check("sos6 zi zf", rel(zf, zref), 1e-5)
    # up / dn
    for L in (2, 4, 12):
        y, zf, ref, zref = run("sos6", 1 << 18, L=L)
        check("sos6 up(%d) n=2^18" % L, rel(y, ref), 2e-6)
    for M in (2, 4, 12):
        y, zf, ref, zref = run("sos6", (1 << 21) + 5, M=M)
        check("sos6 dn(%d) n=2^21+5" % M, rel(y, ref), 2e-6)
    # scan-kernel path for comparison
    y, zf, ref, zref = run("sos6", 1 << 22, variant=1)
    check("sos6 n=2^22 scan kernels", rel(y, ref), 1e-4)

    if not quick:
        plan = _engine.SosPlan(F["sos6"])
        x = torch.randn(1 << 28, dtype=torch.float32, device="cuda")
        for variant in (1, 2, 0):
            lib.b200dsp_set_sos_variant(variant)
            for _ in range(3):
                y = _engine.sos_filter(plan, x)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                y = _engine.sos_filter(plan, x)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            print("cfg4 sos6 2^28 f32 variant %d: %.4f ms  %.1f GS/s  %.1f GB/s algorithmic" %
                  (variant, ms, (1 << 28) / ms / 1e6, 8 * (1 << 28) / ms / 1e6), flush=True)
        lib.b200dsp_set_sos_variant(0)
        # parity on the big stream: windows vs the oracle fed the same samples (cascade memory ~ 2000 samples)
        xh = x[:1 << 21].cpu().numpy()
        yh = y[:1 << 21].cpu().numpy()
        ref = oracle.sos_filter(F["sos6"], xh.astype(np.float64))
        check("cfg4 head window 2^21", rel(yh, ref), 2e-6)
        o = (1 << 27) - 12345
        xh = x[o:o + (1 << 21)].cpu().numpy()
        yh = y[o:o + (1 << 21)].cpu().numpy()
        ref = oracle.sos_filter(F["sos6"], xh.astype(np.float64))
        check("cfg4 middle window (after 2^16 settle)", rel(yh[1 << 16:], ref[1 << 16:]), 2e-6)
    print("FAILS", fails)
    sys.exit(1 if fails else 0)
