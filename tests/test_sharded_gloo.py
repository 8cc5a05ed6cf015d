"""world_size-2 (and 3) gloo tests of the sharded overlap-save HOST logic on CPU.

The arithmetic callable is injected (the oracle acts as the checker's kernel); what is under
test is segmentation, halo send/recv plumbing, complex halo packing, ordering and the
"segment shorter than the filter memory" error -- the parts of sharded.py that run on the host.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dtype_name, n_total, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from sk_dsp_comm_b200.sharded import ShardedFIR, segment_bounds

        b = np.load(os.path.join(GOLDEN, "filters.npz"))["b256"]
        rng = np.random.default_rng(5)
        xg = rng.standard_normal(n_total)
        if "complex" in dtype_name:
            xg = xg + 1j * rng.standard_normal(n_total)
        xg = xg.astype(dtype_name)

        def compute(x, hist):
            y = oracle.fir_filter(b, x.numpy(), hist=None if hist is None else hist.numpy(),
                                  backend="numpy")
            return torch.from_numpy(y.astype(dtype_name))

        sh = ShardedFIR(b, compute=compute)
        lo, hi = segment_bounds(n_total, world, rank)
        y_local = sh.filter(torch.from_numpy(xg[lo:hi].copy()))
        y_ref = oracle.fir_filter(b, xg, backend="numpy")[lo:hi]
        err = float(np.abs(y_local.numpy() - y_ref).max())
        scale = float(np.abs(y_ref).max())
        # too-short segment must raise on every rank that has to send a halo
        raised = False
        try:
            sh.exchange_halo(torch.zeros(10, dtype=y_local.dtype))
        except ValueError:
            raised = True
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, err, scale, raised, None))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, None, None, None, traceback.format_exc()))


@pytest.mark.parametrize("world,dtype_name", [(2, "float64"), (2, "complex64"), (3, "complex128")])
def test_sharded_fir_matches_monolithic(world, dtype_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_total = 3001
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype_name, n_total, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, scale, raised, tb in res:
        assert tb is None, tb
        tol = 1e-6 if dtype_name == "complex64" else 1e-12
        assert err <= tol * scale, (rank, err, scale)
        assert raised


def _iir_worker(rank, world, port, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from sk_dsp_comm_b200.sharded import ShardedIIR, segment_bounds

        sos = np.load(os.path.join(GOLDEN, "filters.npz"))["sos6"]
        n_total = 20011
        xg = np.random.default_rng(11).standard_normal(n_total)

        def compute(x, zi):
            y, zf = oracle.sos_filter(sos, x.numpy(), zi=None if zi is None else zi.numpy(), return_zf=True)
            return torch.from_numpy(y), torch.from_numpy(zf)

        sh = ShardedIIR(sos, compute=compute)
        lo, hi = segment_bounds(n_total, world, rank)
        y_local = sh.filter(torch.from_numpy(xg[lo:hi].copy()))
        y_ref = oracle.sos_filter(sos, xg)[lo:hi]
        err = float(np.abs(y_local.numpy() - y_ref).max())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, err, float(np.abs(y_ref).max()), None))
    except Exception:      # pragma: no cover
        import traceback
        q.put((rank, None, None, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_iir_state_chain(world):
    """Sharded IIR: zero-state carries + host powers of the state matrix reproduce the monolithic sosfilt."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_iir_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, scale, tb in res:
        assert tb is None, tb
        assert err <= 1e-11 * scale, (rank, err, scale)


def test_sos_state_matrix_matches_oracle():
    from sk_dsp_comm_b200.sharded import sos_state_matrix
    import oracle
    sos = np.load(os.path.join(GOLDEN, "filters.npz"))["sos6"]
    A = sos_state_matrix(sos)
    rng = np.random.default_rng(0)
    z0 = rng.standard_normal((6, 2))
    _, zf = oracle.sos_filter(sos, np.zeros(37), zi=z0, return_zf=True)      # 37 zero-input steps
    assert np.allclose(np.linalg.matrix_power(A, 37) @ z0.reshape(-1), zf.reshape(-1), rtol=0, atol=1e-12)


def _updn_worker(rank, world, port, dtype_name, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from sk_dsp_comm_b200.sharded import ShardedFIR, segment_bounds

        b = np.load(os.path.join(GOLDEN, "filters.npz"))["b256"]
        n_total = 6007
        rng = np.random.default_rng(9)
        xg = rng.standard_normal(n_total)
        if "complex" in dtype_name:
            xg = xg + 1j * rng.standard_normal(n_total)
        xg = xg.astype(dtype_name)

        def up(x, hist, L, out):
            # the checker has no history argument for up(): prepend the halo and drop its outputs
            if hist is None:
                out.copy_(torch.from_numpy(oracle.fir_up(b, x.numpy(), L)))
            else:
                xe = np.concatenate([hist.numpy(), x.numpy()])
                out.copy_(torch.from_numpy(oracle.fir_up(b, xe, L)[L * len(hist):]))

        def dn(x, hist, M, out):
            if hist is None:
                out.copy_(torch.from_numpy(oracle.fir_dn(b, x.numpy(), M)))
            else:
                # keep phase 0 on x[0]: pad the history on the left to a multiple of M
                h = hist.numpy()
                pad = (-len(h)) % M
                xe = np.concatenate([np.zeros(pad, h.dtype), h, x.numpy()])
                out.copy_(torch.from_numpy(oracle.fir_dn(b, xe, M)[(pad + len(h)) // M:]))

        sh = ShardedFIR(b, compute_up=up, compute_dn=dn)
        errs = []
        for L in (4, 3, 12):
            lo, hi = segment_bounds(n_total, world, rank)
            y = sh.up(torch.from_numpy(xg[lo:hi].copy()), L).numpy()
            ref = oracle.fir_up(b, xg, L)[L * lo:L * hi]
            errs.append(("up%d" % L, float(np.abs(y - ref).max()), float(np.abs(ref).max())))
        for M in (4, 3, 12):
            lo, hi = segment_bounds(n_total, world, rank, align=M)
            y = sh.dn(torch.from_numpy(xg[lo:hi].copy()), M).numpy()
            ref = oracle.fir_dn(b, xg, M)[lo // M:lo // M + (hi - lo) // M]
            assert y.shape == ref.shape, (y.shape, ref.shape)
            errs.append(("dn%d" % M, float(np.abs(y - ref).max()), float(np.abs(ref).max())))
        # a segment that breaks the phase rule must be refused on every rank that is not the last
        raised = rank + 1 == world                 # (the last rank may hold a ragged tail: it is not called)
        if not raised:
            try:
                sh.dn(torch.zeros(1001, dtype=torch.float64), 4)
            except ValueError:
                raised = True
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, errs, raised, None))
    except Exception:      # pragma: no cover
        import traceback
        q.put((rank, None, None, traceback.format_exc()))


@pytest.mark.parametrize("world,dtype_name", [(2, "float64"), (2, "complex128"), (3, "float64")])
def test_sharded_fir_up_dn_match_monolithic(world, dtype_name):
    """Sharded .up(L) / .dn(M): halo lengths, split points, decimation-phase alignment."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_updn_worker, args=(r, world, port, dtype_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, errs, raised, tb in res:
        assert tb is None, tb
        assert raised
        for name, err, scale in errs:
            assert err <= 1e-12 * scale, (rank, name, err, scale)
