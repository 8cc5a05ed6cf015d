"""Overlap-save sharding of one long stream across GPUs.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).  Rank r owns
the contiguous segment ``[r*n_local, (r+1)*n_local)`` of the global stream.  A causal FIR with K
taps needs, besides its own segment, the LAST K-1 INPUT samples of rank r-1 (rank 0: the
reference's zero initial state) -- the classic overlap-save halo.  That is the only exchange
step of the path: one neighbour send/recv of ``(K-1)*itemsize`` bytes (2040 B for the 256-tap
complex64 config), no all-reduce / all-gather.  The result is bit-identical to filtering the
whole stream on one device because the halo reproduces the filter state exactly.

Overlap: the interior outputs ``y[K-1:]`` depend only on local samples, so their kernel is
launched first; the halo travels meanwhile; the first outputs (up to the next 64-sample boundary
after ``K-1``) are computed by a second, tiny launch once the halo has landed.

``halo="peer"`` replaces the NCCL message by a direct NVLink read: the neighbour's tail is
exposed through symmetric memory and handed to the kernel as its ``hist`` pointer, so the FIR
kernel itself pulls the 2 KB halo over NVSwitch (no extra launch, no host round trip).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _engine


def segment_bounds(n_total: int, world: int, rank: int, align: int = 1):
    """Contiguous segment of rank ``rank``: boundaries are multiples of ``align`` (use M for a
    decimating filter so phase 0 stays global); the last rank takes the remainder."""
    per = (n_total // world) // align * align
    lo = rank * per
    hi = n_total if rank == world - 1 else lo + per
    return lo, hi


class ShardedFIR:
    """``multirate_FIR`` semantics over a stream sharded across the ranks of a process group.

    compute : callables used for the arithmetic.  The default is the CUDA engine; the CPU
              (gloo) tests inject a checker so the HOST-side logic (segmentation, halo plumbing,
              ordering) can be exercised without a GPU.  The product path never passes one.
    """

    def __init__(self, b, group=None, compute=None):
        self.plan = _engine.FirPlan(b)
        self.group = group
        self.k1 = self.plan.ntaps - 1
        self._fir = compute or (lambda x, hist: _engine.fir_filter(self.plan, x, hist=hist))
        self._symm = None

    # -- plumbing -------------------------------------------------------------------------
    def _rank_world(self):
        return dist.get_rank(self.group), dist.get_world_size(self.group)

    def exchange_halo(self, x_local: torch.Tensor):
        """Send my last K-1 samples to rank+1, receive rank-1's.  Returns (halo or None, works)."""
        rank, world = self._rank_world()
        k1 = self.k1
        if k1 == 0 or world == 1:
            return None, []
        if x_local.numel() < k1:
            raise ValueError("segment shorter than the filter memory (%d < %d): use fewer ranks"
                             % (x_local.numel(), k1))
        ops = []
        halo = None
        tail = x_local[-k1:].contiguous()
        if x_local.is_complex():                      # send as real pairs (gloo has no complex send)
            tail = torch.view_as_real(tail)
        if rank + 1 < world:
            ops.append(dist.P2POp(dist.isend, tail, rank + 1, self.group))
        if rank > 0:
            shape = (k1, 2) if x_local.is_complex() else (k1,)
            rdt = tail.dtype
            halo = torch.empty(shape, dtype=rdt, device=x_local.device)
            ops.append(dist.P2POp(dist.irecv, halo, rank - 1, self.group))
        works = dist.batch_isend_irecv(ops) if ops else []
        self._keepalive = tail
        if halo is not None and x_local.is_complex():
            halo = torch.view_as_complex(halo)
        return halo, works

    # -- the sharded filter -----------------------------------------------------------------
    def filter(self, x_local: torch.Tensor) -> torch.Tensor:
        """y_local = this rank's segment of lfilter(b,[1],x_global)."""
        rank, world = self._rank_world()
        k1 = self.k1
        n = x_local.numel()
        if world == 1 or k1 == 0:
            return self._fir(x_local, None)
        on_gpu = x_local.is_cuda
        if on_gpu:
            comm = self._comm_stream(x_local.device)
            cur = torch.cuda.current_stream(x_local.device)
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                halo, works = self.exchange_halo(x_local)
                for w in works:
                    w.wait()               # enqueues the dependency on `comm`, does not block the host
                ev = torch.cuda.Event()
                ev.record(comm)
        else:
            halo, works = self.exchange_halo(x_local)
        y = torch.empty_like(x_local)
        # interior first: needs no halo, overlaps the exchange.  The split point is a multiple of 64
        # samples so both launches see 16-byte aligned streams (tensor-core path).
        sp = min((k1 + 63) // 64 * 64, n)
        if n > sp:
            hist_i = x_local[sp - k1:sp]
            if on_gpu:
                _engine.fir_filter(self.plan, x_local[sp:], hist=hist_i.contiguous(), out=y[sp:])
            else:
                y[sp:] = self._fir(x_local[sp:], hist_i)
        if on_gpu:
            cur.wait_event(ev)
        else:
            for w in works:
                w.wait()
        head = x_local[:sp]
        if on_gpu:
            _engine.fir_filter(self.plan, head, hist=halo, out=y[:sp])
        else:
            y[:sp] = self._fir(head, halo)
        return y

    _comm_streams = {}

    def _comm_stream(self, device):
        s = ShardedFIR._comm_streams.get(device)
        if s is None:
            s = torch.cuda.Stream(device)
            ShardedFIR._comm_streams[device] = s
        return s

    # -- peer-memory halo (NVLink direct read by the FIR kernel) -------------------------------
    def attach_symmetric(self, n_local: int, dtype, device):
        """Allocate this rank's segment in symmetric memory so the left neighbour's tail can be
        read by the kernel over NVLink.  Returns the local tensor to fill with input samples."""
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(n_local, dtype=dtype, device=device)
        group_name = (self.group or dist.group.WORLD).group_name
        hdl = symm_mem.rendezvous(t, group_name)
        self._symm = (t, hdl, n_local)
        return t

    def filter_peer(self) -> torch.Tensor:
        """Filter the symmetric-memory segment; the halo is read from rank-1's memory by the
        kernel itself (``hist`` = peer pointer).  The caller guarantees every rank has finished
        writing its segment (``barrier()``) before calling."""
        t, hdl, n = self._symm
        rank, world = self._rank_world()
        k1 = self.k1
        hist = None
        if rank > 0 and k1 > 0:
            peer = hdl.get_buffer(rank - 1, (n,), t.dtype)
            hist = peer[n - k1:]
        return _engine.fir_filter(self.plan, t, hist=hist)
