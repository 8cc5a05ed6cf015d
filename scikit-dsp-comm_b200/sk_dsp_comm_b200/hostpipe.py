"""Host-buffer FIR path: chunked, three-stream pipeline (H2D | kernel | D2H overlapped).

This is what ``multirate_FIR.filter`` runs when it is handed a long HOST array (a torch CPU
tensor, ideally pinned): the stream is cut into chunks, every chunk carries the previous
``ntaps-1`` samples as its overlap-save halo (``hist`` argument of ``b200dsp_fir_filter``), and
copies of chunk i+1 / i-1 overlap the kernel of chunk i.  The halo reproduces the filter state
exactly: for float64 / complex128 the result is bit-identical to the monolithic device call; the
float32 / complex64 tensor-core kernels anchor their tile grid (block scales, summation order) at each
launch's first sample, so there a different cut changes the last bits (both within 1e-6 max|y| of the oracle).

bench.py's ``e2e`` number is measured through this path (host buffers in, host buffers out).
"""
from __future__ import annotations

import torch

from . import _engine

DEFAULT_CHUNK = 1 << 24      # samples per chunk (128 MiB of complex64)
NBUF = 3


def bind_to_gpu_numa(device_index: int):
    """Pin the calling process to the CPU cores of the NUMA node its GPU hangs off (PCI topology in sysfs), so that
    pinned staging buffers are first-touched in that node's memory and the copy threads run next to the GPU.
    One process per GPU (torchrun): without this every rank allocates on node 0 and the host<->device copies of
    several ranks share one memory controller / one PCIe root.  Returns a dict describing what was done."""
    import os
    info = {"node": None, "cpus": None}
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"node": node, "cpus": len(allowed)}
    except Exception:           # noqa: BLE001 -- topology files missing (containers): leave the affinity alone
        pass
    return info


class _Pipe:
    """Device staging buffers + streams, cached per (device, dtype, chunk, ntaps)."""

    def __init__(self, dev, dtype, chunk, k1):
        self.k1 = k1
        # the chunk itself starts at a 64-sample boundary so the tensor-core path (16-byte aligned
        # bulk-TMA loads) is taken; the halo sits right in front of it
        self.ka = (k1 + 63) // 64 * 64
        self.xbuf = [torch.empty(self.ka + chunk, dtype=dtype, device=dev) for _ in range(NBUF)]
        self.ybuf = [torch.empty(chunk, dtype=dtype, device=dev) for _ in range(NBUF)]
        self.s_h2d = torch.cuda.Stream(dev)
        self.s_cmp = torch.cuda.Stream(dev)
        self.s_d2h = torch.cuda.Stream(dev)


_pipes = {}          # insertion-ordered: least recently used first
_MAX_PIPES = 4       # bounded: each pipe owns 6 device buffers of ~chunk samples


def _pipe(dev, dtype, chunk, k1):
    """Staging buffers are keyed on a chunk size rounded up to a power of two (callers slice what they need), and
    the cache is a small LRU: streams of varying length no longer grow device memory without bound."""
    cap = 1 << max(int(chunk - 1).bit_length(), 10)
    key = (dev.index, dtype, cap, k1)
    p = _pipes.pop(key, None)
    if p is None:
        while len(_pipes) >= _MAX_PIPES:
            _pipes.pop(next(iter(_pipes)))          # drop the least recently used set of buffers
        p = _Pipe(dev, dtype, cap, k1)
    _pipes[key] = p
    return p


def fir_filter_host(plan: _engine.FirPlan, x: torch.Tensor, out=None, chunk: int = DEFAULT_CHUNK,
                    device=None) -> torch.Tensor:
    """y = FIR(x) for a 1-D host tensor x; returns a pinned host tensor (or fills ``out``)."""
    if x.is_cuda or x.dim() != 1:
        raise ValueError("fir_filter_host expects a 1-D CPU tensor")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = x.numel()
    k1 = plan.ntaps - 1
    xs = x.contiguous()
    y = out if out is not None else torch.empty(n, dtype=x.dtype, pin_memory=True)
    if n == 0:
        return y
    chunk = max(min(chunk, n), 1)
    p = _pipe(dev, x.dtype, chunk, k1)
    ka = p.ka
    cur = torch.cuda.current_stream(dev)
    for s in (p.s_h2d, p.s_cmp, p.s_d2h):
        s.wait_stream(cur)
    ev_free_x = [None] * NBUF       # kernel that last read xbuf[j] has finished
    ev_free_y = [None] * NBUF       # D2H that last read ybuf[j] has finished
    i = 0
    for c0 in range(0, n, chunk):
        ln = min(chunk, n - c0)
        j = i % NBUF
        xb, yb = p.xbuf[j], p.ybuf[j]
        with torch.cuda.stream(p.s_h2d):
            if ev_free_x[j] is not None:
                p.s_h2d.wait_event(ev_free_x[j])
            xb[ka:ka + ln].copy_(xs[c0:c0 + ln], non_blocking=True)
            if c0 > 0 and k1 > 0:
                xb[ka - k1:ka].copy_(xs[c0 - k1:c0], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(p.s_h2d)
        with torch.cuda.stream(p.s_cmp):
            p.s_cmp.wait_event(ev_in)
            if ev_free_y[j] is not None:
                p.s_cmp.wait_event(ev_free_y[j])
            hist = xb[ka - k1:ka] if (c0 > 0 and k1 > 0) else None
            _engine.fir_filter(plan, xb[ka:ka + ln], hist=hist, out=yb[:ln])
            ev_k = torch.cuda.Event()
            ev_k.record(p.s_cmp)
            ev_free_x[j] = ev_k
        with torch.cuda.stream(p.s_d2h):
            p.s_d2h.wait_event(ev_k)
            y[c0:c0 + ln].copy_(yb[:ln], non_blocking=True)
            ev_o = torch.cuda.Event()
            ev_o.record(p.s_d2h)
            ev_free_y[j] = ev_o
        i += 1
    cur.wait_stream(p.s_d2h)
    p.s_d2h.synchronize()          # the caller gets a host tensor: it must be complete
    return y


def bytes_per_call(n: int, itemsize: int, ntaps: int, chunk: int = DEFAULT_CHUNK):
    """(h2d_bytes, d2h_bytes) moved by fir_filter_host for an n-sample stream."""
    nchunks = (n + chunk - 1) // chunk
    return n * itemsize + max(nchunks - 1, 0) * (ntaps - 1) * itemsize, n * itemsize
