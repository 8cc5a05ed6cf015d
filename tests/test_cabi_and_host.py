"""CPU-only checks: the C-ABI library loads and exports every symbol include/b200dsp.h declares,
host-side error contracts fire before any CUDA call, and the product package has no CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200dsp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200dsp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import sk_dsp_comm_b200._cabi as cabi
    syms = _header_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert sorted(cabi.SYMBOLS) == syms
    assert lib.b200dsp_version() == 200


def test_library_is_sm100a_native():
    """The shipped library must carry sm_100a SASS (no PTX-JIT of some other arch)."""
    import subprocess
    import sk_dsp_comm_b200._cabi as cabi
    out = subprocess.run(["cuobjdump", "--list-elf", cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_product_does_not_import_oracle_or_scipy():
    pkg = os.path.join(ROOT, "scikit-dsp-comm_b200", "sk_dsp_comm_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            s = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", s, flags=re.M), f
            assert not re.search(r"^\s*(import|from)\s+scipy\b", s, flags=re.M), f
            assert "/root/reference" not in s, f


def test_no_cpu_path_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sk_dsp_comm_b200.multirate_helper as mrh
    import sk_dsp_comm_b200.sigsys as ss
    with pytest.raises(RuntimeError, match="no CPU path"):
        mrh.multirate_FIR(np.ones(4)).filter(np.ones(8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ss.upsample(np.ones(8), 2)


def test_error_contracts_before_any_cuda_call(filters):
    import sk_dsp_comm_b200.multirate_helper as mrh
    import sk_dsp_comm_b200.sigsys as ss
    # /root/reference/tests/test_sigsys.py:665-668
    with pytest.raises(TypeError, match="M must be an int"):
        ss.downsample(np.zeros(0), 3.0)
    with pytest.raises(IndexError):
        ss.downsample(np.zeros(9), 3, 3)             # p >= M  (sigsys.py:3082)
    with pytest.raises(AttributeError):
        ss.upsample([1.0, 2.0], 2)                   # list has no .reshape (sigsys.py:3051)
    with pytest.raises(ValueError):
        ss.upsample(np.zeros((2, 3)), 2)             # 2-D reshape error
    with pytest.raises(TypeError, match="M must be an int"):
        mrh.multirate_FIR(filters["b7"]).dn(np.zeros(8), 2.0)
    # scipy _validate_sos contract, raised at call time like the reference
    bad = filters["sos6"].copy()
    bad[0, 3] = 2.0
    with pytest.raises(ValueError, match="all ones"):
        mrh.multirate_IIR(bad).filter(np.zeros(4))
    with pytest.raises(ValueError, match="n_sections, 6"):
        mrh.multirate_IIR(np.zeros((2, 5))).filter(np.zeros(4))
    with pytest.raises(ValueError):
        mrh.multirate_FIR(np.zeros((2, 2)))


def test_attributes_and_logging(filters, caplog):
    import logging
    import sk_dsp_comm_b200.multirate_helper as mrh
    with caplog.at_level(logging.INFO, logger="sk_dsp_comm_b200.multirate_helper"):
        f = mrh.multirate_FIR(filters["b256"])
        g = mrh.multirate_IIR(filters["sos6"])
    assert f.N_forder == 256 and f.b is filters["b256"]
    assert g.N_forder == 12 and g.sos is filters["sos6"]
    assert "FIR filter taps = 256" in caplog.text          # multirate_helper.py:101
    assert "IIR filter order = 12" in caplog.text          # multirate_helper.py:166
    import inspect
    assert str(inspect.signature(f.up)) == "(x, L_change=12)"
    assert str(inspect.signature(f.dn)) == "(x, M_change=12)"
    assert str(inspect.signature(g.up)) == "(x, L_change=12)"
    assert str(inspect.signature(g.dn)) == "(x, M_change=12)"
    import sk_dsp_comm_b200.sigsys as ss
    assert str(inspect.signature(ss.downsample)) == "(x, M, p=0)"
    assert str(inspect.signature(ss.upsample)) == "(x, L)"


def test_segment_bounds():
    from sk_dsp_comm_b200.sharded import segment_bounds
    n, w = 1003, 4
    segs = [segment_bounds(n, w, r, align=4) for r in range(w)]
    assert segs[0][0] == 0 and segs[-1][1] == n
    for (a, b), (c, d) in zip(segs, segs[1:]):
        assert b == c and a % 4 == 0 and c % 4 == 0


def test_hostpipe_staging_cache_is_bounded(monkeypatch):
    """hostpipe keeps its device staging buffers in a small LRU keyed on a power-of-two chunk size (ADVICE.md round 1:
    one set of buffers per distinct stream length grew device memory without bound).  Host logic only: the buffer
    class is replaced by a stub, no CUDA call is made."""
    import torch
    from sk_dsp_comm_b200 import hostpipe

    made = []

    class FakePipe:
        def __init__(self, dev, dtype, chunk, k1):
            self.chunk, self.k1 = chunk, k1
            made.append(chunk)

    monkeypatch.setattr(hostpipe, "_Pipe", FakePipe)
    monkeypatch.setattr(hostpipe, "_pipes", {})
    dev = torch.device("cuda", 0)
    a = hostpipe._pipe(dev, torch.complex64, 1000000, 255)
    b = hostpipe._pipe(dev, torch.complex64, 1048576, 255)
    assert a is b and a.chunk == 1 << 20 and made == [1 << 20]            # same power-of-two bucket: one set of buffers
    for e in range(11, 11 + 2 * hostpipe._MAX_PIPES):
        hostpipe._pipe(dev, torch.complex64, 1 << e, 255)
        assert len(hostpipe._pipes) <= hostpipe._MAX_PIPES
    keys = list(hostpipe._pipes)
    hostpipe._pipe(dev, torch.complex64, keys[0][2], 255)                   # touching the oldest makes it the newest
    assert list(hostpipe._pipes)[-1] == keys[0]


def test_numa_binding_is_a_noop_without_topology(monkeypatch):
    """bind_to_gpu_numa must leave the process alone where the PCI topology is hidden (containers, the VMs the GPU
    boxes are) instead of raising"""
    import torch
    from sk_dsp_comm_b200 import hostpipe

    def boom(_):
        raise RuntimeError("no CUDA here")

    monkeypatch.setattr(torch.cuda, "get_device_properties", boom)
    before = os.sched_getaffinity(0)
    assert hostpipe.bind_to_gpu_numa(0) == {"node": None, "cpus": None}
    assert os.sched_getaffinity(0) == before
