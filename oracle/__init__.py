"""CPU oracle for the FIR / SOS-IIR / integer multirate hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package;
the product package (``scikit-dsp-comm_b200/``) never does and fails loudly when its
CUDA library is missing instead of falling back to anything here.

What is restated, and from where (reference = mwickert/scikit-dsp-comm at
``/root/reference``; arithmetic lives in the un-vendored third-party dependency
``scipy`` -- ``requirements.txt:3`` says ``scipy>=1.1.0`` (unpinned); the dev container
and the GPU boxes carry scipy 1.18.1 / numpy 2.3.5):

* ``fir_filter``  -- ``multirate_FIR.filter``  ``src/sk_dsp_comm/multirate_helper.py:104-109``
  = ``scipy.signal.lfilter(b,[1],x)``; for ``len(a)==1`` scipy evaluates
  ``np.convolve(b, x)[:len(x)]`` in ``dtype = np.result_type(b, a, x)``.
* ``fir_up`` / ``fir_dn`` -- ``multirate_helper.py:112-127``.
* ``sos_filter`` / ``sos_up`` / ``sos_dn`` -- ``multirate_helper.py:169-192``
  = ``scipy.signal.sosfilt`` (direct-form-II-transposed cascade, zero initial state).
* ``upsample`` / ``downsample`` -- ``src/sk_dsp_comm/sigsys.py:3031-3083``.

Pinning (see ``tests/test_oracle_golden.py``): the oracle is checked against
(1) the reference's own literal known-answer vectors that touch this arithmetic
(``tests/test_sigsys.py:28-34`` ten-band biquad cascade, ``:688-706`` boxcar FIR,
``:655-668`` up/downsample shape + TypeError contract) and (2) outputs of the
UNMODIFIED reference classes run in the dev container and committed as fixtures
(``tests/golden/*.npz`` written by ``tests/golden/make_golden.py``).

Two back-ends with identical semantics:
* numpy (``np.convolve`` -- the very call scipy makes for the FIR branch), and
* ``oracle.c`` (plain C, doubles, OpenMP over outputs) loaded through ctypes when
  ``oracle/_build/liboracle.so`` exists; it is required for SOS inputs longer than a
  few thousand samples (the pure-Python biquad loop is only for small cases).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c (gcc + OpenMP).  Building the checker is not using it."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/liboracle.so"])
    return _SO


def _clib():
    global _lib
    if _lib is None and os.path.exists(_SO):
        lib = ctypes.CDLL(_SO)
        P, I64, I32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        lib.oracle_num_threads.restype = ctypes.c_int
        lib.oracle_fir_f64.argtypes = [P, I32, P, P, I64, I32, P]
        lib.oracle_fir_up_f64.argtypes = [P, I32, P, I64, I32, I32, P]
        lib.oracle_fir_dn_f64.argtypes = [P, I32, P, I64, I32, I32, P]
        lib.oracle_sosfilt_f64.argtypes = [P, I32, P, I64, I32, P, P, P]
        lib.oracle_lfilter_ba_f64.argtypes = [P, P, I32, P, I64, P]
        lib.oracle_upsample_bytes.argtypes = [P, I64, I32, I32, P]
        lib.oracle_downsample_bytes.argtypes = [P, I64, I32, I32, I32, P]
        _lib = lib
    return _lib


def have_c() -> bool:
    return _clib() is not None


def num_threads() -> int:
    lib = _clib()
    return int(lib.oracle_num_threads()) if lib else 1


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _as_real_view(x, dt):
    """Return (float64 2-D-ish flat view, nch) for a 1-D real/complex array in dtype dt."""
    x = np.ascontiguousarray(x, dtype=dt)
    if np.iscomplexobj(x):
        return x, x.view(np.float64), 2
    return x, x, 1


def result_dtype(coeffs, x):
    """scipy's promotion rule for lfilter/sosfilt: np.result_type(coeffs, [1], x),
    restricted to dtype chars 'fdgFDGO' (ints promote through the float64 coefficients)."""
    dt = np.result_type(np.asarray(coeffs).dtype, np.asarray(x).dtype, np.float64)
    if dt.char not in "dD":
        dt = np.dtype(np.complex128) if dt.kind == "c" else np.dtype(np.float64)
    return dt


# --------------------------------------------------------------------------- FIR
def fir_filter(b, x, hist=None, backend="numpy"):
    """``multirate_FIR.filter`` (multirate_helper.py:104-109).

    ``hist`` (optional) holds the ``len(b)-1`` samples that precede ``x`` -- the
    overlap-save halo used by the sharded path; ``None`` = zeros, which is
    the reference's stateless behaviour.
    """
    b = np.asarray(b)
    # complex taps (lfilter accepts them): numpy path only, the result is complex128
    b = b.astype(np.complex128) if np.iscomplexobj(b) else b.astype(np.float64)
    x = np.asarray(x)
    if x.ndim != 1:
        return np.apply_along_axis(lambda v: fir_filter(b, v, None, backend), -1, x)
    dt = result_dtype(b, x)
    if np.iscomplexobj(b):
        backend = "numpy"
    n, K = x.shape[0], b.shape[0]
    if backend == "c" and have_c():
        xs, xv, nch = _as_real_view(x, dt)
        y = np.empty(n, dtype=dt)
        h = None
        if hist is not None:
            h = np.ascontiguousarray(hist, dtype=dt)
            assert h.shape[0] == K - 1
        _clib().oracle_fir_f64(_ptr(b), K, _ptr(xv), _ptr(h) if h is not None else None,
                               n, nch, _ptr(y))
        return y
    xe = x.astype(dt)
    if hist is not None:
        hist = np.asarray(hist, dtype=dt)
        assert hist.shape[0] == K - 1
        full = np.convolve(b, np.concatenate([hist, xe]))
        return full[K - 1:K - 1 + n].astype(dt)
    if n == 0:
        return np.zeros(0, dtype=dt)
    return np.convolve(b, xe)[:n].astype(dt)


def upsample(x, L):
    """``sigsys.upsample`` (sigsys.py:3031-3053): zero-stuff; output dtype >= float64
    because the zeros come from ``np.zeros`` (float64) -- restated literally."""
    N_input = len(x)
    y = np.hstack((x.reshape(N_input, 1), np.zeros((N_input, int(L - 1)))))
    return y.flatten()


def downsample(x, M, p=0):
    """``sigsys.downsample`` (sigsys.py:3056-3083): keep every M-th sample, phase p."""
    if not isinstance(M, int):
        raise TypeError("M must be an int")
    x = x[0:int(np.floor(len(x) / M)) * M]
    x = x.reshape((int(np.floor(len(x) / M)), M))
    return x[:, p]


def fir_up(b, x, L, backend="numpy"):
    """``multirate_FIR.up`` (multirate_helper.py:112-118)."""
    b = np.asarray(b)
    if np.iscomplexobj(b):
        return fir_filter(b, L * upsample(x, L))
    b = b.astype(np.float64)
    x = np.asarray(x)
    if backend == "c" and have_c():
        dt = result_dtype(b, x)
        xs, xv, nch = _as_real_view(x, dt)
        y = np.empty(x.shape[0] * int(L), dtype=dt)
        _clib().oracle_fir_up_f64(_ptr(b), b.shape[0], _ptr(xv), x.shape[0], nch, int(L), _ptr(y))
        return y
    return fir_filter(b, L * upsample(x, L))


def fir_dn(b, x, M, backend="numpy"):
    """``multirate_FIR.dn`` (multirate_helper.py:121-127)."""
    b = np.asarray(b)
    if np.iscomplexobj(b):
        return downsample(fir_filter(b, x), M)
    b = b.astype(np.float64)
    x = np.asarray(x)
    if backend == "c" and have_c():
        dt = result_dtype(b, x)
        xs, xv, nch = _as_real_view(x, dt)
        y = np.empty(x.shape[0] // int(M), dtype=dt)
        _clib().oracle_fir_dn_f64(_ptr(b), b.shape[0], _ptr(xv), x.shape[0], nch, int(M), _ptr(y))
        return y
    return downsample(fir_filter(b, x), M)


# --------------------------------------------------------------------------- SOS IIR
def validate_sos(sos):
    """scipy ``_validate_sos``: (n_sections, 6) with sos[:,3] == 1, else ValueError."""
    sos = np.atleast_2d(np.asarray(sos))
    if sos.ndim != 2:
        raise ValueError("sos array must be 2D")
    n_sections, m = sos.shape
    if m != 6:
        raise ValueError("sos array must be shape (n_sections, 6)")
    if not (sos[:, 3] == 1).all():
        raise ValueError("sos[:, 3] should be all ones")
    return sos, n_sections


def _sosfilt_py(sos, x, zi=None):
    """Pure-Python DF-II-T loop -- small cases only."""
    nsec = sos.shape[0]
    z = np.zeros((nsec, 2), dtype=x.dtype) if zi is None else np.array(zi, dtype=x.dtype)
    y = np.empty_like(x)
    for n in range(x.shape[0]):
        x_cur = x[n]
        for s in range(nsec):
            x_new = sos[s, 0] * x_cur + z[s, 0]
            z[s, 0] = sos[s, 1] * x_cur - sos[s, 4] * x_new + z[s, 1]
            z[s, 1] = sos[s, 2] * x_cur - sos[s, 5] * x_new
            x_cur = x_new
        y[n] = x_cur
    return y, z


def sos_filter(sos, x, zi=None, return_zf=False):
    """``multirate_IIR.filter`` (multirate_helper.py:169-174) = ``sosfilt(sos, x)``.

    ``zi``/``zf`` have shape (n_sections, 2) like scipy's; ``None`` = zero state
    (the reference never passes ``zi``).
    """
    sos, nsec = validate_sos(sos)
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    x = np.asarray(x)
    if x.ndim != 1:
        assert zi is None and not return_zf
        return np.apply_along_axis(lambda v: sos_filter(sos, v), -1, x)
    dt = result_dtype(sos, x)
    xs, xv, nch = _as_real_view(x, dt)
    lib = _clib()
    if lib is not None and nsec <= 64:
        y = np.empty(xs.shape[0], dtype=dt)
        zi_a = None
        if zi is not None:
            zi_a = np.ascontiguousarray(zi, dtype=dt)
            assert zi_a.shape == (nsec, 2)
        zf = np.empty((nsec, 2), dtype=dt)
        lib.oracle_sosfilt_f64(_ptr(sos), nsec, _ptr(xv), xs.shape[0], nch,
                               _ptr(zi_a) if zi_a is not None else None, _ptr(zf), _ptr(y))
    else:
        if xs.shape[0] > 20000:
            raise RuntimeError("oracle: build oracle/_build/liboracle.so for long SOS inputs")
        y, zf = _sosfilt_py(sos, xs, zi)
    return (y, zf) if return_zf else y


def sos_up(sos, x, L):
    """``multirate_IIR.up`` (multirate_helper.py:177-183)."""
    return sos_filter(sos, L * upsample(np.asarray(x), L))


def sos_dn(sos, x, M):
    """``multirate_IIR.dn`` (multirate_helper.py:186-192)."""
    return downsample(sos_filter(sos, x), M)


def lfilter_ba(b, a, x):
    """``scipy.signal.lfilter(b, a, x)`` for real float64 data (direct form II transposed, zero state):
    the arithmetic behind ``rate_change`` (multirate_helper.py:73-74,81-82) and ``interp24/deci24``."""
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    nc = max(len(b), len(a))
    bp = np.zeros(nc)
    ap = np.zeros(nc)
    bp[:len(b)] = b
    ap[:len(a)] = a
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib = _clib()
    if lib is not None and nc <= 64:
        lib.oracle_lfilter_ba_f64(_ptr(bp), _ptr(ap), nc, _ptr(x), x.shape[0], _ptr(y))
        return y
    z = np.zeros(nc)
    for i in range(x.shape[0]):                      # small cases only
        yi = bp[0] / ap[0] * x[i] + z[0]
        for k in range(1, nc):
            z[k - 1] = bp[k] / ap[0] * x[i] - ap[k] / ap[0] * yi + (z[k] if k + 1 < nc else 0.0)
        y[i] = yi
    return y


def rate_change_up(b, a, x, M):
    """``rate_change.up`` (multirate_helper.py:69-75)."""
    return lfilter_ba(b, a, M * upsample(np.asarray(x), M))


def rate_change_dn(b, a, x, M):
    """``rate_change.dn`` (multirate_helper.py:77-83)."""
    return downsample(lfilter_ba(b, a, x), M)


def iir_order(sos):
    """``multirate_IIR.__init__`` order bookkeeping (multirate_helper.py:163-164)."""
    sos = np.asarray(sos)
    return np.sum(np.sign(np.abs(sos[:, 2]))) + np.sum(np.sign(np.abs(sos[:, 1])))


# --------------------------------------------------------------------------- CPU baseline helpers
# (bench.py cpu_baseline / --impl reference legs only)
_BASE_CACHE = {}


def _baseline_chunk(task):
    """Worker: one halo-chunk of the reference FIR call on its own seeded complex64 data.
    The input is generated once per worker and cached so that timed passes measure
    ``np.convolve`` (what lfilter's FIR branch executes, in complex128) and nothing else."""
    seed, n, b_bytes = task[:3]
    kind = task[3] if len(task) > 3 else "port"
    b = np.frombuffer(b_bytes, dtype=np.float64)
    key = (seed, n)
    x = _BASE_CACHE.get(key)
    if x is None:
        rng = np.random.default_rng(seed)
        x = (rng.standard_normal(n + len(b) - 1, dtype=np.float32)
             + 1j * rng.standard_normal(n + len(b) - 1, dtype=np.float32)).astype(np.complex64)
        _BASE_CACHE.clear()
        _BASE_CACHE[key] = x
    K = len(b)
    if kind == "scipy":
        # the reference's literal call (multirate_helper.py:108) on the chunk with its halo in front
        from scipy import signal
        y = signal.lfilter(b, [1], x)[K - 1:]
    else:
        y = fir_filter(b, x[K - 1:], hist=x[:K - 1])     # np.convolve in complex128
    return float(np.abs(y[:16]).sum())


class FirCpuBaseline:
    """All-cores run of the reference's FIR arithmetic on halo-chunks:
    ``cores`` worker processes, each filtering ``chunk`` complex64 samples per pass."""

    def __init__(self, b, cores=None, chunk=1 << 21, kind="auto"):
        import multiprocessing as mp
        import os
        if kind == "auto":
            try:
                from scipy import signal  # noqa: F401
                kind = "scipy"
            except Exception:
                kind = "port"
        self.kind = kind              # "scipy": scipy.signal.lfilter(b,[1],x) itself; "port": the np.convolve restatement
        self.cores = cores or len(os.sched_getaffinity(0))
        self.chunk = chunk
        self.b_bytes = np.asarray(b, dtype=np.float64).tobytes()
        self.pool = mp.get_context("fork").Pool(self.cores)      # created before any CUDA init
        self.tasks = [(1000 + i, chunk, self.b_bytes, kind) for i in range(self.cores)]
        self.samples_per_pass = self.cores * chunk
        self.run_pass()                                    # generate inputs, warm numpy

    def run_pass(self):
        return self.pool.map(_baseline_chunk, self.tasks, chunksize=1)

    def close(self):
        self.pool.close()
        self.pool.join()
