"""sk_dsp_comm_b200 -- B200-native (sm_100a) FIR / SOS-IIR / integer multirate engine behind the
Python surface of mwickert/scikit-dsp-comm's ``multirate_helper`` and ``sigsys.upsample`` /
``downsample``.

    import sk_dsp_comm_b200.multirate_helper as mrh     # was: sk_dsp_comm.multirate_helper
    import sk_dsp_comm_b200.sigsys as ss                # was: sk_dsp_comm.sigsys (up/downsample)

Importing the package loads ``_lib/libb200dsp.so`` (hand-written CUDA, built by
``scikit-dsp-comm_b200/build.py``) and raises ImportError if it is missing: there is no CPU,
PyTorch or scipy fallback anywhere in this package.
"""
from . import _cabi          # noqa: F401  (fails loudly when the CUDA library is absent)
from . import sigsys         # noqa: F401
from . import multirate_helper  # noqa: F401
from .sigsys import upsample, downsample          # noqa: F401
from .multirate_helper import multirate_FIR, multirate_IIR   # noqa: F401

__version__ = "0.1.0"
