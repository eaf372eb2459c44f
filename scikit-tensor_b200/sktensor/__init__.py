"""scikit-tensor's CP-ALS / MTTKRP / HOOI path on NVIDIA B200 (sm_100a).

Drop-in for the ``sktensor`` package of mnick/scikit-tensor: the same names
(``dtensor``, ``sptensor``, ``ktensor``, ``cp_als``, ``tucker_hooi`` ...), the same
signatures, defaults and quirks, with the numerical work done by hand-written CUDA kernels
in ``libsktensor_b200.so`` (C ABI: ``include/sktensor_b200.h``).  There is no CPU fallback.
"""
from .version import __version__

from .utils import *          # noqa: F401,F403
from .core import *           # noqa: F401,F403

from .sptensor import sptensor, unfolded_sptensor
from .dtensor import dtensor, unfolded_dtensor
from .ktensor import ktensor

from .cp import als as cp_als
from .tucker import hooi as tucker_hooi
# the reference exports hooi under this name as well (sktensor/__init__.py:14); kept for parity
from .tucker import hooi as tucker_hosvd
