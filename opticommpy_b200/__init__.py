"""B200-native (sm_100a) drop-in for OptiCommPy's two data-parallel hot loops.

    from opticommpy_b200.channels import ssfm, manakovSSF
    from opticommpy_b200.equalization import edc, mimoAdaptEqualizer, manakovDBP
    from opticommpy_b200.carrierRecovery import bps, cpr
    from opticommpy_b200.utils import parameters

Host code is Python (this package); all arithmetic runs in hand-written CUDA kernels reached
through the C-ABI of ``libopticomm_b200.so`` (``include/opticomm_b200.h``).  There is no CPU
fallback: importing the compute modules works anywhere, calling them needs a B200.
"""
from .utils import parameters  # noqa: F401

__version__ = "0.1.0"
__all__ = ["parameters", "channels", "equalization", "carrierRecovery", "modulation", "sharding"]
