"""B200-native (sm_100a) drop-in for OptiCommPy's two data-parallel hot loops.

    from opticommpy_b200.channels import ssfm, manakovSSF
    from opticommpy_b200.equalization import edc, mimoAdaptEqualizer, manakovDBP
    from opticommpy_b200.carrierRecovery import bps, cpr
    from opticommpy_b200.core import firFilter, decimate, pnorm, symbolSync, delaySignal
    from opticommpy_b200.devices import pdmCoherentReceiver, basicLaserModel
    from opticommpy_b200.tx import simpleWDMTx, pulseShape
    from opticommpy_b200.metrics import fastBERcalc
    from opticommpy_b200.utils import parameters
    from opticommpy_b200.sharding import shard_units, run_concurrent, gather_device     # units over GPUs / in flight per GPU

Host code is Python (this package); all arithmetic runs in hand-written CUDA kernels reached
through the C-ABI of ``libopticomm_b200.so`` (``include/opticomm_b200.h``).  There is no CPU
fallback: importing the compute modules works anywhere, calling them needs a B200.
"""
from .utils import parameters  # noqa: F401

__version__ = "0.1.0"
__all__ = ["parameters", "channels", "equalization", "carrierRecovery", "core", "devices", "tx", "metrics", "modulation",
           "rxchain", "pipelines", "sharding"]
