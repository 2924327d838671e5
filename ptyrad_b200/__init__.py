"""ptyrad_b200 -- B200-native implementation of PtyRAD's per-batch multislice hot path.

Drop-in for ``ptyrad.models.PtychoAD`` and ``ptyrad.losses.CombinedLoss`` (same names, arguments and behaviour);
the arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI (``include/ptyrad_b200.h``).
"""
from .models import PtychoAD, LazyPatches            # noqa: F401
from .losses import CombinedLoss, MeasurementView    # noqa: F401

__version__ = "0.1.0"
