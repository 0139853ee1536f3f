"""gpsat_b200: B200-native batched engine for GPSat's local-expert optimal-interpolation hot path.

The numerical path is hand-written sm_100a CUDA in libgpsat_b200.so (C ABI in include/gpsat_b200.h);
this package is the Python host side that mirrors the reference's BaseGPRModel / LocalExpertOI
interfaces.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import KERNEL_IDS, OPT_STATUS, GpsatError  # noqa: F401


def get_engine(device: int = 0):
    """Process-wide Engine per device (created on first use)."""
    from .engine import Engine
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = Engine(device)
    return eng


_ENGINES = {}
