"""cffm_b200 -- B200-native implementation of the CFFM training / scoring hot path.

Host side in Python (mirrors the reference's ``CFFM`` class, ``LoadData`` and CLI); all
arithmetic in ``libcffm_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/cffm.h``).  There is no CPU fallback."""
from ._lib import CffmError, build, device_available, load  # noqa: F401
from .engine import Engine, comm_unique_id  # noqa: F401
from .model import CFFM  # noqa: F401
from .data import LoadData  # noqa: F401

__all__ = ["CFFM", "LoadData", "Engine", "CffmError", "build", "load", "device_available", "comm_unique_id"]
