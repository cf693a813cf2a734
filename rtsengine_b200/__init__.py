"""B200-native agent-update hot path of RTSEngine behind a C ABI (include/rts_gpu.h)."""
from . import abi  # noqa: F401

__all__ = ["abi", "engine", "scenario"]
