from .runtime import B200Runtime, B200RuntimeConfig
from .store import FrameStore

__all__ = ["B200Runtime", "B200RuntimeConfig", "FrameStore"]
