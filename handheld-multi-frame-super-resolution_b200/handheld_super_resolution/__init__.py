"""B200-native drop-in for the hot path of Jamy-L/Handheld-Multi-Frame-Super-Resolution.

Same package name and function surface as the reference (`from handheld_super_resolution import process`,
handheld_super_resolution/__init__.py:8); every stage runs hand-written sm_100a CUDA from libhhsr.so."""
from .super_resolution import main, process  # noqa: F401
from .config import Config, load_config  # noqa: F401

__all__ = ["process", "main", "Config", "load_config"]
