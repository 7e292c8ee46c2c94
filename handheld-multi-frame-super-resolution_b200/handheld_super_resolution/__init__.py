"""B200-native drop-in for the hot path of Jamy-L/Handheld-Multi-Frame-Super-Resolution.

Same package name and function surface as the reference (`from handheld_super_resolution import process`,
handheld_super_resolution/__init__.py:8); every stage runs hand-written sm_100a CUDA from libhhsr.so."""
import os as _os

# The pipeline runs six streams at once (compute, two alignment chains, upload, result copy of the caller, NCCL).  With
# the default of 8 hardware queues, PyTorch's stream pools (32 streams per priority) share queues at random and a copy can
# end up queued behind another stream's waiting copies; more queues make that unlikely.  Only effective when set before
# the CUDA context exists, never overrides the user's choice.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .super_resolution import main, process  # noqa: F401,E402
from .config import Config, load_config  # noqa: F401,E402

__all__ = ["process", "main", "Config", "load_config"]
