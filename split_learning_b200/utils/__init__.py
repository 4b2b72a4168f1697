"""Small shared helpers."""
from .timing import CudaTimer, Watchdog

__all__ = ["CudaTimer", "Watchdog"]
