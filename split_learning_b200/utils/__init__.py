"""Small shared helpers."""
from .timing import CudaTimer, Watchdog, capture_graph

__all__ = ["CudaTimer", "Watchdog", "capture_graph"]
