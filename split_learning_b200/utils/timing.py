"""Timing / liveness helpers used by trainers, benchmarks and tools."""
from __future__ import annotations

import threading
import time
from typing import Callable


class CudaTimer:
    """CUDA-event stopwatch on the current stream (``with CudaTimer() as t: ...; t.ms``)."""

    def __init__(self):
        import torch
        self._torch = torch
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)
        self.ms = 0.0

    def __enter__(self):
        self.e0.record()
        return self

    def __exit__(self, *exc):
        self.e1.record()
        self.e1.synchronize()
        self.ms = self.e0.elapsed_time(self.e1)
        return False


class Watchdog:
    """Fires ``on_timeout`` when ``kick()`` has not been called for ``seconds`` (failure detection:
    the reference dead-locks silently when a client dies)."""

    def __init__(self, seconds: float, on_timeout: Callable[[], None]):
        self.seconds, self.on_timeout = seconds, on_timeout
        self._last = time.monotonic()
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def start(self):
        self._t.start()
        return self

    def kick(self):
        self._last = time.monotonic()

    def stop(self):
        self._stop.set()

    def _run(self):
        while not self._stop.wait(min(1.0, self.seconds / 4)):
            if time.monotonic() - self._last > self.seconds:
                self.on_timeout()
                return


_CAPTURE_LOCK = threading.Lock()


def capture_graph(stream, fn):
    """Capture ``fn()`` (kernel launches on ``stream``) into a CUDA graph.

    Unlike ``torch.cuda.graph`` this does NOT ``torch.cuda.synchronize()`` the whole device first: a sibling stage
    sharing the GPU (threads in one process) may have a flag-spinning kernel in flight that only *our* next launch can
    release, so a device-wide sync here would dead-lock.  Captures are serialised process-wide and use thread-local
    capture mode so other threads may keep launching."""
    import torch
    g = torch.cuda.CUDAGraph()
    # No stream/device synchronisation while holding the lock: a sibling thread may need the lock to *launch* the very
    # work our pending kernels are waiting for (observed dead-lock: stage 1 draining B(k) <- L(k) not yet captured).
    with _CAPTURE_LOCK:
        with torch.cuda.stream(stream):
            g.capture_begin(capture_error_mode="thread_local")
            try:
                fn()
            finally:
                g.capture_end()
    return g
