"""Logging: ANSI console helper + file logger (API of reference src/Log.py:4-50 and the
2LS extras: ``minimal`` mode, log-dir creation, no duplicate handlers, other/2LS/src/Log.py:17-37)."""
from __future__ import annotations

import logging
import os
import sys
import threading

_ANSI = {"header": "\033[95m", "blue": "\033[94m", "green": "\033[92m", "yellow": "\033[93m",
         "red": "\033[91m", "end": "\033[0m"}
_QUIET = os.environ.get("SLB200_QUIET", "0") == "1"
_lock = threading.Lock()


def set_quiet(q: bool) -> None:
    global _QUIET
    _QUIET = bool(q)


def print_with_color(text, color: str = "end") -> None:
    if _QUIET:
        return
    code = _ANSI.get(str(color).lower(), _ANSI["end"])
    with _lock:
        sys.stdout.write(f"{code}{text}{_ANSI['end']}\n")
        sys.stdout.flush()


class Logger:
    def __init__(self, log_path: str = "app.log", debug_mode: bool = False, minimal: bool = False,
                 name: str = "split_learning_b200"):
        d = os.path.dirname(os.path.abspath(log_path))
        os.makedirs(d, exist_ok=True)
        self.debug_mode, self.minimal = debug_mode, minimal
        self.logger = logging.getLogger(f"{name}:{os.path.abspath(log_path)}")
        self.logger.setLevel(logging.DEBUG)
        self.logger.propagate = False
        for h in list(self.logger.handlers):
            self.logger.removeHandler(h)
        fh = logging.FileHandler(log_path)
        fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        self.logger.addHandler(fh)

    def log_info(self, message):
        if not self.minimal and not _QUIET:
            print(f"[INFO] {message}")
        self.logger.info(message)

    def log_warning(self, message):
        print_with_color(f"[WARN] {message}", "yellow")
        self.logger.warning(message)

    def log_error(self, message):
        print_with_color(f"[ERROR] {message}", "red")
        self.logger.error(message)

    def log_debug(self, message):
        if self.debug_mode:
            if not self.minimal:
                print_with_color(f"[DEBUG] {message}", "green")
            self.logger.debug(message)
