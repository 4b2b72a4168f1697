"""split_learning_b200 — a Blackwell-native split-learning engine."""
__version__ = "0.1.0"

import os as _os

# Kernels that spin on mailbox flags must never wait for a *lazily loaded* sibling kernel
# (CUDA's default lazy module loading can block a first launch until running kernels drain).
_os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
