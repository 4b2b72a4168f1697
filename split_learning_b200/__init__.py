"""split_learning_b200 — a Blackwell-native split-learning engine."""
__version__ = "0.1.0"

# NOTE on CUDA lazy module loading: a kernel that spins on a mailbox flag must not be the reason a
# sibling kernel's *first* launch blocks (lazy loading may wait for running kernels).  The pipelines
# therefore run every program once, sequentially, before any overlapped/graph execution
# (``DeviceStage._exec`` eager first pass).  Forcing CUDA_MODULE_LOADING=EAGER instead costs minutes
# of start-up (torch's fatbins) — measured, rejected.
