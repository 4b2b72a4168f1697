"""Data-plane trainers and stage executors (reference L2)."""
from .dataplane import HostDataPlane, QueueGrammar
from .executor import StageExecutor, TorchExecutor, make_executor
from .trainer import StageTrainer

__all__ = ["HostDataPlane", "QueueGrammar", "StageExecutor", "TorchExecutor", "make_executor", "StageTrainer"]
