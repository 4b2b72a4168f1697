"""split_learning_b200 — a Blackwell-native split-learning engine."""
__version__ = "0.1.0"
