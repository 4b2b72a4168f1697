"""Topology planning: clustering, device selection, cut-point search."""
from .cluster import clustering_algorithm, kmeans
from .partition import partition, partition_multi
from .selection import auto_threshold, gmm_1d

__all__ = ["clustering_algorithm", "kmeans", "partition", "partition_multi", "auto_threshold", "gmm_1d"]
