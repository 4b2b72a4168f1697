"""Label-histogram clustering of first-layer devices.

Behaviour of reference src/Cluster.py:5-21: L1-normalise each device's label counts, run
k-means, return (labels, [[count]] per cluster).  The reference delegates to sklearn's
``KMeans(random_state=42)``; ``kmeans`` below is a dependency-free Lloyd/k-means++
implementation with multiple restarts.  ``backend="sklearn"`` reproduces the reference's
exact assignment when scikit-learn is importable (used for parity tests).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def _l1_normalize(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    n = np.abs(x).sum(axis=1, keepdims=True)
    n[n == 0] = 1.0
    return x / n


def kmeans(x: np.ndarray, k: int, seed: int = 42, n_init: int = 10, iters: int = 300) -> Tuple[np.ndarray, np.ndarray]:
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    k = min(k, n)
    rng = np.random.RandomState(seed)
    best = (np.inf, None, None)
    for _ in range(n_init):
        # k-means++ seeding
        centers = [x[rng.randint(n)]]
        for _c in range(1, k):
            d2 = np.min([((x - c) ** 2).sum(1) for c in centers], axis=0)
            tot = d2.sum()
            centers.append(x[rng.randint(n)] if tot <= 0 else x[np.searchsorted(np.cumsum(d2 / tot), rng.rand())
                                                                 .clip(0, n - 1)])
        centers = np.stack(centers)
        labels = np.zeros(n, dtype=np.int64)
        for _it in range(iters):
            d = ((x[:, None, :] - centers[None]) ** 2).sum(-1)
            new = d.argmin(1)
            for c in range(k):
                if (new == c).any():
                    centers[c] = x[new == c].mean(0)
            if (new == labels).all() and _it > 0:
                break
            labels = new
        inertia = ((x - centers[labels]) ** 2).sum()
        if inertia < best[0] - 1e-12:
            best = (inertia, labels.copy(), centers.copy())
    return best[1], best[2]


def clustering_algorithm(label_counts, num_cluster: int, backend: str = "auto"):
    x = _l1_normalize(np.asarray(label_counts))
    labels = None
    if backend in ("auto", "sklearn"):
        try:
            from sklearn.cluster import KMeans  # parity with the reference when available
            labels = KMeans(n_clusters=num_cluster, random_state=42).fit(x).labels_
        except Exception:
            if backend == "sklearn":
                raise
    if labels is None:
        labels, _ = kmeans(x, num_cluster, seed=42)
    counts = np.bincount(labels, minlength=num_cluster)
    return np.asarray(labels), [[int(c)] for c in counts]
