"""Per-client label-count allocation (reference src/Server.py:87-101).

IID: ``num-sample // num-label`` of every label for every first-layer client.
non-IID: one Dirichlet(alpha) draw per client scaled by ``num-sample`` and truncated.
FLEX ships a fixed 9x4 "non-iid-rate" style matrix (other/FLEX/src/Server.py:79-93); the
``non_iid_rate`` mode here generalises it: a fraction ``rate`` of each client's samples is
concentrated on its own dominant labels, the rest is spread uniformly.
Unlike the reference (quirk C4) the numpy RNG *is* seeded from ``random-seed``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def preset_matrix(name: str, num_clients: int, num_label: int) -> np.ndarray:
    """Fixed non-IID label-probability matrices of the variants, generalised to any client / label count.
    ``flex`` (other/FLEX/src/Server.py:79-93): 75 % of one of the first L-1 labels + 25 % of the last label, the
    dominant label following the reference's client order 0,1,0,1,2,1,2,0,2.  ``2ls`` (other/2LS/src/Server.py:84-103):
    contiguous label blocks with in-block weights .3/.3/.3/.1."""
    m = np.zeros((num_clients, num_label))
    if name == "flex":
        order = [0, 1, 0, 1, 2, 1, 2, 0, 2]
        for c in range(num_clients):
            m[c, order[c % len(order)] % max(num_label - 1, 1)] = 0.75
            m[c, num_label - 1] += 0.25
    elif name == "2ls":
        blk = [0.3, 0.3, 0.3, 0.1]
        n_blocks = max(1, (num_label - 1) // 3)
        for c in range(num_clients):
            b0 = (c * n_blocks // max(num_clients, 1)) % n_blocks * 3
            for j, w in enumerate(blk):
                m[c, min(b0 + j, num_label - 1)] += w
    else:
        raise ValueError(f"unknown label-matrix preset {name}")
    return m


def label_counts(num_clients: int, num_label: int, num_sample: int, non_iid: bool = False,
                 alpha: float = 1.0, seed: Optional[int] = None, non_iid_rate: Optional[float] = None,
                 matrix=None) -> np.ndarray:
    """``matrix``: explicit per-client label probabilities (rows cycle) or a preset name (``flex`` / ``2ls``)."""
    if num_clients <= 0:
        return np.zeros((0, num_label), dtype=np.int64)
    if not non_iid:
        return np.full((num_clients, num_label), num_sample // num_label, dtype=np.int64)
    if matrix is not None:
        m = preset_matrix(matrix, num_clients, num_label) if isinstance(matrix, str) else np.asarray(matrix, dtype=float)
        rows = np.stack([m[c % len(m)] for c in range(num_clients)])
        return (rows[:, :num_label] * num_sample).astype(np.int64)
    rng = np.random.RandomState(seed if seed is not None else None)
    if non_iid_rate is not None:
        rate = float(non_iid_rate)
        out = np.zeros((num_clients, num_label), dtype=np.int64)
        dom = max(1, num_label // max(num_clients, 1))
        for c in range(num_clients):
            major = [(c * dom + j) % num_label for j in range(dom)]
            base = int(num_sample * (1.0 - rate)) // num_label
            out[c, :] = base
            extra = int(num_sample * rate) // len(major)
            for m in major:
                out[c, m] += extra
        return out
    dist = rng.dirichlet([alpha] * num_label, num_clients)
    return (dist * num_sample).astype(np.int64)
