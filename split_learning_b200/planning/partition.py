"""Cut-point search.

``partition`` reproduces reference src/Partition.py:2-21 for two stages: every candidate
cut ``c`` scores min(stage-1 rate, stage-2 rate) where a device's rate is
1 / (its compute time for its side + size_data[c] / its bandwidth); returns ``[c + 1]``
(1-based layer index).  Prefix sums make it O(L·D) instead of the reference's O(L²·D).
``partition_multi`` generalises to N stages by DP (our own contract — the reference's
auto-mode is hard-wired to 2 layers, src/Server.py:343).
"""
from __future__ import annotations

from itertools import accumulate
from typing import List, Sequence


def _prefix(xs: Sequence[float]) -> List[float]:
    return [0.0] + list(accumulate(xs))


def partition(exe_time_layer_1, net_layer_1, exe_time_layer_2, net_layer_2, size_data) -> List[int]:
    best, best_rate = 0, 0.0
    pre1 = [_prefix(e) for e in exe_time_layer_1]
    pre2 = [_prefix(e) for e in exe_time_layer_2]
    for c, size in enumerate(size_data):
        r1 = sum(1.0 / (p[min(c + 1, len(p) - 1)] + size / net) for p, net in zip(pre1, net_layer_1))
        r2 = sum(1.0 / ((p[-1] - p[min(c + 1, len(p) - 1)]) + size / net) for p, net in zip(pre2, net_layer_2))
        rate = min(r1, r2)
        if rate > best_rate:
            best, best_rate = c + 1, rate
    return [best]


def partition_multi(exe_times: Sequence[Sequence[Sequence[float]]], nets: Sequence[Sequence[float]],
                    size_data: Sequence[float]) -> List[int]:
    """N-stage generalisation: ``exe_times[s]`` = per-device layer-time vectors of stage s,
    ``nets[s]`` = per-device bandwidths.  Maximises the minimum stage rate; returns the
    N-1 increasing 1-based cut indices."""
    n_stage, n_layer = len(exe_times), len(size_data)
    if n_stage == 1:
        return []
    if n_stage == 2:
        return partition(exe_times[0], nets[0], exe_times[1], nets[1], size_data)
    pres = [[_prefix(e) for e in stage] for stage in exe_times]

    def rate(s, a, b):  # stage s runs layers (a, b]
        tot = 0.0
        for p, net in zip(pres[s], nets[s]):
            comm = 0.0
            if b < n_layer:
                comm += size_data[b - 1] / net
            if a > 0:
                comm += size_data[a - 1] / net
            t = p[min(b, len(p) - 1)] - p[min(a, len(p) - 1)] + comm
            tot += 1.0 / max(t, 1e-30)
        return tot

    NEG = -1.0
    # dp[s][b] = best min-rate using stages 0..s covering layers (0, b]
    dp = [[NEG] * (n_layer + 1) for _ in range(n_stage)]
    arg = [[0] * (n_layer + 1) for _ in range(n_stage)]
    for b in range(1, n_layer + 1):
        dp[0][b] = rate(0, 0, b)
    for s in range(1, n_stage):
        for b in range(s + 1, n_layer + 1):
            for a in range(s, b):
                if dp[s - 1][a] < 0:
                    continue
                v = min(dp[s - 1][a], rate(s, a, b))
                if v > dp[s][b]:
                    dp[s][b], arg[s][b] = v, a
    cuts, b = [], n_layer
    for s in range(n_stage - 1, 0, -1):
        b = arg[s][b]
        cuts.append(b)
    return cuts[::-1]
