"""Topology plan: which layers each stage owns, who belongs to which cluster, and which
rank/GPU hosts which client.

Layer-range arithmetic is the reference's (src/Server.py:222-228): with cut list ``c`` and
``L`` stages, stage 1 owns ``[0, c[0]]``, the last ``[c[-1], -1]``, stage k ``[c[k-2], c[k-1]]``;
``c == [0]`` means "whole model on stage 1" → ``[0, 0]``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


def stage_layers(layer_id: int, num_stages: int, cut_layers: Sequence[int]) -> List[int]:
    """``layers=[a, b]`` for a 1-based ``layer_id`` (b == -1 → to the end)."""
    if num_stages == 1:
        return [0, -1]
    if layer_id == 1:
        return [0, int(cut_layers[0])]
    if layer_id == num_stages:
        return [int(cut_layers[-1]), -1]
    return [int(cut_layers[layer_id - 2]), int(cut_layers[layer_id - 1])]


def resolve_range(layers: Sequence[int], total_layers: int) -> Tuple[int, int]:
    """Concrete ``(start, end]`` for a START ``layers`` pair."""
    a, b = int(layers[0]), int(layers[1])
    if b == 0:
        return 0, total_layers
    return a, total_layers if b == -1 else b


@dataclass
class ClientInfo:
    client_id: str
    layer_id: int
    profile: Optional[dict] = None
    cluster: int = 0
    label_counts: List[int] = field(default_factory=list)
    train: bool = True
    rank: Optional[int] = None          # process rank / GPU ordinal hosting this client
    idx: Optional[int] = None           # 2LS device index
    extras: Dict = field(default_factory=dict)


@dataclass
class ClusterPlan:
    cluster_id: int
    cut_layers: List[int]
    members: List[List[str]]            # members[stage-1] = client ids


@dataclass
class Topology:
    """Static plan for one round (what ``cluster_and_selection`` produces in the reference)."""
    num_stages: int
    clusters: List[ClusterPlan]

    def infor_cluster(self) -> List[List[int]]:
        return [[len(m) for m in c.members] for c in self.clusters]

    def cut_layers(self) -> List[List[int]]:
        return [c.cut_layers for c in self.clusters]

    def layers_for(self, cluster: int, layer_id: int) -> List[int]:
        return stage_layers(layer_id, self.num_stages, self.clusters[cluster].cut_layers)

    def peers(self, cluster: int, layer_id: int) -> List[str]:
        return list(self.clusters[cluster].members[layer_id - 1])


def rank_assignment(clients_per_stage: Sequence[int], infor_cluster: Optional[Sequence[Sequence[int]]] = None
                    ) -> List[Tuple[int, int, int]]:
    """Deterministic ``rank -> (layer_id, cluster, index_in_cluster_stage)`` map used by
    ``launch.py``/``bench.py``: ranks are laid out cluster-major, stage-major so that the two
    ends of a cut edge sit on neighbouring GPUs."""
    out = []
    if infor_cluster is None:
        infor_cluster = [list(clients_per_stage)]
    for c, info in enumerate(infor_cluster):
        for s, n in enumerate(info):
            for i in range(n):
                out.append((s + 1, c, i))
    return out
