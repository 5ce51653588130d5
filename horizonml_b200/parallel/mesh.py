"""Process mesh: lets the three strategies compose (DP × PP × TP).

The reference runs every strategy on the single WORLD group (SURVEY §2.3 "Hybrid … NO (despite the
'hybrid' tagline, README.md:5)"); this mesh is what its tagline promises.  Rank layout on one NVSwitch
box: tensor-parallel ranks are adjacent (they exchange tiles inside fused kernels), pipeline stages
next, data-parallel replicas outermost:

    rank = (dp_index * pp + pp_index) * tp + tp_index

Every rank creates every sub-group in the same order (a ``torch.distributed.new_group`` requirement).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch.distributed as dist


@dataclass(frozen=True)
class MeshCoord:
    dp: int
    pp: int
    tp: int


class DeviceMesh:
    def __init__(self, world: int, rank: int, dp: int = 1, pp: int = 1, tp: int = 1, create_groups: bool = True):
        if dp * pp * tp != world:
            raise ValueError(f"mesh dp={dp} x pp={pp} x tp={tp} != world_size {world}")
        self.world, self.rank = world, rank
        self.dp, self.pp, self.tp = dp, pp, tp
        self.coord = self.coord_of(rank)
        self.dp_group = self.pp_group = self.tp_group = None
        if create_groups and world > 1 and dist.is_available() and dist.is_initialized():
            self._make_groups()

    # ---- layout ---------------------------------------------------------------------------------
    def coord_of(self, rank: int) -> MeshCoord:
        tp_i = rank % self.tp
        pp_i = (rank // self.tp) % self.pp
        dp_i = rank // (self.tp * self.pp)
        return MeshCoord(dp_i, pp_i, tp_i)

    def rank_of(self, dp: int, pp: int, tp: int) -> int:
        return (dp * self.pp + pp) * self.tp + tp

    def dp_ranks(self, pp: Optional[int] = None, tp: Optional[int] = None) -> List[int]:
        c = self.coord
        return [self.rank_of(d, c.pp if pp is None else pp, c.tp if tp is None else tp) for d in range(self.dp)]

    def pp_ranks(self, dp: Optional[int] = None, tp: Optional[int] = None) -> List[int]:
        c = self.coord
        return [self.rank_of(c.dp if dp is None else dp, s, c.tp if tp is None else tp) for s in range(self.pp)]

    def tp_ranks(self, dp: Optional[int] = None, pp: Optional[int] = None) -> List[int]:
        c = self.coord
        return [self.rank_of(c.dp if dp is None else dp, c.pp if pp is None else pp, t) for t in range(self.tp)]

    def pp_neighbours(self) -> Tuple[Optional[int], Optional[int]]:
        """(global rank of the previous stage, of the next stage) inside this rank's pipeline."""
        c = self.coord
        prev = self.rank_of(c.dp, c.pp - 1, c.tp) if c.pp > 0 else None
        nxt = self.rank_of(c.dp, c.pp + 1, c.tp) if c.pp < self.pp - 1 else None
        return prev, nxt

    # ---- groups ---------------------------------------------------------------------------------
    def _make_groups(self) -> None:
        for p in range(self.pp):
            for t in range(self.tp):
                ranks = [self.rank_of(d, p, t) for d in range(self.dp)]
                g = dist.new_group(ranks) if self.dp > 1 else None
                if self.rank in ranks:
                    self.dp_group = g
        for d in range(self.dp):
            for t in range(self.tp):
                ranks = [self.rank_of(d, p, t) for p in range(self.pp)]
                g = dist.new_group(ranks) if self.pp > 1 else None
                if self.rank in ranks:
                    self.pp_group = g
        for d in range(self.dp):
            for p in range(self.pp):
                ranks = [self.rank_of(d, p, t) for t in range(self.tp)]
                g = dist.new_group(ranks) if self.tp > 1 else None
                if self.rank in ranks:
                    self.tp_group = g

    def describe(self) -> dict:
        return {"dp": self.dp, "pp": self.pp, "tp": self.tp, "rank": self.rank,
                "coord": [self.coord.dp, self.coord.pp, self.coord.tp]}
