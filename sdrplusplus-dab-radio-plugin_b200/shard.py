"""Stream -> GPU partitioning.  Streams (tuners/ensembles) are independent, so the only multi-GPU logic is which
rank owns which stream; there is no data-path collective."""
from __future__ import annotations

from typing import List


def streams_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Stream s lives on rank s % world (round robin keeps per-GPU load equal when streams come and go)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_streams, world))


def owner_of(stream: int, world: int) -> int:
    return stream % world
