"""Row sharding for multi-GPU fits (host-side planning; no reference counterpart -- SURVEY 8e).

Rows are partitioned by the highest-cardinality fixed effect so that each of its groups lives on exactly
one rank: that FE's tables, its singleton test and (when it is the cluster key) the cluster scores stay
rank-local, and only the other FEs' small tables, the Gram matrix and the meat cross NVLink.
"""
from __future__ import annotations

import numpy as np


def key_range_cuts(key_counts: np.ndarray, world: int) -> np.ndarray:
    """Cut the sorted distinct-key axis into ``world`` contiguous ranges balanced by ROW count.

    ``key_counts[g]`` = rows of the g-th smallest key.  Returns ``cuts`` (world+1,) with rank r owning
    key ranks ``[cuts[r], cuts[r+1])``.
    """
    key_counts = np.asarray(key_counts, dtype=np.int64)
    csum = np.cumsum(key_counts)
    total = int(csum[-1]) if csum.size else 0
    targets = (np.arange(1, world) * total) // world
    inner = np.searchsorted(csum, targets, side="left") + 1 if csum.size else np.zeros(world - 1, dtype=np.int64)
    cuts = np.concatenate(([0], np.minimum(inner, key_counts.size), [key_counts.size])).astype(np.int64)
    return np.maximum.accumulate(cuts)


def shard_rows_by_key(keys: np.ndarray, world: int) -> list[np.ndarray]:
    """Row index arrays, one per rank; all rows of a key land on one rank; row order is preserved."""
    uniq, inv, cnt = np.unique(np.asarray(keys), return_inverse=True, return_counts=True)
    cuts = key_range_cuts(cnt, world)
    owner = np.searchsorted(cuts, inv.reshape(-1), side="right") - 1
    return [np.flatnonzero(owner == r) for r in range(world)]
