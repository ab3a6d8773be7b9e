"""Multi-GPU plumbing.  The path shards by independent units — scenes (accumulators) or
BEV samples — with NO collective on the data path: rank r owns the units
i with i % world == r, its own ring, maps and output planes.  torch.distributed is used
only for the barrier around timed regions and for reducing a small vector of summary
statistics (points, BEVs, seconds) at the end.
"""
from __future__ import annotations

import os


def rank_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def shard_units(n_units: int, rank: int, world: int) -> list:
    """Unit ids owned by `rank` (round robin: equal work under weak scaling)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world {world}')
    return list(range(rank, n_units, world))


def reduce_stats(dist, stats: dict, device=None) -> dict:
    """Whole-job statistics from per-rank ones: keys ending in `_max` (elapsed times) take
    the maximum over ranks, everything else the sum.  `dist` is torch.distributed (already
    initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(stats)
    import torch
    keys = sorted(stats)
    kw = {'device': device} if device is not None else {}
    s = torch.tensor([float(stats[k]) for k in keys if not k.endswith('_max')],
                     dtype=torch.float64, **kw)
    m = torch.tensor([float(stats[k]) for k in keys if k.endswith('_max')],
                     dtype=torch.float64, **kw)
    if s.numel():
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    if m.numel():
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.endswith('_max'):
            out[k] = float(m[mi])
            mi += 1
        else:
            out[k] = float(s[si])
            si += 1
    return out
