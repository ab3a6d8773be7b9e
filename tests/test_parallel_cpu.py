"""N > 1 host logic on CPU: world_size-2 gloo processes shard scenes by rank, never
exchange data, and reduce only summary statistics (sum of counts, max of times)."""
import os
import socket

import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

from pc_accumulation_lib_b200 import parallel  # noqa: E402


def test_shard_units_partition():
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            mine = parallel.shard_units(37, r, world)
            assert all(u % world == r for u in mine)
            seen += mine
        assert sorted(seen) == list(range(37))
    with pytest.raises(ValueError):
        parallel.shard_units(4, 2, 2)


def _worker(rank, world, port, n_scenes, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = parallel.shard_units(n_scenes, rank, world)
        # each "scene" contributes a deterministic amount of work
        stats = {'points': sum(1000 + 7 * s for s in mine), 'bevs': 32 * len(mine),
                 'seconds_max': 1.0 + 0.25 * rank}
        dist.barrier()
        tot = parallel.reduce_stats(dist, stats)
        q.put((rank, mine, tot))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_stats():
    world, n_scenes = 2, 9
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(u for _, mine, _ in res for u in mine)
    assert owned == list(range(n_scenes))                   # a partition: no scene twice
    for _, _, tot in res:                                   # every rank sees the same totals
        assert tot['points'] == sum(1000 + 7 * s for s in range(n_scenes))
        assert tot['bevs'] == 32 * n_scenes
        assert tot['seconds_max'] == 1.25                   # MAX over ranks, not the sum


def test_reduce_stats_single_process():
    assert parallel.reduce_stats(None, {'a': 1, 't_max': 2.0}) == {'a': 1, 't_max': 2.0}
