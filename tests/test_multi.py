"""CPU tests of the multi-GPU plumbing (onesolver_b200/multi.py): sharding by global trajectory
id and the single best-energy gather, exercised with world_size-2 gloo process groups."""
import os
import socket

import numpy as np
import pytest

from onesolver_b200 import multi


def test_shards_cover_the_id_range_exactly():
    for total, world in [(1_048_576, 8), (100, 3), (7, 8), (65536, 4)]:
        spans = [multi.shard(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1


def test_encode_decode_picks_min_energy_then_lowest_id():
    n = 77
    rng = np.random.default_rng(0)
    states = rng.integers(0, 2, size=(4, n)).astype(np.uint8)
    rows = np.stack([multi.encode_best(-3.5, 900, states[0]), multi.encode_best(-7.25, 1500, states[1]),
                     multi.encode_best(-7.25, 1200, states[2]), multi.encode_best(-1.0, 3, states[3])])
    e, idx, s = multi.decode_best(rows, n)
    assert (e, idx) == (-7.25, 1200) and (s == states[2]).all()


def test_ranks_without_trajectories_never_win():
    """num_tries < world: the high ranks get count 0 and contribute an infinite energy with the
    all-ones id; decode_best ignores them."""
    assert [multi.shard(3, 8, r)[1] for r in range(8)] == [1, 1, 1, 0, 0, 0, 0, 0]
    n = 40
    empty = multi.encode_best(np.inf, 2**64 - 1, np.zeros(n, dtype=np.uint8))
    real = multi.encode_best(12.5, 2, np.ones(n, dtype=np.uint8))
    e, idx, s = multi.decode_best(np.stack([empty, real, empty]), n)
    assert (e, idx) == (12.5, 2) and s.all()
    with pytest.raises(ValueError):
        multi.decode_best(np.stack([empty, empty]), n)


def _worker(rank, world, port, n, out_q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    state = rng.integers(0, 2, size=n).astype(np.uint8)
    first, count = multi.shard(1000, world, rank)
    energy = [-5.0, -9.5][rank]
    index = first + 7
    e, idx, s = multi.gather_best(dist, torch, energy, index, state, torch.device("cpu"))
    out_q.put((rank, e, idx, s.tolist(), state.tolist(), first, count))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_best_world_size_2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 4096
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, e0, i0, s0, own0, f0, c0), (r1, e1, i1, s1, own1, f1, c1) = results
    assert (f0, c0, f1, c1) == (0, 500, 500, 500)
    assert e0 == e1 == -9.5 and i0 == i1 == 507      # every rank agrees on the winner (rank 1)
    assert s0 == own1 and s1 == own1
