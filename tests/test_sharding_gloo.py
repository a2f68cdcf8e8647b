"""Multi-rank host logic on CPU: pairs shard across ranks with no data-path collective, and the one
collective of the path (final gather of the ragged index arrays) reassembles them in order.
Runs world_size = 2 over gloo (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from densematcher_b200 import pipeline


def test_shard_pairs_partitions_every_pair_once():
    for n in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            blocks = [pipeline.shard_pairs(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)                       # same "dataset" on every rank
        sizes = rng.integers(5, 40, size=n_pairs)
        full = [rng.integers(0, 1000, size=int(s)) for s in sizes]
        lo, hi = pipeline.shard_pairs(n_pairs, rank, world)
        local = torch.from_numpy(np.concatenate(full[lo:hi]) if hi > lo else np.zeros(0, np.int64))
        counts = [int(sum(sizes[a:b])) for a, b in (pipeline.shard_pairs(n_pairs, r, world) for r in range(world))]
        got = pipeline.gather_results(local, counts)
        ok = np.array_equal(got.numpy(), np.concatenate(full))
        q.put((rank, bool(ok), int(got.numel())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [9, 2])
def test_gather_results_world2_gloo(n_pairs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True] and res[0][2] == res[1][2]


def test_ranks_of_a_category_sorted_bank_prepare_disjoint_mesh_ranges():
    """cfg5 shape: meshes sorted by category, every ordered intra-category pair, pairs sharded in contiguous blocks.
    The range of meshes a rank has to prepare (``pairs_mesh_range``) covers the pairs of its block, the ranges of all
    ranks cover the bank, and neighbouring ranks overlap by at most the one category their boundary cuts through."""
    rng = np.random.default_rng(5000)
    cats = np.sort(rng.integers(0, 24, size=599))
    src, dst = pipeline.intra_category_pairs(cats)
    assert len(src) == sum(c * (c - 1) for c in np.bincount(cats)) and np.all(cats[src] == cats[dst])
    for world in (1, 2, 8):
        ranges = []
        for r in range(world):
            lo, hi = pipeline.shard_pairs(len(src), r, world)
            a, b = pipeline.pairs_mesh_range(src, dst, lo, hi)
            ids = np.concatenate([src[lo:hi], dst[lo:hi]])
            assert a == ids.min() and b == ids.max() + 1
            ranges.append((a, b))
        assert ranges[0][0] == 0 and ranges[-1][1] == 599
        largest = int(np.bincount(cats).max())
        for (a0, b0), (a1, b1) in zip(ranges[:-1], ranges[1:]):
            assert a1 <= b0 and b0 - a1 <= largest          # contiguous cover, overlap within one category
        if world == 8:
            assert sum(b - a for a, b in ranges) <= 599 + 7 * largest
    assert pipeline.pairs_mesh_range(src, dst, 5, 5) == (0, 0)
