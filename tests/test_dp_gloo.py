"""CPU, world_size 2 over gloo: the data-parallel host logic (batch sharding + bucketed averaging
all-reduce of a flat gradient buffer addressed by parameter-name prefixes)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uniception_b200 import dp


def test_shard_batch_partitions_exactly():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert dp.shard_pairs_symmetrized(8, 1, 2) == (4, 8)  # (a,b),(b,a) partners stay on one rank


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    index = {"encoder.a.weight": (0, torch.Size([4, 8])), "encoder.b.bias": (64, torch.Size([8])),
             "info_sharing.x.weight": (128, torch.Size([16, 4])), "head1.linear.bias": (192, torch.Size([5]))}
    flat = torch.full((256,), float(rank + 1))
    sync = dp.GradSync(flat, index, max_bucket_elems=48)
    sync.ready("head1.")           # backward order: heads, decoder, encoder
    sync.ready("info_sharing.")
    sync.finish()                   # everything not announced (encoder + padding) goes here
    expect = sum(range(1, world + 1)) / world
    ok = bool(torch.allclose(flat, torch.full_like(flat, expect)))
    # a second step must work too (state is reset by finish())
    flat.fill_(float(10 * (rank + 1)))
    sync.ready("encoder.")
    sync.finish()
    ok = ok and bool(torch.allclose(flat, torch.full_like(flat, 10 * expect)))
    out[rank] = ok
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert out[0] and out[1]
