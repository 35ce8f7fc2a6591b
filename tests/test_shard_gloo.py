"""CPU suite: the multi-GPU plumbing (partition + gather of the compressed chunks) with world_size 2 over gloo."""
import os
import socket
import importlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("turbo-range-coder_b200.shard")
    rng = np.random.default_rng(100 + rank)
    n = [1000, 37, 0, 5000][rank % 4] if world > 1 else 1000
    payload = torch.from_numpy(rng.integers(0, 256, n, dtype=np.uint8))
    buf, offs = shard.gather_compressed(payload, dst=0)
    if rank == 0:
        ok = True
        for r in range(world):
            exp = np.random.default_rng(100 + r).integers(0, 256, [1000, 37, 0, 5000][r % 4], dtype=np.uint8)
            ok &= np.array_equal(buf[int(offs[r]):int(offs[r + 1])].numpy(), exp)
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_compressed_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_partition():
    shard = importlib.import_module("turbo-range-coder_b200.shard")
    assert shard.partition(125, 8)[0] == (0, 16) and shard.partition(125, 8)[-1][1] == 125
    for nb, w in [(1, 8), (7, 8), (128, 8), (125, 8), (24, 2), (5, 3)]:
        parts = shard.partition(nb, w)
        assert parts[0][0] == 0 and parts[-1][1] == nb
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    sb = shard.shard_bytes(8_000_000_000, 64_000_000, 8)
    assert sb[0] == (0, 1_024_000_000) and sb[-1][1] == 8_000_000_000
