"""Worker of tests/test_multi_gpu.py::test_peer_gather (one process per GPU, torchrun): every rank encodes its own shard, pushes
the packed stream into rank 0's buffer over peer memory, rank 0 waits for the flags ON THE DEVICE and checks every slot against
the oracle; then every rank fetches its own slot back (the scatter) and decodes it."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    trc = importlib.import_module("turbo-range-coder_b200")
    shard = importlib.import_module("turbo-range-coder_b200.shard")
    dg = importlib.import_module("turbo-range-coder_b200.datagen")
    from oracle import cpu
    from helpers import cpu_batch
    trc.lib.trc_set_device(local)
    n, chunk = 4_000_000 + 160 * rank, 1760
    d = dg.zipf(n, seed=100 + rank)
    cdf = cpu.port().cdfini(d)
    t = torch.from_numpy(d).to(dev)
    b = trc.DeviceBatch(trc.RCS2, n, chunk, cdfnum=256, device=dev)
    b.set_cdf(cdf)
    pg = shard.PeerGather(b.out.numel(), dst=0, depth=2, split=3)                  # hinted pushes of >= 1 MiB go out in three pieces
    total_ptr = b.off.data_ptr() + 8 * b.n
    side = torch.cuda.Stream(device=dev)
    for step in range(1, 6):                                 # several rounds through the two slot sets
        b.encode(t)
        ev = torch.cuda.Event(); ev.record()
        side.wait_event(ev)
        with torch.cuda.stream(side):
            seq = pg.push(b.out, total_ptr, side, hint_bytes=(0, 1 << 20, 1 << 30)[step % 3])   # no hint / too short / too long
        assert seq == step
        if rank == 0:
            pg.wait_all(seq)                                  # device-side: the current stream continues only when all streams landed
            lens = pg.read_lens(dev, seq).cpu().numpy()
            got = [pg.read_slot(r, int(lens[r]), dev, seq).cpu().numpy() for r in range(world)]
            assert not pg.overflowed(dev)
            for r in range(world):
                dr = dg.zipf(4_000_000 + 160 * r, seed=100 + r)
                want, _ = cpu_batch(cpu.port(), trc.RCS2, dr[:50 * chunk], chunk, cpu.port().cdfini(dr), 256)
                assert lens[r] > 0 and np.array_equal(got[r][:want.size], want), (step, r)
            pg.ack(seq)                                       # slot set free again: push(seq + depth) may proceed
        # the mirror: fetch the own slot back from rank 0 and decode it
        back = torch.zeros_like(b.out)
        with torch.cuda.stream(side):
            pg.fetch(rank, seq, back, side)
        side.synchronize()
        clen = b.compressed_len()
        assert torch.equal(back[:clen], b.out[:clen])
        assert torch.equal(b.decode(back, b.off), t)
        # (no barrier: push(seq + 2) waits on the device for rank 0's acknowledgement of seq)
    pg.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("peer gather ok")


if __name__ == "__main__":
    main()
