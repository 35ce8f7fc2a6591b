#!/usr/bin/env python
"""Host-side ceiling of the end-to-end path on N GPUs: concurrent pinned H2D + D2H on every GPU's PCIe link (one process per GPU,
torchrun).  Prints one JSON line on rank 0: per-GPU GB/s alone (rank 0 only active) and with all ranks active at once.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 100_000_000
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
h1.zero_(); h2.zero_()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d(); d2h()


def t(f, reps=10):
    f(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - a) / reps


def allmax(x):
    if world == 1:
        return x
    v = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v)


res = {"gpus": world, "bytes_per_copy": n, "host_cpus": os.cpu_count()}
# everyone active
for name, f in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    sec = allmax(t(f))
    res[f"all_{name}_gbs_per_gpu"] = round(n / sec / 1e9, 2)
    res[f"all_{name}_gbs_total"] = round(world * n * (2 if name == "both" else 1) / sec / 1e9, 1)
# rank 0 alone (the others idle at the barrier inside t())
for name, f in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    if rank == 0:
        f(); torch.cuda.synchronize()
        a = time.perf_counter()
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - a) / 10
        res[f"alone_{name}_gbs"] = round(n * (2 if name == "both" else 1) / sec / 1e9, 2)
    if world > 1:
        dist.barrier()
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
