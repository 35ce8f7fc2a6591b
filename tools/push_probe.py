#!/usr/bin/env python
"""Bandwidth of the device-driven peer push (csrc/pack.cuh k_push) on otherwise idle GPUs: every rank != 0 pushes `MB` megabytes into
rank 0's buffer `reps` times.  torchrun --nproc-per-node N tools/push_probe.py [MB]"""
import importlib, json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
trc = importlib.import_module("turbo-range-coder_b200"); shard = importlib.import_module("turbo-range-coder_b200.shard")
trc.lib.trc_set_device(local)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 72
n = mb * 1_000_000 // 16 * 16
src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
ln = torch.tensor([n], dtype=torch.int64, device=dev)
pg = shard.PeerGather(n + 4096, dst=0, depth=16)
res = {}
for active in ("one", "all"):
    dist.barrier()
    reps = 10
    doit = rank != 0 and (active == "all" or rank == 1)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps):
        if doit:
            pg.push(src, ln.data_ptr())
        else:
            pg.seq += 1
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    t = torch.tensor([n / ms / 1e6 if doit else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    # rank 0 acknowledges everything so that later pushes are not held back
    if rank == 0:
        for s in range(pg.seq - reps + 1, pg.seq + 1):
            pg.ack(s)
    torch.cuda.synchronize(); dist.barrier()
    res[f"{active}_pushers_total_gbs"] = round(float(t), 1)
if rank == 0:
    print(json.dumps({"gpus": world, "mb_per_push": mb, "ctas": os.environ.get("TRC_PUSH_CTAS", "32"), **res}))
dist.destroy_process_group()
