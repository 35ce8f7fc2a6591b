#!/usr/bin/env python
"""NVLink ingress into ONE GPU: every rank != 0 moves `MB` megabytes into rank 0's buffer at the same time, (a) with the
device-driven push kernel (csrc/pack.cuh k_push), (b) with copy engines (cudaMemcpyAsync to the peer mapping), whole or split
over several streams.  torchrun --nproc-per-node N tools/push_probe.py [MB]"""
import ctypes, importlib, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
trc = importlib.import_module("turbo-range-coder_b200"); shard = importlib.import_module("turbo-range-coder_b200.shard")
trc.lib.trc_set_device(local)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 72
n = mb * 1_000_000 // 4096 * 4096
src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
ln = torch.tensor([n], dtype=torch.int64, device=dev)
pg = shard.PeerGather(n + 4096, dst=0, depth=64)
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
res = {"gpus": world, "mb_per_push": mb}
reps = 8


def timed(fn):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    t = torch.tensor([ms if rank != 0 else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round((world - 1) * n / float(t) / 1e6, 1)           # total GB/s into rank 0


def push_kernel():
    if rank != 0:
        pg.push(src, ln.data_ptr())
    else:
        pg.seq += 1


def copy_split(k):
    def f():
        if rank == 0:
            return
        part = n // k
        cur = torch.cuda.current_stream()
        for q in range(k):
            st = streams[q] if k > 1 else cur
            if k > 1:
                st.wait_stream(cur)
            trc.lib.trc_memcpy_dev(ctypes.c_void_p(pg.slot_ptr(rank, 1) + q * part), ctypes.c_void_p(src.data_ptr() + q * part), ctypes.c_size_t(part),
                                   ctypes.c_void_p(st.cuda_stream))
    return f


res["push_kernel_ingress_gbs"] = timed(push_kernel)
for k in (1, 2, 4):
    res[f"copy_engine_x{k}_ingress_gbs"] = timed(copy_split(k))
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
