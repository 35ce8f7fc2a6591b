#!/usr/bin/env python
"""TRC_ANS1 (order-1) throughput against the chunk size on 1 GB of order-1 Markov bytes (BASELINE config 4), device resident."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
trc = importlib.import_module("turbo-range-coder_b200")
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
d = bench.markov1_dev(torch, n, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for chunk in (4 << 20, 1 << 20, 256 << 10, 64 << 10):
    b = trc.DeviceBatch(trc.ANS1, n, chunk, device=dev)
    b.encode(d); back = b.decode(); torch.cuda.synchronize()
    assert torch.equal(back, d), chunk
    e, dd = bench.time_batch(trc, torch, b, d, flush, 2, warmup=1)
    print(json.dumps({"codec": "ans1", "bytes": n, "chunk": chunk, "n_calls": b.n, "enc_gbs": round(n / e / 1e6, 3), "dec_gbs": round(n / dd / 1e6, 3),
                      "value": round(n / (e + dd) / 1e6, 3), "ratio": round(b.compressed_len() / n, 5)}))
    del b
