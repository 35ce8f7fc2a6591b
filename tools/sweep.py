#!/usr/bin/env python
"""Device-resident throughput sweep over codecs and chunk sizes (CUDA-event times, L2 flushed between runs).
Usage: python tools/sweep.py [--size N] [--codecs rcs2,ans4s] [--chunks 1024,2048,4096] [--src zipf|bwt] [--reps 5]
Prints one JSON line per (codec, chunk)."""
import argparse, importlib, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trc = importlib.import_module("turbo-range-coder_b200")
dg = importlib.import_module("turbo-range-coder_b200.datagen")
CODECS = {"ans4s": 0, "ans4": 1, "ans": 2, "ans1": 3, "rcs": 4, "rcs2": 5, "rc": 6, "rci": 7, "rc4": 8, "rc4i": 9, "answ": 10, "rc8": 11, "rci8": 12,
          "ansu16": 13, "ansuz16": 14, "ansv16": 15, "ansvz16": 16, "ansv32": 17, "ansvz32": 18, "rcv16": 19, "rcvz16": 20, "rcv32": 21, "rcvz32": 22, "rcu16": 23, "rcu32": 24}

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=100_000_000)
ap.add_argument("--codecs", default="rcs2,rcs,ans4s")
ap.add_argument("--chunks", default="1024,2048,4096,16384,65536")
ap.add_argument("--src", default="zipf")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--nibble", action="store_true", help="mask input to 4 bits (16-symbol alphabet)")
a = ap.parse_args()

data = {"zipf": dg.zipf, "bwt": dg.bwt_shaped, "o1": dg.markov1, "uniform": dg.uniform}[a.src](a.size)
if a.nibble:
    data = data & 15
d_in = torch.from_numpy(data).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cdf_dev, _ = trc.cdfini_dev(d_in, a.size, a.size)
trc.profile_enable(True)
for cname in a.codecs.split(","):
    codec = CODECS[cname]
    if codec in (1, 8, 9) and not a.nibble:
        continue
    for chunk in [int(c) for c in a.chunks.split(",")]:
        static = codec in (0, 4, 5, 10)
        b = trc.DeviceBatch(codec, a.size, chunk, cdfnum=256 if static else 0)
        if static:
            b.cdf = cdf_dev
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        te = td = 0.0
        kms = None
        for r in range(a.reps + 1):
            flush.zero_(); ev[0].record(); b.encode(d_in); ev[1].record()
            ke = trc.profile_read()
            flush.zero_(); ev[2].record(); b.decode(); ev[3].record()
            kd = trc.profile_read()
            torch.cuda.synchronize()
            if r:
                te += ev[0].elapsed_time(ev[1]) / a.reps; td += ev[2].elapsed_time(ev[3]) / a.reps
                kms = [round(x, 4) for x in ke + kd]
        ok = bool(torch.equal(b.dec[:a.size], d_in))
        clen = b.compressed_len()
        print(json.dumps({"codec": cname, "chunk": chunk, "src": a.src, "size": a.size, "ratio": round(clen / a.size, 5), "ok": ok,
                          "enc_ms": round(te, 4), "dec_ms": round(td, 4), "enc_gbs": round(a.size / te / 1e6, 2), "dec_gbs": round(a.size / td / 1e6, 2),
                          "encdec_gbs": round(a.size / (te + td) / 1e6, 2), "kernel_ms[enc,scan,pack,dec]": kms}), flush=True)
        del b
