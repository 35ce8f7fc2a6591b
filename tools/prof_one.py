#!/usr/bin/env python
"""Minimal driver for ncu: encode+decode `--reps` times for one codec/chunk on 100 MB Zipf (device resident)."""
import argparse, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trc = importlib.import_module("turbo-range-coder_b200")
dg = importlib.import_module("turbo-range-coder_b200.datagen")
CODECS = {"ans4s": 0, "ans4": 1, "ans": 2, "ans1": 3, "rcs": 4, "rcs2": 5, "rc": 6, "rci": 7, "rc4": 8, "rc4i": 9, "answ": 10}
ap = argparse.ArgumentParser()
ap.add_argument("--codec", default="rcs2"); ap.add_argument("--chunk", type=int, default=4096)
ap.add_argument("--size", type=int, default=100_000_000); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--src", default="zipf")
a = ap.parse_args()
data = {"zipf": dg.zipf, "bwt": dg.bwt_shaped}[a.src](a.size)
codec = CODECS[a.codec]
if codec in (1, 8, 9): data = data & 15
d_in = torch.from_numpy(data).cuda()
static = codec in (0, 4, 5, 10)
b = trc.DeviceBatch(codec, a.size, a.chunk, cdfnum=256 if static else 0)
if static:
    b.cdf, _ = trc.cdfini_dev(d_in, a.size, a.size)
for _ in range(a.reps):
    b.encode(d_in); b.decode()
torch.cuda.synchronize()
assert torch.equal(b.dec[:a.size], d_in)
print("ok", b.compressed_len())
