#!/usr/bin/env python
"""Driver for compute-sanitizer on the round-2 headline kernels (k_rcs2_enc3, k_rcs2_dec3, tables, container, multi-wave grids):
    compute-sanitizer --tool {memcheck|racecheck|synccheck} python tools/sanitize_r2.py
Small shapes (the tools slow kernels down 10-100x) that still cover: several waves of CTAs through the decoupled look-back
(chunk 48 x 9 MB = 196 608 calls -> 512 CTAs), the short last call, the raw path, TMA and per-lane input, prebuilt tables."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trc = importlib.import_module("turbo-range-coder_b200")
dg = importlib.import_module("turbo-range-coder_b200.datagen")
for n, chunk, src in ((9 << 20, 48, "zipf"), (600_000 + 77, 1760, "zipf"), (300_000, 4096, "uniform"), (500_000, 65536, "zipf")):
    d = dg.zipf(n, seed=5) if src == "zipf" else dg.uniform(n)
    t = torch.from_numpy(d).cuda()
    b = trc.DeviceBatch(trc.RCS2, n, chunk, cdfnum=256)
    b.cdf, _ = trc.cdfini_dev(t, n, n)
    b.encode(t); torch.cuda.synchronize()
    assert torch.equal(b.decode(), t)
    b.prebuild_tables()
    b.encode(t); torch.cuda.synchronize()
    assert torch.equal(b.decode(), t)
    print("ok", n, chunk, src, b.compressed_len())
blob = trc.compress(trc.RCS2, dg.zipf(200_000), 1760)
assert np.array_equal(trc.decompress(blob), dg.zipf(200_000))
print("sanitize_r2 done")
