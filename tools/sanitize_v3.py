#!/usr/bin/env python
"""compute-sanitizer driver for the kernels added this round (adaptive_v3.cuh, vnibble.cuh, the container): small inputs,
two-call-per-warp and ragged geometries, multi-block calls, order 1."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trc = importlib.import_module("turbo-range-coder_b200"); dg = importlib.import_module("turbo-range-coder_b200.datagen")
bad = 0
for src in (dg.bwt_shaped(40_001, seed=4), dg.zipf(33_333, seed=3)):
    for codec in (trc.ANS, trc.ANS1, trc.RC8, trc.RCI8):
        for chunk in (4096, 1001, 16384, src.size):
            got, off = trc.enc_batch_host(codec, src, chunk)
            back = trc.dec_batch_host(codec, got, off, src.size, chunk)
            bad += not np.array_equal(back, src)
    for codec in (trc.RCS2, trc.ANS, trc.RCI8):
        blob = trc.compress(codec, src, 4096)
        bad += not np.array_equal(trc.decompress(blob), src)
big = dg.bwt_shaped((1 << 22) + 5000, seed=9)                 # two blocks of one call (anscdfenc re-initialises per 4 MiB)
for codec in (trc.ANS, trc.ANS1):
    got, off = trc.enc_batch_host(codec, big, big.size)
    bad += not np.array_equal(trc.dec_batch_host(codec, got, off, big.size, big.size), big)
print("sanitize_v3 driver done, mismatches:", bad)
