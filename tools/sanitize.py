#!/usr/bin/env python
"""compute-sanitizer driver: every codec, aligned and ragged geometries, through the host batch API and the drop-ins."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
trc = importlib.import_module("turbo-range-coder_b200"); dg = importlib.import_module("turbo-range-coder_b200.datagen")
from oracle import cpu
from helpers import CODECS, cpu_batch
port = cpu.port()
bad = 0
for src in (dg.zipf(70_001, seed=3), dg.bwt_shaped(70_001, seed=4), dg.uniform(20_000, seed=5)):
    for codec, (enc, dec, need_cdf, nib) in CODECS.items():
        x = dg.nibbles(src) if nib else src
        cdf = port.cdfini(x) if need_cdf else None
        num = int(x.max()) + 1 if need_cdf else 0
        for chunk in (4096, 1000, 16384, x.size):
            got, off = trc.enc_batch_host(codec, x, chunk, cdf=cdf, cdfnum=num)
            want, woff = cpu_batch(port, codec, x, chunk, cdf, num)
            ok = np.array_equal(off, woff) and np.array_equal(got, want)
            back = trc.dec_batch_host(codec, got, off, x.size, chunk, cdf=cdf, cdfnum=num)
            bad += (not ok)
    c = trc.cdfini(src)
    l, s = trc.dropin_enc("anscdfenc", src); trc.dropin_dec("anscdfdec", s, src.size) if l < src.size else None
    l, s = trc.dropin_enc("rccdfs2enc", src, c, 256); trc.dropin_dec("rccdfsb2dec", s, src.size, c, 256) if l < src.size else None
print("sanitize driver done, mismatches:", bad)
