"""Regenerates tests/golden/golden_v1.npz from the COMPILED REFERENCE (oracle/_ref/libtrcref.so).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
The fixtures pin the oracle (and the GPU path) to bytes the reference itself produced, and travel to
the GPU box where /root/reference does not exist.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cpu                      # noqa: E402
from helpers import CODECS                  # noqa: E402

dg = importlib.import_module("turbo-range-coder_b200.datagen")
R = cpu.ref()
assert R is not None, "build oracle/_ref first (make -C oracle ref)"

out = {}
srcs = {"zipf": dg.zipf(20000, seed=11), "bwt": dg.bwt_shaped(20000, seed=12), "o1": dg.markov1(20000, seed=13)}
for sname, src in srcs.items():
    for n in (8, 100, 1001, 4096, 20000):
        d = src[:n]
        dn = dg.nibbles(d)
        out[f"in/{sname}/{n}"] = d
        cdf, cdfn = R.cdfini(d), R.cdfini(dn)
        out[f"cdf/{sname}/{n}"] = cdf
        out[f"cdfn/{sname}/{n}"] = cdfn
        for codec, (enc, dec, need_cdf, nib) in CODECS.items():
            x = dn if nib else d
            tab = (cdfn if nib else cdf) if need_cdf else None
            num = int(x.max()) + 1 if need_cdf else None
            l, s = R.enc(enc, x, tab, num)
            out[f"enc/{enc}/{sname}/{n}"] = s
            out[f"len/{enc}/{sname}/{n}"] = np.array([l], np.int64)
            if l < n and not np.array_equal(s, x[:l]):
                out[f"dec/{dec}/{sname}/{n}"] = R.dec(dec, s, n, tab, num)
        l, s = R.enc("anscdf4senc", d, cdf)          # byte alphabet through the static rANS encoder
        if l < n:
            out[f"enc/anscdf4senc.bytes/{sname}/{n}"] = s
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
print("wrote", len(out), "arrays")
