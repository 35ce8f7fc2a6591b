"""Regenerates tests/golden/golden_vlc_v1.npz from the COMPILED REFERENCE (oracle/_ref/libtrcref.so): the VLC-over-CDF
integer codecs of SURVEY.md section 8f.2.  Run in the build container:  python tests/golden/make_golden_vlc.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cpu                      # noqa: E402
from vlc_data import sources                # noqa: E402

R = cpu.ref()
assert R is not None, "build oracle/_ref first (make -C oracle ref)"
out = {}
for fam, w in cpu.VLC_CODECS:
    for cnt in (1, 3, 64, 1000, 6000):
        for sname, a in sources(w, cnt).items():
            x = a.view(np.uint8)
            out.setdefault(f"in/{w}/{sname}/{cnt}", x)
            l, s = R.enc(f"{fam}enc{w}", x)
            out[f"enc/{fam}enc{w}/{sname}/{cnt}"] = s
            out[f"len/{fam}enc{w}/{sname}/{cnt}"] = np.array([l], np.int64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_vlc_v1.npz"), **out)
print("wrote", len(out), "arrays")
