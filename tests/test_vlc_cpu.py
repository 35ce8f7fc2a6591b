"""CPU suite for the VLC-over-CDF integer codecs: the oracle restatement against the compiled
reference (when oracle/_ref was built) and against committed golden vectors the reference produced."""
import os

import numpy as np
import pytest

from oracle import cpu
from vlc_data import sources

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_vlc_v1.npz"))


@pytest.mark.parametrize("fam,w", cpu.VLC_CODECS)
def test_vlc_port_matches_reference(port, ref, fam, w):
    enc, dec = f"{fam}enc{w}", f"{fam}dec{w}"
    for cnt in (1, 2, 3, 7, 64, 1000, 20000, 150001):
        for sname, a in sources(w, cnt).items():
            x = a.view(np.uint8)
            lp, op = port.enc(enc, x)
            lr, orf = ref.enc(enc, x)
            assert lp == lr and np.array_equal(op, orf), (enc, sname, cnt, lp, lr)
            if lr < x.size:
                a1, b1 = port.dec(dec, orf, x.size), ref.dec(dec, orf, x.size)
                assert np.array_equal(a1, b1) and np.array_equal(b1, x), (dec, sname, cnt)


@pytest.mark.parametrize("fam,w", cpu.VLC_CODECS)
def test_vlc_port_matches_golden(port, fam, w):
    enc, dec = f"{fam}enc{w}", f"{fam}dec{w}"
    keys = [k for k in G.files if k.startswith(f"enc/{enc}/")]
    assert keys
    for k in keys:
        _, _, sname, cnt = k.split("/")
        x = G[f"in/{w}/{sname}/{cnt}"]
        l, s = port.enc(enc, x)
        assert l == int(G[f"len/{enc}/{sname}/{cnt}"][0]) and np.array_equal(s, G[k]), k
        if l < x.size:
            assert np.array_equal(port.dec(dec, s, x.size), x), k
