"""CPU suite: the oracle restatement against the committed golden vectors (bytes the compiled reference
produced, tests/golden/make_golden.py).  Runs where /root/reference does not exist."""
import os

import numpy as np
import pytest

from helpers import CODECS

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
KEYS = sorted(k for k in G.files if k.startswith("in/"))


@pytest.mark.parametrize("key", KEYS)
def test_port_matches_golden(port, key, dg):
    _, sname, n = key.split("/")
    d = G[key]
    dn = dg.nibbles(d)
    cdf, cdfn = G[f"cdf/{sname}/{n}"], G[f"cdfn/{sname}/{n}"]
    assert np.array_equal(port.cdfini(d), cdf) and np.array_equal(port.cdfini(dn), cdfn)
    for codec, (enc, dec, need_cdf, nib) in CODECS.items():
        x = dn if nib else d
        tab = (cdfn if nib else cdf) if need_cdf else None
        num = int(x.max()) + 1 if need_cdf else None
        l, s = port.enc(enc, x, tab, num)
        assert l == int(G[f"len/{enc}/{sname}/{n}"][0]), (enc, key)
        assert np.array_equal(s, G[f"enc/{enc}/{sname}/{n}"]), (enc, key)
        dk = f"dec/{dec}/{sname}/{n}"
        if dk in G.files:
            assert np.array_equal(port.dec(dec, s, d.size, tab, num), G[dk]), (dec, key)
    bk = f"enc/anscdf4senc.bytes/{sname}/{n}"
    if bk in G.files:
        l, s = port.enc("anscdf4senc", d, cdf)
        assert np.array_equal(s, G[bk])
        assert np.array_equal(port.dec("ans_sdec_n", s, d.size, cdf, 256), d)
