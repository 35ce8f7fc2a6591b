"""GPU parity for the VLC-over-CDF integer codecs (vlc.cuh): every chunk must be byte-for-byte the
reference call on those integers (oracle port, itself pinned against the compiled reference in test_vlc_cpu.py), golden
vectors produced by the compiled reference must be reproduced, and the drop-in symbols must behave like the reference's."""
import os

import numpy as np
import pytest

from vlc_data import VLC_IDS, sources

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_vlc_v1.npz"))


def _cpu_batch(port, enc, x, chunk):
    parts, off = [], [0]
    for s in range(0, x.size, chunk):
        l, o = port.enc(enc, x[s:s + chunk])
        parts.append(o); off.append(off[-1] + l)
    return np.concatenate(parts), np.array(off, np.uint64)


@pytest.mark.gpu
@pytest.mark.parametrize("codec", sorted(VLC_IDS))
def test_vlc_batch_parity(trc, port, codec):
    fam, w = VLC_IDS[codec]
    enc, esz = f"{fam}enc{w}", w // 8
    for cnt, chunk_el in [(1, 1), (3, 3), (777, 100), (5000, 5000), (40_000, 2048), (150_000, 32768)]:
        for sname, a in sources(w, cnt).items():
            x = a.view(np.uint8)
            chunk = chunk_el * esz
            want, woff = _cpu_batch(port, enc, x, chunk)
            got, goff = trc.enc_batch_host(codec, x, chunk)
            assert np.array_equal(goff, woff), (enc, sname, cnt, chunk_el, "offsets")
            assert np.array_equal(got, want), (enc, sname, cnt, chunk_el, "bytes")
            back = trc.dec_batch_host(codec, got, goff, x.size, chunk)
            assert np.array_equal(back, x), (enc, sname, cnt, chunk_el, "round trip")


@pytest.mark.gpu
@pytest.mark.parametrize("codec", sorted(VLC_IDS))
def test_vlc_golden_and_dropin(trc, codec):
    fam, w = VLC_IDS[codec]
    enc, dec = f"{fam}enc{w}", f"{fam}dec{w}"
    for k in sorted(k for k in G.files if k.startswith(f"enc/{enc}/")):
        _, _, sname, cnt = k.split("/")
        x = G[f"in/{w}/{sname}/{cnt}"]
        want_len = int(G[f"len/{enc}/{sname}/{cnt}"][0])
        got, off = trc.enc_batch_host(codec, x, x.size)
        assert int(off[1]) == want_len and np.array_equal(got, G[k]), k
        l, s = trc.dropin_enc(enc, x)                                   # the reference's own symbol name
        assert l == want_len and np.array_equal(s, G[k]), ("drop-in", k)
        if l < x.size:
            assert np.array_equal(trc.dropin_dec(dec, s, x.size), x), ("drop-in", k)


def test_vlc_rejects_partial_elements(trc):
    """CPU: lengths that are not whole elements are an argument error (the reference would read past the buffer)."""
    assert trc.lib.trc_enc_scratch_bytes(trc.ANSV32, 1002, 1002) == 0
    assert trc.lib.trc_enc_scratch_bytes(trc.RCU16, 1001, 1001) == 0
    assert trc.lib.trc_enc_scratch_bytes(trc.RCU16, 1000, 333) == 0
    assert trc.lib.trc_enc_scratch_bytes(trc.RCU16, 1000, 100) > 0
