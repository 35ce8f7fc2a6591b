"""Self-describing container (include/trc_b200.h): header checks on the CPU; on the GPU the payload
must be the batch layer's packed stream (== the oracle's per-chunk reference calls), the tables the oracle's cdfini per
block, and decompress must need nothing but the container."""
import struct

import numpy as np
import pytest

from helpers import CODECS, cpu_batch


def _header(codec=5, total=1000, chunk=100, cdf_block=0, n=10, ntab=1, cdfnum=256, payload=0, magic=0x42435254, version=1):
    return struct.pack("<IHBBQQQQIIQQ", magic, version, codec, 0, total, chunk, cdf_block, n, ntab, cdfnum, payload, 0)


def test_container_header_checks(trc):
    """CPU: trc_container_info validates magic, version, geometry and that directory + payload fit the buffer."""
    body = np.zeros(64 + 520 + 40 + 16, np.uint8)
    good = np.frombuffer(_header(), np.uint8)
    blob = np.concatenate([good, body[64:]])
    info = trc.container_info(blob)
    assert info == {"codec": 5, "total_len": 1000, "chunk_len": 100, "n_chunks": 10}
    for bad in (_header(magic=0x12345678), _header(version=2), _header(codec=99), _header(n=11), _header(ntab=2),
                _header(cdfnum=16), _header(payload=1 << 40), _header(total=0), _header(chunk=0), _header(cdf_block=150)):
        with pytest.raises(trc.TrcError):
            trc.container_info(np.concatenate([np.frombuffer(bad, np.uint8), body[64:]]))
    with pytest.raises(trc.TrcError):
        trc.container_info(blob[:40])                                   # shorter than a header
    with pytest.raises(trc.TrcError):
        trc.container_info(blob[:100])                                  # tables / directory cut off
    assert trc.lib.trc_container_bound(5, 1000, 100, 0) >= 64 + 514 + 40 + 1000
    assert trc.lib.trc_container_bound(5, 1000, 100, 150) == 0          # a table must cover whole chunks
    assert trc.lib.trc_container_bound(99, 1000, 100, 0) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("codec", sorted(CODECS))
def test_container_roundtrip_and_payload(trc, port, dg, codec):
    enc, dec, need_cdf, nib = CODECS[codec]
    for n, chunk, cdf_block in [(70_001, 4096, 0), (300_000, 16384, 65536), (5_000, 5_000, 0)]:
        src = dg.bwt_shaped(n) if codec in (2, 3, 6, 7) else dg.zipf(n)
        d = dg.nibbles(src) if nib else src
        if not need_cdf:
            cdf_block = 0
        blob = trc.compress(codec, d, chunk, cdf_block)
        info = trc.container_info(blob)
        assert info == {"codec": codec, "total_len": n, "chunk_len": chunk, "n_chunks": -(-n // chunk)}
        back = trc.decompress(blob)
        assert np.array_equal(back, d), (enc, n, chunk)
        # the payload is the reference's bytes, chunk by chunk; the tables are cdfini of each block
        nt = (-(-n // cdf_block) if cdf_block else 1) if need_cdf else 0
        cdfnum = 16 if codec == 0 else 256
        tabs = None
        if need_cdf:
            blk = cdf_block or n
            tabs = np.stack([port.cdfini(d[s:s + blk], cdfnum) for s in range(0, n, blk)])
            got_tabs = np.frombuffer(blob[64:64 + nt * 514].tobytes(), np.uint16).reshape(nt, 257)
            assert np.array_equal(got_tabs[:, :cdfnum + 1], tabs[:, :cdfnum + 1])
        want, woff = cpu_batch(port, codec, d, chunk, tabs, cdfnum if need_cdf else 0, (cdf_block // chunk) if cdf_block else 0)
        dir_off = 64 + ((nt * 514 + 7) & ~7)
        clen = np.frombuffer(blob[dir_off:dir_off + 4 * info["n_chunks"]].tobytes(), np.uint32)
        assert np.array_equal(clen, np.diff(woff).astype(np.uint32))
        pay_off = (dir_off + 4 * info["n_chunks"] + 15) & ~15
        assert np.array_equal(blob[pay_off:], want)


@pytest.mark.gpu
def test_container_rejects_corrupt_directory(trc, dg):
    d = dg.zipf(50_000)
    blob = trc.compress(trc.RCS2, d, 4096).copy()
    dir_off = 64 + ((514 + 7) & ~7)
    blob[dir_off:dir_off + 4] = np.frombuffer(struct.pack("<I", 1 << 30), np.uint8)   # a chunk longer than its input
    with pytest.raises(trc.TrcError):
        trc.decompress(blob)


@pytest.mark.gpu
def test_container_ans4s_rejects_bytes_outside_the_alphabet(trc, dg):
    """TRC_ANS4S codes a 16-symbol alphabet (anscdf.c:57; the harness guards it at the caller, turborc.c:535 `if(m<16)`): a byte
    >= 16 must be reported, not silently coded with an empty table entry."""
    d = dg.nibbles(dg.zipf(50_000, seed=4))
    blob = trc.compress(trc.ANS4S, d, 4096)
    assert np.array_equal(trc.decompress(blob), d)
    # the table rows travel in the container: entries past the alphabet are defined (zero), so the bytes are reproducible
    assert np.array_equal(blob, trc.compress(trc.ANS4S, d, 4096))
    bad = d.copy(); bad[1234] = 200
    with pytest.raises(trc.TrcError):
        trc.compress(trc.ANS4S, bad, 4096)


@pytest.mark.gpu
def test_container_corrupt_payload_does_not_crash(trc, dg):
    """A damaged payload decodes to garbage or is rejected, but never reads or writes out of bounds (ring decoders clamp)."""
    d = dg.zipf(300_000, seed=8)
    for codec in (trc.RCS2, trc.ANS4S, trc.RCS, trc.ANS, trc.RC):
        x = dg.nibbles(d) if codec == trc.ANS4S else d
        blob = trc.compress(codec, x, 1760 if codec == trc.RCS2 else 4096).copy()
        rng = np.random.default_rng(codec)
        hdr = trc.CONTAINER_HEADER
        for _ in range(3):
            b = blob.copy()
            pos = rng.integers(blob.size // 2, blob.size, 64)
            b[pos] ^= rng.integers(1, 256, 64).astype(np.uint8)
            try:
                out = trc.decompress(b)
                assert out.size == x.size
            except trc.TrcError:
                pass
        t = blob[: hdr + (blob.size - hdr) // 2]                   # truncated
        with pytest.raises(trc.TrcError):
            trc.decompress(t)
