"""GPU suite at BASELINE.json's full sizes (100 MB): exact parity against the oracle per chunk where the CPU
finishes in seconds, plus size-independent properties (round trip, length bookkeeping)."""
import hashlib

import numpy as np
import pytest

from helpers import cpu_batch, CODECS

pytestmark = pytest.mark.gpu
N = 100_000_000


@pytest.fixture(scope="module")
def zipf100(dg):
    z = dg.zipf(N)
    assert dg.sha16(z) == "e9e9669ae62e03c9"          # SURVEY.md section 8c input hash
    return z


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_config2_static_100mb(trc, port, zipf100):
    """BASELINE config 2: 100 MB Zipf(1.1), static CDF from cdfini on the whole buffer, batch of 4 KiB chunks."""
    cdf = trc.cdfini(zipf100)
    assert _sha(cdf.view(np.uint8)) == "31279976b59a2da3"
    for codec in (trc.RCS2, trc.ANS4S):
        got, off = trc.enc_batch_host(codec, zipf100, 4096, cdf=cdf, cdfnum=256)
        want, woff = cpu_batch(port, codec, zipf100, 4096, cdf, 256)
        assert np.array_equal(off, woff)
        assert _sha(got) == _sha(want)
        back = trc.dec_batch_host(codec, got, off, N, 4096, cdf=cdf, cdfnum=256)
        assert _sha(back) == "e9e9669ae62e03c9"


def _oracle_batch(port, codec, data, chunk, cdf=None, cdfnum=0):
    """Checker's packed stream for the batch, multi-threaded (oracle/cpu_bench.c): compiled reference when built, else the port."""
    from oracle import cpu
    return cpu.batch_enc(cpu.ref() or port, CODECS[codec][0], data, chunk, cdf, cdfnum)


@pytest.mark.parametrize("chunk", ["bench", 65536, 1 << 20])
def test_headline_chunks_100mb(trc, port, zipf100, chunk):
    """The configuration bench.py times: TRC_RCS2 over 100 MB Zipf(1.1) at bench.py's default chunk (1760 B on a B200),
    and at 64 KiB / 1 MiB: packed stream, offsets and round trip against the oracle, byte for byte."""
    if chunk == "bench":
        import argparse, os, sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        chunk = bench.default_chunk(argparse.Namespace(chunk=0, size=N, codec="rcs2"))
        assert chunk == 1760
    cdf = trc.cdfini(zipf100)
    got, off = trc.enc_batch_host(trc.RCS2, zipf100, chunk, cdf=cdf, cdfnum=256)
    want, woff = _oracle_batch(port, trc.RCS2, zipf100, chunk, cdf, 256)
    assert np.array_equal(off, woff)
    assert got.size == want.size and np.array_equal(got, want)
    if chunk == 1760:
        assert got.size == 72_551_464                   # the size both bench arms print
    back = trc.dec_batch_host(trc.RCS2, got, off, N, chunk, cdf=cdf, cdfnum=256)
    assert _sha(back) == "e9e9669ae62e03c9"


@pytest.mark.parametrize("chunk", [1328, 880, 512, 256])
def test_rcs2_launch_shapes_100mb(trc, port, zipf100, chunk):
    """Device-resident 100 MB in ONE batch, so that the kernels see every launch shape of trc_b200.cu's e3_shape / lpc_shape:
    one 1024-thread CTA per SM (1328-byte chunks: 509 calls per SM), two / three / six balanced waves of two CTAs per SM
    (880, 512, 256).  Packed stream, offsets and round trip against the oracle."""
    import torch
    t = torch.from_numpy(zipf100).cuda()
    b = trc.DeviceBatch(trc.RCS2, N, chunk, cdfnum=256)
    b.set_cdf(trc.cdfini(zipf100))
    b.prebuild_tables()
    b.encode(t); torch.cuda.synchronize()
    n = b.compressed_len()
    want, woff = _oracle_batch(port, trc.RCS2, zipf100, chunk, trc.cdfini(zipf100), 256)
    assert n == want.size
    assert np.array_equal(b.off.cpu().numpy().view(np.uint64), woff)
    assert np.array_equal(b.out[:n].cpu().numpy(), want)
    assert torch.equal(b.decode()[:N], t)
    b.drop_tables()


def test_config3_adaptive_100mb(trc, port, dg):
    """BASELINE config 3: 100 MB BWT-shaped stream through the adaptive byte rANS (-e56) and RC (-e46), 64 KiB chunks."""
    b = dg.bwt_shaped(N)
    assert dg.sha16(b) == "eb7a495148b2b3e2"
    for codec in (trc.ANS, trc.RC):
        got, off = trc.enc_batch_host(codec, b, 65536)
        want, woff = cpu_batch(port, codec, b, 65536)
        assert np.array_equal(off, woff)
        assert _sha(got) == _sha(want)
        back = trc.dec_batch_host(codec, got, off, N, 65536)
        assert _sha(back) == "eb7a495148b2b3e2"
