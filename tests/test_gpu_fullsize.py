"""GPU suite at BASELINE.json's full sizes (100 MB): exact parity against the oracle per chunk where the CPU
finishes in seconds, plus size-independent properties (round trip, length bookkeeping)."""
import hashlib

import numpy as np
import pytest

from helpers import cpu_batch

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
