"""Shared helpers for the parity tests: run the CPU checker chunk by chunk like the batch API does."""
import numpy as np

# codec id -> (encoder, decoder, needs cdf, nibble alphabet)
CODECS = {
    0: ("anscdf4senc", "anscdf4sdec", True, True),
    1: ("anscdf4enc", "anscdf4dec", False, True),
    2: ("anscdfenc", "anscdfdec", False, False),
    3: ("anscdf1enc", "anscdf1dec", False, False),
    4: ("rccdfsenc", "rccdfsbdec", True, False),
    5: ("rccdfs2enc", "rccdfsb2dec", True, False),
    6: ("rccdfenc", "rccdfdec", False, False),
    7: ("rccdfienc", "rccdfidec", False, False),
    8: ("rccdf4enc", "rccdf4dec", False, True),
    9: ("rccdf4ienc", "rccdf4idec", False, True),
    11: ("rccdfenc8", "rccdfdec8", False, False),      # vnibble; id 10 is TRC_ANSW (own format, own test)
    12: ("rccdfienc8", "rccdfidec8", False, False),
}


def chunks(n, chunk_len):
    return [(s, min(chunk_len, n - s)) for s in range(0, n, chunk_len)]


def cpu_batch(lib, codec, data, chunk_len, cdf=None, cdfnum=None, chunks_per_cdf=0):
    """Reference semantics of the batch: encoder called per chunk, results packed back to back."""
    enc = CODECS[codec][0]
    parts, off = [], [0]
    for c, (s, l) in enumerate(chunks(data.size, chunk_len)):
        tab = None
        if cdf is not None:
            t = c // chunks_per_cdf if chunks_per_cdf else 0
            tab = cdf.reshape(-1)[t * 257:(t + 1) * 257]
        r, out = lib.enc(enc, data[s:s + l], tab, cdfnum)
        if r > out.size:      # rccdf4ienc on < 4 bytes returns 4: bytes past the raw copy are undefined, we define 0
            assert enc == "rccdf4ienc" and l < 4
            out = np.concatenate([out, np.zeros(r - out.size, np.uint8)])
        assert out.size == r, (enc, l, r, out.size)
        parts.append(out)
        off.append(off[-1] + r)
    return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), np.array(off, np.uint64)


def first_diff(a, b):
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if d.size else (n if a.size != b.size else -1)
