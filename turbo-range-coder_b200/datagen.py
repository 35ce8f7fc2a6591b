"""Deterministic synthetic byte streams for parity tests and bench.py.

No file or network access: every workload is regenerated from a seed.  ``sha16`` is the check the
survey quotes for the 100 MB files (first 16 hex digits of sha-256).
"""
import hashlib
import numpy as np

ZIPF_SEED, BWT_SEED, O1_SEED = 20261017, 20261018, 20261019


def sha16(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint8).tobytes()).hexdigest()[:16]


def uniform(n: int, seed: int = 1) -> np.ndarray:
    """BASELINE config 1: incompressible bytes (raw-copy path)."""
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def zipf(n: int, alpha: float = 1.1, seed: int = ZIPF_SEED) -> np.ndarray:
    """BASELINE config 2/5: Zipf(alpha) over 256 symbols through a fixed random permutation."""
    rng = np.random.default_rng(seed)
    k = np.arange(1, 257.0)
    p = k ** -alpha
    p /= p.sum()
    perm = rng.permutation(256).astype(np.uint8)
    x = rng.choice(256, size=n, p=p).astype(np.uint8)
    return perm[x]


def bwt_shaped(n: int, seed: int = BWT_SEED) -> np.ndarray:
    """BASELINE config 3: runs of a few symbols per geometric-length segment (post-BWT look)."""
    rng = np.random.default_rng(seed)
    out = np.empty(n, np.uint8)
    pos = 0
    while pos < n:
        seglen = int(rng.geometric(1 / 4000))
        ksz = int(rng.integers(2, 12))
        syms = rng.choice(256, ksz, replace=False).astype(np.uint8)
        pk = np.arange(1, ksz + 1.0) ** -2.0
        pk /= pk.sum()
        nsym = max(1, seglen // 3)
        s = syms[rng.choice(ksz, nsym, p=pk)]
        runs = rng.geometric(1 / 3, nsym)
        seg = np.repeat(s, runs)[:seglen]
        m = min(seg.size, n - pos)
        out[pos:pos + m] = seg[:m]
        pos += m
    return out


def markov1(n: int, alpha: float = 1.1, seed: int = O1_SEED) -> np.ndarray:
    """BASELINE config 4: order-1 source, rank ~ Zipf(alpha) emitted as perm[prev][rank]."""
    rng = np.random.default_rng(seed)
    k = np.arange(1, 257.0)
    p = k ** -alpha
    p /= p.sum()
    perms = np.stack([rng.permutation(256) for _ in range(256)]).astype(np.uint8)
    ranks = rng.choice(256, size=n, p=p).astype(np.intp)
    out = np.empty(n, np.uint8)
    prev = 0
    # the chain is serial by construction; vectorise over independent 4 KiB lanes instead
    lanes = max(1, min(n // 4096, 4096))
    per = -(-n // lanes)
    prevs = np.zeros(lanes, np.intp)
    idx0 = np.arange(lanes) * per
    for j in range(per):
        idx = idx0 + j
        ok = idx < n
        ii = idx[ok]
        v = perms[prevs[ok], ranks[ii]]
        out[ii] = v
        prevs[ok] = v
    del prev
    return out


def nibbles(a: np.ndarray) -> np.ndarray:
    """Low nibble of every byte: the input class of the 16-symbol codecs (turborc -e65 / xnibble)."""
    return (a & 15).astype(np.uint8)
