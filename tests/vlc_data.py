"""Integer-shaped sources for the VLC-over-CDF codecs: little-endian 16/32-bit values."""
import numpy as np


def sources(width, count, seed=3):
    """-> {name: uint16/uint32 array}: small values, a random walk (for the zigzag-delta codecs), geometric, log-normal, full range"""
    rng = np.random.default_rng(seed + width + count % 1000)
    dt = np.uint16 if width == 16 else np.uint32
    hi = 65535 if width == 16 else 2 ** 32 - 1
    return {
        "small": rng.integers(0, 6, count).astype(dt),
        "walk": (np.cumsum(rng.integers(-40, 41, count)) + 30000).astype(dt),
        "geom": np.minimum(rng.geometric(0.02, count), hi).astype(dt),
        "lognorm": np.minimum(np.exp(rng.normal(6, 3, count)), hi).astype(dt),
        "wide": rng.integers(0, hi, count, dtype=np.uint64).astype(dt),
    }


# codec id (include/trc_b200.h) -> (family, width); order == enum trc_codec 13..24
VLC_IDS = {13: ("anscdfu", 16), 14: ("anscdfuz", 16), 15: ("anscdfv", 16), 16: ("anscdfvz", 16), 17: ("anscdfv", 32), 18: ("anscdfvz", 32),
           19: ("rccdfv", 16), 20: ("rccdfvz", 16), 21: ("rccdfv", 32), 22: ("rccdfvz", 32), 23: ("rccdfu", 16), 24: ("rccdfu", 32)}
