"""Invariants of the adaptive 16-symbol CDF (cdf16upd, cdf_.h:46-50) the GPU kernels rely on.

k_ans_code3 (adaptive_v3.cuh) divides by the adaptive frequency with a reciprocal table and leaves out the f == 1 special
case of the static encoder's table entries, and every kernel uses the update in the form m' = (127 m + T) >> 7.  Both rest
on: adjacent entries stay >= 10 apart (so every frequency is >= 9 once the implicit entry 16 = 32768 is counted), entry 0
stays 0 and entry 15 never exceeds 32759.  The argument is in the kernel's comment; this test replays it numerically on
random and adversarial symbol sequences, and checks the algebraic form against the reference's signed 16-bit arithmetic."""
import numpy as np

MIX, IC = 32736, 10


def upd_ref(m, x):
    """cdf16upd as the reference computes it: signed 16-bit lanes, arithmetic shift (cdf_.h:46-50)."""
    i = np.arange(16, dtype=np.int32)
    t = IC * i + np.where(m > m[x], MIX, 0)
    d = (t - m).astype(np.int16).astype(np.int32)          # the difference fits 16 bits: no wrap in the SIMD lanes
    assert np.array_equal(d, t - m)
    return m + (d >> 7)


def upd_gpu(m, x):
    i = np.arange(16, dtype=np.int32)
    return (127 * m + IC * i + np.where(i > x, MIX, 0)) >> 7


def _run(symbols):
    m = (np.arange(16, dtype=np.int32) << 11)
    fmin = 1 << 15
    for x in symbols:
        a, b = upd_ref(m, x), upd_gpu(m, x)
        assert np.array_equal(a, b)
        m = a
        gaps = np.diff(np.append(m, 32768))
        fmin = min(fmin, int(gaps.min()))
        assert m[0] == 0 and m[15] <= 32759 and gaps[:-1].min() >= 10
    return fmin


def test_update_forms_agree_and_frequencies_stay_large():
    rng = np.random.default_rng(5)
    seqs = [rng.integers(0, 16, 20000), np.zeros(20000, np.int64), np.full(20000, 15), np.tile([0, 15], 10000),
            np.repeat(np.arange(16), 2000), np.repeat(np.arange(15, -1, -1), 2000),
            rng.choice(16, 20000, p=np.r_[0.97, np.full(15, 0.002)])]
    assert min(_run(s) for s in seqs) >= 9
