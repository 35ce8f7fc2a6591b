"""CPU suite: the C-ABI library loads and exports every symbol include/trc_b200.h declares; host-only logic."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "trc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_symbols_exported(trc):
    syms = declared_symbols()
    assert len(syms) >= 50, syms
    missing = [s for s in syms if not hasattr(trc.lib, s)]
    assert not missing, missing
    for s in trc.DROPIN_SYMBOLS:
        assert s in syms


def test_host_selftest(trc):
    """Division-by-reciprocal table exact for all 2^15 frequencies; chunk/unit geometry."""
    assert trc.lib.trc_selftest_host() == 0


def test_geometry_and_arg_errors(trc):
    assert trc.num_chunks(100_000_000, 4096) == 24415
    assert trc.num_chunks(10, 4096) == 1
    assert trc.lib.trc_enc_scratch_bytes(trc.RCS2, 100_000_000, 4096) > 100_000_000
    assert trc.lib.trc_enc_scratch_bytes(99, 1000, 100) == 0          # bad codec
    assert trc.lib.trc_enc_scratch_bytes(trc.ANS, 0, 100) == 0          # empty input
    # argument errors are reported before anything touches CUDA
    rc = trc.lib.trc_enc_batch_dev(trc.RCS, None, 1000, 100, None, 256, 0, None, None, None, 0, None)
    assert rc == trc.E_ARG
    rc = trc.lib.trc_dec_batch_dev(99, None, None, None, 1000, 100, None, 0, 0, 0, None)
    assert rc == trc.E_ARG


def test_no_cpu_fallback_in_product(trc):
    """The product path must not depend on the oracle: no include/import/symbol use in the package sources (comments may
    point at the specification), nothing in the shared library."""
    pkg = os.path.dirname(trc.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith((".cu", ".cuh", ".h")):
                code = open(path, errors="ignore").read()
                code = re.sub(r"/\*.*?\*/", "", code, flags=re.S)
                code = re.sub(r"//[^\n]*", "", code)
                assert "oracle" not in code and "orc_" not in code, path
            elif f.endswith(".py"):
                code = open(path, errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", code, flags=re.M), path
                assert "libtrc_oracle" not in code and "libtrcref" not in code, path
            elif f == "Makefile":
                assert "oracle" not in open(path).read(), path
    blob = open(trc.LIB_PATH, "rb").read()
    assert b"orc_" not in blob and b"libtrc_oracle" not in blob and b"libtrcref" not in blob


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_gpu(trc):
    import numpy as np
    with pytest.raises(trc.TrcError):
        trc.enc_batch_host(trc.RC, np.zeros(1000, np.uint8), 100)


def test_rcs2_launch_shapes(trc):
    """Host logic of the TRC_RCS2 launch shapes (trc_b200.cu: e3_shape / lpc_shape) for a 148-SM part: every call is covered, CTAs
    stay within what the kernels were built for, one-wave batches get one encoder CTA (two decoder CTAs) per SM, bigger batches
    whole warps of calls per CTA and -- up to four waves -- equally sized CTAs that fill every wave."""
    fn = trc.lib.trc_debug_rcs2_shapes
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_uint)]
    sm = 148

    def shapes(n, cpc=0):
        a = (ctypes.c_uint * 4)()
        assert fn(n, cpc, sm, a) == 0
        return tuple(a)

    assert fn(0, 0, sm, (ctypes.c_uint * 4)()) != 0 and fn(10, 0, 0, (ctypes.c_uint * 4)()) != 0
    for n in [1, 15, 16, 17, 147, 148, 149, 4766, 24415, 56819, 56832, 75302, 75776, 75777, 113637, 151552, 195313, 390625, 610081, 6_100_000]:
        ec, en, dc, dn = shapes(n)
        assert ec * en >= n and ec * (en - 1) < n, (n, ec, en)
        assert dc * dn >= n and dc * (dn - 1) < n, (n, dc, dn)
        assert 16 <= ec <= 512 and 1 <= dc <= 256, (n, ec, dc)
        per_sm = -(-n // sm)
        if per_sm <= 512:                                  # one wave: one encoder CTA per SM
            assert en <= sm and ec == max(16, per_sm), (n, ec, en)
        else:                                              # waves of two CTAs per SM, whole warps
            assert ec % 16 == 0 and ec <= 256, (n, ec)
            waves = -(-en // (2 * sm))
            if waves <= 4:
                assert en > (waves - 1) * 2 * sm and (2 * sm * waves - en) * ec < 2 * sm * waves * 16 + ec, (n, ec, en)   # every wave full up to the rounding to whole warps
        if 64 < per_sm <= 768:                             # decoder: one wave of k <= 3 equal CTAs per SM
            assert dn <= 3 * sm, (n, dc, dn)
    assert shapes(56819) == (384, 148, 192, 296)            # the headline batch: 100 MB at 1760-byte chunks
    assert shapes(610081)[0] == 256 and shapes(610081)[2] <= 256
    # per-block tables (chunks_per_cdf != 0): decoder CTAs never straddle a table group
    for n, cpc in [(262144, 16384), (24415, 128), (100000, 256)]:
        _, _, dc, dn = shapes(n, cpc)
        assert cpc % dc == 0 and dc * dn >= n, (n, cpc, dc)
