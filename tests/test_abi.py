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
