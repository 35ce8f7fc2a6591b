"""The reference's OWN benchmark harness (turborc.c, unmodified, built by oracle/Makefile with its -D_EXT hook picking up
host/xturborc.h + host/xturborc.c and linked against libtrc_b200.so) runs CPU and GPU ids on the same buffers.
bench() calls memcheck() after every decode (turborc.c:576), so a wrong GPU round trip shows up as an ERROR line; the
drop-in ids must report exactly the compressed size of the CPU id they replace, the batch ids the size the oracle port
gives for the same chunking."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "turborc_gpu")


def _rows(out):
    """-> {id: compressed size} from the harness's table ('  2883620  72.09%  ...  45:cdfsb ...')."""
    rows = {}
    for line in out.replace("\b", " ").splitlines():
        m = re.match(r"\s*(\d+)\s+[\d.]+%.*?\s(\d+):\S", line)
        if m:
            rows[int(m.group(2))] = int(m.group(1))
    return rows


def test_harness_links_library():
    """CPU: the harness binary was linked against the product library and still carries the reference's own codecs."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/turborc_gpu not built (needs /root/reference at build time)")
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libtrc_b200.so" in out and "not found" not in out, out
    und = subprocess.run(["nm", "-D", "--undefined-only", BIN], capture_output=True, text=True).stdout
    assert "trc_enc_batch_host" in und and "trc_dec_batch_host" in und
    defined = subprocess.run(["nm", "--defined-only", BIN], capture_output=True, text=True).stdout
    assert " T rccdfs2enc" in defined and " T anscdfenc" in defined      # CPU rows stay the reference's


def test_harness_cpu_ids_run(tmp_path, dg):
    """CPU: the reference ids still work in the extended binary (no GPU touched before a GPU id is asked for)."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/turborc_gpu not built")
    src = tmp_path / "z.bin"
    dg.zipf(200_000).tofile(src)
    r = subprocess.run([BIN, "-I1", "-J1", "-e45,56", str(src)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ERROR" not in r.stdout.upper(), r.stdout + r.stderr
    rows = _rows(r.stdout)
    assert set(rows) == {45, 56}, r.stdout


@pytest.mark.gpu
def test_harness_gpu_rows(tmp_path, dg, port):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/turborc_gpu not built")
    from helpers import cpu_batch
    n = 3_000_001
    cases = {"zipf": (dg.zipf(n), "45,90,91,96,56,92,97,46,93,47,95,48,98,49,99"), "o1": (dg.markov1(n), "64,94")}
    for name, (data, ids) in cases.items():
        src = tmp_path / f"{name}.bin"
        data.tofile(src)
        r = subprocess.run([BIN, "-I1", "-J1", "-e" + ids, str(src)], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "ERROR" not in (r.stdout + r.stderr).upper(), r.stdout + r.stderr    # memcheck(in, n, cpy) after every id
        rows = _rows(r.stdout)
        assert set(rows) == {int(x) for x in ids.split(",")}, r.stdout
        if name == "zipf":
            assert rows[96] == rows[45] and rows[97] == rows[56]                  # drop-in symbols: the reference's own sizes
            cdf = port.cdfini(data)
            num = int(data.max()) + 1
            for ident, codec, chunk in ((90, 5, 4096), (91, 4, 4096), (92, 2, 65536), (93, 6, 65536), (95, 7, 65536), (98, 11, 4096), (99, 12, 4096)):   # ids of helpers.CODECS
                _, off = cpu_batch(port, codec, data, chunk, cdf if codec in (4, 5) else None, num if codec in (4, 5) else 0)
                assert rows[ident] == int(off[-1]), (ident, rows[ident], int(off[-1]))
        else:
            _, off = cpu_batch(port, 3, data, 4194304, None, 0)
            assert rows[94] == int(off[-1]) == rows[64]                              # one 4 MiB chunk == the whole 3 MB call


@pytest.mark.gpu
def test_harness_gpu_vlc_rows(tmp_path, port):
    """GPU batch rows for the VLC-over-CDF integer ids (70-77; 80-87 are taken by the reference's transforms) next to the reference's 50-53 / 60-63: round trip checked by the
    harness's own memcheck, compressed sizes equal to the oracle's per-chunk calls."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/turborc_gpu not built")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vlc_data import sources
    from oracle import cpu
    pairs16 = {70: "rccdfuenc16", 72: "rccdfvenc16", 73: "rccdfvzenc16", 74: "anscdfuenc16", 75: "anscdfuzenc16", 76: "anscdfvenc16", 77: "anscdfvzenc16"}
    pairs32 = {70: "rccdfuenc32", 72: "rccdfvenc32", 73: "rccdfvzenc32", 76: "anscdfvenc32", 77: "anscdfvzenc32"}
    for width, flag, pairs, ref_ids in ((16, "-Os", pairs16, "50,52,53,60,61,62,63"), (32, "-Ou", pairs32, "50,52,53,62,63")):
        data = sources(width, 400_000)["walk"]
        src = tmp_path / f"w{width}.bin"
        data.tofile(src)
        ids = ",".join(str(k) for k in sorted(pairs))
        r = subprocess.run([BIN, "-I1", "-J1", flag, "-e" + ref_ids + "," + ids, str(src)], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "ERROR" not in (r.stdout + r.stderr).upper(), r.stdout + r.stderr
        rows = _rows(r.stdout)
        raw = data.view(np.uint8)
        for ident, enc in pairs.items():
            assert ident in rows, (ident, r.stdout)
            _, off = cpu.batch_enc(cpu.port(), enc, raw, 4096)
            assert rows[ident] == int(off[-1]), (width, ident, rows[ident], int(off[-1]))
