"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden vectors.

Everything here is bit-exact integer/byte work: compressed bytes, lengths, offsets and decoded bytes must be
identical to the reference semantics (oracle per chunk == one reference call per chunk)."""
import os

import numpy as np
import pytest

from helpers import CODECS, chunks, cpu_batch, first_diff

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))


def _tabs(port, codec, d):
    need_cdf = CODECS[codec][2]
    if not need_cdf:
        return None, 0
    return port.cdfini(d), int(d.max()) + 1


def _check_batch(trc, port, codec, d, chunk_len, label):
    enc, dec, need_cdf, nib = CODECS[codec]
    cdf, num = _tabs(port, codec, d)
    want, woff = cpu_batch(port, codec, d, chunk_len, cdf, num)
    got, goff = trc.enc_batch_host(codec, d, chunk_len, cdf=cdf, cdfnum=num)
    assert np.array_equal(goff, woff), (label, enc, chunk_len, "offsets", first_diff(goff, woff))
    assert np.array_equal(got, want), (label, enc, chunk_len, "bytes differ at", first_diff(got, want))
    # decode what we produced
    def _quirk(c, s, l):
        r = int(woff[c + 1] - woff[c])
        k = min(r, l)
        return r != l and np.array_equal(want[int(woff[c]):int(woff[c]) + k], d[s:s + k])
    quirk = codec == 9 and any(_quirk(c, s, l) for c, (s, l) in enumerate(chunks(d.size, chunk_len)))
    if quirk:                                   # rccdf4ienc raw-with-short-length: not decodable, by the reference either
        return
    back = trc.dec_batch_host(codec, got, goff, d.size, chunk_len, cdf=cdf, cdfnum=num)
    assert np.array_equal(back, d), (label, dec, chunk_len, "round trip differs at", first_diff(back, d))
    if codec in (0, 1):                          # reference-compatible tail handling == oracle's faithful decoders
        back = trc.dec_batch_host(codec, got, goff, d.size, chunk_len, cdf=cdf, cdfnum=num, flags=trc.F_REF_TAIL)
        for c, (s, l) in enumerate(chunks(d.size, chunk_len)):
            a, b = int(goff[c]), int(goff[c + 1])
            if b - a == l:
                assert np.array_equal(back[s:s + l], d[s:s + l])
            else:
                exp = port.dec(dec, got[a:], l, cdf, num)     # garbage tails over-read: give both the same bytes
                assert np.array_equal(back[s:s + l], exp), (label, dec, "ref tail", c)


@pytest.mark.parametrize("codec", sorted(CODECS))
def test_batch_parity_small(trc, port, sources, dg, codec):
    """Ragged and tiny geometries: tails of 1-3 bytes, odd lengths, chunks that hit the raw-copy rules."""
    nib = CODECS[codec][3]
    for sname, src in sources.items():
        for n, chunk_len in [(8, 8), (100, 33), (1001, 1001), (1003, 250), (4099, 1000), (20000, 4096), (65537, 65537), (70001, 16384)]:
            d = src[:n]
            if nib:
                d = dg.nibbles(d)
            _check_batch(trc, port, codec, d, chunk_len, f"{sname}/{n}")


@pytest.mark.parametrize("codec", sorted(CODECS))
def test_batch_parity_300k(trc, port, sources, dg, codec):
    nib = CODECS[codec][3]
    for sname in ("zipf", "bwt"):
        d = sources[sname]
        if nib:
            d = dg.nibbles(d)
        for chunk_len in (4096, 65536):
            _check_batch(trc, port, codec, d, chunk_len, sname)


@pytest.mark.parametrize("codec", sorted(CODECS))
def test_golden_vectors(trc, dg, codec):
    """Bytes produced by the compiled reference itself (tests/golden/make_golden.py)."""
    enc, dec, need_cdf, nib = CODECS[codec]
    for key in sorted(k for k in G.files if k.startswith("in/")):
        _, sname, n = key.split("/")
        d = G[key]
        x = dg.nibbles(d) if nib else d
        cdf = (G[f"cdfn/{sname}/{n}"] if nib else G[f"cdf/{sname}/{n}"]) if need_cdf else None
        num = int(x.max()) + 1 if need_cdf else 0
        got, off = trc.enc_batch_host(codec, x, x.size, cdf=cdf, cdfnum=num)
        assert int(off[1]) == int(G[f"len/{enc}/{sname}/{n}"][0]), (enc, key)
        assert np.array_equal(got, G[f"enc/{enc}/{sname}/{n}"]), (enc, key)
        dk = f"dec/{dec}/{sname}/{n}"
        if dk in G.files:
            back = trc.dec_batch_host(codec, got, off, x.size, x.size, cdf=cdf, cdfnum=num, flags=trc.F_REF_TAIL)
            assert np.array_equal(back, G[dk]), (dec, key)


def test_static_rans_byte_alphabet(trc, port, dg):
    """256-symbol static rANS: encoder bit-exact with the reference's anscdf4senc run on bytes, decoder inverts it."""
    for key in sorted(k for k in G.files if k.startswith("enc/anscdf4senc.bytes/")):
        _, _, sname, n = key.split("/")
        d, cdf = G[f"in/{sname}/{n}"], G[f"cdf/{sname}/{n}"]
        got, off = trc.enc_batch_host(trc.ANS4S, d, d.size, cdf=cdf, cdfnum=256)
        assert np.array_equal(got, G[key]), key
        assert np.array_equal(trc.dec_batch_host(trc.ANS4S, got, off, d.size, d.size, cdf=cdf, cdfnum=256), d)
    d = dg.zipf(1_000_003)
    cdf = port.cdfini(d)
    for chunk_len in (4096, 100_000):
        want, woff = cpu_batch(port, trc.ANS4S, d, chunk_len, cdf, 256)
        got, goff = trc.enc_batch_host(trc.ANS4S, d, chunk_len, cdf=cdf, cdfnum=256)
        assert np.array_equal(goff, woff) and np.array_equal(got, want)
        assert np.array_equal(trc.dec_batch_host(trc.ANS4S, got, goff, d.size, chunk_len, cdf=cdf, cdfnum=256), d)


def test_uniform_is_raw(trc, dg):
    """BASELINE config 1: 1 MiB uniform bytes through -e45 (rccdfs2enc) is a raw copy, l == n."""
    d = dg.uniform(1 << 20)
    cdf = trc.cdfini(d)
    l, out = trc.dropin_enc("rccdfs2enc", d, cdf, 256)
    assert l == d.size and np.array_equal(out, d)
    got, off = trc.enc_batch_host(trc.RCS2, d, 65536, cdf=cdf, cdfnum=256)
    assert int(off[-1]) == d.size and np.array_equal(got, d)
    assert np.array_equal(trc.dec_batch_host(trc.RCS2, got, off, d.size, 65536, cdf=cdf, cdfnum=256), d)


def test_cdfini(trc, port, sources, dg):
    for sname, src in sources.items():
        for n in (1, 7, 1000, 300000):
            d = src[:n]
            assert np.array_equal(trc.cdfini(d), port.cdfini(d)), (sname, n)
            dn = dg.nibbles(d)
            assert np.array_equal(trc.cdfini(dn), port.cdfini(dn)), (sname, n)


def test_per_block_tables(trc, port, dg):
    """One cdfini table per group of chunks (BASELINE config 5 shape: table per 64 MB block of small chunks)."""
    import torch
    d = dg.zipf(200_000)
    d[100_000:] = dg.bwt_shaped(100_000)
    chunk_len, cpc = 4096, 8
    nt = -(-trc.num_chunks(d.size, chunk_len) // cpc)
    t = torch.from_numpy(d).cuda()
    cdf_dev, status = trc.cdfini_dev(t, d.size, chunk_len * cpc)
    assert int(status.abs().sum().item()) == 0
    cdf = cdf_dev.cpu().numpy().view(np.uint16).reshape(nt, 257)
    for k in range(nt):
        assert np.array_equal(cdf[k], port.cdfini(d[k * chunk_len * cpc:(k + 1) * chunk_len * cpc]))
    for codec in (trc.ANS4S, trc.RCS, trc.RCS2):
        want, woff = cpu_batch(port, codec, d, chunk_len, cdf, 256, cpc)
        got, goff = trc.enc_batch_host(codec, d, chunk_len, cdf=cdf, cdfnum=256, chunks_per_cdf=cpc)
        assert np.array_equal(goff, woff) and np.array_equal(got, want), codec
        back = trc.dec_batch_host(codec, got, goff, d.size, chunk_len, cdf=cdf, cdfnum=256, chunks_per_cdf=cpc)
        assert np.array_equal(back, d), codec
    # groups that do not align with the CTA size exercise the per-thread global-table path
    for codec in (trc.ANS4S, trc.RCS2):
        cpc2 = 3
        nt2 = -(-trc.num_chunks(d.size, chunk_len) // cpc2)
        cdf2 = np.stack([port.cdfini(d[k * chunk_len * cpc2:(k + 1) * chunk_len * cpc2]) for k in range(nt2)])
        want, woff = cpu_batch(port, codec, d, chunk_len, cdf2, 256, cpc2)
        got, goff = trc.enc_batch_host(codec, d, chunk_len, cdf=cdf2, cdfnum=256, chunks_per_cdf=cpc2)
        assert np.array_equal(goff, woff) and np.array_equal(got, want), codec
        assert np.array_equal(trc.dec_batch_host(codec, got, goff, d.size, chunk_len, cdf=cdf2, cdfnum=256, chunks_per_cdf=cpc2), d)


def test_dropin_symbols(trc, port, dg):
    """The reference-named whole-buffer entry points (host pointers), as the turborc harness calls them."""
    d = dg.bwt_shaped(50_001)
    dn = dg.nibbles(d)
    cdf, cdfn = trc.cdfini(d), trc.cdfini(dn)
    assert np.array_equal(cdf, port.cdfini(d))
    for codec, (enc, dec, need_cdf, nib) in CODECS.items():
        x = dn if nib else d
        tab = (cdfn if nib else cdf) if need_cdf else None
        num = (int(x.max()) + 1) if need_cdf else None
        l, s = trc.dropin_enc(enc, x, tab, num)
        lw, sw = port.enc(enc, x, tab, num)
        assert l == lw and np.array_equal(s, sw), enc
        if l < x.size:
            back = trc.dropin_dec(dec, s, x.size, tab, num)
            assert np.array_equal(back, port.dec(dec, sw, x.size, tab, num)), dec
    # the per-ISA aliases the harness names (turborc.c:516-521) resolve to the same path
    l, s = trc.dropin_enc("anscdfencx", d)
    assert (l, s.tobytes()) == (lambda r: (r[0], r[1].tobytes()))(port.enc("anscdfenc", d))
    assert np.array_equal(trc.dropin_dec("anscdfdecs", s, d.size), d)


def test_multiblock_whole_call(trc, port, dg):
    """inlen > 4 MiB: blocks are coded independently and concatenated; order-1 carries cx across blocks;
    the decoder walks the blocks sequentially (no block directory in the format)."""
    d = dg.zipf((1 << 22) + 4097)
    for codec in (trc.ANS, trc.ANS1):
        enc, dec = CODECS[codec][0], CODECS[codec][1]
        want = port.enc(enc, d)
        got, off = trc.enc_batch_host(codec, d, d.size)
        assert int(off[1]) == want[0] and np.array_equal(got, want[1]), enc
        assert np.array_equal(trc.dec_batch_host(codec, got, off, d.size, d.size), d), dec


def test_device_api_matches_host_api(trc, port, dg):
    import torch
    d = dg.zipf(1_000_000)
    cdf = port.cdfini(d)
    for codec in (trc.ANS4S, trc.RCS2, trc.ANS, trc.RC):
        b = trc.DeviceBatch(codec, d.size, 4096, cdfnum=256 if codec in trc.STATIC else 0)
        if codec in trc.STATIC:
            b.set_cdf(cdf)
        t = torch.from_numpy(d).cuda()
        b.encode(t)
        torch.cuda.synchronize()
        n = b.compressed_len()
        got = b.out[:n].cpu().numpy()
        off = b.off.cpu().numpy().astype(np.uint64)
        want, woff = trc.enc_batch_host(codec, d, 4096, cdf=cdf if codec in trc.STATIC else None, cdfnum=256 if codec in trc.STATIC else 0)
        assert np.array_equal(off, woff) and np.array_equal(got, want)
        back = b.decode().cpu().numpy()
        assert np.array_equal(back, d)


def test_rare_redo_paths(port, dg):
    """TRC_FORCE_REDO=1 makes the kernels take the walk-back redo path (a pending word wrapping under a carry,
    p ~ 2^-32 per word in real data) for every call; results must not change.  Runs in a subprocess because the hook
    is read once at library load."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import importlib, sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        trc = importlib.import_module("turbo-range-coder_b200"); dg = importlib.import_module("turbo-range-coder_b200.datagen")
        from oracle import cpu
        from helpers import cpu_batch
        port = cpu.port()
        d = dg.bwt_shaped(150_001)
        for codec in (trc.RC, trc.RCI):
            for chunk in (65536, 20000):
                want, woff = cpu_batch(port, codec, d, chunk)
                got, off = trc.enc_batch_host(codec, d, chunk)
                assert np.array_equal(off, woff) and np.array_equal(got, want), (codec, chunk)
                assert np.array_equal(trc.dec_batch_host(codec, got, off, d.size, chunk), d)
        z = dg.zipf(300_007)                       # TRC_RCS2, lane-per-coder encoder: every call redone by the walk-back coder
        cdf = port.cdfini(z)
        for chunk in (1760, 4096, 304):
            want, woff = cpu_batch(port, trc.RCS2, z, chunk, cdf, 256)
            got, off = trc.enc_batch_host(trc.RCS2, z, chunk, cdf=cdf, cdfnum=256)
            assert np.array_equal(off, woff) and np.array_equal(got, want), ("rcs2", chunk)
            assert np.array_equal(trc.dec_batch_host(trc.RCS2, got, off, z.size, chunk, cdf=cdf, cdfnum=256), z)
        print("redo ok")
    ''') % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, TRC_FORCE_REDO="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "redo ok" in r.stdout, r.stdout + r.stderr


def test_randomized_geometries(trc, port, dg):
    """Seeded random (codec, source, length, chunk length) cases incl. degenerate ones (1-byte chunks, chunk > data,
    lengths around the 16/32-byte kernel block sizes)."""
    rng = np.random.default_rng(20261017)
    srcs = [dg.zipf(400_000, seed=21), dg.bwt_shaped(400_000, seed=22), dg.markov1(400_000, seed=23), dg.uniform(400_000, seed=24),
            np.zeros(400_000, np.uint8), (np.arange(400_000) % 251).astype(np.uint8)]
    for case in range(200):
        codec = int(rng.integers(0, 10))
        src = srcs[int(rng.integers(0, len(srcs)))]
        n = int(rng.choice([1, 2, 3, 5, 15, 16, 17, 31, 32, 33, 47, 63, 64, 65, 127, 129, 255, 1000, 4097, 20_001, int(rng.integers(1, 400_000))]))
        chunk = int(rng.choice([1, 2, 3, 4, 7, 16, 32, 48, 100, 256, 1024, 4096, 4100, 65536, n, n + 5, int(rng.integers(1, 70_000))]))
        if n // max(chunk, 1) > 60_000:          # keep the oracle loop short
            chunk = max(chunk, n // 50_000 + 1)
        off0 = int(rng.integers(0, 400_000 - n + 1))
        d = np.ascontiguousarray(src[off0:off0 + n])
        if CODECS[codec][3]:
            d = dg.nibbles(d)
        if CODECS[codec][2]:
            try:
                port.cdfini(d)
            except ValueError:                    # degenerate table: the reference die()s
                continue
        _check_batch(trc, port, codec, d, chunk, f"case{case}/n{n}/chunk{chunk}")


def test_wide_rans_answ(trc, port, dg):
    """TRC_ANSW, the 32-way warp-interleaved static rANS (a NEW format, parity unpinned): GPU bytes == the format
    specification in oracle/trc_oracle.c, exact round trip, and size within 4*32 bytes + 0.01 % of the static range
    coder's on the same table (SURVEY.md section 8c acceptance)."""
    for sname, n, chunk in [("zipf", 300_000, 4096), ("zipf", 300_000, 65536), ("bwt", 100_003, 16384), ("zipf", 1000, 1000),
                            ("uniform", 50_000, 4096), ("zipf", 129, 128), ("zipf", 5, 4), ("o1", 262_144 + 77, 262_144)]:
        d = {"zipf": dg.zipf, "bwt": dg.bwt_shaped, "uniform": dg.uniform, "o1": dg.markov1}[sname](n)
        cdf = port.cdfini(d)
        parts, offs = [], [0]
        for s in range(0, n, chunk):
            r, o = port.enc("answenc", d[s:s + chunk], cdf, 256)
            parts.append(o); offs.append(offs[-1] + r)
        want, woff = np.concatenate(parts), np.array(offs, np.uint64)
        got, goff = trc.enc_batch_host(trc.ANSW, d, chunk, cdf=cdf, cdfnum=256)
        assert np.array_equal(goff, woff), (sname, n, chunk, first_diff(goff, woff))
        assert np.array_equal(got, want), (sname, n, chunk, first_diff(got, want))
        back = trc.dec_batch_host(trc.ANSW, got, goff, n, chunk, cdf=cdf, cdfnum=256)
        assert np.array_equal(back, d), (sname, n, chunk, first_diff(back, d))
    d = dg.zipf(4_000_000)
    cdf = port.cdfini(d)
    got, goff = trc.enc_batch_host(trc.ANSW, d, 65536, cdf=cdf, cdfnum=256)
    rc_len = port.enc("rccdfsenc", d, cdf, 256)[0]
    n_chunks = trc.num_chunks(d.size, 65536)
    assert int(goff[-1]) <= rc_len * 1.0001 + n_chunks * 4 * 32, (int(goff[-1]), rc_len)
    assert np.array_equal(trc.dec_batch_host(trc.ANSW, got, goff, d.size, 65536, cdf=cdf, cdfnum=256), d)
    with pytest.raises(trc.TrcError):                      # our own format: calls must start 4-byte aligned
        trc.enc_batch_host(trc.ANSW, d[:10_000], 1001, cdf=cdf, cdfnum=256)


@pytest.mark.gpu
def test_fused_encoder_many_waves(trc, port, dg):
    """k_rcs2_enc_fused on a grid of several waves (262 144 calls -> 4096 CTAs): offsets from the decoupled look-back and the
    in-kernel layout must equal the reference calls packed back to back."""
    d = dg.zipf(32 << 20, seed=21)
    cdf = port.cdfini(d)
    chunk = 128
    got, off = trc.enc_batch_host(trc.RCS2, d, chunk, cdf=cdf, cdfnum=256)
    want, woff = cpu_batch(port, trc.RCS2, d[:1 << 20], chunk, cdf, 256)           # the oracle on the first 8192 calls ...
    k = (1 << 20) // chunk
    assert np.array_equal(off[:k + 1], woff) and np.array_equal(got[:int(woff[-1])], want)
    assert np.all(np.diff(off.astype(np.int64)) > 0) and int(off[-1]) == got.size    # ... and every call through its round trip
    back = trc.dec_batch_host(trc.RCS2, got, off, d.size, chunk, cdf=cdf, cdfnum=256)
    assert np.array_equal(back, d)


def test_table_handles(trc, port, dg):
    """Prebuilt coding tables (trc_tables_create_dev + *_tab calls) give the bytes of the per-call build, for one table and for
    a table per group of chunks."""
    import torch
    d = dg.zipf(1_000_000 + 48, seed=3)
    d[500_000:] = dg.bwt_shaped(d.size - 500_000)
    t = torch.from_numpy(d).cuda()
    for codec in (trc.RCS2, trc.RCS, trc.ANS4S, trc.ANSW):
        for chunk, cpc in ((4096, 0), (1760, 0), (512, 128)):
            if codec == trc.ANSW and chunk % 4:
                continue
            a = trc.DeviceBatch(codec, d.size, chunk, cdfnum=256, chunks_per_cdf=cpc)
            a.cdf, status = trc.cdfini_dev(t, d.size, chunk * cpc if cpc else d.size)
            assert int(status.abs().sum().item()) == 0
            a.encode(t); torch.cuda.synchronize()
            n = a.compressed_len()
            want, woff = a.out[:n].clone(), a.off.clone()
            a.prebuild_tables()
            a.out.zero_(); a.off.zero_()
            a.encode(t); torch.cuda.synchronize()
            assert a.compressed_len() == n and torch.equal(a.off, woff) and torch.equal(a.out[:n], want), (codec, chunk, cpc)
            assert torch.equal(a.decode(), t), (codec, chunk, cpc)
            if codec == trc.RCS2 and cpc == 0:           # the checker's bytes, too
                cdf = a.cdf.cpu().numpy().view(np.uint16)[:257]
                ow, ooff = cpu_batch(port, codec, d, chunk, cdf, 256)
                assert np.array_equal(want.cpu().numpy(), ow) and np.array_equal(woff.cpu().numpy().view(np.uint64), ooff)
            a.drop_tables()


def test_rcs2_ab_switches(port, dg):
    """The A/B switches of the TRC_RCS2 kernels stay bit-exact: per-lane loads instead of TMA tiles (TRC_ENC_TMA=0), the
    previous decoder generation (TRC_DEC3=0), the three-kernel encoder (TRC_FUSED=0).  Subprocesses: read at library load."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import importlib, sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        trc = importlib.import_module("turbo-range-coder_b200"); dg = importlib.import_module("turbo-range-coder_b200.datagen")
        from oracle import cpu
        from helpers import cpu_batch
        port = cpu.port()
        z = dg.zipf(2_000_003); u = dg.uniform(100_000)
        for d in (z, u, z[:17], z[:1760], z[:1761]):
            cdf = port.cdfini(d)
            for chunk in (1760, 4096, 65536, 48):
                want, woff = cpu_batch(port, trc.RCS2, d, chunk, cdf, 256)
                got, off = trc.enc_batch_host(trc.RCS2, d, chunk, cdf=cdf, cdfnum=256)
                assert np.array_equal(off, woff) and np.array_equal(got, want), (d.size, chunk)
                assert np.array_equal(trc.dec_batch_host(trc.RCS2, got, off, d.size, chunk, cdf=cdf, cdfnum=256), d)
        print("ab ok")
    ''') % (ROOT, os.path.join(ROOT, "tests"))
    for var in ({}, {"TRC_ENC_TMA": "0"}, {"TRC_DEC3": "0"}, {"TRC_FUSED": "0"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **var), timeout=900)
        assert r.returncode == 0 and "ab ok" in r.stdout, (var, r.stdout[-2000:] + r.stderr[-2000:])


def test_order1_many_calls(trc, port, sources, dg):
    """TRC_ANS1 with hundreds of calls takes the half-warp-per-call decoder whose low-nibble tables live in global memory
    (k_ans1_dec_g); it must decode the oracle's streams byte for byte -- ragged last call, an odd number of calls, raw calls."""
    for name, n, chunk in (("o1", 3_100_003, 4096), ("bwt", 1_600_000, 2048), ("uniform", 800_000, 1024), ("zipf", 6_100_000, 8192 + 1)):
        d = (dg.markov1(n) if name == "o1" else dg.bwt_shaped(n) if name == "bwt" else dg.uniform(n) if name == "uniform" else dg.zipf(n))
        assert trc.num_chunks(n, chunk) >= 5 * 148
        want, woff = cpu_batch(port, trc.ANS1, d, chunk)
        got, off = trc.enc_batch_host(trc.ANS1, d, chunk)
        assert np.array_equal(off, woff) and np.array_equal(got, want), (name, chunk)
        back = trc.dec_batch_host(trc.ANS1, want, woff, n, chunk)
        assert np.array_equal(back, d), (name, chunk, first_diff(back, d))


def test_rcs2_extreme_tables(trc, port, dg):
    """TRC_RCS2 (lane-per-coder kernels) with tables that are NOT the data's own: many frequency-1 symbols (15 bits per symbol: the
    most words a block can emit, streams that expand until the raw rule fires mid-chunk), two-symbol alphabets (frequency 32767: almost
    no output, long stretches without renormalisation), mismatched skew -- every call against the oracle, encode and decode."""
    rng = np.random.default_rng(77)
    n = 600_000 + 112
    cases = []
    # (a) uniform table over 256 symbols, skewed data and the reverse
    flat = (np.arange(257) * 128).astype(np.uint16)
    cases.append(("flat table / zipf data", flat, 256, dg.zipf(n, seed=31)))
    z = dg.zipf(n, seed=32)
    cases.append(("zipf table / uniform data", port.cdfini(z), 256, dg.uniform(n, seed=33)))
    # (b) 200 symbols of frequency 1, the rest shares 32768 - 200; data lives on the rare symbols half of the time
    f = np.ones(256, np.int64); f[200:] = 0; rest = 32768 - 200
    f[200:] = rest // 56; f[255] += rest - 56 * (rest // 56)
    rare = np.concatenate([[0], np.cumsum(f)]).astype(np.uint16)
    d = np.where(rng.random(n) < 0.5, rng.integers(0, 200, n), rng.integers(200, 256, n)).astype(np.uint8)
    cases.append(("rare symbols", rare, 256, d))
    cases.append(("rare symbols only", rare, 256, rng.integers(0, 200, n).astype(np.uint8)))
    # (c) two symbols, 32767 : 1
    two = np.zeros(257, np.uint16); two[1] = 32767; two[2] = 32768
    d2 = (rng.random(n) < 0.001).astype(np.uint8)
    cases.append(("two symbols", two, 2, d2))
    cases.append(("two symbols, all zero", two, 2, np.zeros(n, np.uint8)))
    for name, cdf, num, data in cases:
        for chunk in (1760, 4096, 48, 65536):
            want, woff = cpu_batch(port, trc.RCS2, data, chunk, cdf, num)
            got, off = trc.enc_batch_host(trc.RCS2, data, chunk, cdf=cdf, cdfnum=num)
            assert np.array_equal(off, woff), (name, chunk, first_diff(off, woff))
            assert np.array_equal(got, want), (name, chunk, first_diff(got, want))
            back = trc.dec_batch_host(trc.RCS2, got, off, data.size, chunk, cdf=cdf, cdfnum=num)
            assert np.array_equal(back, data), (name, chunk, first_diff(back, data))


def test_rcs2_device_shape_boundaries(trc, port, dg):
    """ONE device batch around the launch-shape boundaries of the TRC_RCS2 kernels (trc_b200.cu e3_shape / lpc_shape: 512 calls
    per SM = one wave, then waves of two CTAs per SM, balanced up to four waves), with tiny chunks so that the oracle stays fast:
    packed stream, offsets and round trip, with and without a ragged last call."""
    import torch
    from oracle import cpu
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    src = dg.zipf(12_000_000, seed=77)
    cdf = port.cdfini(src)
    lib = cpu.ref() or port
    cases = []
    for chunk in (16, 48, 128):
        for calls in (sm * 512 - 1, sm * 512, sm * 512 + 1, sm * 2 * 256 * 2 + 5, sm * 2 * 256 * 4 + 17, sm * 2 * 256 * 5 - 3):
            for tail in (0, 7):
                n = calls * chunk + tail
                if n <= src.size:
                    cases.append((n, chunk))
    assert len(cases) >= 12
    for n, chunk in cases:
        d = np.ascontiguousarray(src[:n])
        t = torch.from_numpy(d).cuda()
        b = trc.DeviceBatch(trc.RCS2, n, chunk, cdfnum=256)
        b.set_cdf(cdf)
        b.encode(t); torch.cuda.synchronize()
        want, woff = cpu.batch_enc(lib, CODECS[trc.RCS2][0], d, chunk, cdf, 256)
        m = b.compressed_len()
        assert m == want.size, (n, chunk, m, want.size)
        assert np.array_equal(b.off.cpu().numpy().view(np.uint64), woff), (n, chunk)
        assert np.array_equal(b.out[:m].cpu().numpy(), want), (n, chunk)
        assert torch.equal(b.decode()[:n], t), (n, chunk)


def test_rcs2_fused_per_group_tables(trc, port, dg):
    """The fused TRC_RCS2 encoder and the round-2 decoder with a table per group of calls (BASELINE config 5's shape), in ONE device
    batch big enough for the multi-wave launch shapes: groups of 256 calls (256-call CTAs), of 384 (128-call CTAs) and of 128."""
    import torch
    from oracle import cpu
    lib = cpu.ref() or port
    d = dg.zipf(16 << 20, seed=9)
    d[5 << 20:9 << 20] = dg.bwt_shaped(4 << 20, seed=10)
    t = torch.from_numpy(d).cuda()
    chunk = 128
    for cpc in (256, 384, 128):
        b = trc.DeviceBatch(trc.RCS2, d.size, chunk, cdfnum=256, chunks_per_cdf=cpc)
        b.cdf, status = trc.cdfini_dev(t, d.size, chunk * cpc)
        assert int(status.abs().sum().item()) == 0
        nt = -(-b.n // cpc)
        cdf = b.cdf.cpu().numpy().view(np.uint16).reshape(-1, 257)[:nt]
        for k in (0, nt // 2, nt - 1):
            assert np.array_equal(cdf[k], port.cdfini(d[k * chunk * cpc:(k + 1) * chunk * cpc]))
        want, woff = cpu.batch_enc(lib, CODECS[trc.RCS2][0], d, chunk, cdf, 256, cpc)
        for pre in (False, True):
            if pre:
                b.prebuild_tables()
            b.out.zero_(); b.off.zero_()
            b.encode(t); torch.cuda.synchronize()
            m = b.compressed_len()
            assert m == want.size, (cpc, pre, m, want.size)
            assert np.array_equal(b.off.cpu().numpy().view(np.uint64), woff), (cpc, pre)
            assert np.array_equal(b.out[:m].cpu().numpy(), want), (cpc, pre)
            assert torch.equal(b.decode()[:d.size], t), (cpc, pre)
        b.drop_tables()
