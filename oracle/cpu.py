"""ctypes front-end to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * ``port()``  -> oracle/libtrc_oracle.so, our scalar C restatement (symbols ``orc_<name>``)
  * ``ref()``   -> oracle/_ref/libtrcref.so, the unmodified reference compiled by oracle/Makefile
                   (symbols carry the reference's own names); None when it was never built.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  Nothing under turbo-range-coder_b200/ does.

Both objects expose the same two calls so tests can run either through the same code:
    enc(name, data, cdf=None, cdfnum=None) -> (returned length, bytes written (out[:len]))
    dec(name, stream, outlen, cdf=None, cdfnum=None) -> decoded bytes
Buffers are laid out ``[in | pad | out]`` in ONE allocation so that ``out`` lies above ``in``: the
reference's anscdf4senc compares its output cursor with the input pointer (anscdf.c:63,66) and returns
a raw copy whenever ``out < in``.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (kind, needs_cdf, needs_cdfnum)
ENCODERS = {
    "anscdf4senc": (True, False), "anscdf4enc": (False, False), "anscdfenc": (False, False),
    "anscdf1enc": (False, False), "rccdfsenc": (True, True), "rccdfs2enc": (True, True),
    "rccdfenc": (False, False), "rccdfienc": (False, False), "rccdf4enc": (False, False),
    "rccdf4ienc": (False, False), "rccdfenc8": (False, False), "rccdfienc8": (False, False),
    "answenc": (True, True),      # port-only: this repository's 32-way interleaved static rANS (parity unpinned)
}
DECODERS = {
    "anscdf4sdec": (True, False), "anscdf4dec": (False, False), "anscdfdec": (False, False),
    "anscdf1dec": (False, False), "rccdfsbdec": (True, True), "rccdfsb2dec": (True, True),
    "rccdfsldec": (True, True), "rccdfsl2dec": (True, True),
    "rccdfdec": (False, False), "rccdfidec": (False, False), "rccdf4dec": (False, False),
    "rccdf4idec": (False, False), "rccdfdec8": (False, False), "rccdfidec8": (False, False),
    # port-only (no reference counterpart): true inverses with the tail state fixed / wide alphabet
    "ans_sdec_n": (True, True), "anscdf4dec_fix": (False, False), "answdec": (True, True),
}
# VLC-over-CDF integer codecs: (family, element bits); input/output are little-endian integers
VLC_CODECS = [("anscdfu", 16), ("anscdfuz", 16), ("anscdfv", 16), ("anscdfvz", 16), ("anscdfv", 32), ("anscdfvz", 32),
              ("rccdfv", 16), ("rccdfvz", 16), ("rccdfv", 32), ("rccdfvz", 32), ("rccdfu", 16), ("rccdfu", 32)]
for _n, _w in VLC_CODECS:
    ENCODERS[f"{_n}enc{_w}"] = (False, False)
    DECODERS[f"{_n}dec{_w}"] = (False, False)

PAIRS = {  # encoder -> decoder the reference harness pairs it with (turborc.c:495-536)
    "anscdf4senc": "anscdf4sdec", "anscdf4enc": "anscdf4dec", "anscdfenc": "anscdfdec",
    "anscdf1enc": "anscdf1dec", "rccdfsenc": "rccdfsbdec", "rccdfs2enc": "rccdfsb2dec",
    "rccdfenc": "rccdfdec", "rccdfienc": "rccdfidec", "rccdf4enc": "rccdf4dec",
    "rccdf4ienc": "rccdf4idec", "rccdfenc8": "rccdfdec8", "rccdfienc8": "rccdfidec8",
}


class _Lib:
    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.path = path

    def _fn(self, name, restype=ctypes.c_size_t):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def has(self, name):
        return hasattr(self.lib, self.prefix + name)

    def cdfini(self, data, cdfnum=256):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cdf = np.zeros(257, np.uint16)
        r = self._fn("cdfini", ctypes.c_int)(ctypes.c_void_p(data.ctypes.data), ctypes.c_size_t(data.size),
                                              ctypes.c_void_p(cdf.ctypes.data), ctypes.c_uint(cdfnum))
        if r < 0:
            raise ValueError("cdfini: degenerate distribution")
        return cdf

    @staticmethod
    def _extra(needs, cdf, cdfnum, keep):
        args = []
        if needs[0]:
            c = np.ascontiguousarray(cdf, dtype=np.uint16)
            keep.append(c)
            args.append(ctypes.c_void_p(c.ctypes.data))
        if needs[1]:
            args.append(ctypes.c_uint(int(cdfnum)))
        return args

    def enc(self, name, data, cdf=None, cdfnum=None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = data.size
        cap = n + n // 2 + 4096
        buf = np.zeros(n + 64 + cap, np.uint8)
        buf[:n] = data
        keep = []
        args = self._extra(ENCODERS[name], cdf, cdfnum, keep)
        pin = buf.ctypes.data
        pout = pin + n + 64
        r = self._fn(name)(ctypes.c_void_p(pin), ctypes.c_size_t(n), ctypes.c_void_p(pout), *args)
        m = min(int(r), n) if r <= cap else 0
        return int(r), buf[n + 64: n + 64 + max(m, 0)].copy()

    def dec(self, name, stream, outlen, cdf=None, cdfnum=None):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        buf = np.zeros(stream.size + 64 + outlen + 64, np.uint8)     # decoders over-read <= 4 B
        buf[:stream.size] = stream
        keep = []
        args = self._extra(DECODERS[name], cdf, cdfnum, keep)
        pin = buf.ctypes.data
        pout = pin + stream.size + 64
        self._fn(name)(ctypes.c_void_p(pin), ctypes.c_size_t(outlen), ctypes.c_void_p(pout), *args)
        return buf[stream.size + 64: stream.size + 64 + outlen].copy()

    def raw_fn(self, name):
        """Bare ctypes function (for timing loops that manage their own buffers)."""
        return self._fn(name)


def build(ref_too=True):
    """(Re)build the checkers with oracle/Makefile.  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    if ref_too:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_port = _ref = None


def port():
    global _port
    if _port is None:
        p = os.path.join(_HERE, "libtrc_oracle.so")
        if not os.path.exists(p):
            build(ref_too=False)
        _port = _Lib(p, "orc_")
    return _port


def ref():
    """The compiled reference, or None if oracle/_ref/libtrcref.so was never built."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libtrcref.so")
        if not os.path.exists(p):
            return None
        _ref = _Lib(p, "")
    return _ref


def cpu_bench(use_ref, enc, dec, data, chunk, cdf=None, cdfnum=0, threads=None, reps=1):
    """Time encoder+decoder per chunk on `threads` host threads (oracle/cpu_bench.c).
    -> dict(enc_s, dec_s, clen, ok, threads, kind)."""
    p = os.path.join(_HERE, "libtrc_cpubench.so")
    if not os.path.exists(p):
        build(ref_too=False)
    lib = ctypes.CDLL(p)
    refp = os.path.join(_HERE, "_ref", "libtrcref.so")
    if use_ref and os.path.exists(refp):
        libpath, prefix, kind = refp, "", "reference"
    else:
        port()
        libpath, prefix, kind = os.path.join(_HERE, "libtrc_oracle.so"), "orc_", "port"
    threads = threads or os.cpu_count() or 1
    data = np.ascontiguousarray(data, dtype=np.uint8)
    sig = 2 if ENCODERS[enc][1] else (1 if ENCODERS[enc][0] else 0)
    tab = np.zeros(257, np.uint16)
    if cdf is not None:
        c = np.ascontiguousarray(cdf, dtype=np.uint16).reshape(-1)
        tab[:min(257, c.size)] = c[:257]
    es, ds = ctypes.c_double(0), ctypes.c_double(0)
    cl, ok = ctypes.c_size_t(0), ctypes.c_int(0)
    lib.orc_cpu_bench.restype = ctypes.c_int
    rc = lib.orc_cpu_bench(libpath.encode(), (prefix + enc).encode(), (prefix + dec).encode(), ctypes.c_int(sig),
                           ctypes.c_void_p(data.ctypes.data), ctypes.c_size_t(data.size), ctypes.c_size_t(chunk),
                           ctypes.c_void_p(tab.ctypes.data), ctypes.c_uint(cdfnum), ctypes.c_int(threads), ctypes.c_int(reps),
                           ctypes.byref(es), ctypes.byref(ds), ctypes.byref(cl), ctypes.byref(ok))
    if rc:
        raise RuntimeError(f"orc_cpu_bench failed: {rc}")
    return dict(enc_s=es.value, dec_s=ds.value, clen=cl.value, ok=bool(ok.value), threads=threads, kind=kind)


def batch_enc(lib, enc, data, chunk, cdf=None, cdfnum=0, chunks_per_cdf=0, threads=None):
    """The checker's statement of the batch layer: `enc` called once per chunk by `threads` host threads, results packed
    back to back (oracle/cpu_bench.c orc_batch_enc).  `lib` = port() or ref().  -> (packed bytes, offsets[n+1])."""
    p = os.path.join(_HERE, "libtrc_cpubench.so")
    if not os.path.exists(p):
        build(ref_too=False)
    drv = ctypes.CDLL(p)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = -(-data.size // chunk)
    sig = 2 if ENCODERS[enc][1] else (1 if ENCODERS[enc][0] else 0)
    tabs = None
    if cdf is not None:
        tabs = np.ascontiguousarray(cdf, dtype=np.uint16).reshape(-1)
        nt = -(-n // chunks_per_cdf) if chunks_per_cdf else 1
        if tabs.size < nt * 257:
            tabs = np.concatenate([tabs, np.zeros(nt * 257 - tabs.size, np.uint16)])
    cap = data.size + 4 * n + 64
    out = np.empty(cap, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    drv.orc_batch_enc.restype = ctypes.c_int
    rc = drv.orc_batch_enc(lib.path.encode(), (lib.prefix + enc).encode(), ctypes.c_int(sig), ctypes.c_void_p(data.ctypes.data),
                           ctypes.c_size_t(data.size), ctypes.c_size_t(chunk), ctypes.c_void_p(tabs.ctypes.data if tabs is not None else None),
                           ctypes.c_uint(cdfnum), ctypes.c_size_t(chunks_per_cdf), ctypes.c_int(threads or os.cpu_count() or 1),
                           ctypes.c_void_p(out.ctypes.data), ctypes.c_size_t(cap), ctypes.c_void_p(off.ctypes.data))
    if rc:
        raise RuntimeError(f"orc_batch_enc failed: {rc}")
    return out[:int(off[n])], off
