"""turbo-range-coder_b200 -- B200-native CDF entropy path (rANS + range coder) behind the TurboRC API.

This module is only a ctypes view of ``libtrc_b200.so`` (built from ``csrc/`` for sm_100a) mirroring the
reference's interface for this path: the same function names (``anscdfenc`` .. ``rccdfs2enc``, ``cdfini``),
argument meaning and raw-copy rule, plus the batch entry points.  There is no Python or CPU implementation
behind it: if the shared library is missing the import fails loudly.

    import importlib; trc = importlib.import_module("turbo-range-coder_b200")
    out, off = trc.enc_batch_host(trc.RCS2, data, 4096, cdf=cdf, cdfnum=256)
    back     = trc.dec_batch_host(trc.RCS2, out, off, data.size, 4096, cdf=cdf, cdfnum=256)
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrc_b200.so")

# enum trc_codec (include/trc_b200.h)
ANS4S, ANS4, ANS, ANS1, RCS, RCS2, RC, RCI, RC4, RC4I, ANSW, RC8, RCI8 = range(13)
(ANSU16, ANSUZ16, ANSV16, ANSVZ16, ANSV32, ANSVZ32, RCV16, RCVZ16, RCV32, RCVZ32, RCU16, RCU32) = range(13, 25)   # VLC-over-CDF integer codecs
CODEC_NAMES = ["ANS4S", "ANS4", "ANS", "ANS1", "RCS", "RCS2", "RC", "RCI", "RC4", "RC4I", "ANSW", "RC8", "RCI8",
               "ANSU16", "ANSUZ16", "ANSV16", "ANSVZ16", "ANSV32", "ANSVZ32", "RCV16", "RCVZ16", "RCV32", "RCVZ32", "RCU16", "RCU32"]
#: codec id -> (reference encoder, reference decoder)
REF_NAMES = {
    ANS4S: ("anscdf4senc", "anscdf4sdec"), ANS4: ("anscdf4enc", "anscdf4dec"), ANS: ("anscdfenc", "anscdfdec"),
    ANS1: ("anscdf1enc", "anscdf1dec"), RCS: ("rccdfsenc", "rccdfsbdec"), RCS2: ("rccdfs2enc", "rccdfsb2dec"),
    RC: ("rccdfenc", "rccdfdec"), RCI: ("rccdfienc", "rccdfidec"), RC4: ("rccdf4enc", "rccdf4dec"),
    RC4I: ("rccdf4ienc", "rccdf4idec"), RC8: ("rccdfenc8", "rccdfdec8"), RCI8: ("rccdfienc8", "rccdfidec8"),
    ANSU16: ("anscdfuenc16", "anscdfudec16"), ANSUZ16: ("anscdfuzenc16", "anscdfuzdec16"), ANSV16: ("anscdfvenc16", "anscdfvdec16"),
    ANSVZ16: ("anscdfvzenc16", "anscdfvzdec16"), ANSV32: ("anscdfvenc32", "anscdfvdec32"), ANSVZ32: ("anscdfvzenc32", "anscdfvzdec32"),
    RCV16: ("rccdfvenc16", "rccdfvdec16"), RCVZ16: ("rccdfvzenc16", "rccdfvzdec16"), RCV32: ("rccdfvenc32", "rccdfvdec32"),
    RCVZ32: ("rccdfvzenc32", "rccdfvzdec32"), RCU16: ("rccdfuenc16", "rccdfudec16"), RCU32: ("rccdfuenc32", "rccdfudec32"),
}
STATIC = (ANS4S, RCS, RCS2, ANSW)
F_REF_TAIL = 1
CDF_STRIDE = 257
OK, E_ARG, E_CUDA, E_NOMEM = 0, -1, -2, -3


class TrcError(RuntimeError):
    pass


def build():
    """Compile csrc/ for sm_100a into libtrc_b200.so (nvcc cross-compiles without a GPU)."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)


if not os.path.exists(LIB_PATH):
    try:                                       # a fresh checkout: compile once (needs nvcc; no GPU required)
        build()
    except Exception as e:                     # noqa: BLE001
        raise ImportError(f"{LIB_PATH} is missing and could not be built ({e}): run `make -C {_HERE}/csrc` "
                          "(or __graft_entry__.build()).  There is no CPU fallback.") from e
lib = ctypes.CDLL(LIB_PATH)

_sz, _vp, _u, _i = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int
lib.trc_version.restype = ctypes.c_char_p
lib.trc_last_error.restype = ctypes.c_char_p
lib.trc_launch_count.restype = ctypes.c_ulonglong
lib.trc_num_chunks.restype = _sz; lib.trc_num_chunks.argtypes = [_sz, _sz]
lib.trc_enc_bound.restype = _sz; lib.trc_enc_bound.argtypes = [_sz, _sz]
lib.trc_enc_scratch_bytes.restype = _sz; lib.trc_enc_scratch_bytes.argtypes = [_i, _sz, _sz]
lib.trc_enc_batch_dev.restype = _i
lib.trc_enc_batch_dev.argtypes = [_i, _vp, _sz, _sz, _vp, _u, _sz, _vp, _vp, _vp, _sz, _vp]
lib.trc_dec_batch_dev.restype = _i
lib.trc_dec_batch_dev.argtypes = [_i, _vp, _vp, _vp, _sz, _sz, _vp, _u, _sz, _u, _vp]
lib.trc_tables_create_dev.restype = _i
lib.trc_tables_create_dev.argtypes = [_vp, _u, _sz, _vp, _vp]
lib.trc_tables_destroy.restype = None; lib.trc_tables_destroy.argtypes = [_vp]
lib.trc_enc_batch_dev_tab.restype = _i
lib.trc_enc_batch_dev_tab.argtypes = [_i, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]
lib.trc_dec_batch_dev_tab.restype = _i
lib.trc_dec_batch_dev_tab.argtypes = [_i, _vp, _vp, _vp, _sz, _sz, _vp, _sz, _u, _vp]
lib.trc_enc_batch_host.restype = _i
lib.trc_enc_batch_host.argtypes = [_i, _vp, _sz, _sz, _vp, _u, _sz, _vp, _vp, _vp]
lib.trc_dec_batch_host.restype = _i
lib.trc_dec_batch_host.argtypes = [_i, _vp, _vp, _vp, _sz, _sz, _vp, _u, _sz, _u]
lib.trc_cdfini_batch_dev.restype = _i
lib.trc_cdfini_batch_dev.argtypes = [_vp, _sz, _sz, _vp, _u, _vp, _vp]
lib.trc_set_device.restype = _i; lib.trc_set_device.argtypes = [_i]
lib.trc_device_count.restype = _i
lib.cdfini.restype = _i; lib.cdfini.argtypes = [_vp, _sz, _vp, _u]


def _check(rc, what):
    if rc != OK:
        raise TrcError(f"{what} failed with {rc}: {lib.trc_last_error().decode()}")


def version():
    return lib.trc_version().decode()


def launch_count():
    """Kernels launched by the library so far in this process."""
    return int(lib.trc_launch_count())


def profile_enable(on=True):
    lib.trc_profile_enable(1 if on else 0)


def profile_read():
    """Milliseconds of each kernel of the most recent batch call (needs profile_enable())."""
    buf = (ctypes.c_float * 8)()
    n = lib.trc_profile_read(buf, 8)
    return [float(buf[i]) for i in range(n)]


def num_chunks(total_len, chunk_len):
    return int(lib.trc_num_chunks(total_len, chunk_len))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _cdfarg(cdf):
    if cdf is None:
        return None, None
    c = np.ascontiguousarray(cdf, dtype=np.uint16)
    return c, ctypes.c_void_p(c.ctypes.data)


# ----------------------------------------------------------------------------------------------------------
# batch layer, host pointers (H2D / D2H inside the call)
# ----------------------------------------------------------------------------------------------------------
def enc_batch_host(codec, data, chunk_len, cdf=None, cdfnum=0, chunks_per_cdf=0, out=None, off=None):
    """-> (packed bytes, offsets[n+1]); chunk c == reference encoder called on data[c*chunk_len:...]"""
    data = _u8(data)
    n = num_chunks(data.size, chunk_len)
    if out is None:
        out = np.empty(int(lib.trc_enc_bound(data.size, chunk_len)), np.uint8)
    if off is None:
        off = np.empty(n + 1, np.uint64)
    keep, cp = _cdfarg(cdf)
    olen = ctypes.c_size_t(0)
    rc = lib.trc_enc_batch_host(codec, data.ctypes.data, data.size, chunk_len, cp, cdfnum, chunks_per_cdf,
                                out.ctypes.data, off.ctypes.data, ctypes.addressof(olen))
    _check(rc, "trc_enc_batch_host")
    return out[:olen.value], off


def dec_batch_host(codec, stream, off, total_len, chunk_len, cdf=None, cdfnum=0, chunks_per_cdf=0, flags=0, out=None):
    stream = _u8(stream)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    if out is None:
        out = np.empty(total_len, np.uint8)
    keep, cp = _cdfarg(cdf)
    rc = lib.trc_dec_batch_host(codec, stream.ctypes.data, off.ctypes.data, out.ctypes.data, total_len, chunk_len,
                                cp, cdfnum, chunks_per_cdf, flags)
    _check(rc, "trc_dec_batch_host")
    return out


lib.trc_enc_batch_host_multi.restype = _i
lib.trc_enc_batch_host_multi.argtypes = [_i, _vp, _i, _vp, _sz, _sz, _vp, _u, _sz, _vp, _vp, _vp]
lib.trc_dec_batch_host_multi.restype = _i
lib.trc_dec_batch_host_multi.argtypes = [_i, _vp, _i, _vp, _vp, _vp, _sz, _sz, _vp, _u, _sz, _u]


def enc_batch_host_multi(codec, devs, data, chunk_len, cdf=None, cdfnum=0, chunks_per_cdf=0, out=None, off=None):
    """enc_batch_host with the chunks sharded over the GPUs `devs` (one process, one host thread per device)."""
    data = _u8(data)
    n = num_chunks(data.size, chunk_len)
    if out is None:
        out = np.empty(int(lib.trc_enc_bound(data.size, chunk_len)), np.uint8)
    if off is None:
        off = np.empty(n + 1, np.uint64)
    keep, cp = _cdfarg(cdf)
    dv = (ctypes.c_int * len(devs))(*devs)
    olen = ctypes.c_size_t(0)
    rc = lib.trc_enc_batch_host_multi(codec, dv, len(devs), data.ctypes.data, data.size, chunk_len, cp, cdfnum, chunks_per_cdf,
                                      out.ctypes.data, off.ctypes.data, ctypes.addressof(olen))
    _check(rc, "trc_enc_batch_host_multi")
    return out[:olen.value], off


def dec_batch_host_multi(codec, devs, stream, off, total_len, chunk_len, cdf=None, cdfnum=0, chunks_per_cdf=0, flags=0, out=None):
    stream = _u8(stream)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    if out is None:
        out = np.empty(total_len, np.uint8)
    keep, cp = _cdfarg(cdf)
    dv = (ctypes.c_int * len(devs))(*devs)
    rc = lib.trc_dec_batch_host_multi(codec, dv, len(devs), stream.ctypes.data, off.ctypes.data, out.ctypes.data, total_len, chunk_len,
                                      cp, cdfnum, chunks_per_cdf, flags)
    _check(rc, "trc_dec_batch_host_multi")
    return out


# ----------------------------------------------------------------------------------------------------------
# self-describing container (include/trc_b200.h): tables computed on the GPU, directory + payload in one buffer
# ----------------------------------------------------------------------------------------------------------
lib.trc_container_bound.restype = _sz; lib.trc_container_bound.argtypes = [_i, _sz, _sz, _sz]
lib.trc_compress_host.restype = _i; lib.trc_compress_host.argtypes = [_i, _vp, _sz, _sz, _sz, _vp, _sz, _vp]
lib.trc_decompress_host.restype = _i; lib.trc_decompress_host.argtypes = [_vp, _sz, _vp, _sz, _vp]
lib.trc_container_info.restype = _i; lib.trc_container_info.argtypes = [_vp, _sz, _vp, _vp, _vp, _vp]
CONTAINER_HEADER = 64


def compress(codec, data, chunk_len, cdf_block=0):
    """-> container bytes (header, static tables made by the GPU's cdfini, per-chunk lengths, packed reference streams)"""
    data = _u8(data)
    cap = int(lib.trc_container_bound(codec, data.size, chunk_len, cdf_block))
    if cap == 0:
        raise TrcError("trc_container_bound: bad arguments")
    out = np.empty(cap, np.uint8)
    olen = ctypes.c_size_t(0)
    _check(lib.trc_compress_host(codec, data.ctypes.data, data.size, chunk_len, cdf_block, out.ctypes.data, cap, ctypes.addressof(olen)),
           "trc_compress_host")
    return out[:olen.value]


def container_info(blob):
    """-> dict(codec, total_len, chunk_len, n_chunks); raises TrcError on a malformed header (host only, no GPU)"""
    blob = _u8(blob)
    codec = ctypes.c_int(0)
    tl, cl, n = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = lib.trc_container_info(blob.ctypes.data if blob.size else None, blob.size, ctypes.addressof(codec), ctypes.addressof(tl),
                                ctypes.addressof(cl), ctypes.addressof(n))
    if rc != OK:
        raise TrcError(f"trc_container_info failed ({rc}): not a TRCB container")
    return {"codec": codec.value, "total_len": tl.value, "chunk_len": cl.value, "n_chunks": n.value}


def decompress(blob):
    blob = _u8(blob)
    info = container_info(blob)
    out = np.empty(info["total_len"], np.uint8)
    olen = ctypes.c_size_t(0)
    _check(lib.trc_decompress_host(blob.ctypes.data, blob.size, out.ctypes.data, out.size, ctypes.addressof(olen)), "trc_decompress_host")
    return out[:olen.value]


# ----------------------------------------------------------------------------------------------------------
# batch layer, device pointers (torch tensors only carry the memory and the stream)
# ----------------------------------------------------------------------------------------------------------
class DeviceBatch:
    """Device-resident encode/decode of one geometry; buffers are allocated once and reused."""

    def __init__(self, codec, total_len, chunk_len, cdfnum=0, chunks_per_cdf=0, device="cuda:0"):
        import torch
        self.torch = torch
        self.codec, self.total_len, self.chunk_len = codec, int(total_len), int(chunk_len)
        self.cdfnum, self.cpc = int(cdfnum), int(chunks_per_cdf)
        self.n = num_chunks(total_len, chunk_len)
        self.device = torch.device(device)
        sb = int(lib.trc_enc_scratch_bytes(codec, total_len, chunk_len))
        if sb == 0:
            raise TrcError("bad geometry")
        self.scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
        self.out = torch.empty(int(lib.trc_enc_bound(total_len, chunk_len)), dtype=torch.uint8, device=self.device)
        self.off = torch.empty(self.n + 1, dtype=torch.int64, device=self.device)
        self.dec = torch.empty(total_len + 64, dtype=torch.uint8, device=self.device)
        self.cdf = None
        self.tables = None                    # trc_tables handle (prebuilt coding tables), see prebuild_tables()

    def set_cdf(self, cdf):
        c = np.ascontiguousarray(cdf, dtype=np.uint16).reshape(-1)
        self.cdf = self.torch.from_numpy(c.view(np.int16).copy()).to(self.device)
        self.drop_tables()

    def prebuild_tables(self):
        """Build the coding tables of self.cdf once (trc_tables_create_dev); encode()/decode() then skip the per-call build."""
        self.drop_tables()
        n_tab = -(-self.n // self.cpc) if self.cpc else 1
        h = ctypes.c_void_p()
        _check(lib.trc_tables_create_dev(self.cdf.data_ptr(), self.cdfnum, n_tab, self._stream(), ctypes.byref(h)), "trc_tables_create_dev")
        self.tables = h

    def borrow_tables(self, other):
        """Use the prebuilt tables of another DeviceBatch on the same device and cdf (not owned: `other` must outlive this)."""
        self.drop_tables()
        self.tables, self._borrowed = other.tables, True

    def drop_tables(self):
        if getattr(self, "tables", None) and not getattr(self, "_borrowed", False):
            self.torch.cuda.synchronize(self.device)
            lib.trc_tables_destroy(self.tables)
        self.tables, self._borrowed = None, False

    def __del__(self):
        try:
            self.drop_tables()
        except Exception:      # noqa: BLE001  (interpreter shutdown)
            pass

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def encode(self, d_in):
        """d_in: uint8 cuda tensor of total_len bytes.  Fills self.out / self.off (asynchronous)."""
        if self.tables:
            rc = lib.trc_enc_batch_dev_tab(self.codec, d_in.data_ptr(), self.total_len, self.chunk_len, self.tables, self.cpc,
                                           self.out.data_ptr(), self.off.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(),
                                           self._stream())
            return _check(rc, "trc_enc_batch_dev_tab")
        cp = self.cdf.data_ptr() if self.cdf is not None else None
        rc = lib.trc_enc_batch_dev(self.codec, d_in.data_ptr(), self.total_len, self.chunk_len, cp, self.cdfnum, self.cpc,
                                   self.out.data_ptr(), self.off.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(),
                                   self._stream())
        _check(rc, "trc_enc_batch_dev")

    def decode(self, d_stream=None, d_off=None, flags=0):
        cp = self.cdf.data_ptr() if self.cdf is not None else None
        s = self.out if d_stream is None else d_stream
        o = self.off if d_off is None else d_off
        if self.tables:
            rc = lib.trc_dec_batch_dev_tab(self.codec, s.data_ptr(), o.data_ptr(), self.dec.data_ptr(), self.total_len, self.chunk_len,
                                           self.tables, self.cpc, flags, self._stream())
            _check(rc, "trc_dec_batch_dev_tab")
            return self.dec[: self.total_len]
        rc = lib.trc_dec_batch_dev(self.codec, s.data_ptr(), o.data_ptr(), self.dec.data_ptr(), self.total_len, self.chunk_len,
                                   cp, self.cdfnum, self.cpc, flags, self._stream())
        _check(rc, "trc_dec_batch_dev")
        return self.dec[: self.total_len]

    def compressed_len(self):
        return int(self.off[self.n].item())


def cdfini_dev(d_in, total_len, chunk_len, cdfnum=256):
    """Per-chunk static tables computed on the device -> (int16-view cuda tensor [n*257], status tensor)."""
    import torch
    n = num_chunks(total_len, chunk_len)
    cdf = torch.zeros(n * CDF_STRIDE, dtype=torch.int16, device=d_in.device)
    status = torch.zeros(n, dtype=torch.int32, device=d_in.device)
    rc = lib.trc_cdfini_batch_dev(d_in.data_ptr(), total_len, chunk_len, cdf.data_ptr(), cdfnum, status.data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream(d_in.device).cuda_stream))
    _check(rc, "trc_cdfini_batch_dev")
    return cdf, status


# ----------------------------------------------------------------------------------------------------------
# drop-in layer: the reference's own names (host buffers, whole-buffer calls)
# ----------------------------------------------------------------------------------------------------------
_ENC3 = ["anscdf4enc", "anscdfenc", "anscdf1enc", "rccdfenc", "rccdfienc", "rccdf4enc", "rccdf4ienc", "rccdfenc8", "rccdfienc8",
         "anscdf4encs", "anscdf4encx", "anscdfencs", "anscdfencx", "anscdf1encs", "anscdf1encx"]
_DEC3 = ["anscdf4dec", "anscdfdec", "anscdf1dec", "rccdfdec", "rccdfidec", "rccdf4dec", "rccdf4idec", "rccdfdec8", "rccdfidec8",
         "anscdf4decs", "anscdf4decx", "anscdfdecs", "anscdfdecx", "anscdf1decs", "anscdf1decx"]
_ENC3 += [e for c, (e, d) in sorted(REF_NAMES.items()) if c >= ANSU16]
_DEC3 += [d for c, (e, d) in sorted(REF_NAMES.items()) if c >= ANSU16]
_ENC4 = ["anscdf4senc", "anscdf4sencs", "anscdf4sencx"]
_DEC4 = ["anscdf4sdec", "anscdf4sdecs", "anscdf4sdecx"]
_ENC5 = ["rccdfsenc", "rccdfs2enc"]
_DEC5 = ["rccdfsbdec", "rccdfsldec", "rccdfsb2dec", "rccdfsl2dec", "rccdfsvbdec", "rccdfsvldec"]
DROPIN_SYMBOLS = _ENC3 + _DEC3 + _ENC4 + _DEC4 + _ENC5 + _DEC5 + ["cdfini", "anscdfini"]


def _dropin(name):
    f = getattr(lib, name)
    f.restype = ctypes.c_size_t
    return f


def dropin_enc(name, data, cdf=None, cdfnum=None):
    """Call the drop-in encoder `name` like the reference harness does -> (returned length, out[:length])."""
    data = _u8(data)
    out = np.zeros(data.size + data.size // 3 + 1024, np.uint8)          # OSIZE(n) = n*4/3 (turborc.c:418)
    args = [ctypes.c_void_p(data.ctypes.data), ctypes.c_size_t(data.size), ctypes.c_void_p(out.ctypes.data)]
    keep, cp = _cdfarg(cdf)
    if name in _ENC4 or name in _ENC5:
        args.append(cp)
    if name in _ENC5:
        args.append(ctypes.c_uint(int(cdfnum)))
    r = int(_dropin(name)(*args))
    return r, out[: min(r, out.size)].copy()


def dropin_dec(name, stream, outlen, cdf=None, cdfnum=None):
    stream = _u8(stream)
    buf = np.zeros(max(stream.size, outlen) + 64, np.uint8)               # the decoder ships outlen bytes of `in`
    buf[: stream.size] = stream
    out = np.zeros(outlen, np.uint8)
    args = [ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(outlen), ctypes.c_void_p(out.ctypes.data)]
    keep, cp = _cdfarg(cdf)
    if name in _DEC4 or name in _DEC5:
        args.append(cp)
    if name in _DEC5:
        args.append(ctypes.c_uint(int(cdfnum)))
    _dropin(name)(*args)
    return out


def cdfini(data, cdfnum=256):
    """Drop-in cdfini (rccdf.c:50): histogram + normalisation on the GPU -> cdf_t[257]."""
    data = _u8(data)
    cdf = np.zeros(CDF_STRIDE, np.uint16)
    lib.cdfini(data.ctypes.data, data.size, cdf.ctypes.data, cdfnum)
    return cdf
