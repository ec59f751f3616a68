"""ctypes front-end of oracle/hdlz_oracle.c — TEST INFRASTRUCTURE ONLY.

`compress(data)` restates the reference's FAST+MATCH10 static-tree compressor
(deflate.py:734-1016, see the C file for line-level citations); `inflate()`
restates its decoder; `batch()` runs either (or host zlib) over many items on
several threads for the CPU baseline of bench.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_PORT_COMPRESS, KIND_ZLIB_DEFLATE, KIND_ZLIB_INFLATE, KIND_PORT_INFLATE, KIND_FAST_COMPRESS = 0, 1, 2, 3, 4
STATUS_NAMES = ("OK", "SHORT_INPUT", "BAD_BTYPE", "BAD_CODE", "DIST_TOO_FAR", "TRUNCATED",
                "OUT_OVERFLOW", "BAD_STORED", "BAD_HEADER", "BAD_ADLER")


def build(force=False):
    so = os.path.join(_HERE, "libhdlz_oracle.so")
    src = os.path.join(_HERE, "hdlz_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libhdlz_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u8p, u32p, u64p = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        L.hdlz_oracle_compress.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32, u32p]
        L.hdlz_oracle_compress_ex.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32, u32p,
                                              ctypes.c_uint, ctypes.c_uint]
        L.hdlz_oracle_compress_fast.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32, u32p, ctypes.c_uint]
        L.hdlz_oracle_inflate.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32, u32p, ctypes.c_uint32]
        L.hdlz_oracle_parse.argtypes = [u8p, ctypes.c_uint32, u32p, ctypes.c_uint32]
        L.hdlz_oracle_parse.restype = ctypes.c_uint32
        L.hdlz_oracle_batch.argtypes = [ctypes.c_int, u8p, u64p, u32p, u8p, u64p, ctypes.c_uint32, u32p, u32p,
                                        ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _LIB = L
    return _LIB


def _buf(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return np.ascontiguousarray(a)


def compress(data, cwindow=32, maxlen=10):
    """-> (status, bytes). Reference compress contract."""
    a = _buf(data)
    out = np.empty(2 + (3 + 9 * len(a) + 7 + 7) // 8 + 4 + 8, dtype=np.uint8)
    n = ctypes.c_uint32(0)
    st = lib().hdlz_oracle_compress_ex(a.ctypes.data, len(a), out.ctypes.data, len(out), ctypes.byref(n),
                                       cwindow, maxlen)
    return st, out[:n.value].tobytes()


def compress_fast(data, maxlen=10):
    """-> (status, bytes).  The tuned CPU arm (SSE2 compares, 64-bit bit buffer): same bytes as compress()."""
    a = _buf(data)
    out = np.empty(2 + (3 + 9 * len(a) + 7 + 7) // 8 + 4 + 8, dtype=np.uint8)
    n = ctypes.c_uint32(0)
    st = lib().hdlz_oracle_compress_fast(a.ctypes.data, len(a), out.ctypes.data, len(out), ctypes.byref(n), maxlen)
    return st, out[:n.value].tobytes()


def inflate(stream, cap, flags=0):
    """-> (status, bytes)."""
    a = _buf(stream)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = ctypes.c_uint32(0)
    st = lib().hdlz_oracle_inflate(a.ctypes.data, len(a), out.ctypes.data, cap, ctypes.byref(n), flags)
    return st, out[:n.value].tobytes()


def parse(data):
    """Token trace [(pos, len, dist)] of the reference's greedy parse."""
    a = _buf(data)
    tok = np.empty(len(a) + 1, dtype=np.uint32)
    n = lib().hdlz_oracle_parse(a.ctypes.data, len(a), tok.ctypes.data, len(tok))
    t = tok[:n]
    return [(int(v & 0xFFFF), int(v >> 24), int((v >> 16) & 0xFF)) for v in t]


def batch(kind, inp, in_off, in_len, out, out_off, out_cap, nthreads=1, level=6, strategy=0):
    """Run `kind` over len(in_len) items; returns (out_len[u32], status[u32])."""
    n = len(in_len)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint32)
    out_len = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.uint32)
    lib().hdlz_oracle_batch(kind, inp.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, out.ctypes.data,
                            out_off.ctypes.data, out_cap, out_len.ctypes.data, status.ctypes.data, n,
                            nthreads, level, strategy)
    return out_len, status
