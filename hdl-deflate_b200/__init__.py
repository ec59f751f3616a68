"""hdl-deflate_b200 — B200-native deflate engine behind the HDL-deflate interface.

Host side of the hot path only:
  * `Engine`      thin object over the C ABI of include/hdlz.h (ctypes, libhdlz.so);
  * `dropin/`     `deflate` + `myhdl` modules with the reference's names, so the
                  reference's own test bench drives the GPU engine
                  (put hdl-deflate_b200/dropin on PYTHONPATH);
  * `workload`    the synthetic block generator of the benchmark configs.

The compute lives in csrc/*.cu (sm_100a).  There is no CPU implementation in this
package: without libhdlz.so or without a GPU every call raises.
"""
import ctypes

import numpy as np

from . import _native
from . import workload  # noqa: F401

# constants of the reference module (deflate.py:18, 56-89) at the BASELINE sizes
IDLE, WRITE, READ, STARTC, STARTD = range(5)
CWINDOW = 32
IBSIZE = 2048
OBSIZE = 32768
LMAX = 24
MIN_INPUT = 5

F_VERIFY_HEADER = 1
F_VERIFY_ADLER = 2     # the container checksum: Adler-32 (zlib) or CRC-32 + ISIZE (gzip)
F_RAW = 4              # decompress: bare RFC 1951 stream
F_GZIP = 8             # decompress: one RFC 1952 member
F_PERSIST_TABLES = 16  # decompress: keep the dynamic-block decode tables in the L2 (persisting carve-out)
CONTAINER_ZLIB, CONTAINER_RAW, CONTAINER_GZIP = 0, 1, 2

STATUS_NAMES = ("OK", "SHORT_INPUT", "BAD_BTYPE", "BAD_CODE", "DIST_TOO_FAR", "TRUNCATED",
                "OUT_OVERFLOW", "BAD_STORED", "BAD_HEADER", "BAD_ADLER", "BAD_CRC", "NO_CODE")

# the message the reference raises for the condition (deflate.py:721, 1140, 1508, 1539, 1560)
REFERENCE_MESSAGES = {
    1: "input shorter than 5 bytes: the engine never starts (isize < 4)",
    2: "Bad method",
    3: "Invalid data",
    4: "distance too big",
    5: "NO EOF!",
    6: "output buffer too small",
    7: "Invalid data",
    8: "unexpected mode",
    9: "Invalid data",
    11: "the tree has no code for a symbol of this stream",
}


class HdlzError(RuntimeError):
    """A C-ABI call failed (negative hdlz_error)."""


class StreamError(ValueError):
    """A stream finished with a non-zero hdlz_status."""

    def __init__(self, status):
        self.status = int(status)
        name = STATUS_NAMES[self.status] if self.status < len(STATUS_NAMES) else "UNKNOWN"
        ValueError.__init__(self, "%s (%s)" % (REFERENCE_MESSAGES.get(self.status, "error"), name))


def compress_bound(n, container=CONTAINER_ZLIB):
    return int(_native.load().hdlz_compress_bound_ex(int(n), int(container)))


def _ptr(a):
    """Device/host address of a numpy array, torch tensor or raw int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return int(a.ctypes.data)


class Engine(object):
    """One hdlz context (one GPU).  Mirrors one instantiated `deflate()` block."""

    def __init__(self, device=0):
        self._lib = _native.load()
        self._ctx = ctypes.c_void_p()
        self.device = int(device)
        self._check(self._lib.hdlz_create(self.device, ctypes.byref(self._ctx)))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            for ptr, _ in getattr(self, "_pin", {}).values():
                self._lib.hdlz_host_free_pinned(self._ctx, ptr)
            self._pin = {}
            self._lib.hdlz_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def _pinned(self, which, nbytes):
        """A pinned staging buffer of the engine (grown on demand): single-stream calls copy through it, so the
        device copies run at the link's rate and no fresh pages are touched per call."""
        if not hasattr(self, "_pin"):
            self._pin = {}
        ptr, cap = self._pin.get(which, (None, 0))
        if cap < nbytes:
            if ptr is not None:
                self._lib.hdlz_host_free_pinned(self._ctx, ptr)
            cap = max(1 << 16, nbytes + nbytes // 4)
            ptr = ctypes.c_void_p()
            self._check(self._lib.hdlz_host_alloc_pinned(self._ctx, cap, ctypes.byref(ptr)))
            self._pin[which] = (ptr, cap)
        return ptr

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise HdlzError("hdlz error %d: %s" % (rc, self._lib.hdlz_last_error().decode("utf-8", "replace")))

    @property
    def launch_count(self):
        return int(self._lib.hdlz_launch_count(self._ctx))

    @property
    def match10(self):
        """The reference's MATCH10 switch (deflate.py:34-35): True = matches up to 10 bytes (default),
        False = up to 5."""
        return bool(self._lib.hdlz_get_match10(self._ctx))

    @match10.setter
    def match10(self, on):
        self._check(self._lib.hdlz_set_match10(self._ctx, 1 if on else 0))

    @property
    def fast(self):
        """The reference's FAST switch (deflate.py:36-37): True = 32-byte window (default), False = the non-FAST
        engine with CWINDOW = 256."""
        return bool(self._lib.hdlz_get_fast(self._ctx))

    @fast.setter
    def fast(self, on):
        self._check(self._lib.hdlz_set_fast(self._ctx, 1 if on else 0))

    @property
    def container(self):
        """Framing the compressor writes around the deflate body: CONTAINER_ZLIB (the reference's,
        default), CONTAINER_RAW or CONTAINER_GZIP."""
        return int(self._lib.hdlz_get_container(self._ctx))

    @container.setter
    def container(self, kind):
        self._check(self._lib.hdlz_set_container(self._ctx, int(kind)))

    # ---- the code the compressor writes with (README.md:43-45 "dedicated pre-computed Huffman tree") ----
    @property
    def tree(self):
        """None (the reference's fixed code) or (lit_len uint8[286], dist_len uint8[30])."""
        lit, dist = np.zeros(286, np.uint8), np.zeros(30, np.uint8)
        if not self._lib.hdlz_get_tree(self._ctx, _ptr(lit), _ptr(dist)):
            return None
        return lit, dist

    def set_tree(self, lit_len=None, dist_len=None):
        """Install code lengths (hdlz_set_tree); no arguments = back to the fixed code."""
        if lit_len is None and dist_len is None:
            self._check(self._lib.hdlz_set_tree(self._ctx, None, None))
            return
        lit = np.ascontiguousarray(lit_len, dtype=np.uint8)
        dist = np.ascontiguousarray(dist_len, dtype=np.uint8)
        if lit.shape != (286,) or dist.shape != (30,):
            raise ValueError("lit_len needs 286 entries and dist_len 30")
        self._check(self._lib.hdlz_set_tree(self._ctx, _ptr(lit), _ptr(dist)))

    def train_tree(self, blocks, lens=None):
        """Count the symbols of the parse of `blocks` (uint8 [n, in_stride], host) on the GPU and install the
        optimal length-limited code for them (hdlz_train_tree).  -> the installed (lit_len, dist_len)."""
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        n, in_stride = blocks.shape
        if in_stride % 16:
            raise ValueError("in_stride must be a multiple of 16")
        d_in, d_len = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self._lib.hdlz_dev_alloc(self._ctx, blocks.nbytes, ctypes.byref(d_in)))
        try:
            self._check(self._lib.hdlz_copy_h2d(self._ctx, d_in, _ptr(blocks), blocks.nbytes, None))
            if lens is not None:
                lens = np.ascontiguousarray(lens, dtype=np.uint32)
                self._check(self._lib.hdlz_dev_alloc(self._ctx, lens.nbytes, ctypes.byref(d_len)))
                self._check(self._lib.hdlz_copy_h2d(self._ctx, d_len, _ptr(lens), lens.nbytes, None))
            self.train_tree_device(d_in.value, in_stride, d_len.value, in_stride, n)
        finally:
            self._lib.hdlz_dev_free(self._ctx, d_in)
            if d_len.value:
                self._lib.hdlz_dev_free(self._ctx, d_len)
        return self.tree

    def train_tree_device(self, d_in, in_stride, d_in_len, uniform_len, n, stream=0):
        self._check(self._lib.hdlz_train_tree(self._ctx, _ptr(d_in), in_stride, _ptr(d_in_len), uniform_len, n,
                                              stream or None))

    def bound(self, n):
        """Slot size a stream of n input bytes needs with the current container and code."""
        return int(self._lib.hdlz_compress_bound_tree(self._ctx, int(n)))

    # ---- one stream: a STARTC / STARTD job ---------------------------------------------
    def compress(self, data):
        """zlib stream of `data`, bit-identical to the reference's FAST+MATCH10 output."""
        data = bytes(data)
        cap = self.bound(len(data))
        src, out = self._pinned("in", max(len(data), 16)), self._pinned("out", cap)
        ctypes.memmove(src, data, len(data))
        n, st = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._check(self._lib.hdlz_compress_stream(self._ctx, src, len(data), out, cap, ctypes.byref(n), ctypes.byref(st)))
        if st.value:
            raise StreamError(st.value)
        return ctypes.string_at(out, n.value)

    def compress_dynamic(self, data):
        """One stream coded with its own Huffman tree (hdlz_compress_stream_dyn): statistics counted on the GPU over
        the stream's 2 KiB blocks, one BTYPE = 10 block, the reference's parse."""
        data = bytes(data)
        cap = compress_bound(len(data), self.container) + 2 * len(data) + 512      # a 15-bit code is the worst case
        src, out = self._pinned("in", max(len(data), 16)), self._pinned("out", cap)
        ctypes.memmove(src, data, len(data))
        n, st = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._check(self._lib.hdlz_compress_stream_dyn(self._ctx, src, len(data), out, cap, ctypes.byref(n), ctypes.byref(st)))
        if st.value:
            raise StreamError(st.value)
        return ctypes.string_at(out, n.value)

    def decompress(self, data, max_out=None, flags=0):
        """Inflate one zlib stream.  `max_out=None` grows the buffer until it fits (< 2^LMAX)."""
        data = bytes(data)
        src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
        cap = int(max_out) if max_out is not None else max(1 << 16, 8 * len(data))
        while True:
            out = np.empty(max(cap, 1), dtype=np.uint8)
            n, st = ctypes.c_uint32(0), ctypes.c_uint32(0)
            self._check(self._lib.hdlz_decompress_stream(self._ctx, src.ctypes.data, len(data), out.ctypes.data, cap,
                                                         ctypes.byref(n), ctypes.byref(st), flags))
            if st.value == 6 and max_out is None and cap < (1 << LMAX):
                cap = min(cap * 4, 1 << LMAX)
                continue
            if st.value:
                raise StreamError(st.value)
            return out[:n.value].tobytes()

    # ---- batches in HOST memory (numpy) --------------------------------------------------
    def compress_host(self, blocks, lens=None, out_stride=None):
        """blocks: uint8 [n, in_stride] (C-contiguous, in_stride % 16 == 0).
        -> (out uint8 [n, out_stride], out_len uint32 [n], status uint32 [n])."""
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        n, in_stride = blocks.shape
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.uint32)
            maxlen = int(lens.max()) if n else 0
        else:
            maxlen = in_stride
        if out_stride is None:
            out_stride = self.bound(maxlen)
        out = np.empty((n, out_stride), dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint32)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.hdlz_compress_host(self._ctx, _ptr(blocks), in_stride, _ptr(lens), in_stride,
                                                 _ptr(out), out_stride, _ptr(out_len), _ptr(status), n))
        return out, out_len, status

    def compress_host_packed(self, blocks, lens=None):
        """As compress_host, but the streams come back packed (starts 4-byte aligned).
        -> (packed uint8 [total], off uint64 [n], out_len uint32 [n], status uint32 [n])."""
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        n, in_stride = blocks.shape
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.uint32)
            maxlen = int(lens.max()) if n else 0
        else:
            maxlen = in_stride
        cap = n * self.bound(maxlen)
        out = np.empty(max(cap, 16), dtype=np.uint8)
        off = np.zeros(n, dtype=np.uint64)
        out_len = np.zeros(n, dtype=np.uint32)
        status = np.zeros(n, dtype=np.uint32)
        total = ctypes.c_uint64(0)
        self._check(self._lib.hdlz_compress_host_packed(self._ctx, _ptr(blocks), in_stride, _ptr(lens), in_stride,
                                                        _ptr(out), cap, _ptr(off), _ptr(out_len), _ptr(status), n,
                                                        ctypes.byref(total)))
        return out[:total.value], off, out_len, status

    def pack_batch(self, d_slots, stride, d_len, d_packed, d_off, d_total, n, stream=0):
        self._check(self._lib.hdlz_pack_batch(self._ctx, _ptr(d_slots), stride, _ptr(d_len), _ptr(d_packed),
                                              _ptr(d_off), _ptr(d_total), n, stream or None))

    def decompress_host(self, data, in_len, out_cap, in_off=None, in_stride=0, out_stride=None, flags=0):
        """data: uint8 buffer holding the streams (packed with in_off, or [n, in_stride]).
        -> (out uint8 [n, out_stride], out_len, status)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint32)
        n = len(in_len)
        if in_off is not None:
            in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        elif data.ndim == 2:
            in_stride = data.shape[1]
        if out_stride is None:
            out_stride = (int(out_cap) + 15) & ~15
        out = np.empty((n, out_stride), dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint32)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.hdlz_decompress_host(self._ctx, _ptr(data), _ptr(in_off), in_stride, _ptr(in_len),
                                                   _ptr(out), out_stride, int(out_cap), _ptr(out_len), _ptr(status),
                                                   n, flags))
        return out, out_len, status

    # ---- batches in DEVICE memory (pointers or torch tensors; asynchronous on `stream`) ----
    def compress_batch(self, d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n,
                       stream=0):
        self._check(self._lib.hdlz_compress_batch(self._ctx, _ptr(d_in), in_stride, _ptr(d_in_len), uniform_len,
                                                  _ptr(d_out), out_stride, _ptr(d_out_len), _ptr(d_status), n,
                                                  stream or None))

    def decompress_batch(self, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len, d_status,
                         n, flags=0, stream=0):
        self._check(self._lib.hdlz_decompress_batch(self._ctx, _ptr(d_in), _ptr(d_in_off), in_stride, _ptr(d_in_len),
                                                    _ptr(d_out), out_stride, out_cap, _ptr(d_out_len), _ptr(d_status),
                                                    n, flags, stream or None))

    def generate_blocks(self, d_out, stride, length, n, seed=workload.DEFAULT_SEED, first_block=0, stream=0):
        self._check(self._lib.hdlz_generate_blocks(self._ctx, _ptr(d_out), stride, length, n, seed, first_block,
                                                   stream or None))

    def compress_stream(self):
        """A CompressStream on this engine."""
        return CompressStream(self)

    def decompress_stream(self, max_out=1 << 20, flags=0):
        """A DecompressStream on this engine."""
        return DecompressStream(self, max_out, flags)

    def sync(self, stream=0):
        self._check(self._lib.hdlz_stream_sync(self._ctx, stream or None))


class CompressStream(object):
    """One compress stream fed in pieces (hdlz_cstream_*): the port protocol's WRITE ... WRITE ... IDLE with the
    engine running alongside (deflate.py:459-461, 768).  feed() returns the stream bytes completed so far,
    finish() the rest; joined they equal Engine.compress() of the whole input."""

    def __init__(self, engine):
        self._eng = engine
        self._lib = engine._lib
        self._st = ctypes.c_void_p()
        engine._check(self._lib.hdlz_cstream_begin(engine._ctx, ctypes.byref(self._st)))
        self.in_progress = 0           # input position up to which the stream is encoded (o_iprogress)

    def feed(self, data):
        data = bytes(data)
        src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
        cap = compress_bound(len(data) + 2048)
        out = np.empty(cap, dtype=np.uint8)
        n, prog = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._eng._check(self._lib.hdlz_cstream_feed(self._st, src.ctypes.data, len(data), out.ctypes.data, cap,
                                                     ctypes.byref(n), ctypes.byref(prog)))
        self.in_progress = int(prog.value)
        return out[:n.value].tobytes()

    def finish(self):
        cap = compress_bound((1 << 20) + 8192)
        out = np.empty(cap, dtype=np.uint8)
        n, st = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._eng._check(self._lib.hdlz_cstream_finish(self._st, out.ctypes.data, cap, ctypes.byref(n), ctypes.byref(st)))
        if st.value:
            raise StreamError(st.value)
        return out[:n.value].tobytes()

    def close(self):
        if self._st.value:
            self._lib.hdlz_cstream_end(self._st)
            self._st = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DecompressStream(object):
    """One decompress stream fed in pieces (hdlz_dstream_*): the reference's decoder running while the host is
    still writing (deflate.py:1529) and the host reading as o_oprogress advances.  feed() returns the output bytes
    that became available (at most `room` of them: the rest waits on the device), finish() the rest; joined they
    equal Engine.decompress() of the whole stream."""

    def __init__(self, engine, max_out=1 << 20, flags=0):
        self._eng = engine
        self._lib = engine._lib
        self._st = ctypes.c_void_p()
        self.max_out = int(max_out)
        engine._check(self._lib.hdlz_dstream_begin(engine._ctx, self.max_out, int(flags), ctypes.byref(self._st)))
        self.in_progress = 0           # input byte position the decoder has reached (o_iprogress)

    def feed(self, data, room=None):
        data = bytes(data)
        src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
        cap = self.max_out if room is None else int(room)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n, prog = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._eng._check(self._lib.hdlz_dstream_feed(self._st, src.ctypes.data, len(data), out.ctypes.data, cap,
                                                     ctypes.byref(n), ctypes.byref(prog)))
        self.in_progress = int(prog.value)
        return out[:n.value].tobytes()

    def finish(self, room=None):
        """-> the remaining output (all of it unless `room` is given; then call again while .remaining)."""
        cap = self.max_out if room is None else int(room)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n, rem, st = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._eng._check(self._lib.hdlz_dstream_finish(self._st, out.ctypes.data, cap, ctypes.byref(n), ctypes.byref(rem),
                                                       ctypes.byref(st)))
        self.remaining = int(rem.value)
        if st.value:
            raise StreamError(st.value)
        return out[:n.value].tobytes()

    def close(self):
        if self._st.value:
            self._lib.hdlz_dstream_end(self._st)
            self._st = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_engine = None


def default_engine():
    """Process-wide engine on cuda:0 (what the drop-in `deflate` module uses)."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine
