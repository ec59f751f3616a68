"""Drop-in `deflate` module: the reference's module interface on the B200 engine.

Same public names as tomtor/HDL-deflate's deflate.py — the mode codes and size
constants the test bench imports (test_deflate.py:11-13) and the block factory
`deflate(i_mode, o_done, i_data, o_iprogress, o_oprogress, o_byte, i_waddr, i_raddr,
clk, reset)` (deflate.py:219-221) — so `test_deflate.py` runs unchanged with this
directory on PYTHONPATH.  The clocked FSM is gone: a small host model keeps the port
protocol and hands each START job to the CUDA engine through the C ABI
(include/hdlz.h, hdlz_compress_stream / hdlz_decompress_stream).

Port protocol kept (one command per rising clock edge):
  WRITE  iram[i_waddr] = i_data; isize = i_waddr           (deflate.py:602-605)
  READ / any mode: o_byte <= oram[i_raddr], valid one clock later  (deflate.py:601)
  STARTC / STARTD  honoured only when idle: clear o_done and both progress
         counters (deflate.py:616-654)
  IDLE   after a START means "no more input" (deflate.py:768, 1529): the job runs on
         bytes 0..isize and then o_oprogress = length, o_done = 1
  while input is still arriving o_iprogress follows isize (compress: isize - 10, the
  reference's stall point deflate.py:768; decompress: isize - 4, deflate.py:1529), which
  is what the host's `o_iprogress > i - CWINDOW` flow control needs
  (test_deflate.py:159, 250); nothing happens while isize < 4 (deflate.py:429-432).
A non-zero stream status is raised as myhdl.Error with the reference's message.
"""
from math import log2

from myhdl import always, block, Error

IDLE, WRITE, READ, STARTC, STARTD = range(5)

# feature flags of the reference (deflate.py:20-41): the configuration this engine implements.
# MATCH10 and FAST may be set to False before a block is instantiated (as with the reference, where they
# are module globals read at elaboration): MATCH10 = False stops matches at 5 bytes (deflate.py:913-924),
# FAST = False selects the non-FAST engine with CWINDOW = 256 (deflate.py:56-59, 996-1062) — set CWINDOW
# = 256 with it, as the reference derives it.
LOWLUT = False
COMPRESS = True
DECOMPRESS = True
DYNAMIC = True
MATCH10 = True
FAST = True
ONEBLOCK = False

CWINDOW = 32      # search window (deflate.py:56-57)
OBSIZE = 32768    # "You need 32768 to decompress ALL valid deflate streams" (README.md:20-21)
IBSIZE = 2048     # README.md:23
LMAX = 24         # deflate.py:73-76

if OBSIZE > IBSIZE:
    LBSIZE = int(log2(OBSIZE))
else:
    LBSIZE = int(log2(IBSIZE))
LIBSIZE = int(log2(IBSIZE))
LOBSIZE = int(log2(OBSIZE))
IBS = (1 << LIBSIZE) - 1
OBS = (1 << LOBSIZE) - 1

_MAX_STREAM = 1 << LMAX

_backend = None


def _get_backend():
    """The engine that runs START jobs: hdl_deflate_b200.default_engine() (CUDA, no fallback)."""
    global _backend
    if _backend is None:
        import os
        import sys
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        if root not in sys.path:
            sys.path.insert(0, root)
        import hdl_deflate_b200
        _backend = hdl_deflate_b200.default_engine()
    return _backend


def set_backend(engine):
    """Use `engine` (anything with compress(bytes)->bytes and decompress(bytes)->bytes)."""
    global _backend
    _backend = engine


@block
def deflate(i_mode, o_done, i_data, o_iprogress, o_oprogress, o_byte,
            i_waddr, i_raddr, clk, reset):
    """ Deflate (de)compress — same ports as the reference block. """

    ring = bytearray(IBSIZE)       # mirror of iram: what a byte address holds if never rewritten
    lin = bytearray()              # bytes by full address since the last write at address 0
    st = {"isize": 0, "job": IDLE, "out": b"", "stream": None, "fed": 0}
    match10 = bool(MATCH10)        # read at elaboration, like every configuration global of the reference
    fast = bool(FAST)

    def feed_stream():
        """STARTC with input still arriving (the UnitTest flow, test_deflate.py:216-260): every IBSIZE bytes the
        host has written in order go to the engine's stream (hdlz_cstream_feed) while it keeps writing — the
        reference consumes its iram the same way (deflate.py:459-461, 768) — and the stream bytes completed so
        far become readable: o_oprogress follows the engine's real output, as it does in the reference."""
        eng = _get_backend()
        if st["stream"] is None:
            if not fast or not hasattr(eng, "compress_stream") or getattr(eng, "container", 0) == 2 or \
                    getattr(eng, "tree", None) is not None:
                return                     # these jobs run in one piece when the host goes IDLE
            if hasattr(eng, "match10"):
                eng.match10 = match10
            if hasattr(eng, "fast"):
                eng.fast = True
            st["stream"] = eng.compress_stream()
            st["out"] = b""
        piece = bytes(lin[st["fed"]:st["fed"] + IBSIZE])
        st["fed"] += IBSIZE
        st["out"] = st["out"] + st["stream"].feed(piece)

    def feed_dstream():
        """STARTD with input still arriving (test_deflate.py:136-169): the same for the decoder (hdlz_dstream_feed) —
        the reference inflates as far as the bytes received allow and waits at `di >= isize - 4` (deflate.py:1529);
        what it has produced is readable through o_oprogress."""
        eng = _get_backend()
        if st["stream"] is None:
            if not hasattr(eng, "decompress_stream"):
                return
            st["stream"] = eng.decompress_stream((1 << LMAX) - 1)
            st["out"] = b""
        piece = bytes(lin[st["fed"]:st["fed"] + IBSIZE])
        st["fed"] += IBSIZE
        st["out"] = st["out"] + st["stream"].feed(piece)

    def run_job():
        data = bytes(lin[:st["isize"] + 1])
        eng = _get_backend()
        try:
            if st["stream"] is not None:
                stream, st["stream"] = st["stream"], None
                try:
                    return st["out"] + stream.feed(data[st["fed"]:]) + stream.finish()
                finally:
                    stream.close()
            if st["job"] == STARTC:
                if hasattr(eng, "match10"):
                    eng.match10 = match10
                if hasattr(eng, "fast"):
                    eng.fast = fast
                return eng.compress(data)
            return eng.decompress(data)
        except ValueError as e:            # StreamError: non-zero hdlz_status
            raise Error(str(e).split(" (")[0])

    @always(clk.posedge)
    def port():
        # io_logic (deflate.py:599-605)
        ra = int(i_raddr)
        out = st["out"]
        o_byte.next = out[ra] if ra < len(out) else 0
        mode = int(i_mode)
        if mode == WRITE:
            a = int(i_waddr)
            if a == 0:
                del lin[:]
            if a > len(lin):               # gap: those addresses keep the ring's stale contents
                for q in range(len(lin), a):
                    lin.append(ring[q & IBS])
            b = int(i_data)
            if a < len(lin):
                lin[a] = b
            else:
                lin.append(b)
            ring[a & IBS] = b
            st["isize"] = a

        # logic (deflate.py:607-654)
        if reset:
            st["job"] = IDLE
            o_done.next = False
        elif st["job"] == IDLE:
            if mode == STARTC or mode == STARTD:
                st["job"] = mode
                st["fed"] = 0
                if st["stream"] is not None:
                    st["stream"].close()
                    st["stream"] = None
                o_done.next = False
                o_iprogress.next = 0
                o_oprogress.next = 0
        else:
            isize = st["isize"]
            if isize >= 4 and len(lin) > isize:
                if mode == IDLE:
                    res = run_job()
                    st["out"] = res
                    st["job"] = IDLE
                    o_iprogress.next = isize
                    o_oprogress.next = len(res)
                    o_done.next = True
                else:
                    slack = 10 if st["job"] == STARTC else 4
                    o_iprogress.next = isize - slack if isize > slack else 0
                    # bytes written in order since the START, a whole IBSIZE of them beyond what the engine has
                    # (plus its 32-byte look-ahead): hand them over while the host keeps writing
                    if mode == WRITE and int(i_waddr) == isize and isize + 1 - st["fed"] >= IBSIZE + 2 * CWINDOW:
                        try:
                            if st["job"] == STARTC:
                                feed_stream()
                            else:
                                feed_dstream()
                        except ValueError as e:
                            raise Error(str(e).split(" (")[0])
                        if st["stream"] is not None:
                            o_oprogress.next = len(st["out"])

    return port
