"""Host logic of the drop-in `deflate` module (port protocol), exercised on the CPU with the
oracle test double answering the START jobs.  The GPU twin is tests/test_gpu_port.py."""
import os
import random
import subprocess
import sys
import zlib

import pytest

from oracle_engine import OracleEngine
from port_driver import Port, import_dropin
from oracle import hdlz_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TEST = "/root/reference/test_deflate.py"


def reference_style_data(mode, tlen, rnd):
    """Same six data modes as test_deflate.py:38-66, seeded."""
    if mode == 0:
        return " ".join("Hello World! 1 " for _ in range(tlen)).encode()
    if mode == 1:
        return " ".join("   Hello World! %d     " % i for i in range(tlen)).encode()
    if mode == 2:
        return " ".join("Hi: %d " % rnd.randrange(0, 0x1000) for _ in range(tlen)).encode()
    if mode == 3:
        return bytes(rnd.randrange(256) for _ in range(tlen))
    if mode == 4:
        return "".join(str(rnd.randrange(2)) for _ in range(tlen)).encode()
    return b""


def test_module_surface():
    my, m = import_dropin()
    for name in ("IDLE", "WRITE", "READ", "STARTC", "STARTD", "LBSIZE", "IBSIZE", "CWINDOW", "COMPRESS",
                 "DECOMPRESS", "OBSIZE", "LMAX", "LIBSIZE", "DYNAMIC", "LOBSIZE", "LOWLUT", "deflate"):
        assert hasattr(m, name), name                       # test_deflate.py:11-13, 21
    assert (m.IDLE, m.WRITE, m.READ, m.STARTC, m.STARTD) == (0, 1, 2, 3, 4)   # deflate.py:18
    assert (m.CWINDOW, m.IBSIZE, m.OBSIZE, m.LMAX) == (32, 2048, 32768, 24)
    assert (m.LIBSIZE, m.LOBSIZE, m.LBSIZE) == (11, 15, 15) and 9 <= m.LOBSIZE <= 15
    assert m.COMPRESS and m.DECOMPRESS and m.DYNAMIC and not m.LOWLUT


@pytest.mark.parametrize("mode", range(6))
def test_streaming_round_trip(mode):
    """test_deflate.py testMain: streaming decompress of a zlib stream, then streaming compress of
    10000 bytes taken cyclically, on the same DUT."""
    rnd = random.Random(mode)
    be = OracleEngine()
    p = Port(be)
    if mode == 0:
        p.pulse_reset()
    b_data = reference_style_data(mode, 600, rnd)
    zl = zlib.compress(b_data, 6)
    out, _ = p.stream(p.m.STARTD, zl)
    assert out == b_data
    slen = 10000
    src = bytes(b_data[i % len(b_data)] for i in range(slen)) if b_data else b""
    res, _ = p.stream(p.m.STARTC, src)
    rlen = min(len(b_data), slen)
    assert zlib.decompress(res)[:rlen] == b_data[:rlen]          # test_deflate.py:285
    if b_data:
        assert res == hdlz_oracle.compress(src)[1]
        assert be.jobs[-1] == ("C", slen, len(res))
    else:
        assert be.jobs[-1][1] == 5                               # short-input quirk: L = 5


def test_preload_flow_and_read_latency():
    """test_deflate_bench: preload, START, IDLE, wait for o_done, READ with one clock of latency;
    then feed the compressed bytes back (CRESULT -> VDECOMPRESS)."""
    p = Port(OracleEngine())
    p.pulse_reset()
    data = " ".join("   Hello World! %d     " % i for i in range(100)).encode()[:2034]
    comp = p.preload(p.m.STARTC, data)
    assert comp == hdlz_oracle.compress(data)[1]
    back = p.preload(p.m.STARTD, comp)
    assert back == data


def test_match10_switch_is_read_at_elaboration():
    """MATCH10 is a module global of the reference (deflate.py:34-35) read when the block is built."""
    from port_driver import import_dropin
    _, m = import_dropin()
    data = " ".join("   Hello World! %d     " % i for i in range(100)).encode()[:1000]
    m.MATCH10 = False
    try:
        p = Port(OracleEngine())
        p.pulse_reset()
        comp5 = p.preload(p.m.STARTC, data)
    finally:
        m.MATCH10 = True
    p = Port(OracleEngine())
    p.pulse_reset()
    comp10 = p.preload(p.m.STARTC, data)
    assert comp5 == hdlz_oracle.compress(data, maxlen=5)[1]
    assert comp10 == hdlz_oracle.compress(data)[1]
    assert comp5 != comp10 and zlib.decompress(comp5) == data


def test_start_ignored_while_busy_and_below_four_bytes():
    p = Port(OracleEngine())
    m = p.m
    p.pulse_reset()
    p.i_mode.next = m.WRITE; p.i_waddr.next = 0; p.i_data.next = 65; p.clock()
    p.i_mode.next = m.STARTC; p.clock()
    p.i_mode.next = m.IDLE
    for _ in range(20):
        p.clock()
    assert not p.o_done                       # isize < 4: the engine idles (deflate.py:429-432)
    for a in range(1, 8):
        p.i_mode.next = m.WRITE; p.i_waddr.next = a; p.i_data.next = 65 + a; p.clock()
    p.i_mode.next = m.STARTD; p.clock()       # ignored: a job is active (deflate.py:616-654)
    assert not p.o_done
    p.i_mode.next = m.IDLE; p.clock()
    assert p.o_done and p.o_oprogress == len(hdlz_oracle.compress(b"ABCDEFGH")[1])
    p.reset.next = 1; p.clock(); p.reset.next = 0; p.clock()
    assert not p.o_done                       # reset clears o_done (deflate.py:609-612)


def test_progress_keeps_host_writing():
    """o_iprogress must advance while input is still arriving or the reference's host loop
    (`o_iprogress > i - CWINDOW`, test_deflate.py:250) dead-locks."""
    p = Port(OracleEngine())
    data = bytes(random.Random(1).randrange(256) for _ in range(3000))
    out, waits = p.stream(p.m.STARTC, data)
    assert out == hdlz_oracle.compress(data)[1]
    assert waits <= 2 * len(data)


def test_bad_stream_raises_reference_error():
    my, m = import_dropin()

    class Bad(object):
        def decompress(self, data, **kw):
            import hdl_deflate_b200
            raise hdl_deflate_b200.StreamError(2)
    p = Port(Bad())
    with pytest.raises(my.Error, match="Bad method"):
        p.stream(m.STARTD, b"\x78\x9c\x07\x00\x00\x00\x00\x01")


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="reference sources not present (GPU box)")
def test_unchanged_reference_unittest_drives_the_host_model():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_unittest.py"), "--oracle"],
                       capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert "Ran 1 test" in tail and "OK" in tail, tail


def test_fast_switch_is_read_at_elaboration():
    """FAST is a module global of the reference (deflate.py:36-37) read when the block is built: False selects
    the 256-byte window of the non-FAST engine."""
    from port_driver import import_dropin
    _, m = import_dropin()
    base = bytes(range(7, 107))
    data = (base * 5)[:450]                 # repeats at distance 100: invisible to the 32-byte window
    m.FAST = False
    try:
        p = Port(OracleEngine())
        p.pulse_reset()
        slow = p.preload(p.m.STARTC, data)
    finally:
        m.FAST = True
    p = Port(OracleEngine())
    p.pulse_reset()
    fast = p.preload(p.m.STARTC, data)
    assert slow == hdlz_oracle.compress(data, cwindow=256)[1]
    assert fast == hdlz_oracle.compress(data)[1]
    assert len(slow) < len(fast) and zlib.decompress(slow) == data
