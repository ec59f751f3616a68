"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference engine on the CPU.

Loads `/root/reference/deflate.py` from where it lies (never copied) under the
repo's MyHDL-compat layer (hdl-deflate_b200/dropin/myhdl) and clocks its
`deflate()` block through the port protocol the way the reference's own test
bench does (test_deflate.py:92-288: WRITE@0 clear, STARTC/STARTD, stream bytes
with the `o_iprogress > i - CWINDOW` flow control, READ bytes while
`ri < o_oprogress`, stop on `o_done and o_oprogress == ri`).

This is the strongest oracle available: the reference's own implementation
executing.  It exists only in the build container (`/root/reference` is absent
on the GPU box), so it is used by `oracle/make_golden.py` to produce the
committed fixtures under tests/golden/ and by the CPU tests that cross-check
the C restatement (oracle/hdlz_oracle.c) — never by the product path.
"""

import contextlib
import importlib.util
import io
import math
import os
import sys

REFERENCE_DIR = os.environ.get("HDLZ_REFERENCE_DIR", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DROPIN = os.path.join(_REPO, "hdl-deflate_b200", "dropin")

_ref = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "deflate.py"))


def _compat_myhdl():
    """Import the compat `myhdl` (by path, so a foreign one is never picked up)."""
    if "myhdl" in sys.modules and getattr(sys.modules["myhdl"], "__file__", "").startswith(_DROPIN):
        return sys.modules["myhdl"]
    spec = importlib.util.spec_from_file_location(
        "myhdl", os.path.join(_DROPIN, "myhdl", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["myhdl"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference(ibsize=2048, obsize=32768, fast=True, match10=True, cwindow=None):
    """Import reference deflate.py (once) and set its size globals.

    The sizes are module globals read at elaboration time (deflate.py:61-89),
    so they can be overridden after import without touching the file
    (SURVEY.md §8(c)).  Output bytes do not depend on IBSIZE/OBSIZE.
    """
    global _ref
    _compat_myhdl()
    if _ref is None:
        spec = importlib.util.spec_from_file_location(
            "_hdlz_reference_deflate", os.path.join(REFERENCE_DIR, "deflate.py"))
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        _ref = mod
    m = _ref
    m.FAST = fast
    m.MATCH10 = match10
    m.CWINDOW = cwindow if cwindow is not None else (32 if fast else 256)
    m.IBSIZE = ibsize
    m.OBSIZE = obsize
    m.LIBSIZE = int(math.log2(ibsize))
    m.LOBSIZE = int(math.log2(obsize))
    m.LBSIZE = max(m.LIBSIZE, m.LOBSIZE)
    m.IBS = (1 << m.LIBSIZE) - 1
    m.OBS = (1 << m.LOBSIZE) - 1
    return m


class RefDut(object):
    """One instance of the reference engine plus a clock-level driver."""

    def __init__(self, **cfg):
        self.m = m = load_reference(**cfg)
        my = _compat_myhdl()
        S, intbv, modbv = my.Signal, my.intbv, my.modbv
        self.my = my
        self.i_mode = S(intbv(0)[3:])
        self.o_done = S(bool(0))
        self.i_data = S(intbv()[8:])
        self.o_byte = S(intbv()[8:])
        self.o_iprogress = S(intbv()[m.LMAX:])
        self.o_oprogress = S(intbv()[m.LMAX:])
        self.i_waddr = S(modbv()[m.LMAX:])
        self.i_raddr = S(modbv()[m.LMAX:])
        self.clk = S(bool(0))
        self.reset = my.ResetSignal(0, 1, True)
        with contextlib.redirect_stdout(io.StringIO()):
            self.dut = m.deflate(self.i_mode, self.o_done, self.i_data, self.o_iprogress,
                                 self.o_oprogress, self.o_byte, self.i_waddr, self.i_raddr,
                                 self.clk, self.reset)
        self.sim = my.Simulation(self.dut)
        self.sim._start()
        self.cycles = 0
        self._sink = io.StringIO()
        self._pulse_reset()

    def _clock(self):
        """One full clock period (rising edge first, as test_deflate.py:95-97,124-127)."""
        with contextlib.redirect_stdout(self._sink):
            self.clk.next = not self.clk
            self.sim._settle()
            self.clk.next = not self.clk
            self.sim._settle()
        self._sink.seek(0)
        self._sink.truncate()
        self.cycles += 1

    def _pulse_reset(self):
        self.reset.next = 1
        self._clock()
        self.reset.next = 0
        self._clock()

    def _run(self, start_mode, data, max_cycles):
        m = self.m
        data = bytes(data)
        self.i_mode.next = m.WRITE
        self.i_waddr.next = 0
        self.i_raddr.next = 0
        self._clock()
        self.i_mode.next = start_mode
        self._clock()
        i = ri = 0
        out = bytearray()
        start = self.cycles
        while True:
            did_read = False
            if ri < self.o_oprogress:
                did_read = True
                self.i_mode.next = m.READ
                self.i_raddr.next = ri
                self._clock()
                ri += 1
            if i < len(data):
                if self.o_iprogress > i - m.CWINDOW:
                    self.i_mode.next = m.WRITE
                    self.i_waddr.next = i
                    self.i_data.next = data[i]
                    i += 1
            else:
                self.i_mode.next = m.IDLE
            self._clock()
            if did_read:
                out.append(int(self.o_byte))
            if self.o_done and self.o_oprogress == ri:
                break
            if self.cycles - start > max_cycles:
                raise RuntimeError("reference engine did not finish in %d cycles" % max_cycles)
        self.i_mode.next = m.IDLE
        self._clock()
        return bytes(out), self.cycles - start

    def compress(self, data, max_cycles=None):
        """Reference compress of `data` (len >= 5, deflate.py:429). -> (bytes, cycles)."""
        if len(data) < 5:
            raise ValueError("the reference engine needs at least 5 input bytes (isize >= 4)")
        return self._run(self.m.STARTC, data, max_cycles or (40 * len(data) + 10000))

    def decompress(self, stream, max_cycles=None):
        """Reference decompress of a zlib stream. -> (bytes, cycles)."""
        return self._run(self.m.STARTD, stream, max_cycles or (400000 + 200 * len(stream)))


_shared = {}


def ref_compress(data, **cfg):
    key = tuple(sorted(cfg.items()))
    d = _shared.get(key)
    if d is None:
        d = _shared[key] = RefDut(**cfg)
    return d.compress(data)


def ref_decompress(stream, **cfg):
    key = tuple(sorted(cfg.items()))
    d = _shared.get(key)
    if d is None:
        d = _shared[key] = RefDut(**cfg)
    return d.decompress(stream)


if __name__ == "__main__":
    import hashlib
    import time
    t0 = time.time()
    for inp in (b"abcde", b"a" * 12, b"abcabcabcabcabcabc"):
        o, c = ref_compress(inp)
        print(inp, len(o), o.hex(), c)
    text = " ".join("   Hello World! %d     " % i for i in range(100)).encode()
    cyc = (text * 3)[:2048]
    for inp in (text[:50], text[:498], cyc, bytes(2048), bytes(range(256)) * 8):
        o, c = ref_compress(inp)
        print(len(inp), len(o), hashlib.sha256(o).hexdigest()[:16], c, "%.2f cyc/B" % (c / len(inp)))
    print("%.1fs" % (time.time() - t0))
