"""Clock-level driver for a `deflate()` block, written after the reference's own test bench
(test_deflate.py:92-288 streaming flow, :421-452 / :513-545 preload flow).  Works for the drop-in
module of this repo; oracle/ref_sim.py holds the twin that drives the reference engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "hdl-deflate_b200", "dropin")


def import_dropin():
    """-> (myhdl compat module, drop-in deflate module), imported the way a user would:
    hdl-deflate_b200/dropin on sys.path, `import myhdl`, `import deflate`."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import myhdl
    import deflate
    assert os.path.dirname(os.path.abspath(deflate.__file__)) == DROPIN, deflate.__file__
    return myhdl, deflate


class Port(object):
    def __init__(self, backend=None):
        self.my, self.m = import_dropin()
        if backend is not None:
            self.m.set_backend(backend)
        my, m = self.my, self.m
        S, intbv, modbv = my.Signal, my.intbv, my.modbv
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
        self.dut = m.deflate(self.i_mode, self.o_done, self.i_data, self.o_iprogress, self.o_oprogress,
                             self.o_byte, self.i_waddr, self.i_raddr, self.clk, self.reset)
        self.sim = my.Simulation(self.dut)
        self.sim._start()
        self.cycles = 0

    def clock(self):
        self.clk.next = not self.clk
        self.sim._settle()
        self.clk.next = not self.clk
        self.sim._settle()
        self.cycles += 1

    def pulse_reset(self):
        self.reset.next = 1
        self.clock()
        self.reset.next = 0
        self.clock()

    def stream(self, start_mode, data, max_cycles=None):
        """test_deflate.py streaming flow: clear, START, interleave READ / flow-controlled WRITE."""
        m = self.m
        data = bytes(data)
        max_cycles = max_cycles or (40 * len(data) + 400000)
        self.i_mode.next = m.WRITE
        self.i_waddr.next = 0
        self.i_raddr.next = 0
        self.clock()
        self.i_mode.next = start_mode
        self.clock()
        i = ri = 0
        out = bytearray()
        waits = 0
        start = self.cycles
        while True:
            did_read = False
            if ri < self.o_oprogress:
                did_read = True
                self.i_mode.next = m.READ
                self.i_raddr.next = ri
                self.clock()
                ri += 1
            if len(data) < 4 and i == 0 and start_mode == m.STARTC:
                self.i_mode.next = m.WRITE          # "SHORT INPUT" quirk, test_deflate.py:239-248
                self.i_waddr.next = 4
                self.i_data.next = 0
                i = 1
            elif i < len(data) and not (len(data) < 4 and start_mode == m.STARTC):
                if self.o_iprogress > i - m.CWINDOW:
                    self.i_mode.next = m.WRITE
                    self.i_waddr.next = i
                    self.i_data.next = data[i]
                    i += 1
                else:
                    waits += 1
            else:
                self.i_mode.next = m.IDLE
            self.clock()
            if did_read:
                out.append(int(self.o_byte))
            if self.o_done and self.o_oprogress == ri:
                break
            assert self.cycles - start < max_cycles, "engine did not finish (deadlock?)"
        self.i_mode.next = m.IDLE
        self.clock()
        return bytes(out), waits

    def preload(self, start_mode, data, max_wait=1000):
        """test_deflate_bench flow: write everything, IDLE, START, IDLE until o_done, pipelined READ."""
        m = self.m
        for a, b in enumerate(bytes(data)):
            self.i_mode.next = m.WRITE
            self.i_waddr.next = a
            self.i_data.next = b
            self.clock()
        self.i_mode.next = m.IDLE
        self.clock()
        self.i_mode.next = start_mode
        self.clock()
        self.i_mode.next = m.IDLE
        n = 0
        while not self.o_done:
            self.clock()
            n += 1
            assert n < max_wait, "o_done never rose"
        total = int(self.o_oprogress)
        out = bytearray()
        self.i_mode.next = m.READ
        for a in range(total):
            self.i_raddr.next = a
            self.clock()                 # o_byte is registered: valid after this edge (deflate.py:601)
            out.append(int(self.o_byte))
        self.i_mode.next = m.IDLE
        self.clock()
        return bytes(out)
