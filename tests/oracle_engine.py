"""Test double for the host-logic tests: an object with the Engine's stream methods whose
answers come from the CPU oracle.  Lives in tests/ only; the product has no such path."""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hdlz_oracle  # noqa: E402


class OracleStreamError(ValueError):
    pass


class OracleEngine(object):
    def __init__(self):
        self.jobs = []
        self.match10 = True
        self.fast = True

    def compress(self, data):
        st, out = hdlz_oracle.compress(data, cwindow=32 if self.fast else 256, maxlen=10 if self.match10 else 5)
        self.jobs.append(("C", len(data), len(out)))
        if st:
            raise OracleStreamError("status %d" % st)
        return out

    def decompress(self, data, max_out=None, flags=0):
        out = zlib.decompress(data)
        self.jobs.append(("D", len(data), len(out)))
        return out
