"""Generate tests/golden/compress_golden_w256.json by EXECUTING the unmodified reference engine in its
non-FAST configuration (deflate.py:36-37, 56-59: FAST = False => CWINDOW = 256; SEARCH walks cur_search
down from di - 1, :996-1016, SEARCH10 extends, :1018-1062; distance codes up to 15 with the `outcarry` split,
:875-880), for both MATCH10 settings.  Build container only:
    python oracle/make_golden_window256.py
"""
import hashlib
import json
import os
import random
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hdl_deflate_b200  # noqa: E402,F401
from hdl_deflate_b200 import workload  # noqa: E402
from oracle import ref_sim, hdlz_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    t0 = time.time()
    rnd = random.Random(20261019)
    text = " ".join("   Hello World! %d     " % i for i in range(100)).encode()   # test_deflate.py:45-47
    duts = {}
    cases = []

    def add(name, data, match10=True, recipe=None):
        # the reference's switches are module globals read while the engine runs: set them for THIS run
        ref_sim.load_reference(fast=False, match10=match10)
        if match10 not in duts:
            duts[match10] = ref_sim.RefDut(fast=False, match10=match10)
        out, cycles = duts[match10].compress(data, max_cycles=400 * len(data) + 20000)
        st, oc = hdlz_oracle.compress(data, 256, 10 if match10 else 5)
        assert st == 0 and oc == out, "C restatement (window 256) differs from the reference on %s" % name
        assert zlib.decompress(out) == data
        c = {"name": name, "len": len(data), "match10": match10, "in_sha256": sha(data), "out_len": len(out),
             "out_sha256": sha(out), "cycles": cycles}
        if recipe is not None:
            c["recipe"] = recipe
        else:
            c["in_hex"] = data.hex()
        if len(out) <= 64:
            c["out_hex"] = out.hex()
        cases.append(c)
        print(name, len(data), len(out), cycles, "%.0fs" % (time.time() - t0), flush=True)

    for m10 in (True, False):
        tag = "" if m10 else "_m5"
        add("abcde" + tag, b"abcde", m10)
        add("a12" + tag, b"a" * 12, m10)
        add("abc6" + tag, b"abcabcabcabcabcabc", m10)
        add("text498" + tag, text[:498], m10)
        for n in (5, 6, 7, 33, 64, 255, 256, 257, 258, 259, 300, 513, 700):
            add("wl_len%d%s" % (n, tag), workload.block(3000 + n, n), m10, {"index": 3000 + n, "length": n, "seed": workload.DEFAULT_SEED})
        # matches at distances 33 .. 256: what the 32-byte window of the FAST engine cannot see
        for period in (40, 100, 193, 255, 256):
            base = bytes(rnd.randrange(256) for _ in range(period))
            add("period%d%s" % (period, tag), (base * 6)[:period * 3 + 17], m10)
        add("far_and_near" + tag, bytes(rnd.choice(b"abcdefgh") for _ in range(900)), m10)
        add("rand600" + tag, bytes(rnd.randrange(256) for _ in range(600)), m10)
        add("zeros600" + tag, bytes(600), m10)
    add("wl_blk0_2048", workload.block(0, 2048), True, {"index": 0, "length": 2048, "seed": workload.DEFAULT_SEED})
    add("text2048", (text * 2)[:2048], True)
    add("wl_blk1_2048_m5", workload.block(1, 2048), False, {"index": 1, "length": 2048, "seed": workload.DEFAULT_SEED})
    add("wl_multi2300", workload.block(78, 2300), True, {"index": 78, "length": 2300, "seed": workload.DEFAULT_SEED})
    with open(os.path.join(GOLD, "compress_golden_w256.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden_window256.py", "reference": "deflate.py FAST=False CWINDOW=256, both MATCH10 settings",
                   "cases": cases}, f, indent=0)
    print("compress cases (FAST=False):", len(cases), "%.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
