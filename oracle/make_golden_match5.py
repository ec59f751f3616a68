"""Generate tests/golden/compress_golden_match5.json by EXECUTING the unmodified reference engine
with MATCH10 = False (deflate.py:34-35: matches of 3..5 bytes; SEARCHF :913-924), see
oracle/make_golden.py for the default configuration.  Build container only:
    python oracle/make_golden_match5.py
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
    rnd = random.Random(20261018)
    text = " ".join("   Hello World! %d     " % i for i in range(100)).encode()   # test_deflate.py:45-47
    cases = []

    def add(name, data, recipe=None):
        out, cycles = ref_sim.ref_compress(data, match10=False)
        st, oc = hdlz_oracle.compress(data, maxlen=5)
        assert st == 0 and oc == out, "C restatement (maxlen 5) differs from the reference on %s" % name
        assert zlib.decompress(out) == data
        c = {"name": name, "len": len(data), "in_sha256": sha(data), "out_len": len(out),
             "out_sha256": sha(out), "cycles": cycles}
        if recipe is not None:
            c["recipe"] = recipe
        else:
            c["in_hex"] = data.hex()
        if len(out) <= 64:
            c["out_hex"] = out.hex()
        cases.append(c)

    add("abcde", b"abcde")
    add("a12", b"a" * 12)
    add("abc6", b"abcabcabcabcabcabc")
    add("text50", text[:50])
    add("text498", text[:498])
    add("text2048", (text * 2)[:2048])
    add("zeros2048", bytes(2048))
    add("ramp2048", bytes(range(256)) * 8)
    for n in list(range(5, 40)) + [63, 64, 65, 1023, 1024, 1025, 2046, 2047]:
        add("wl_len%d" % n, workload.block(2000 + n, n), {"index": 2000 + n, "length": n, "seed": workload.DEFAULT_SEED})
    for i in range(12):
        add("wl_blk%d" % i, workload.block(i, 2048), {"index": i, "length": 2048, "seed": workload.DEFAULT_SEED})
    add("ab2048", b"ab" * 1024)
    add("rand2048", bytes(rnd.randrange(256) for _ in range(2048)))
    add("wl_multi4133", workload.block(77, 4133), {"index": 77, "length": 4133, "seed": workload.DEFAULT_SEED})
    add("runs3000", b"".join(bytes([rnd.randrange(256)]) * rnd.randrange(1, 40) for _ in range(200))[:3000])
    with open(os.path.join(GOLD, "compress_golden_match5.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden_match5.py", "reference": "deflate.py FAST=True MATCH10=False CWINDOW=32",
                   "cases": cases}, f, indent=0)
    print("compress cases (MATCH10=False):", len(cases), "%.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
