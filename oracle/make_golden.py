"""Generate tests/golden/*.json by EXECUTING the unmodified reference engine
(/root/reference/deflate.py under the MyHDL-compat layer, see oracle/ref_sim.py).

Run in the build container only (the reference is absent on the GPU box):
    python oracle/make_golden.py
The committed fixtures pin (a) the C restatement oracle, (b) the CUDA path, to what the
reference FSM itself produces.  Inputs are stored either inline (hex) or as a workload
recipe (hdl-deflate_b200/workload.py: block index, length, seed) plus their sha256.
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
import hdl_deflate_b200  # noqa: E402
from hdl_deflate_b200 import workload  # noqa: E402
from oracle import ref_sim, hdlz_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    os.makedirs(GOLD, exist_ok=True)
    t0 = time.time()
    rnd = random.Random(20261017)
    text = " ".join("   Hello World! %d     " % i for i in range(100)).encode()   # test_deflate.py:45-47
    cases = []

    def add(name, data, recipe=None):
        out, cycles = ref_sim.ref_compress(data)
        st, oc = hdlz_oracle.compress(data)
        assert st == 0 and oc == out, "C restatement differs from the reference on %s" % name
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

    # SURVEY.md 8(a) vectors
    add("abcde", b"abcde")
    add("a12", b"a" * 12)
    add("abc6", b"abcabcabcabcabcabc")
    add("text50", text[:50])
    add("text498", text[:498])
    add("text2048", (text * 2)[:2048])
    add("zeros2048", bytes(2048))
    add("ramp2048", bytes(range(256)) * 8)
    # edge lengths, workload recipe
    for n in list(range(5, 71)) + [127, 128, 129, 255, 256, 257, 1023, 1024, 1025, 2040, 2046, 2047]:
        add("wl_len%d" % n, workload.block(1000 + n, n), {"index": 1000 + n, "length": n, "seed": workload.DEFAULT_SEED})
    # config-2 blocks
    for i in range(48):
        add("wl_blk%d" % i, workload.block(i, 2048), {"index": i, "length": 2048, "seed": workload.DEFAULT_SEED})
    # two-symbol / run-heavy / multi-tile
    add("ab2048", b"ab" * 1024)
    add("abc2049", (b"abc" * 700)[:2049])
    add("bin0_1500", bytes(rnd.choice(b"01") for _ in range(1500)))
    add("rand2048", bytes(rnd.randrange(256) for _ in range(2048)))
    add("zeros5000", bytes(5000))
    add("text10000", (text * 5)[:10000])
    add("wl_multi4133", workload.block(77, 4133), {"index": 77, "length": 4133, "seed": workload.DEFAULT_SEED})
    add("runs3000", b"".join(bytes([rnd.randrange(256)]) * rnd.randrange(1, 40) for _ in range(200))[:3000])
    with open(os.path.join(GOLD, "compress_golden.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "reference": "deflate.py FAST=MATCH10=True CWINDOW=32",
                   "cases": cases}, f, indent=0)
    print("compress cases:", len(cases), "%.1fs" % (time.time() - t0))

    # decompress: zlib streams the reference engine itself inflates correctly (OBSIZE=32768)
    dcases = []

    def addd(name, plain, level, strategy, wbits=15):
        co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
        z = co.compress(plain) + co.flush()
        out, cycles = ref_sim.ref_decompress(z)
        assert out == plain, name
        st, oo = hdlz_oracle.inflate(z, len(plain), 3)
        assert st == 0 and oo == plain
        dcases.append({"name": name, "stream_hex": z.hex(), "out_len": len(plain), "out_sha256": sha(plain),
                       "cycles": cycles, "level": level, "strategy": strategy})

    addd("empty", b"", 6, 0)
    addd("text_fixed", text[:700], 6, zlib.Z_FIXED)
    addd("text_dynamic", text, 6, 0)
    addd("stored", bytes(rnd.randrange(256) for _ in range(600)), 0, 0)
    addd("rand_dynamic", bytes(rnd.choice(b"abcdefgh") for _ in range(3000)), 9, 0)
    addd("wl_fixed", workload.block(5, 2048), 6, zlib.Z_FIXED)
    addd("wl_dynamic", workload.block(6, 2048), 6, 0)
    with open(os.path.join(GOLD, "decompress_golden.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "cases": dcases}, f, indent=0)
    print("decompress cases:", len(dcases), "%.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
