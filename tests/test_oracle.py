"""The oracle itself: C restatement vs the committed golden vectors (produced by the
executing reference), vs the live reference where it exists, and inflate vs zlib."""
import hashlib
import random
import zlib

import pytest

from conftest import load_golden, golden_input
from oracle import hdlz_oracle, ref_sim


def test_compress_restatement_matches_golden():
    cases = load_golden("compress_golden.json")
    assert len(cases) >= 100
    for c in cases:
        data = golden_input(c)
        st, out = hdlz_oracle.compress(data)
        assert st == 0
        assert len(out) == c["out_len"], c["name"]
        assert hashlib.sha256(out).hexdigest() == c["out_sha256"], c["name"]
        if "out_hex" in c:
            assert out.hex() == c["out_hex"]
        assert zlib.decompress(out) == data          # the reference's own check (test_deflate.py:285)


def test_compress_restatement_matches_golden_match5():
    """MATCH10 = False (deflate.py:34-35, 913-924): fixtures of oracle/make_golden_match5.py."""
    cases = load_golden("compress_golden_match5.json")
    assert len(cases) >= 60
    differs = 0
    for c in cases:
        data = golden_input(c)
        st, out = hdlz_oracle.compress(data, maxlen=5)
        assert st == 0
        assert len(out) == c["out_len"], c["name"]
        assert hashlib.sha256(out).hexdigest() == c["out_sha256"], c["name"]
        assert zlib.decompress(out) == data
        differs += out != hdlz_oracle.compress(data)[1]
    assert differs > 10            # the switch changes the streams


def test_compress_restatement_matches_golden_window256():
    """FAST = False (deflate.py:36-37, 56-59: CWINDOW = 256, SEARCH / SEARCH10, distance codes up to 15 with the
    `outcarry` split): fixtures of oracle/make_golden_window256.py, both MATCH10 settings."""
    cases = load_golden("compress_golden_w256.json")
    assert len(cases) >= 50
    far = 0
    for c in cases:
        data = golden_input(c)
        st, out = hdlz_oracle.compress(data, cwindow=256, maxlen=10 if c["match10"] else 5)
        assert st == 0
        assert len(out) == c["out_len"], c["name"]
        assert hashlib.sha256(out).hexdigest() == c["out_sha256"], c["name"]
        assert zlib.decompress(out) == data
        far += len(out) < len(hdlz_oracle.compress(data, maxlen=10 if c["match10"] else 5)[1])
    assert far > 10                # the wider window finds matches the FAST engine cannot


def test_tuned_cpu_arm_equals_the_restatement():
    """hdlz_oracle_compress_fast (the CPU baseline bench.py times) must produce the restatement's bytes:
    golden vectors of both MATCH10 settings plus seeded fuzz around the 32-byte window edge."""
    from hdl_deflate_b200 import workload
    for name, ml in (("compress_golden.json", 10), ("compress_golden_match5.json", 5)):
        for c in load_golden(name):
            data = golden_input(c)
            st, out = hdlz_oracle.compress_fast(data, ml)
            assert st == 0 and hashlib.sha256(out).hexdigest() == c["out_sha256"], c["name"]
    rnd = random.Random(5)
    for t in range(1500):
        n = rnd.choice([5, 6, 31, 32, 33, 34, 35, 36, 64, 65, 300, 2047, 2048, 4100])
        d = [bytes(rnd.randrange(256) for _ in range(n)), bytes(rnd.choice(b"ab") for _ in range(n)),
             workload.block(t, n), bytes([rnd.randrange(256)]) * n][t % 4]
        for ml in (10, 5):
            assert hdlz_oracle.compress_fast(d, ml) == hdlz_oracle.compress(d, 32, ml), (t, n, ml)
    for n in range(5):
        assert hdlz_oracle.compress_fast(bytes(n)) == (1, b"")


def test_survey_known_answers():
    # vectors produced by the reference FSM
    assert hdlz_oracle.compress(b"abcde")[1].hex() == "789c4b4c4a4e49050005c801f0"
    assert hdlz_oracle.compress(b"a" * 12)[1].hex() == "789c4b8483c444001d9a048d"
    assert hdlz_oracle.compress(b"abcabcabcabcabcabc")[1].hex() == "789c4b4c4a4620204a4a0600417c06e5"
    assert len(hdlz_oracle.compress(bytes(2048))[1]) == 318
    assert len(hdlz_oracle.compress(bytes(range(256)) * 8)[1]) == 2168


def test_short_input_status():
    for n in range(5):
        st, out = hdlz_oracle.compress(bytes(n))
        assert st == 1 and out == b""


def test_worst_case_size():
    data = bytes([200]) + bytes(random.Random(1).randrange(144, 256) for _ in range(2047))
    st, out = hdlz_oracle.compress(data)
    assert st == 0 and len(out) <= 2312


@pytest.mark.skipif(not ref_sim.available(), reason="reference sources not present (GPU box)")
def test_restatement_matches_live_reference_fuzz():
    rnd = random.Random(99)
    for t in range(24):
        n = rnd.choice([5, 6, 9, 17, 33, 64, 100, 257, 700])
        kind = t % 4
        if kind == 0:
            data = bytes(rnd.randrange(256) for _ in range(n))
        elif kind == 1:
            data = bytes(rnd.choice(b"ab") for _ in range(n))
        elif kind == 2:
            data = bytes(rnd.choice(b"abcd") for _ in range(n))
        else:
            data = (bytes(rnd.randrange(256) for _ in range(7)) * (n // 7 + 1))[:n]
        ref, _ = ref_sim.ref_compress(data)      # one shared DUT: stale state carries over
        st, out = hdlz_oracle.compress(data)
        assert st == 0 and out == ref, (t, n, kind)


@pytest.mark.skipif(not ref_sim.available(), reason="reference sources not present (GPU box)")
def test_restatement_matches_live_reference_fuzz_match5():
    """The same against the reference built with MATCH10 = False (deflate.py:34-35, 913-924)."""
    rnd = random.Random(7)
    for t in range(12):
        n = rnd.choice([5, 8, 21, 64, 130, 400])
        kind = t % 3
        if kind == 0:
            data = bytes(rnd.choice(b"ab") for _ in range(n))
        elif kind == 1:
            data = bytes(rnd.choice(b"abcdefgh") for _ in range(n))
        else:
            data = (bytes(rnd.randrange(256) for _ in range(5)) * (n // 5 + 1))[:n]
        ref, _ = ref_sim.ref_compress(data, match10=False)
        st, out = hdlz_oracle.compress(data, maxlen=5)
        assert st == 0 and out == ref, (t, n, kind)


def test_inflate_restatement_matches_zlib():
    rnd = random.Random(5)
    text = " ".join("   Hello World! %d     " % i for i in range(300)).encode()
    for t in range(60):
        n = rnd.randrange(0, 40000)
        kind = t % 4
        if kind == 0:
            data = bytes(rnd.randrange(256) for _ in range(n))
        elif kind == 1:
            data = bytes(rnd.choice(b"abc ") for _ in range(n))
        elif kind == 2:
            data = (text * (n // len(text) + 1))[:n]
        else:
            data = bytes(rnd.randrange(256) for _ in range(n // 50 + 1)) * 50
        for lvl, strat in ((6, 0), (6, zlib.Z_FIXED), (0, 0), (9, zlib.Z_FILTERED)):
            co = zlib.compressobj(lvl, zlib.DEFLATED, 15, 8, strat)
            z = co.compress(data) + co.flush()
            st, out = hdlz_oracle.inflate(z, len(data), flags=3)
            assert st == 0 and out == data


def test_inflate_golden_and_errors():
    for c in load_golden("decompress_golden.json"):
        z = bytes.fromhex(c["stream_hex"])
        st, out = hdlz_oracle.inflate(z, c["out_len"], flags=3)
        assert st == 0 and hashlib.sha256(out).hexdigest() == c["out_sha256"]
    z = zlib.compress(b"hello hello hello hello", 6)
    assert hdlz_oracle.inflate(z[:-5], 100)[0] == 5                      # truncated
    assert hdlz_oracle.inflate(z[:2] + bytes([z[2] | 6]) + z[3:], 100)[0] == 2   # BTYPE=3
    assert hdlz_oracle.inflate(z, 5)[0] == 6                             # overflow
    bad = bytearray(z); bad[-1] ^= 1
    assert hdlz_oracle.inflate(bytes(bad), 100, flags=2)[0] == 9         # adler
    assert hdlz_oracle.inflate(b"\x79\x9c" + z[2:], 100, flags=1)[0] == 8


def test_kernel_decomposition_model():
    """tests/kernel_model.py restates the CUDA kernel's phases lane by lane; it must agree
    with the oracle (design check for hdlz_compress.cu)."""
    import kernel_model
    from hdl_deflate_b200 import workload
    rnd = random.Random(3)
    cases = [b"abcde", bytes(2048), b"ab" * 1024, bytes(range(256)) * 8]
    cases += [workload.block(i, n) for i, n in enumerate([5, 6, 31, 32, 33, 63, 64, 65, 66, 67, 200, 2047, 2048, 2049,
                                                          2080, 2081, 4100])]
    cases += [bytes(rnd.choice(b"ab") for _ in range(n)) for n in (70, 2100)]
    cases.append(workload.block(880395, 2048))          # a lane emits 289 bits (> 32 * 9): needs 10 private words
    for d in cases:
        assert kernel_model.compress(d) == hdlz_oracle.compress(d)[1], len(d)
