"""Parity of the CUDA inflater (through the C ABI) with zlib — the reference's own decompress
check (test_deflate.py:194) — on fixed, dynamic, stored and multi-block streams, plus the error
statuses that stand in for the reference's `raise Error(...)` sites."""
import hashlib
import random
import zlib

import numpy as np
import pytest

from conftest import load_golden
import hdl_deflate_b200 as hz
from hdl_deflate_b200 import workload

pytestmark = pytest.mark.gpu


def zl(data, level=6, strategy=0, wbits=15):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    return co.compress(data) + co.flush()


def test_golden_streams(engine):
    for c in load_golden("decompress_golden.json"):
        out = engine.decompress(bytes.fromhex(c["stream_hex"]), flags=3)
        assert len(out) == c["out_len"] and hashlib.sha256(out).hexdigest() == c["out_sha256"], c["name"]


def test_reference_test_modes(engine):
    """The six data modes of test_deflate.py:38-66 at the reference's tlen=2500, zlib level 6, wbits=LOBSIZE."""
    rnd = random.Random(11)
    modes = [" ".join("Hello World! 1 " for _ in range(2500)).encode(),
             " ".join("   Hello World! %d     " % i for i in range(2500)).encode(),
             " ".join("Hi: %d " % rnd.randrange(0x1000) for _ in range(2500)).encode(),
             bytes(rnd.randrange(256) for _ in range(2500)),
             "".join(str(rnd.randrange(2)) for _ in range(2500)).encode(),
             b""]
    for b_data in modes:
        assert engine.decompress(zl(b_data), flags=3) == b_data


def test_mixed_batch_matches_zlib(engine):
    rnd = random.Random(3)
    text = " ".join("   Hello World! %d     " % i for i in range(3000)).encode()
    plains = []
    for t in range(96):
        n = rnd.choice([0, 1, 2, 3, 100, 2048, 5000, 32768, 40000])
        kind = t % 5
        if kind == 0:
            d = bytes(rnd.randrange(256) for _ in range(n))
        elif kind == 1:
            d = bytes(rnd.choice(b"abc ") for _ in range(n))
        elif kind == 2:
            d = text[:n]
        elif kind == 3:
            d = workload.block(t, max(n, 1))[:n]
        else:
            d = bytes([rnd.randrange(256)]) * n
        plains.append(d)
    streams = []
    for i, d in enumerate(plains):
        lvl, strat = [(6, 0), (6, zlib.Z_FIXED), (0, 0), (9, 0), (1, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)][i % 6]
        streams.append(zl(d, lvl, strat))
    # packed layout with an offset array (arbitrary, unaligned offsets)
    off = np.zeros(len(streams), dtype=np.uint64)
    pos = 0
    for i, s in enumerate(streams):
        off[i] = pos
        pos += len(s) + (i % 3)
    buf = np.zeros(pos + 16, dtype=np.uint8)
    for i, s in enumerate(streams):
        buf[int(off[i]):int(off[i]) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    for route in (0, FORCE_LANES, FORCE_LANES | NO_LANE_SCRATCH):
        out, out_len, status = engine.decompress_host(buf, lens, 40000, in_off=off, flags=3 | route)
        assert not status.any(), (route, status)
        for i, d in enumerate(plains):
            assert out[i, :out_len[i]].tobytes() == d, (route, i)


def test_multi_block_and_window(engine):
    """Several deflate blocks per stream (Z_FULL_FLUSH), distances up to 32 KiB."""
    rnd = random.Random(8)
    base = bytes(rnd.randrange(256) for _ in range(32768))
    data = base + bytes(rnd.choice(b"xyz") for _ in range(1000)) + base[:20000] + base[100:5000]
    co = zlib.compressobj(9, zlib.DEFLATED, 15, 9)
    z = co.compress(data[:30000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(data[30000:60000]) + \
        co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[60000:]) + co.flush()
    assert engine.decompress(z, flags=3) == data


FORCE_GENERAL, FORCE_LANES, NO_LANE_SCRATCH, NO_SPLIT = 0x100, 0x200, 0x400, 0x1000   # internal routing flags (csrc/hdlz_common.cuh)


@pytest.mark.parametrize("route", [0, FORCE_GENERAL, FORCE_LANES])
def test_config3_zfixed_blocks(engine, route):
    """BASELINE config 3 shape: 2 KiB blocks as zlib Z_FIXED streams, packed + offsets (unaligned) and
    4-byte aligned offsets; through the lane-per-stream kernel, the warp-per-stream kernel, and the default."""
    n = 2048
    blocks = workload.blocks(900, n, 2048)
    streams = [zl(b, 6, zlib.Z_FIXED) for b in blocks]
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    for align in (1, 4):
        padded = [s + bytes((-len(s)) % align) for s in streams]
        off = np.cumsum([0] + [len(s) for s in padded[:-1]]).astype(np.uint64)
        buf = np.frombuffer(b"".join(padded) + bytes(16), dtype=np.uint8)
        out, out_len, status = engine.decompress_host(buf, lens, 2048, in_off=off, flags=3 | route)
        assert not status.any() and (out_len == 2048).all()
        assert out.tobytes() == b"".join(blocks)


@pytest.mark.parametrize("route", [FORCE_GENERAL, FORCE_LANES, FORCE_LANES | NO_SPLIT, FORCE_LANES | NO_LANE_SCRATCH])
def test_routes_agree_on_mixed_and_corrupt_streams(engine, route):
    """Fixed, stored, dynamic (handed over by the lane kernel), long-distance and corrupted streams:
    both routes must give zlib's bytes or an error status, never differ on valid streams."""
    rnd = random.Random(77 + route)
    text = " ".join("   Hello World! %d     " % i for i in range(400)).encode()
    plains, streams = [], []
    for t in range(1200):
        n = rnd.choice([0, 1, 5, 64, 700, 2048, 6000])
        kind = t % 4
        d = [bytes(rnd.randrange(256) for _ in range(n)), text[:n], workload.block(t, max(n, 1))[:n],
             bytes(rnd.choice(b"ab") for _ in range(n))][kind]
        lvl, strat = [(6, zlib.Z_FIXED), (0, 0), (6, 0), (1, zlib.Z_FIXED), (9, zlib.Z_FIXED)][t % 5]
        z = bytearray(zl(d, lvl, strat))
        if t % 7 == 0 and len(z) > 8:
            z[rnd.randrange(2, len(z))] ^= 1 << rnd.randrange(8)
        plains.append(d)
        streams.append(bytes(z))
    stride = (max(len(s) for s in streams) + 15) & ~15
    buf = np.zeros((len(streams), stride), dtype=np.uint8)
    for i, s in enumerate(streams):
        buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    out, out_len, status = engine.decompress_host(buf, lens, 6000, flags=3 | route)
    for i, s in enumerate(streams):
        try:
            want = zlib.decompress(s)
        except zlib.error:
            want = None
        if status[i] == 0:
            assert want is not None and out[i, :out_len[i]].tobytes() == want, i
        else:
            assert want is None or len(want) > 6000, (i, status[i])


@pytest.mark.parametrize("route,n", [(0, 64), (FORCE_LANES, 64), (0, 1500), (NO_SPLIT, 1500), (NO_SPLIT | hz.F_PERSIST_TABLES, 1500)])
def test_config4_dynamic_32k(engine, route, n):
    """BASELINE config 4 shape: 32 KiB plain, zlib level 6 dynamic trees, OBSIZE = 32768; through the
    warp-per-stream kernel (few streams) and the lane-per-stream kernel with per-lane tables (many)."""
    rnd = np.random.default_rng(4)
    plains, streams = [], []
    for i in range(n):
        sym = rnd.zipf(1.3, 32768) % 64 + 32
        d = bytes(sym.astype(np.uint8))
        d = d[:20000] + d[3000:15768]
        plains.append(d)
        z = zl(d, 6)
        assert (z[2] >> 1) & 3 == 2              # BTYPE = 2
        streams.append(z)
    stride = (max(len(s) for s in streams) + 15) & ~15
    buf = np.zeros((n, stride), dtype=np.uint8)
    for i, s in enumerate(streams):
        buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    out, out_len, status = engine.decompress_host(buf, lens, 32768, flags=3 | route)
    assert not status.any() and (out_len == 32768).all()
    for i in range(n):
        assert out[i].tobytes() == plains[i]


def test_error_statuses(engine):
    z = zl(b"hello hello hello hello hello", 6)
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(z[:-5])
    assert e.value.status == 5                                   # "NO EOF!"
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(z[:2] + bytes([z[2] | 6]) + z[3:])
    assert e.value.status == 2 and "Bad method" in str(e.value)  # BTYPE = 3
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(z, max_out=5)
    assert e.value.status == 6
    bad = bytearray(z)
    bad[-1] ^= 1
    assert engine.decompress(bytes(bad)) == b"hello hello hello hello hello"   # reference never checks Adler
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(bytes(bad), flags=hz.F_VERIFY_ADLER)
    assert e.value.status == 9
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(b"\x79\x9c" + z[2:], flags=hz.F_VERIFY_HEADER)
    assert e.value.status == 8
    # distance beyond the start of the output: fixed block, match (len 3, dist 1) as first token
    bits = 0b011 | (64 << 3)      # BFINAL=1 BTYPE=01, symbol 257 (code 0000001, MSB first), distance code 0, EOB
    raw = bytes([0x78, 0x9c]) + bits.to_bytes(4, "little") + bytes(4)
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(raw)
    assert e.value.status == 4 and "distance too big" in str(e.value)
    # the oracle agrees on every status above
    from oracle import hdlz_oracle
    assert hdlz_oracle.inflate(raw, 100)[0] == 4
    st = hdlz_oracle.inflate(zl(bytes(300), 0)[:2] + b"\x01\x05\x00\x00\x00", 100)[0]
    with pytest.raises(hz.StreamError) as e:
        engine.decompress(zl(bytes(300), 0)[:2] + b"\x01\x05\x00\x00\x00")
    assert e.value.status == st == 7                             # LEN != ~NLEN


def test_corrupted_streams_never_crash(engine):
    """Bit flips anywhere: the status is an error or the output equals zlib's; never a fault."""
    rnd = random.Random(21)
    data = " ".join("Hi: %d " % rnd.randrange(0x1000) for _ in range(400)).encode()
    z = bytearray(zl(data, 6))
    n = 256
    stride = (len(z) + 15) & ~15
    buf = np.zeros((n, stride), dtype=np.uint8)
    for i in range(n):
        c = bytearray(z)
        for _ in range(1 + i % 3):
            c[rnd.randrange(2, len(c))] ^= 1 << rnd.randrange(8)
        buf[i, :len(c)] = np.frombuffer(bytes(c), dtype=np.uint8)
    lens = np.full(n, len(z), dtype=np.uint32)
    out, out_len, status = engine.decompress_host(buf, lens, 2 * len(data), flags=3)
    for i in range(n):
        try:
            want = zlib.decompress(buf[i, :len(z)].tobytes())
        except zlib.error:
            want = None
        if status[i] == 0:
            assert want is not None and out[i, :out_len[i]].tobytes() == want, i
        else:
            assert want is None or len(want) > 2 * len(data), (i, status[i])


class _Bits(object):
    """LSB-first bit writer for hand-made deflate blocks."""

    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, val, nbits):
        self.v |= val << self.n
        self.n += nbits

    def code(self, c, nbits):                   # Huffman codes go MSB-first
        self.put(int(format(c, "0%db" % nbits)[::-1], 2), nbits)

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def _length_symbol_with_empty_distance_code():
    """Dynamic block whose distance code has no symbols at all (legal for literal-only blocks) but
    which then uses length symbol 257: zlib says 'invalid distance code'."""
    b = _Bits()
    b.put(1, 1); b.put(2, 2)                    # BFINAL, BTYPE = 2
    b.put(1, 5); b.put(0, 5); b.put(12, 4)      # HLIT = 258, HDIST = 1, HCLEN = 16
    cl = {0: 2, 2: 2, 17: 2, 18: 2}
    for s in (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2):
        b.put(cl.get(s, 0), 3)
    clcode = {0: 0, 2: 1, 17: 2, 18: 3}
    def zeros(n):
        b.code(clcode[18], 2); b.put(n - 11, 7)
    zeros(97)
    b.code(clcode[2], 2); b.code(clcode[2], 2)          # 'a', 'b'
    zeros(138); zeros(19)
    b.code(clcode[2], 2); b.code(clcode[2], 2)          # 256, 257
    b.code(clcode[0], 2)                                # the one distance length: 0
    lit = {97: 0, 98: 1, 256: 2, 257: 3}
    b.code(lit[97], 2)
    b.code(lit[257], 2)                                 # length 3 ... and no distance code exists
    b.put(0, 16)
    b.code(lit[256], 2)
    body = b.bytes()
    return b"\x78\x9c" + body + zlib.adler32(b"aaaa").to_bytes(4, "big")


@pytest.mark.parametrize("flags", [0, 3])
def test_empty_distance_code_is_rejected_on_every_route(engine, flags):
    """A code without symbols must not be decoded with the previous table's resume point
    (round-1 advisor finding): all routes and zlib reject the stream, with and without checksums."""
    bad = _length_symbol_with_empty_distance_code()
    with pytest.raises(zlib.error):
        zlib.decompress(bad)
    good = zl(bytes(np.random.default_rng(1).integers(97, 101, 3000, dtype=np.uint8)), 6)   # builds real tables first
    n = 1100
    streams = [good if i % 2 == 0 else bad for i in range(n)]
    stride = (max(len(s) for s in streams) + 15) & ~15
    buf = np.zeros((n, stride), dtype=np.uint8)
    for i, s in enumerate(streams):
        buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    for route in (0, FORCE_GENERAL, FORCE_LANES, FORCE_LANES | NO_SPLIT):
        out, out_len, status = engine.decompress_host(buf, lens, 4096, flags=flags | route)
        assert (status[0::2] == 0).all() and (status[1::2] == 3).all(), (route, status[:8])


def test_two_phase_route_shapes(engine):
    """The decode -> resolve route (hdlz_inflate_split.cu) on what stresses its token format: literal runs far
    beyond 255, copies of every length up to 258 at distances 1 .. 32768 (overlapping and not), stored and
    fixed blocks mixed into dynamic streams, outputs that are not multiples of 16, raw containers, the pool
    growing from call to call.  zlib is the expected value."""
    rnd = random.Random(99)
    rng = np.random.default_rng(99)
    plains = []
    for t in range(1100):
        kind = t % 11
        n = rnd.choice([17, 333, 4099, 20000, 32768, 32767, 31001])
        if kind == 0:
            d = bytes(rng.integers(0, 256, n, dtype=np.uint8))                      # incompressible: stored blocks / long literal runs
        elif kind == 1:
            d = bytes([rnd.randrange(256)]) * n                                      # distance 1, length 258 chains
        elif kind == 2:
            p = bytes(rng.integers(97, 123, rnd.choice([2, 3, 5, 31, 33, 257, 259, 4000])))
            d = (p * (n // len(p) + 1))[:n]                                          # periodic: overlapping copies
        elif kind == 3:
            base = bytes(rng.integers(0, 256, 3000, dtype=np.uint8))
            d = (base + bytes(rng.integers(0, 4, n, dtype=np.uint8)))[:n]
            d = d + d[:max(0, 32768 - len(d))]                                       # far copies (distance up to 32 KiB)
            d = d[:32768]
        elif kind == 4:
            d = bytes((rng.zipf(1.3, n) % 64 + 32).astype(np.uint8))
        elif kind == 5:
            d = " ".join("   Hello World! %d     " % i for i in range(2000)).encode()[:n]
        elif kind == 6:
            d = bytes(rng.integers(0, 256, n // 2, dtype=np.uint8)) + bytes(n - n // 2)
        elif kind == 7:
            d = workload.block(t, n)
        elif kind == 8:
            d = bytes(rng.choice(np.frombuffer(b"ab", dtype=np.uint8), n))
        elif kind == 9:
            d = b""
        else:
            d = bytes(rng.integers(0, 256, 5, dtype=np.uint8)) * (n // 5)
        plains.append(d)
    for wbits, fl in ((15, 0), (-15, hz.F_RAW)):
        streams = []
        for i, d in enumerate(plains):
            lvl, strat = [(6, 0), (9, 0), (1, 0), (6, zlib.Z_FILTERED), (6, zlib.Z_RLE), (6, zlib.Z_HUFFMAN_ONLY)][i % 6]
            co = zlib.compressobj(lvl, zlib.DEFLATED, wbits, 8, strat)
            if i % 5 == 0 and len(d) > 100:                                          # several blocks, a stored one in the middle
                z = co.compress(d[:len(d) // 3]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(d[len(d) // 3:]) + co.flush()
            else:
                z = co.compress(d) + co.flush()
            streams.append(z)
        stride = (max(len(s) for s in streams) + 15) & ~15
        buf = np.zeros((len(streams), stride), dtype=np.uint8)
        for i, s in enumerate(streams):
            buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        lens = np.array([len(s) for s in streams], dtype=np.uint32)
        for rep in range(2):                                                         # second call: the pool has grown to the demand
            out, out_len, status = engine.decompress_host(buf, lens, 32768, flags=fl | (hz.F_VERIFY_ADLER if rep else 0))
            assert not status.any(), (wbits, rep, np.nonzero(status)[0][:8], status[status != 0][:8])
            for i, d in enumerate(plains):
                assert out_len[i] == len(d) and out[i, :len(d)].tobytes() == d, (wbits, rep, i)
    # a wrong Adler-32 is caught by phase 2, a too small out_cap by phase 1
    z = bytearray(zl(plains[4], 6))
    z[-1] ^= 0x55
    bad = np.zeros((1100, (len(z) + 15) & ~15), dtype=np.uint8)
    bad[:, :len(z)] = np.frombuffer(bytes(z), dtype=np.uint8)
    blens = np.full(1100, len(z), dtype=np.uint32)
    _, _, status = engine.decompress_host(bad, blens, 32768, flags=hz.F_VERIFY_ADLER)
    assert (status == 9).all()
    _, _, status = engine.decompress_host(bad, blens, len(plains[4]) - 1, flags=0)
    assert (status == 6).all()


def test_decompress_stream_fed_in_pieces(engine):
    """hdlz_dstream_*: a stream fed in pieces inflates as far as the input received allows (the reference's decoder
    waiting at `di >= isize - 4`, deflate.py:1529) and the pieces of output, joined, equal the one-shot result:
    the state kept on the device between calls (bit cursor, output cursor, header position of the current block)
    survives any cut — inside a dynamic block header, a stored block, a symbol."""
    import gzip
    rnd = random.Random(99)
    text = b"".join(b"line %d: %s\n" % (i, bytes(rnd.choice(b"abcdefghij ") for _ in range(rnd.randrange(10, 60))))
                    for i in range(6000))                       # ~260 kB, several dynamic blocks at level 6
    noise = bytes(rnd.randrange(256) for _ in range(70000))      # stored blocks
    co = zlib.compressobj(6)
    mixed = co.compress(text[:50000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(noise[:30000]) + \
        co.flush(zlib.Z_SYNC_FLUSH) + co.compress(text[50000:90000]) + co.flush()
    cases = [
        (zlib.compress(text, 6), text, 0),
        (zlib.compress(text, 9), text, hz.F_VERIFY_ADLER),
        (zlib.compress(noise, 0), noise, hz.F_VERIFY_ADLER),
        (mixed, text[:50000] + noise[:30000] + text[50000:90000], hz.F_VERIFY_ADLER),
        (engine.compress(text[:100000]), text[:100000], hz.F_VERIFY_ADLER),            # the reference's fixed-block format
        (gzip.compress(text[:80000], 6), text[:80000], hz.F_GZIP | hz.F_VERIFY_ADLER),
        (zlib.compress(b"", 6), b"", hz.F_VERIFY_ADLER),
        (zlib.compress(b"abc", 6), b"abc", 0),
    ]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    cases.append((co.compress(text[:60000]) + co.flush(), text[:60000], hz.F_RAW))
    for ci, (stream, plain, flags) in enumerate(cases):
        for sizes in ([1], [7, 100, 3], [2048], [1, 2, 7, 33, 100, 1024, 1100, 4096, 20000], [len(stream) + 1]):
            if sizes == [1] and len(stream) > 3000:
                continue                                         # byte-wise only for the small streams
            s = engine.decompress_stream(max_out=len(plain) + 64, flags=flags)
            got, pos, progress = bytearray(), 0, 0
            while pos < len(stream):
                n = rnd.choice(sizes)
                got += s.feed(stream[pos:pos + n])
                pos += n
                assert progress <= s.in_progress <= pos          # o_iprogress never runs ahead of the input
                progress = s.in_progress
                assert bytes(got) == plain[:len(got)]            # what has been handed out is final
            if len(stream) > 20000 and max(sizes) < 5000:
                assert len(got) > 0.8 * len(plain), (ci, sizes)  # most of the output was out before finish()
            got += s.finish()
            s.close()
            assert bytes(got) == plain, (ci, sizes)
        assert engine.decompress(stream, flags=flags) == plain
    # back-pressure: the caller takes 1000 bytes per call, the rest waits on the device
    stream, plain = cases[0][0], cases[0][1]
    s = engine.decompress_stream(max_out=len(plain))
    got = bytearray()
    for pos in range(0, len(stream), 5000):
        piece = s.feed(stream[pos:pos + 5000], room=1000)
        assert len(piece) <= 1000
        got += piece
    while True:
        got += s.finish(room=1000)
        if not s.remaining:
            break
    s.close()
    assert bytes(got) == plain
    # errors: a stream cut short, a damaged stream, an output that does not fit
    s = engine.decompress_stream(max_out=len(plain))
    s.feed(stream[:len(stream) // 2])
    with pytest.raises(hz.StreamError) as e:
        s.finish()
    assert e.value.status == 5                                   # TRUNCATED ("NO EOF!")
    s.close()
    bad = bytearray(stream)
    bad[-2] ^= 0x55
    s = engine.decompress_stream(max_out=len(plain), flags=hz.F_VERIFY_ADLER)
    for pos in range(0, len(bad), 4096):
        s.feed(bytes(bad[pos:pos + 4096]))
    with pytest.raises(hz.StreamError) as e:
        s.finish()
    assert e.value.status == 9                                   # BAD_ADLER
    s.close()
    s = engine.decompress_stream(max_out=1000)
    with pytest.raises(hz.StreamError) as e:
        for pos in range(0, len(stream), 4096):
            s.feed(stream[pos:pos + 4096])
        s.finish()
    assert e.value.status == 6                                   # OUT_OVERFLOW
    s.close()
