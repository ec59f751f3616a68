"""Container framing around the deflate body: zlib is the reference's
(deflate.py:753-757, :788-814); raw deflate and gzip wrap the same body.  Compress: the body is the
oracle's bits, the frame is what host zlib reads back.  Decompress: byte-exact with host zlib on raw /
gzip streams, fixed and dynamic, with optional gzip header fields, through every kernel route."""
import gzip
import io
import random
import zlib

import numpy as np
import pytest

import hdl_deflate_b200 as hz
from hdl_deflate_b200 import workload
from oracle import hdlz_oracle

pytestmark = pytest.mark.gpu

FORCE_GENERAL, FORCE_LANES = 0x100, 0x200


def zl(data, wbits, level=6, strategy=0):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    return co.compress(data) + co.flush()


def gz_with_fields(data, name=b"block.bin", level=6):
    """gzip member with FNAME (and MTIME) set, as gzip.GzipFile writes it."""
    buf = io.BytesIO()
    with gzip.GzipFile(filename=name.decode(), mode="wb", fileobj=buf, compresslevel=level, mtime=1700000000) as f:
        f.write(data)
    return buf.getvalue()


def batch(engine, streams, out_cap, flags):
    stride = (max(len(s) for s in streams) + 15) & ~15
    arr = np.zeros((len(streams), stride), dtype=np.uint8)
    for i, s in enumerate(streams):
        arr[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    lens = np.array([len(s) for s in streams], dtype=np.uint32)
    return engine.decompress_host(arr, lens, out_cap, flags=flags)


def test_compress_containers_same_body(engine):
    rnd = random.Random(5)
    datas = [b"abcde", b"a" * 12, workload.block(3, 2048), workload.block(4, 1000), bytes(5000),
             bytes(rnd.randrange(256) for _ in range(3333)), workload.block(9, 4133)]
    try:
        for data in datas:
            st, ref = hdlz_oracle.compress(data)
            assert st == 0
            body = ref[2:-4]
            engine.container = hz.CONTAINER_ZLIB
            assert engine.compress(data) == ref
            engine.container = hz.CONTAINER_RAW
            raw = engine.compress(data)
            assert raw == body
            assert zlib.decompress(raw, -15) == data
            engine.container = hz.CONTAINER_GZIP
            gz = engine.compress(data)
            assert gz[:10] == bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF]) and gz[10:-8] == body
            assert int.from_bytes(gz[-8:-4], "little") == zlib.crc32(data)
            assert int.from_bytes(gz[-4:], "little") == len(data)
            assert gzip.decompress(gz) == data and zlib.decompress(gz, 31) == data
            # and our own inflater reads all three back
            assert engine.decompress(gz, flags=hz.F_GZIP | hz.F_VERIFY_ADLER) == data
            assert engine.decompress(raw, flags=hz.F_RAW) == data
    finally:
        engine.container = hz.CONTAINER_ZLIB


def test_compress_gzip_batch(engine):
    n = 4096
    arr = np.frombuffer(b"".join(workload.blocks(700, n, 2048)), dtype=np.uint8).reshape(n, 2048)
    engine.container = hz.CONTAINER_GZIP
    try:
        assert hz.compress_bound(2048, hz.CONTAINER_GZIP) == 2336
        out, out_len, status = engine.compress_host(arr)
        assert not status.any()
        for i in range(0, n, 37):
            s = out[i, :out_len[i]].tobytes()
            assert zlib.decompress(s, 31) == arr[i].tobytes()
            assert s[10:-8] == hdlz_oracle.compress(arr[i].tobytes())[1][2:-4]
        back, back_len, bst = engine.decompress_host(out, out_len, 2048, flags=hz.F_GZIP | hz.F_VERIFY_ADLER)
        assert not bst.any() and np.array_equal(back, arr)
        packed, off, plen, pst = engine.compress_host_packed(arr)
        assert not pst.any() and np.array_equal(plen, out_len)
        j = n - 1
        assert packed[int(off[j]):int(off[j]) + int(plen[j])].tobytes() == out[j, :out_len[j]].tobytes()
    finally:
        engine.container = hz.CONTAINER_ZLIB


@pytest.mark.parametrize("route", [0, FORCE_GENERAL, FORCE_LANES])
def test_decompress_raw_and_gzip_match_zlib(engine, route):
    rnd = random.Random(11)
    plains = [workload.block(i, rnd.choice([5, 64, 700, 2048, 5000])) for i in range(40)]
    plains += [b"", b"x", bytes(3000), bytes(rnd.choice(b"abcdefgh") for _ in range(4000))]
    raws, gzs = [], []
    for k, p in enumerate(plains):
        strategy = zlib.Z_FIXED if k % 3 == 0 else 0
        raws.append(zl(p, -15, 6 if k % 5 else 0, strategy))
        gzs.append(gz_with_fields(p) if k % 2 else zl(p, 31, 6, strategy))
    for streams, flag in ((raws, hz.F_RAW), (gzs, hz.F_GZIP), (gzs, hz.F_GZIP | hz.F_VERIFY_ADLER)):
        out, out_len, status = batch(engine, streams, 5008, flag | route)
        assert not status.any(), status
        for i, p in enumerate(plains):
            assert out[i, :out_len[i]].tobytes() == p, i


def test_gzip_errors(engine):
    data = workload.block(1, 900)
    g = zl(data, 31)
    bad_magic = b"\x1f\x8c" + g[2:]
    bad_crc = g[:-8] + bytes([g[-8] ^ 1]) + g[-7:]
    bad_isize = g[:-1] + bytes([g[-1] ^ 1])
    cut = g[:-3]
    streams = [g, bad_magic, bad_crc, bad_isize, cut] * 8
    for route in (0, FORCE_GENERAL, FORCE_LANES):
        _, out_len, status = batch(engine, streams, 1024, hz.F_GZIP | hz.F_VERIFY_ADLER | route)
        want = [0, 8, 10, 10, 5] * 8
        assert status.tolist() == want, (route, status.tolist())
        assert out_len[0] == 900 and out_len[1] == 0
        _, _, status = batch(engine, streams, 1024, hz.F_GZIP | route)          # no checksum requested
        assert status.tolist() == [0, 8, 0, 0, 5] * 8
    assert hz.STATUS_NAMES[10] == "BAD_CRC"
