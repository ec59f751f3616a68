"""Tree-coded compressor mode on the GPU (hdlz_set_tree / hdlz_train_tree) against the restatement in
oracle/tree_oracle.py: same bytes, valid for zlib, and back through the engine's own dynamic-block inflater."""
import random
import zlib

import numpy as np
import pytest

import hdl_deflate_b200 as hz
from hdl_deflate_b200 import workload
from oracle import hdlz_oracle as O
from oracle import tree_oracle as T

pytestmark = pytest.mark.gpu


def text_blocks(rnd, n, length, alphabet=b"abcdefgh   xyz"):
    return [bytes(rnd.choice(alphabet) for _ in range(length)) for _ in range(n)]


@pytest.fixture()
def eng(engine):
    yield engine
    engine.set_tree()
    engine.container = hz.CONTAINER_ZLIB
    engine.match10 = True


def test_train_equals_restatement_and_streams_bit_exact(eng):
    rnd = random.Random(21)
    n, length = 96, 2048
    blocks = text_blocks(rnd, n // 2, length) + list(workload.blocks(500, n // 2, length))
    arr = np.frombuffer(b"".join(blocks), dtype=np.uint8).reshape(n, length)
    assert eng.tree is None
    lit, dist = eng.train_tree(arr)
    wlit, wdist = T.train(blocks)
    assert list(lit) == wlit and list(dist) == wdist
    assert eng.tree is not None
    out, out_len, status = eng.compress_host(arr)
    assert not status.any()
    assert out.shape[1] == eng.bound(length)
    for i in range(n):
        st, want = T.compress(blocks[i], wlit, wdist)
        got = out[i, :out_len[i]].tobytes()
        assert st == 0 and got == want, i
        assert zlib.decompress(got) == blocks[i]
    # back through the engine's inflater (dynamic-block route), checksums verified
    back, back_len, bst = eng.decompress_host(out, out_len, length, flags=hz.F_VERIFY_ADLER)
    assert not bst.any() and (back_len == length).all() and np.array_equal(back, arr)
    # the text half must have gained on the fixed code
    fixed = sum(len(O.compress(b)[1]) for b in blocks[:n // 2])
    assert int(out_len[:n // 2].sum()) < 0.7 * fixed
    # and the fixed code comes back unchanged
    eng.set_tree()
    assert eng.tree is None
    assert eng.compress(b"a" * 12).hex() == "789c4b8483c444001d9a048d"


def test_ragged_lengths_containers_match5(eng):
    rnd = random.Random(22)
    lens = [5, 6, 7, 31, 32, 33, 64, 100, 1023, 1024, 1025, 2047, 2048, 3000, 4096, 5000]
    blocks = [bytes(rnd.choice(b"hello world 0123") for _ in range(l)) for l in lens]
    stride = 5008
    arr = np.zeros((len(lens), stride), dtype=np.uint8)
    for i, b in enumerate(blocks):
        arr[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    la = np.array(lens, dtype=np.uint32)
    for match10 in (True, False):
        eng.match10 = match10
        maxlen = 10 if match10 else 5
        eng.train_tree(arr, la)
        lit, dist = eng.tree
        wl, wd = T.train(blocks, maxlen=maxlen)
        assert list(lit) == wl and list(dist) == wd
        for cont, wbits, flag in ((hz.CONTAINER_ZLIB, 15, 0), (hz.CONTAINER_RAW, -15, hz.F_RAW), (hz.CONTAINER_GZIP, 31, hz.F_GZIP)):
            eng.container = cont
            out, out_len, status = eng.compress_host(arr, la)
            assert not status.any()
            for i, b in enumerate(blocks):
                st, want = T.compress(b, wl, wd, container=cont, maxlen=maxlen)
                got = out[i, :out_len[i]].tobytes()
                assert st == 0 and got == want, (match10, cont, i)
                assert zlib.decompress(got, wbits) == b
            back, back_len, bst = eng.decompress_host(out, out_len, 5000, flags=flag | hz.F_VERIFY_ADLER)
            assert not bst.any() and list(back_len) == lens
            # single-stream entry point, packed entry point
            assert eng.compress(blocks[-1]) == out[-1, :out_len[-1]].tobytes()
            packed, off, plen, pst = eng.compress_host_packed(arr, la)
            assert not pst.any() and np.array_equal(plen, out_len)
            for i in (0, 7, 15):
                assert packed[int(off[i]):int(off[i]) + int(plen[i])].tobytes() == out[i, :out_len[i]].tobytes()


def test_application_tree_and_missing_code(eng):
    """A code over the byte values the application knows it has (README.md:43-45); a stream with another byte
    ends with NO_CODE, its neighbours are coded."""
    keep = [ord("a"), ord("b")] + list(range(256, 265))
    lit = T.limited_lengths([1 if s in keep else 0 for s in range(286)], 15)
    dist = T.limited_lengths([1] * 10 + [0] * 20, 15)
    eng.set_tree(lit, dist)
    good, bad = b"ababbbabaabab" * 30, b"ababcbabaabab" * 30
    arr = np.frombuffer(good + bad + good, dtype=np.uint8).reshape(3, len(good))
    arr = np.ascontiguousarray(np.pad(arr, ((0, 0), (0, 400 - len(good)))))
    out, out_len, status = eng.compress_host(arr, np.full(3, len(good), np.uint32))
    assert list(status) == [0, 11, 0] and out_len[1] == 0
    st, want = T.compress(good, lit, dist)
    assert out[0, :out_len[0]].tobytes() == want and out[2, :out_len[2]].tobytes() == want
    assert zlib.decompress(want) == good
    with pytest.raises(hz.StreamError):
        eng.compress(bad)
    # refused lengths: over-subscribed, no end-of-block code
    over = [1] * 5 + [0] * 281
    with pytest.raises(hz.HdlzError):
        eng.set_tree(over, dist)
    noeob = list(lit)
    noeob[256] = 0
    with pytest.raises(hz.HdlzError):
        eng.set_tree(noeob, dist)
    # the slow engine and streams fed in pieces keep the fixed code
    eng.fast = False
    try:
        with pytest.raises(hz.HdlzError):
            eng.compress(good)
    finally:
        eng.fast = True
    with pytest.raises(hz.HdlzError):
        eng.compress_stream()


def test_long_stream_and_worst_case_bound(eng):
    """One stream of many tiles; and a tree whose literals cost 15 bits (the slot bound and the lane-private
    streams at their largest)."""
    rnd = random.Random(23)
    data = b"".join(text_blocks(rnd, 1, 70001))
    eng.train_tree(np.frombuffer(data[:70000], dtype=np.uint8).reshape(1, 70000)[:, :69984])
    lit, dist = eng.tree
    got = eng.compress(data)
    st, want = T.compress(data, list(lit), list(dist))
    assert st == 0 and got == want and zlib.decompress(got) == data
    # Fibonacci counts: the code hits the 15-bit limit
    fib = [1, 1]
    while len(fib) < 60:
        fib.append(fib[-1] + fib[-2])
    cnt = [fib[59 - (s % 60)] for s in range(256)] + [1 << 30] + [1 << 20] * 8 + [0] * 21
    lit = T.limited_lengths(cnt, 15)
    assert max(lit) == 15
    dist = T.limited_lengths([1] * 10 + [0] * 20, 15)
    eng.set_tree(lit, dist)
    noise = bytes(rnd.randrange(1, 256) for _ in range(2048))
    arr = np.frombuffer(noise, dtype=np.uint8).reshape(1, 2048)
    out, out_len, status = eng.compress_host(arr)
    assert not status.any() and out_len[0] <= eng.bound(2048)
    st, want = T.compress(noise, lit, dist)
    assert out[0, :out_len[0]].tobytes() == want and zlib.decompress(want) == noise
    # a slot smaller than the bound is refused, not overrun
    out, out_len, status = eng.compress_host(arr, out_stride=2320)
    assert status[0] == 6 and out_len[0] == 0


def test_long_stream_with_a_tree(eng):
    """The long-stream kernel (one stream over the whole grid) with an installed tree: the same bytes as the
    restatement, and a symbol without a code in a late tile fails the whole stream."""
    rnd = random.Random(24)
    data = b"".join(text_blocks(rnd, 1, 300001, b"ACGTN acgt\n"))
    arr = np.frombuffer(data[:294912], dtype=np.uint8).reshape(144, 2048)
    lit, dist = eng.train_tree(arr)
    got = eng.compress(data)
    st, want = T.compress(data, list(lit), list(dist))
    assert st == 0 and got == want and zlib.decompress(got) == data
    assert len(got) < 0.7 * len(O.compress(data)[1])
    eng.container = hz.CONTAINER_RAW
    st, want = T.compress(data, list(lit), list(dist), container=1)
    assert eng.compress(data) == want
    eng.container = hz.CONTAINER_ZLIB
    # a code without 'Z': one 'Z' far into the stream
    keep = set(b"ACGTN acgt\n") | set(range(256, 265))
    lit2 = T.limited_lengths([1 if s in keep else 0 for s in range(286)], 15)
    eng.set_tree(lit2, list(dist))
    assert zlib.decompress(eng.compress(data)) == data
    bad = bytearray(data)
    bad[250000] = ord("Z")
    with pytest.raises(hz.StreamError) as e:
        eng.compress(bytes(bad))
    assert e.value.status == 11


def test_stream_with_its_own_tree(eng):
    """hdlz_compress_stream_dyn: the statistics of the stream's own 2 KiB blocks, one BTYPE = 10 block, the
    reference's tokens; the context's tree setting is untouched afterwards; short streams keep the fixed code."""
    rnd = random.Random(25)
    data = b"".join(text_blocks(rnd, 1, 150001, b"the quick brown fox\n"))
    nblk = len(data) // 2048
    wl, wd = T.train([data[2048 * i:2048 * (i + 1)] for i in range(nblk)])
    got = eng.compress_dynamic(data)
    st, want = T.compress(data, wl, wd)
    assert st == 0 and got == want and zlib.decompress(got) == data
    assert (got[2] >> 1) & 3 == 2                          # BTYPE = 10
    assert len(got) < 0.75 * len(O.compress(data)[1])
    assert eng.tree is None
    assert eng.decompress(got, flags=hz.F_VERIFY_ADLER) == data
    # a few blocks only (the one-warp kernel), bytes outside the trained blocks' alphabet in the tail
    small = data[:5000] + bytes(range(256))
    wl, wd = T.train([small[:2048], small[2048:4096]])
    got = eng.compress_dynamic(small)
    assert got == T.compress(small, wl, wd)[1] and zlib.decompress(got) == small
    # shorter than one block: the fixed code
    assert eng.compress_dynamic(b"a" * 12).hex() == "789c4b8483c444001d9a048d"
    with pytest.raises(hz.StreamError) as e:               # isize < 5: the engine never starts, as in every mode
        eng.compress_dynamic(b"abcd")
    assert e.value.status == 1                             # HDLZ_ST_SHORT_INPUT
    # an installed tree survives the call
    lit, dist = eng.train_tree(np.frombuffer(data[:4096], dtype=np.uint8).reshape(2, 2048))
    eng.compress_dynamic(data[:30000])
    assert eng.tree is not None and list(eng.tree[0]) == list(lit) and list(eng.tree[1]) == list(dist)
    # other containers
    eng.set_tree()
    eng.container = hz.CONTAINER_GZIP
    assert zlib.decompress(eng.compress_dynamic(data[:40000]), 31) == data[:40000]
