"""Tree-coded compressor mode, the parts that need no GPU: the restatement (oracle/tree_oracle.py) against zlib,
and the device-independent host code of hdlz_tree.cu (code lengths, stream prefix) against the restatement."""
import ctypes
import random
import zlib

import numpy as np

from oracle import hdlz_oracle as O
from oracle import tree_oracle as T


def _lib():
    import __graft_entry__
    __graft_entry__.build()
    from hdl_deflate_b200 import _native
    return _native.load()


def small_alphabet(rnd, n, alphabet=b"abcdefgh   xyz"):
    return bytes(rnd.choice(alphabet) for _ in range(n))


def test_oracle_tree_streams_are_valid_deflate():
    """zlib inflates what the restatement writes (the reference's own check of its compressor,
    test_deflate.py:285), in all three containers, for both MATCH10 settings; the tokens are those of the
    fixed-tree restatement (same parse), so only the coding differs."""
    rnd = random.Random(11)
    for trial in range(12):
        data = small_alphabet(rnd, rnd.choice([5, 6, 40, 700, 2048, 5000]))
        lit, dist = T.train([data])
        for cont, wbits in ((0, 15), (1, -15), (2, 31)):
            for maxlen in (10, 5):
                st, z = T.compress(data, lit, dist, container=cont, maxlen=maxlen)
                assert st == 0 and zlib.decompress(z, wbits) == data
        # same parse as the fixed-tree path: token count and the bytes they cover
        toks = T.parse(data)
        assert sum(m for m, _ in toks) == len(data)
        assert [(m, d) for _, m, d in O.parse(data)] == [(m, d if m > 1 else 0) for m, d in toks]


def test_oracle_tree_beats_fixed_on_few_byte_values():
    """The case the reference's README names (README.md:43-45): data over a small set of byte values."""
    rnd = random.Random(5)
    data = small_alphabet(rnd, 4096, b"ACGT")
    lit, dist = T.train([data])
    st, z = T.compress(data, lit, dist)
    assert st == 0 and zlib.decompress(z) == data
    assert len(z) < 0.6 * len(O.compress(data)[1])


def test_oracle_tree_missing_code():
    lit, dist = T.train([b"aaaaabbbbb" * 20])
    lit = list(lit)
    # a code over 'a', 'b', EOB and the length symbols only
    keep = [ord("a"), ord("b")] + list(range(256, 265))
    cnt = [1 if s in keep else 0 for s in range(286)]
    lit = T.limited_lengths(cnt, 15)
    st, z = T.compress(b"ababbbabaabab" * 9, lit, dist)
    assert st == 0 and zlib.decompress(z) == b"ababbbabaabab" * 9
    st, z = T.compress(b"ababcbabaabab" * 9, lit, dist)
    assert st == 11 and z == b""


def test_host_lengths_equal_restatement():
    L = _lib()
    rnd = random.Random(3)
    for trial in range(300):
        n = rnd.choice([19, 30, 286])
        mb = 7 if n == 19 else 15
        kind = trial % 4
        if kind == 0:
            f = [rnd.randrange(0, 1000) for _ in range(n)]
        elif kind == 1:                                   # Fibonacci counts: the unlimited code would be n - 1 deep
            a, b, f = 1, 1, []
            for _ in range(n):
                f.append(min(a, 2 ** 50))
                a, b = b, a + b
            rnd.shuffle(f)
        elif kind == 2:
            f = [rnd.choice([0, 0, 0, 1, 2, 5]) for _ in range(n)]
        else:
            f = [int(2 ** rnd.uniform(0, 30)) for _ in range(n)]
        cnt = np.array(f, dtype=np.uint64)
        out = np.zeros(n, np.uint8)
        assert L.hdlz_tree_lengths(cnt.ctypes.data, n, mb, out.ctypes.data) == 0
        want = T.limited_lengths(f, mb)
        assert list(out) == want, trial
        used = [l for l in want if l]
        if len(used) > 1:
            assert abs(sum(2.0 ** -l for l in used) - 1) < 1e-12 and max(used) <= mb
    # invalid arguments
    cnt = np.ones(20, dtype=np.uint64)
    out = np.zeros(20, np.uint8)
    assert L.hdlz_tree_lengths(cnt.ctypes.data, 20, 4, out.ctypes.data) == -1      # 20 symbols do not fit 4 bits
    assert L.hdlz_tree_lengths(None, 20, 15, out.ctypes.data) == -1


def test_host_prefix_equals_restatement():
    L = _lib()
    rnd = random.Random(8)
    for trial in range(30):
        data = small_alphabet(rnd, 1500, rnd.choice([b"abcdefgh   xyz", b"01", bytes(range(256))]))
        lit, dist = T.train([data])
        for cont in (0, 1, 2):
            b = T._Bits()
            if cont == 0:
                b.put(0x78, 8)
                b.put(0x9C, 8)
            elif cont == 2:
                for v in (0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF):
                    b.put(v, 8)
            T._header(b, lit, dist)
            la, da = np.array(lit, np.uint8), np.array(dist, np.uint8)
            out = np.zeros(400, np.uint8)
            nb = ctypes.c_uint32(0)
            assert L.hdlz_tree_header(la.ctypes.data, da.ctypes.data, cont, out.ctypes.data, 400, ctypes.byref(nb)) == 0
            assert nb.value == b.n and out[:(b.n + 7) // 8].tobytes() == b.tobytes()
    # over-subscribed / incomplete lengths are refused
    bad = np.zeros(286, np.uint8)
    bad[:5] = 1
    da = np.zeros(30, np.uint8)
    out = np.zeros(400, np.uint8)
    nb = ctypes.c_uint32(0)
    assert L.hdlz_tree_header(bad.ctypes.data, da.ctypes.data, 0, out.ctypes.data, 400, ctypes.byref(nb)) == -1
    bad[:] = 0
    bad[0], bad[256] = 2, 2
    assert L.hdlz_tree_header(bad.ctypes.data, da.ctypes.data, 0, out.ctypes.data, 400, ctypes.byref(nb)) == -1
