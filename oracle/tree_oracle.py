"""tree_oracle.py — TEST INFRASTRUCTURE ONLY: CPU restatement of the tree-coded compressor mode
(hdlz_set_tree / hdlz_train_tree, include/hdlz.h).

The reference codes with the fixed tree only (deflate.py:112-149, 1064-1076) and names "a dedicated
pre-computed Huffman tree" as the next step (README.md:43-45); there is no reference implementation of
this mode to execute, so its parity is pinned differently from the fixed-tree path:

  * the TOKENS are the reference's: they come from hdlz_oracle_parse_ex (oracle/hdlz_oracle.c, the
    restatement of SEARCH / SEARCHF, deflate.py:899-1016, itself pinned by the executing reference);
  * the CODING is RFC 1951 3.2.7, restated here from the RFC (canonical codes, HLIT / HDIST / HCLEN,
    run-length coded lengths) independently of hdl-deflate_b200/csrc/hdlz_tree.cu, following the
    deterministic choices documented there (package-merge tie order, run-length rule);
  * validity is checked by zlib (the reference's own check of its compressor, test_deflate.py:285):
    zlib.decompress(stream) == input.

Pure Python: small cases only.
"""
import ctypes
import zlib

import numpy as np

from . import hdlz_oracle as O

DIST_BASE = (1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
             2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577)
CL_ORDER = (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15)


def parse(data, cwindow=32, maxlen=10):
    """[(length, distance)] of the reference's greedy parse; length 1 = literal."""
    L = O.lib()
    L.hdlz_oracle_parse_ex.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
                                       ctypes.c_uint, ctypes.c_uint]
    L.hdlz_oracle_parse_ex.restype = ctypes.c_uint32
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    tok = np.empty(2 * (len(a) + 1), dtype=np.uint32)
    n = L.hdlz_oracle_parse_ex(a.ctypes.data, len(a), tok.ctypes.data, len(a) + 1, cwindow, maxlen)
    t = tok[:2 * n].reshape(n, 2)
    return [(int(m), int(d)) for m, d in t]


def limited_lengths(freq, maxbits):
    """Package-merge: optimal code lengths <= maxbits.  Leaves ordered by (count, symbol); in a merge a leaf
    goes before a package of the same weight."""
    n = len(freq)
    lens = [0] * n
    used = sorted((s for s in range(n) if freq[s]), key=lambda s: (freq[s], s))
    m = len(used)
    if m == 0:
        return lens
    if m == 1:
        lens[used[0]] = 1
        return lens
    lists = [[(freq[s], True) for s in used]]
    for _ in range(1, maxbits):
        prev = lists[-1]
        packages = [prev[2 * i][0] + prev[2 * i + 1][0] for i in range(len(prev) // 2)]
        cur, li, pi = [], 0, 0
        while li < m or pi < len(packages):
            if li < m and (pi >= len(packages) or freq[used[li]] <= packages[pi]):
                cur.append((freq[used[li]], True))
                li += 1
            else:
                cur.append((packages[pi], False))
                pi += 1
        lists.append(cur)
    take = 2 * m - 2
    for level in reversed(lists):
        if take == 0:
            break
        take = min(take, len(level))
        leaves = sum(1 for w, is_leaf in level[:take] if is_leaf)
        for i in range(leaves):
            lens[used[i]] += 1
        take = 2 * (take - leaves)
    return lens


def canonical(lens):
    """RFC 1951 3.2.2 -> list of (code bit-reversed for LSB-first emission, length)."""
    count = [0] * 16
    for l in lens:
        count[l] += 1
    count[0] = 0
    nxt, code = [0] * 16, 0
    for l in range(1, 16):
        code = (code + count[l - 1]) << 1
        nxt[l] = code
    out = []
    for l in lens:
        if l == 0:
            out.append((0, 0))
            continue
        c = nxt[l]
        nxt[l] += 1
        out.append((int(format(c, "0%db" % l)[::-1], 2), l))
    return out


class _Bits(object):
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, v, n):
        self.v |= (v & ((1 << n) - 1)) << self.n
        self.n += n

    def tobytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def _header(b, lit, dist):
    nlit, ndist = 286, 30
    while nlit > 257 and lit[nlit - 1] == 0:
        nlit -= 1
    while ndist > 1 and dist[ndist - 1] == 0:
        ndist -= 1
    seq = list(lit[:nlit]) + list(dist[:ndist])
    rl, i = [], 0
    while i < len(seq):
        v, run = seq[i], 1
        while i + run < len(seq) and seq[i + run] == v:
            run += 1
        i += run
        if v == 0:
            while run >= 11:
                r = min(run, 138)
                rl.append((18, r - 11))
                run -= r
            if run >= 3:
                rl.append((17, run - 3))
                run = 0
            rl += [(0, 0)] * run
        else:
            rl.append((v, 0))
            run -= 1
            while run >= 3:
                r = min(run, 6)
                rl.append((16, r - 3))
                run -= r
            rl += [(v, 0)] * run
    clfreq = [0] * 19
    for s, _ in rl:
        clfreq[s] += 1
    cllen = limited_lengths(clfreq, 7)
    clcode = canonical(cllen)
    ncl = 19
    while ncl > 4 and cllen[CL_ORDER[ncl - 1]] == 0:
        ncl -= 1
    b.put(1, 1)
    b.put(2, 2)
    b.put(nlit - 257, 5)
    b.put(ndist - 1, 5)
    b.put(ncl - 4, 4)
    for k in range(ncl):
        b.put(cllen[CL_ORDER[k]], 3)
    for s, x in rl:
        b.put(*clcode[s])
        if s == 16:
            b.put(x, 2)
        elif s == 17:
            b.put(x, 3)
        elif s == 18:
            b.put(x, 7)


def compress(data, lit, dist, container=0, cwindow=32, maxlen=10):
    """-> (status, bytes): one BTYPE = 10 block coded with (lit, dist); status 11 = a symbol without a code."""
    data = bytes(data)
    if len(data) < 5:
        return 1, b""
    lit, dist = [int(l) for l in lit], [int(l) for l in dist]
    lcode, dcode = canonical(lit), canonical(dist)
    b = _Bits()
    if container == 0:
        b.put(0x78, 8)
        b.put(0x9C, 8)
    elif container == 2:
        for v in (0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF):
            b.put(v, 8)
    _header(b, lit, dist)
    p = 0
    for m, d in parse(data, cwindow, maxlen):
        if m == 1:
            c = lcode[data[p]]
            if c[1] == 0:
                return 11, b""
            b.put(*c)
        else:
            c = lcode[254 + m]                      # lencode = mlength + 254, no extra bits (deflate.py:845-850)
            k = max(i for i in range(30) if DIST_BASE[i] <= d)
            e = dcode[k]
            if c[1] == 0 or e[1] == 0:
                return 11, b""
            b.put(*c)
            b.put(*e)
            b.put(d - DIST_BASE[k], 0 if k < 2 else (k >> 1) - 1)
        p += m
    if lcode[256][1] == 0:
        return 11, b""
    b.put(*lcode[256])
    body = b.tobytes()
    if container == 0:
        body += zlib.adler32(data).to_bytes(4, "big")
    elif container == 2:
        body += zlib.crc32(data).to_bytes(4, "little") + (len(data) & 0xFFFFFFFF).to_bytes(4, "little")
    return 0, body


def train(blocks, cwindow=32, maxlen=10):
    """Counts of the parse over `blocks` (+1 for every symbol the parse can produce) -> (lit, dist) lengths."""
    lf, df = [0] * 286, [0] * 30
    for blk in blocks:
        blk = bytes(blk)
        if len(blk) < 5:
            continue
        p = 0
        for m, d in parse(blk, cwindow, maxlen):
            if m == 1:
                lf[blk[p]] += 1
            else:
                lf[254 + m] += 1
                df[max(i for i in range(30) if DIST_BASE[i] <= d)] += 1
            p += m
    for s in range(256):
        lf[s] += 1
    lf[256] += len(blocks) + 1
    for s in range(257, 265):
        lf[s] += 1
    for c in range(10):
        df[c] += 1
    return limited_lengths(lf, 15), limited_lengths(df, 15)
