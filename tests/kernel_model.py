"""Lane-level Python model of the CUDA compress kernel's decomposition.

Not a product path and not the oracle: a design check.  It restates, in the
same order and with the same bit tricks, what hdl-deflate_b200/csrc/
hdlz_compress.cu does per tile, so that the index arithmetic can be verified
against the oracle on the CPU (tests/test_kernel_model.py) before any GPU time
is spent.  See DESIGN.md "compress kernel" for the derivation.

  phase A  lane<->position : R[p] = 32-bit mask, bit (32-d) set iff x[p-d]==x[p]
                             (match_any inside the 32-chunk + per-value table
                             of the previous chunk, joined by one funnel shift);
                             R[q] = 0 for q >= L-2 encodes every end-of-stream
                             guard of SEARCH/SEARCHF (deflate.py:913-952,975-977)
  phase B  lane<->segment  : M3 = R[p]&R[p+1]&R[p+2]; nearest distance = clz+1;
                             length = 3 + #leading k in 3..9 with bit in R[p+k]
  phase P1 backward DP     : h[j] = exit skip-count of the greedy chain that
                             starts a token at j  (nibble shift register H)
  phase P2                 : entry skip-count of every segment (serial over H)
  phase P3 forward         : mark token starts, sum bits, scan, emit at offsets
"""

TILE = 1024
SEG = 32
MASK32 = 0xFFFFFFFF
PRIV_WORDS = ((SEG - 1) * 9 + 15 + 31) // 32        # 31 nine-bit literals + one 15-bit match token


def rev(v, n):
    r = 0
    for i in range(n):
        r |= ((v >> i) & 1) << (n - 1 - i)
    return r


def lit_token(x):
    if x < 144:
        return rev(0x30 + x, 8), 8
    return rev(0x190 + x - 144, 9), 9


def match_token(d, m):
    """(bits, nbits) of <length symbol 254+m><distance code of d>, LSB-first."""
    code = rev(m - 2, 7)                       # symbol 256+k -> 7-bit code k
    e = d - 1
    if e < 4:
        c, eb, extra = e, 0, 0
    else:
        msb = e.bit_length() - 1
        eb = msb - 1
        c = 2 * msb + ((e >> eb) & 1)
        extra = e & ((1 << eb) - 1)
    dc = rev(c, 5) | (extra << 5)
    return code | (dc << 7), 7 + 5 + eb


def funnelshift_r(lo, hi, s):
    return (((hi << 32) | lo) >> (s & 31)) & MASK32


def funnelshift_l(lo, hi, s):
    s &= 31
    return ((((hi << 32) | lo) << s) >> 32) & MASK32


def clz32(v):
    return 32 - v.bit_length()


def phase_a(x, L, t0):
    """R for positions t0 .. t0+TILE+31 (absolute index q - t0), adler partials."""
    nchunk = TILE // 32 + 1
    R = [0] * (nchunk * 32)
    table = {}
    prev_vals = None
    s1 = s2 = 0
    n_tile = min(TILE, L - t0)
    for c in range(-1, nchunk):
        base = t0 + 32 * c
        vals = [x[base + l] if 0 <= base + l < L else None for l in range(32)]
        # match_any: mask of lanes holding the same value (invalid lanes excluded)
        m_cur = [0] * 32
        for l in range(32):
            if vals[l] is None:
                continue
            for l2 in range(32):
                if vals[l2] == vals[l]:
                    m_cur[l] |= 1 << l2
        m_prev = [table.get(vals[l], 0) if vals[l] is not None else 0 for l in range(32)]
        if prev_vals is not None:
            for v in prev_vals:
                if v is not None:
                    table[v] = 0
        for l in range(32):
            if vals[l] is not None:
                table[vals[l]] = m_cur[l]
        prev_vals = vals
        if c < 0:
            continue
        for l in range(32):
            q = base + l
            r = funnelshift_r(m_prev[l], m_cur[l], l)
            if q >= L - 2:
                r = 0
            R[32 * c + l] = r
            if c < TILE // 32 and q < L:
                s1 += vals[l]
                s2 += vals[l] * (n_tile - (32 * c + l))
    return R, s1, s2, n_tile


def phase_b(x, L, t0, R):
    """tok[i] = (code, nbits, length) for tile positions 0..TILE-1."""
    tok = []
    for i in range(TILE):
        p = t0 + i
        if p >= L:
            tok.append((0, 0, 1))
            continue
        m3 = R[i] & R[i + 1] & R[i + 2]
        if m3:
            cl = clz32(m3)
            d = cl + 1
            t = 0x80000000 >> cl
            c = t
            ln = 3
            for k in range(3, 10):
                c &= R[i + k]
                ln += 1 if c else 0
            code, nb = match_token(d, ln)
            tok.append((code, nb, ln))
        else:
            code, nb = lit_token(x[p])
            tok.append((code, nb, 1))
    return tok


def phase_p1(tok):
    """Per segment: H = nibbles h[0..9] (exit state when a token starts at j)."""
    Hs = []
    for g in range(TILE // SEG):
        H = 0
        for j in range(SEG - 1, -1, -1):
            ln = tok[g * SEG + j][2]
            if j + ln >= SEG:
                hn = j + ln - SEG
            else:
                hn = (H >> (4 * (ln - 1))) & 15
            H = ((H << 4) | hn) & ((1 << 64) - 1)
        Hs.append(H & ((1 << 40) - 1))
    return Hs


def phase_p2(Hs, carry):
    entries = []
    cur = carry
    for H in Hs:
        entries.append(cur)
        if cur < SEG:      # always (cur <= 9)
            cur = (H >> (4 * cur)) & 15
    return entries, cur


def phase_p3(tok, entries, bitbase, words):
    """v2: ONE forward pass per segment writes the segment's started tokens into a lane-private
    word array (flush check every second position: fill < 32 + 2*15 < 64), then a merge pass
    shifts each private stream to its bit offset in the tile's stream (first/last word OR-ed,
    interior words stored).  Returns the new bit position."""
    nseg = TILE // SEG
    priv, nbits = [], []
    for g in range(nseg):
        r = entries[g]
        acc = 0
        fill = 0
        pw = []
        for j in range(SEG):
            code, nb, ln = tok[g * SEG + j]
            if r == 0:
                acc |= code << fill
                fill += nb
                r = ln - 1
            else:
                r -= 1
            if (j & 1) and fill >= 32:
                pw.append(acc & MASK32)
                acc >>= 32
                fill -= 32
        total = 32 * len(pw) + fill
        if fill:
            pw.append(acc & MASK32)          # fill < 32 here
        assert fill < 32
        assert len(pw) <= PRIV_WORDS, (len(pw), total)      # the kernel's kPrivWords
        priv.append(pw)
        nbits.append(total)
    pos = bitbase
    for g in range(nseg):
        sh = pos & 31
        w0 = pos >> 5
        nb = nbits[g]
        if nb:
            nwords_out = (sh + nb + 31) >> 5           # words of the final stream this lane touches
            prev = 0
            for i in range(nwords_out):
                cur = priv[g][i] if i < len(priv[g]) else 0
                val = funnelshift_l(prev, cur, sh)    # (cur << sh) | (prev >> (32 - sh))
                prev = cur
                first = i == 0 and sh != 0
                last = i == nwords_out - 1 and ((sh + nb) & 31) != 0
                if first or last:
                    words[w0 + i] = words.get(w0 + i, 0) | val       # atomicOr
                else:
                    assert words.get(w0 + i, 0) == 0
                    words[w0 + i] = val                               # plain store
        pos += nb
    return pos


def compress(data):
    x = bytes(data)
    L = len(x)
    assert L >= 5
    words = {0: 0x78 | (0x9C << 8) | (3 << 16)}
    bitpos = 19
    carry = 0
    a, b = 1, 0
    t0 = 0
    while t0 < L:
        R, s1, s2, n = phase_a(x, L, t0)
        tok = phase_b(x, L, t0, R)
        Hs = phase_p1(tok)
        entries, carry = phase_p2(Hs, carry)
        bitpos = phase_p3(tok, entries, bitpos, words)
        b = (b + n * a + s2) % 65521
        a = (a + s1) % 65521
        t0 += TILE
    bitpos += 7                                   # EOB
    nbytes = (bitpos + 7) >> 3
    out = bytearray(nbytes + 4)
    for w, v in words.items():
        for k in range(4):
            if 4 * w + k < nbytes:
                out[4 * w + k] = (v >> (8 * k)) & 255
    out[nbytes:nbytes + 4] = bytes([b >> 8, b & 255, a >> 8, a & 255])
    return bytes(out)
