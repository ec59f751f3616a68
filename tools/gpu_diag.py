"""GPU bring-up diagnostics (developer tool): compares the CUDA compressor with the oracle and,
on a mismatch, prints the first differing token of both streams.  Writes gpurun_out/diag.txt."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import hdl_deflate_b200 as hz  # noqa: E402
from hdl_deflate_b200 import workload  # noqa: E402
from oracle import hdlz_oracle  # noqa: E402

LB = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LE = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DB = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
      6145, 8193, 12289, 16385, 24577]


def tokens(stream):
    """Token list [(bitpos, 'L', byte) | (bitpos, 'M', len, dist) | (bitpos, 'E')] of a one-fixed-block stream."""
    bits = int.from_bytes(stream[2:], "little")
    pos = 3
    out = []

    def take(n):
        nonlocal pos
        v = (bits >> pos) & ((1 << n) - 1)
        pos += n
        return v

    def takerev(n):
        v = 0
        for _ in range(n):
            v = (v << 1) | take(1)
        return v
    limit = 8 * (len(stream) - 2)
    while pos < limit:
        p0 = pos
        c = takerev(7)
        if c <= 0x17:
            sym = 256 + c
        else:
            c = (c << 1) | take(1)
            if 0x30 <= c <= 0xBF:
                sym = c - 0x30
            elif 0xC0 <= c <= 0xC7:
                sym = 280 + c - 0xC0
            else:
                c = (c << 1) | take(1)
                sym = 144 + c - 0x190
        if sym < 256:
            out.append((p0, "L", sym))
        elif sym == 256:
            out.append((p0, "E"))
            break
        else:
            t = sym - 257
            ln = LB[t] + take(LE[t])
            dc = takerev(5)
            eb = 0 if dc < 2 else (dc >> 1) - 1
            out.append((p0, "M", ln, DB[dc] + take(eb)))
    return out


def explain(data, got, want, log):
    n = min(len(got), len(want))
    first = next((i for i in range(n) if got[i] != want[i]), n)
    log("  len got %d want %d, first differing byte %d" % (len(got), len(want), first))
    try:
        tg, tw = tokens(got), tokens(want)
    except Exception as e:       # garbage stream
        log("  token decode failed: %r" % (e,))
        return
    ip = 0
    for k in range(min(len(tg), len(tw))):
        if tg[k] != tw[k]:
            log("  token %d at input pos %d: got %r want %r" % (k, ip, tg[k], tw[k]))
            log("  context got : %r" % (tg[max(0, k - 3):k + 4],))
            log("  context want: %r" % (tw[max(0, k - 3):k + 4],))
            log("  input around: %r" % (data[max(0, ip - 34):ip + 12],))
            return
        ip += tw[k][2] if tw[k][1] == "M" else 1
    log("  tokens equal for %d tokens; counts got %d want %d" % (min(len(tg), len(tw)), len(tg), len(tw)))


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    f = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")

    def log(s):
        print(s)
        f.write(s + "\n")
        f.flush()
    eng = hz.Engine(0)
    bad = 0
    cases = [("abcde", b"abcde"), ("a12", b"a" * 12), ("abc6", b"abcabcabcabcabcabc"), ("zeros2048", bytes(2048)),
             ("ramp", bytes(range(256)) * 8), ("ab", b"ab" * 1024)]
    cases += [("wl%d" % n, workload.block(n, n)) for n in (5, 6, 7, 8, 31, 32, 33, 34, 63, 64, 65, 66, 100, 500, 2047,
                                                            2048, 2049, 2050, 2080, 2081, 4096, 4133, 10000, 70000)]
    for name, data in cases:
        want = hdlz_oracle.compress(data)[1]
        try:
            got = eng.compress(data)
        except Exception as e:
            log("FAIL %s: exception %r" % (name, e))
            bad += 1
            continue
        if got != want:
            bad += 1
            log("FAIL %s (L=%d)" % (name, len(data)))
            explain(data, got, want, log)
        else:
            log("ok   %s (L=%d -> %d)" % (name, len(data), len(got)))
    # batch
    small = os.environ.get('HDLZ_DIAG_SMALL') == '1'
    nblk = 64 if small else int(os.environ.get('HDLZ_DIAG_BLOCKS', '4096'))
    if small:
        cases = cases[:12]
    blocks = workload.blocks(0, nblk, 2048)
    arr = np.frombuffer(b"".join(blocks), dtype=np.uint8).reshape(nblk, 2048)
    t = time.time()
    out, out_len, status = eng.compress_host(arr)
    log("batch %d blocks: %.3fs, status nonzero %d" % (nblk, time.time() - t, int((status != 0).sum())))
    nbad = 0
    bad_idx = []
    for i in range(nblk):
        want = hdlz_oracle.compress(blocks[i])[1]
        got = out[i, :out_len[i]].tobytes()
        if got != want:
            nbad += 1
            bad_idx.append(i)
            if nbad <= 4:
                log("FAIL batch block %d" % i)
                explain(blocks[i], got, want, log)
    log("bad block indices (first 40): %r" % (bad_idx[:40],))
    log("batch mismatches: %d / %d" % (nbad, nblk))
    bad += nbad
    # inflate
    import zlib
    ibad = 0
    for name, data in cases:
        for lvl, strat in ((6, 0), (6, zlib.Z_FIXED), (0, 0)):
            co = zlib.compressobj(lvl, zlib.DEFLATED, 15, 8, strat)
            z = co.compress(data) + co.flush()
            try:
                got = eng.decompress(z)
            except Exception as e:
                log("FAIL inflate %s lvl %d strat %d: %r" % (name, lvl, strat, e))
                ibad += 1
                continue
            if got != data:
                ibad += 1
                n = min(len(got), len(data))
                first = next((i for i in range(n) if got[i] != data[i]), n)
                log("FAIL inflate %s lvl %d strat %d: len %d want %d first diff %d" % (name, lvl, strat, len(got),
                                                                                    len(data), first))
    log("inflate mismatches: %d" % ibad)
    log("TOTAL FAILURES: %d" % (bad + ibad))
    return 1 if bad + ibad else 0


if __name__ == "__main__":
    sys.exit(main())
