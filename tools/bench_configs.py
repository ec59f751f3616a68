#!/usr/bin/env python
"""Decompress-side BASELINE configs at full size on one B200:

  config3 : 2^20 x 2 KiB blocks as host-zlib Z_FIXED streams (level 6, wbits 15), packed with an
            offset array (offsets rounded up to 4 bytes), decompress-only, byte-exact vs the blocks.
  config4 : 100 000 x 32 KiB plain, host-zlib level 6 (dynamic trees), fixed-stride slots,
            decompress-only, OBSIZE = 32768.  The 100 000 streams are drawn cyclically from
            --distinct distinct ones (default 16384, ~200 MB compressed: larger than the L2).

Prints one JSON line per config (also what profiles/r01_configs.json holds).  Host zlib is only
the fixture generator here; the timed region is the CUDA decompressor (CUDA events, 3 warm-ups).
"""
import argparse
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
import hdl_deflate_b200 as hz  # noqa: E402
from oracle import hdlz_oracle as O  # noqa: E402  (fixture generation: host zlib through its batch driver)

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0          # fallback of B200_PROFILING.md


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ts = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    return sum(ts) / steps, min(ts)


def host_zlib(plain, level, strategy, cap, threads):
    n, L = plain.shape
    comp = np.empty((n, cap), dtype=np.uint8)
    clen, st = O.batch(O.KIND_ZLIB_DEFLATE, plain, np.arange(n, dtype=np.uint64) * L, np.full(n, L, np.uint32), comp,
                       np.arange(n, dtype=np.uint64) * cap, cap, threads, level, strategy)
    assert not st.any()
    return comp, clen


def config3(eng, args):
    n, L = args.blocks, 2048
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    d_plain = torch.empty(n * L, dtype=torch.uint8, device=dev)
    eng.generate_blocks(d_plain, L, L, n, stream=s)
    torch.cuda.synchronize()
    plain = d_plain.view(n, L).cpu().numpy()
    t0 = time.time()
    comp, clen = host_zlib(plain, 6, zlib.Z_FIXED, 2560, args.threads)
    t_zlib = time.time() - t0
    # pack: offsets rounded up to 4 bytes
    padded = (clen.astype(np.int64) + 3) & ~3
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum(padded[:-1])
    total = int(off[-1] + padded[-1])
    packed = np.zeros(total + 16, dtype=np.uint8)
    idx = np.arange(2560)[None, :]
    mask = idx < clen[:, None]
    packed[(off[:, None].astype(np.int64) + idx)[mask]] = comp[mask]
    d_in = torch.from_numpy(packed).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_len = torch.from_numpy(clen.astype(np.int32)).to(dev)
    d_out = torch.empty(n * L, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)

    def run():
        eng.decompress_batch(d_in, d_off, 0, d_len, d_out, L, L, d_olen, d_st, n, flags=0, stream=s)
    ms, best = timed(run, args.steps)
    assert int(d_st.abs().sum()) == 0 and bool((d_olen == L).all())
    assert torch.equal(d_out, d_plain), "config3: output differs from the original blocks"
    cbytes = int(clen.sum())
    alg = cbytes + n * L
    return {"config": "configs[2]: %d x 2 KiB zlib Z_FIXED streams (host zlib 1.3 level 6), packed, 4-byte aligned "
                      "offsets, decompress-only" % n,
            "decompress_gbps": n * L / (ms * 1e-3) / 1e9, "ms": ms, "ms_best": best,
            "compressed_ratio": cbytes / (n * L),
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": PEAK, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / PEAK, "algorithmic_bytes": alg},
            "byte_exact_vs_original": True,
            "host_zlib_deflate_gbps": n * L / t_zlib / 1e9, "host_threads": args.threads}


def config4(eng, args):
    n, L, nd = args.streams, 32768, args.distinct
    rng = np.random.default_rng(4)
    # Zipf-like bytes over 64 symbols, plus repeats at distances up to 32 KiB
    p = 1.0 / np.arange(1, 65) ** 1.1
    p /= p.sum()
    plain = rng.choice(64, size=(nd, L), p=p).astype(np.uint8) + 32
    for k in range(6):
        src = rng.integers(0, L - 4096, nd)
        dst = rng.integers(0, L - 4096, nd)
        ln = rng.integers(64, 4096, nd)
        for i in range(nd):
            plain[i, dst[i]:dst[i] + ln[i]] = plain[i, src[i]:src[i] + ln[i]].copy()
    t0 = time.time()
    comp, clen = host_zlib(plain, 6, 0, 36864, args.threads)
    t_zlib = time.time() - t0
    btype = (comp[:, 2] >> 1) & 3
    assert (btype == 2).all(), "config4 streams must start with a dynamic block"
    stride = (int(clen.max()) + 15) & ~15
    comp = np.ascontiguousarray(comp[:, :stride])
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    sel = torch.arange(n, device=dev) % nd
    d_in = torch.from_numpy(comp).to(dev)[sel].contiguous()              # n x stride
    d_len = torch.from_numpy(clen.astype(np.int32)).to(dev)[sel].contiguous()
    d_plain = torch.from_numpy(plain).to(dev)
    d_out = torch.empty(n * L, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)

    def run():
        eng.decompress_batch(d_in, None, stride, d_len, d_out, L, L, d_olen, d_st, n, flags=hz.F_PERSIST_TABLES, stream=s)
    ms, best = timed(run, args.steps)
    assert int(d_st.abs().sum()) == 0 and bool((d_olen == L).all())
    assert torch.equal(d_out.view(n, L), d_plain[sel]), "config4: output differs from the original"
    cbytes = int(d_len.sum(dtype=torch.int64))
    alg = cbytes + n * L
    return {"config": "configs[3]: %d x 32 KiB zlib level-6 dynamic-tree streams (drawn cyclically from %d distinct), "
                      "OBSIZE = 32768, decompress-only" % (n, nd),
            "decompress_gbps": n * L / (ms * 1e-3) / 1e9, "ms": ms, "ms_best": best,
            "compressed_ratio": cbytes / (n * L),
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": PEAK, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / PEAK, "algorithmic_bytes": alg},
            "byte_exact_vs_original": True,
            "host_zlib_deflate_gbps": nd * L / t_zlib / 1e9, "host_threads": args.threads}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1 << 20)
    ap.add_argument("--streams", type=int, default=100000)
    ap.add_argument("--distinct", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    eng = hz.Engine(0)
    for name, fn in (("config3", config3), ("config4", config4)):
        if args.only and args.only != name:
            continue
        t0 = time.time()
        r = fn(eng, args)
        r["name"] = name
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
