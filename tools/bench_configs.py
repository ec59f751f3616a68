#!/usr/bin/env python
"""Only the decompress-side BASELINE configs of bench.py (configs[2] Z_FIXED 2 KiB streams, configs[3]
100 000 x 32 KiB level-6 streams) — the same functions bench.py folds into its JSON line, for quick
iteration on the inflate kernels.  One JSON line per config."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import __graft_entry__  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1 << 20)
    ap.add_argument("--streams", type=int, default=100000)
    ap.add_argument("--distinct", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    __graft_entry__.build()
    import hdl_deflate_b200 as hz
    eng = hz.Engine(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    peak, _ = bench.measured_peak()
    if args.only in ("", "config3"):
        t0 = time.time()
        n = args.blocks
        d_plain = torch.empty(n * bench.BLOCK, dtype=torch.uint8, device=dev)
        eng.generate_blocks(d_plain, bench.BLOCK, bench.BLOCK, n, stream=stream)
        torch.cuda.synchronize()
        r = bench.run_config3(eng, hz, d_plain, n, dev, stream, args.steps, args.threads, peak)
        r["name"], r["wall_s"] = "config3", time.time() - t0
        print(json.dumps(r), flush=True)
        del d_plain
    if args.only in ("", "config4"):
        t0 = time.time()
        r = bench.run_config4(eng, hz, args.streams, min(args.distinct, args.streams), dev, stream, args.steps,
                              args.threads, peak)
        r["name"], r["wall_s"] = "config4", time.time() - t0
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
