"""Developer tool: which ingredient of the host pipelines costs PCIe throughput?  Moves the byte counts of one
e2e step (2 GiB of blocks + 1.27 GB of streams, each both ways) with raw cudaMemcpyAsync in several patterns and
prints ms per pattern.  (The e2e leg of bench.py reaches 47 GB/s; raw 48 MiB pieces on two streams 59.)"""
import json
import sys
import time

import torch

dev = torch.device("cuda:0")
BIG, SMALL = 2 << 30, 1266 << 20
h_a = torch.empty(BIG, dtype=torch.uint8, pin_memory=True)
h_b = torch.empty(SMALL, dtype=torch.uint8, pin_memory=True)
h_c = torch.empty(BIG, dtype=torch.uint8, pin_memory=True)
h_d = torch.empty(SMALL, dtype=torch.uint8, pin_memory=True)
h_s = torch.empty(1 << 20, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(BIG, dtype=torch.uint8, device=dev)
d_b = torch.empty(SMALL, dtype=torch.uint8, device=dev)
d_s = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
d_w = torch.empty(64 << 20, dtype=torch.uint8, device=dev)


def run(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"pattern": name, "ms": round(ms, 2), "gbps_equiv": round(2 * BIG / ms / 1e6, 1)}), flush=True)


def two_streams(piece):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def f():
        with torch.cuda.stream(s1):
            for o in range(0, BIG, piece):
                d_a[o:o + piece].copy_(h_a[o:o + piece], non_blocking=True)
            for o in range(0, SMALL, piece):
                d_b[o:o + piece].copy_(h_b[o:o + piece], non_blocking=True)
        with torch.cuda.stream(s2):
            for o in range(0, BIG, piece):
                h_c[o:o + piece].copy_(d_a[o:o + piece], non_blocking=True)
            for o in range(0, SMALL, piece):
                h_d[o:o + piece].copy_(d_b[o:o + piece], non_blocking=True)
    return f


def pipelines(nstreams, nchunks, small_copies, kernel, sync_depth):
    """Two 'calls' (compress-like: big up, small down; decompress-like: small up, big down), each over nchunks chunks
    round-robin on its own nstreams streams; per chunk: up copy, [kernel], [small copies down], down copy."""
    sa = [torch.cuda.Stream() for _ in range(nstreams)]
    sb = [torch.cuda.Stream() for _ in range(nstreams)]
    ca, cb = BIG // nchunks & ~15, SMALL // nchunks & ~15

    def f():
        for k in range(nchunks):
            for streams, up_h, up_d, up_n, dn_h, dn_d, dn_n in ((sa, h_a, d_a, ca, h_d, d_b, cb), (sb, h_b, d_b, cb, h_c, d_a, ca)):
                s = streams[k % nstreams]
                if sync_depth and k >= nstreams:
                    s.synchronize()
                with torch.cuda.stream(s):
                    up_d[k * up_n:(k + 1) * up_n].copy_(up_h[k * up_n:(k + 1) * up_n], non_blocking=True)
                    if kernel:
                        d_w[:8 << 20].add_(1)
                    for j in range(small_copies):
                        h_s[j * 65536:(j + 1) * 65536].copy_(d_s[j * 65536:(j + 1) * 65536], non_blocking=True)
                    dn_h[k * dn_n:(k + 1) * dn_n].copy_(dn_d[k * dn_n:(k + 1) * dn_n], non_blocking=True)
    return f


run("two streams, 48 MiB pieces", two_streams(48 << 20))
run("two streams, 16 MiB pieces", two_streams(16 << 20))
run("two streams, 256 MiB pieces", two_streams(256 << 20))
for ns in (2, 6):
    run("2 x %d streams, 91 chunks" % ns, pipelines(ns, 91, 0, False, False))
    run("2 x %d streams, 91 chunks, bounded depth" % ns, pipelines(ns, 91, 0, False, True))
    run("2 x %d streams, 91 chunks, 4 small copies" % ns, pipelines(ns, 91, 4, False, False))
    run("2 x %d streams, 91 chunks, kernel" % ns, pipelines(ns, 91, 0, True, False))
    run("2 x %d streams, 91 chunks, kernel + small + bounded" % ns, pipelines(ns, 91, 4, True, True))
run("2 x 6 streams, 23 chunks", pipelines(6, 23, 0, False, False))
run("2 x 6 streams, 23 chunks, kernel + small + bounded", pipelines(6, 23, 4, True, True))
run("2 x 3 streams, 364 chunks, kernel + small + bounded", pipelines(3, 364, 4, True, True))
