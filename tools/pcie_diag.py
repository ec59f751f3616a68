"""Developer tool: what the host link gives on this box (pinned H2D / D2H alone and together) next
to the two host-buffer calls of the e2e leg timed separately and overlapped (two contexts, two host
threads).  Writes one JSON line.

Under torchrun (WORLD_SIZE > 1) every rank runs the raw pinned copies on its own GPU at the same time
(after a barrier) and rank 0 prints the aggregate: what N concurrent e2e legs can reach on this host
at most.  python -m torch.distributed.run --nproc-per-node 8 tools/pcie_diag.py --raw-only"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import hdl_deflate_b200 as hz  # noqa: E402

BLOCK = 2048


def raw_multi():
    """All ranks at once: pinned H2D alone, D2H alone, both; aggregate GB/s over the ranks (max time)."""
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    nbytes = 1 << 31
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def wall(fns, reps=3):
        for f in fns:
            f()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    out = {"ranks": world, "bytes_per_rank_per_direction": nbytes, "cpus": os.cpu_count()}
    out["h2d_aggregate_gbps"] = world * nbytes / wall([h2d]) / 1e9
    out["d2h_aggregate_gbps"] = world * nbytes / wall([d2h]) / 1e9
    t = wall([h2d, d2h])
    out["both_aggregate_gbps_per_direction"] = world * nbytes / t / 1e9
    if dist.get_rank() == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    if "--raw-only" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return raw_multi()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    dev = torch.device("cuda", 0)
    eng = hz.Engine(0)
    out = {}
    nbytes = n * BLOCK
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def wall(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    out["h2d_gbps"] = nbytes / wall(lambda: d_a.copy_(h_a, non_blocking=True)) / 1e9
    out["d2h_gbps"] = nbytes / wall(lambda: h_b.copy_(d_b, non_blocking=True)) / 1e9

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
    out["both_each_gbps"] = nbytes / wall(both) / 1e9

    def chunked(csz):
        k = 0
        for o in range(0, nbytes, csz):
            with torch.cuda.stream(s1 if k % 2 == 0 else s2):
                d_a[o:o + csz].copy_(h_a[o:o + csz], non_blocking=True)
                h_b[o:o + csz].copy_(d_a[o:o + csz], non_blocking=True)
            k += 1
    for csz in (16 << 20, 48 << 20, 128 << 20):
        out["chunked_%dMB_roundtrip_gbps" % (csz >> 20)] = nbytes / wall(lambda: chunked(csz)) / 1e9

    # the e2e calls separately
    ostride = hz.compress_bound(BLOCK)
    eng.generate_blocks(d_a, BLOCK, BLOCK, n, first_block=0, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h_in = h_a.view(n, BLOCK)
    h_in.copy_(d_a.view(n, BLOCK))
    h_comp = torch.empty(n * ostride, dtype=torch.uint8, pin_memory=True)
    h_off = torch.zeros(n, dtype=torch.int64, pin_memory=True)
    h_back = h_b.view(n, BLOCK)
    h_clen = torch.zeros(n, dtype=torch.int32, pin_memory=True)
    h_blen = torch.zeros(n, dtype=torch.int32, pin_memory=True)
    h_st = torch.zeros(n, dtype=torch.int32, pin_memory=True)
    lib, ctx = eng._lib, eng._ctx
    total = ctypes.c_uint64(0)

    def comp():
        rc = lib.hdlz_compress_host_packed(ctx, h_in.data_ptr(), BLOCK, None, BLOCK, h_comp.data_ptr(), n * ostride,
                                           h_off.data_ptr(), h_clen.data_ptr(), h_st.data_ptr(), n, ctypes.byref(total))
        assert rc == 0

    def decomp():
        rc = lib.hdlz_decompress_host(ctx, h_comp.data_ptr(), h_off.data_ptr(), 0, h_clen.data_ptr(),
                                      h_back.data_ptr(), BLOCK, BLOCK, h_blen.data_ptr(), h_st.data_ptr(), n, 0)
        assert rc == 0
    tc = wall(comp)
    td = wall(decomp)
    assert torch.equal(h_back, h_in)
    out["compress_host_packed_ms"] = tc * 1e3
    out["decompress_host_ms"] = td * 1e3
    out["packed_bytes"] = int(total.value)
    out["e2e_gbps"] = 2 * nbytes / (tc + td) / 1e9
    # the two calls at the same time: a second context, two host threads (what bench.py's e2e leg does)
    from concurrent.futures import ThreadPoolExecutor
    eng2 = hz.Engine(0)
    h_comp2 = h_comp.clone().pin_memory()
    h_off2 = h_off.clone().pin_memory()
    h_clen2 = h_clen.clone().pin_memory()
    h_st2 = torch.zeros(n, dtype=torch.int32, pin_memory=True)

    def decomp2():
        rc = lib.hdlz_decompress_host(eng2._ctx, h_comp2.data_ptr(), h_off2.data_ptr(), 0, h_clen2.data_ptr(),
                                      h_back.data_ptr(), BLOCK, BLOCK, h_blen.data_ptr(), h_st2.data_ptr(), n, 0)
        assert rc == 0
    pool = ThreadPoolExecutor(2)

    def both_calls():
        fa, fb = pool.submit(comp), pool.submit(decomp2)
        fa.result()
        fb.result()
    tb = wall(both_calls)
    out["both_calls_overlapped_ms"] = tb * 1e3
    out["e2e_overlapped_gbps"] = 2 * nbytes / tb / 1e9
    # raw copies of the same byte counts, the same way: two threads, each H2D + D2H on its own stream pair
    pk = int(total.value)
    s3, s4 = torch.cuda.Stream(), torch.cuda.Stream()

    def raw_a():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_comp[:pk].copy_(d_b[:pk], non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    def raw_b():
        with torch.cuda.stream(s3):
            d_b[:pk].copy_(h_comp2[:pk], non_blocking=True)
        with torch.cuda.stream(s4):
            h_b.copy_(d_a, non_blocking=True)
        s3.synchronize()
        s4.synchronize()

    def raw_both():
        fa, fb = pool.submit(raw_a), pool.submit(raw_b)
        fa.result()
        fb.result()
    tr = wall(raw_both)
    out["raw_copies_same_bytes_overlapped_ms"] = tr * 1e3
    out["raw_copies_same_bytes_gbps_equiv"] = 2 * nbytes / tr / 1e9
    out["cpus"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
