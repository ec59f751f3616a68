"""Developer tool: what the host link gives on this box (pinned H2D / D2H alone and together) next
to the two host-buffer calls of the e2e leg timed separately.  Writes one JSON line."""
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


def main():
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
    out["cpus"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
