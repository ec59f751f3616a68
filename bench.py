#!/usr/bin/env python
"""bench.py — headline measurement of the deflate hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--blocks B]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[1] + its decompress twin configs[2], per GPU; configs[4] when N > 1):
  B = 2^20 independent 2 KiB "random+repeat" blocks (hdl-deflate_b200/workload.py), resident in HBM.
One step = one compress pass over the B blocks (FAST+MATCH10 static-tree format, bit-exact with
deflate.py) followed by one decompress pass over the B streams it produced.
  value [GB/s] = uncompressed bytes through both passes / step time = 2 * B * 2048 * N / t_step
  compress_gbps / decompress_gbps = B * 2048 * N / t_pass, reported beside it.
Timing: CUDA events on the launching stream, max over ranks; each pass touches > 4 GB, far beyond
the 126 MB L2, so no flush is needed between iterations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 2048
METRIC = "deflate GB/s (compress+decompress)"
UNIT = "GB/s"


def workload_name(n_blocks):
    return "%d x 2 KiB random+repeat blocks per GPU (BASELINE configs[1]), compress FAST+MATCH10 static tree, " \
           "then decompress of the produced streams (configs[2]/[4])" % n_blocks


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, n_blocks):
    """dram bytes per launch from the committed ncu capture, if it was taken at this batch size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            e = t.get(kernel)
            if e and int(e["n_blocks"]) == int(n_blocks):
                return float(e["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


# ----------------------------------------------------------------------------------------------
# CPU arms: the oracle port (C restatement of deflate.py's compress + inflate) and host zlib
# ----------------------------------------------------------------------------------------------
def compress_bound_py(n):
    """hdlz_compress_bound restated (2 + ceil((3 + 9n + 7) / 8) + 4, rounded up to 16): the CPU arms must not
    load the product library."""
    return ((6 + (3 + 9 * n + 7 + 7) // 8) + 15) & ~15


def cpu_roundtrip(sample, nthreads, repeat=1):
    """Time the CPU arm on `sample` (uint8 [n, 2048]) with `nthreads` threads: compress = the tuned port
    of deflate.py's compressor (oracle/hdlz_oracle.c hdlz_oracle_compress_fast: SSE2 compares, 64-bit bit
    buffer, zlib's adler32; same bytes as the plain restatement), decompress = the faster of the port's
    inflate and host zlib's inflate.
    -> dict(value, compress_gbps, decompress_gbps, seconds, inflate = which inflater won, ...)."""
    import numpy as np
    from oracle import hdlz_oracle as O
    n = sample.shape[0]
    ostride = compress_bound_py(BLOCK)
    comp = np.zeros((n, ostride), dtype=np.uint8)
    back = np.zeros((n, BLOCK), dtype=np.uint8)
    ioff = np.arange(n, dtype=np.uint64) * BLOCK
    ooff = np.arange(n, dtype=np.uint64) * ostride
    lens = np.full(n, BLOCK, dtype=np.uint32)
    tc = tp = tz = 0.0
    for _ in range(repeat):
        t0 = time.perf_counter()
        clen, st = O.batch(O.KIND_FAST_COMPRESS, sample, ioff, lens, comp, ooff, ostride, nthreads)
        t1 = time.perf_counter()
        blen, st2 = O.batch(O.KIND_PORT_INFLATE, comp, ooff, clen, back, ioff, BLOCK, nthreads)
        t2 = time.perf_counter()
        assert not st.any() and not st2.any() and np.array_equal(back, sample)
        t3 = time.perf_counter()
        blen, st3 = O.batch(O.KIND_ZLIB_INFLATE, comp, ooff, clen, back, ioff, BLOCK, nthreads)
        t4 = time.perf_counter()
        assert not st3.any() and np.array_equal(back, sample)
        tc += t1 - t0
        tp += t2 - t1
        tz += t4 - t3
    td = min(tp, tz)
    byt = n * BLOCK * repeat
    return {"value": 2 * byt / (tc + td) / 1e9, "compress_gbps": byt / tc / 1e9, "decompress_gbps": byt / td / 1e9,
            "seconds": tc + tp + tz, "inflate": "zlib" if tz <= tp else "port",
            "port_inflate_gbps": byt / tp / 1e9, "zlib_inflate_gbps": byt / tz / 1e9,
            "kind": "port+zlib" if tz <= tp else "port"}


def cpu_zlib(sample, nthreads):
    import numpy as np
    import zlib
    from oracle import hdlz_oracle as O
    n = sample.shape[0]
    ostride = 2560
    comp = np.empty((n, ostride), dtype=np.uint8)
    back = np.empty((n, BLOCK), dtype=np.uint8)
    ioff = np.arange(n, dtype=np.uint64) * BLOCK
    ooff = np.arange(n, dtype=np.uint64) * ostride
    lens = np.full(n, BLOCK, dtype=np.uint32)
    t0 = time.perf_counter()
    clen, _ = O.batch(O.KIND_ZLIB_DEFLATE, sample, ioff, lens, comp, ooff, ostride, nthreads, 6, zlib.Z_FIXED)
    t1 = time.perf_counter()
    O.batch(O.KIND_ZLIB_INFLATE, comp, ooff, clen, back, ioff, BLOCK, nthreads)
    t2 = time.perf_counter()
    byt = n * BLOCK
    return {"deflate_zfixed_l6_gbps": byt / (t1 - t0) / 1e9, "inflate_gbps": byt / (t2 - t1) / 1e9,
            "ratio": float(clen.sum()) / byt}


def reference_sim_rate():
    """Cycles per input byte of the reference FSM itself (deflate.py under the MyHDL-compat layer), from the
    `cycles` field oracle/make_golden.py recorded while it produced the 2 KiB fixtures; the simulation cannot
    run on the GPU box.  README.md of the reference quotes 100 MHz for the FPGA."""
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "compress_golden.json")))
        cs = [c for c in g["cases"] if c["len"] == BLOCK and "cycles" in c]
        cyc = sum(c["cycles"] for c in cs)
        byt = sum(c["len"] for c in cs)
        return {"what": "deflate.py FSM, compress, %d golden 2 KiB blocks (tests/golden/compress_golden.json)" % len(cs),
                "cycles_per_byte": cyc / byt, "fpga_100mhz_gbps": 0.1 / (cyc / byt)}
    except Exception:
        return None


def host_sample(n):
    import importlib.util
    import numpy as np
    # the pure-Python workload definition, loaded by path: importing the package would pull in the ctypes binding
    spec = importlib.util.spec_from_file_location("hdlz_workload", os.path.join(ROOT, "hdl-deflate_b200", "workload.py"))
    workload = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(workload)
    return np.frombuffer(b"".join(workload.blocks(0, n, BLOCK)), dtype=np.uint8).reshape(n, BLOCK)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path.  deflate.py is a Python/MyHDL simulation that
    cannot travel to the GPU box (and runs at ~1.5 kB/s), so the arm times the oracle port of it
    (oracle/hdlz_oracle.c, checked byte-for-byte against the executing reference) on all host threads."""
    if rank != 0:
        return
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])     # the CPU arm only: no libhdlz.so here
    cores = os.cpu_count() or 1
    n = args.ref_blocks
    sample = host_sample(min(n, 4096))
    import numpy as np
    if n > sample.shape[0]:
        sample = np.tile(sample, ((n + sample.shape[0] - 1) // sample.shape[0], 1))[:n].copy()
    for _ in range(args.warmup):
        cpu_roundtrip(sample, cores)
    t0 = time.perf_counter()
    tot_c = tot_d = 0.0
    r = None
    for _ in range(args.steps):
        r = cpu_roundtrip(sample, cores)
        tot_c += n * BLOCK / r["compress_gbps"] / 1e9
        tot_d += n * BLOCK / r["decompress_gbps"] / 1e9
    t = time.perf_counter() - t0
    byt = n * BLOCK * args.steps
    value = 2 * byt / (tot_c + tot_d) / 1e9
    sample_desc = "%d of the config's 2 KiB blocks per step (4096 distinct, tiled; bounded sample), %d steps" % (n, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tot_c + tot_d) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.blocks), "sample": sample_desc},
        "compress_gbps": byt / tot_c / 1e9, "decompress_gbps": byt / tot_d / 1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample_desc,
                         "what": "compress: tuned C port of deflate.py's FAST+MATCH10 compressor (bit-identical "
                                 "output); decompress: the faster of the port's inflate and host zlib inflate (%s won: "
                                 "port %.2f, zlib %.2f GB/s)" % (r["inflate"], r["port_inflate_gbps"], r["zlib_inflate_gbps"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# BASELINE configs[2] and [3] at full size (decompress-only), folded into the same JSON line.
# Host zlib (Python's, on a thread pool) is only the fixture generator; the timed region is
# hdlz_decompress_batch on resident device buffers (CUDA events, 3 warm-ups).
# ----------------------------------------------------------------------------------------------
def _zlib_many(rows, level, strategy, threads):
    """Host-zlib streams of the rows of a uint8 [n, L] array, on all cores through tools/zfixture.c (built on
    demand; plumbing, neither product nor oracle) -> (packed uint8 with 4-byte aligned starts, off i64, len u32)."""
    import ctypes
    import numpy as np
    so = os.path.join(ROOT, "tools", "libzfixture.so")
    src = os.path.join(ROOT, "tools", "zfixture.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lz", "-lpthread"])
    lib = ctypes.CDLL(so)
    lib.zfix_deflate_packed.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.POINTER(ctypes.c_uint64)]
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    n, L = rows.shape
    cap = n * (L + L // 8 + 64) + 16
    out = np.empty(cap, dtype=np.uint8)
    off = np.zeros(n, dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint32)
    total = ctypes.c_uint64(0)
    rc = lib.zfix_deflate_packed(rows.ctypes.data, n, L, level, strategy, threads, out.ctypes.data, cap, off.ctypes.data,
                                 lens.ctypes.data, ctypes.byref(total))
    assert rc == 0, "zfixture failed (%d)" % rc
    return out[:total.value + 16], off.astype(np.int64), lens


def _time_launches(fn, steps, warmup=3):
    import torch
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


def run_config3(eng, hz, d_plain, n, dev, stream, steps, threads, peak):
    """configs[2]: n x 2 KiB blocks as host-zlib Z_FIXED streams (level 6), packed with a 4-byte aligned
    offset array, decompress-only; byte-exact against the blocks before anything is timed."""
    import zlib
    import torch
    t0 = time.time()
    plain = d_plain.view(n, BLOCK).cpu().numpy()
    packed, off, lens = _zlib_many(plain, 6, zlib.Z_FIXED, threads)
    t_fix = time.time() - t0
    d_in = torch.from_numpy(packed).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_len = torch.from_numpy(lens.astype("int32")).to(dev)
    d_out = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)

    def run(flags=0):
        eng.decompress_batch(d_in, d_off, 0, d_len, d_out, BLOCK, BLOCK, d_olen, d_st, n, flags=flags, stream=stream)
    run(hz.F_VERIFY_ADLER)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0 and bool((d_olen == BLOCK).all()), "config3: status / length"
    assert torch.equal(d_out, d_plain), "config3: output differs from the original blocks"
    d_out.zero_()
    ms, best = _time_launches(run, steps)
    assert torch.equal(d_out, d_plain)
    ms_v, _ = _time_launches(lambda: run(hz.F_VERIFY_ADLER), steps)
    cbytes = int(lens.sum())
    alg = cbytes + n * BLOCK
    return {"workload": "BASELINE configs[2]: %d x 2 KiB host-zlib Z_FIXED streams (level 6, wbits 15), packed, 4-byte "
                        "aligned offsets, decompress-only" % n,
            "decompress_gbps": n * BLOCK / (ms * 1e-3) / 1e9, "decompress_gbps_verified": n * BLOCK / (ms_v * 1e-3) / 1e9,
            "ms": ms, "ms_best": best, "compressed_ratio": cbytes / (n * BLOCK), "byte_exact_vs_original": True,
            "roofline": {"bound": "hbm", "kernel": "k_inflate_lanes<0>", "achieved": alg / (ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg, "traffic": ncu_traffic("config3", n)},
            "fixture_seconds": t_fix}


def run_tree(eng, hz, d_plain, n, dev, stream, steps):
    """SURVEY 8(f) rank 4 (README.md:43-45): the same blocks, and the same blocks folded onto 16 byte values
    ("data sets with just a small set of used byte values"), compressed with a code trained on the batch
    (hdlz_train_tree: the kernel counts the symbols of its parse, the host turns the counts into code lengths)
    against the reference's fixed code.  Every stream is inflated again by the engine (Adler-32 verified)
    and compared with the blocks before anything is timed."""
    import torch
    out = {}
    few = (d_plain & 15) + 97
    for name, d_src in (("config2_blocks", d_plain), ("sixteen_byte_values", few)):
        eng.set_tree()
        ostride = eng.bound(BLOCK)
        d_comp = torch.empty(n * ostride, dtype=torch.uint8, device=dev)
        d_clen = torch.zeros(n, dtype=torch.int32, device=dev)
        d_st = torch.zeros(n, dtype=torch.int32, device=dev)
        eng.compress_batch(d_src, BLOCK, None, BLOCK, d_comp, ostride, d_clen, d_st, n, stream=stream)
        torch.cuda.synchronize()
        fixed_bytes = int(d_clen.sum())
        t0 = time.time()
        eng.train_tree_device(d_src, BLOCK, None, BLOCK, min(n, 1 << 16), stream=stream)     # trained on the first 65536 blocks
        t_train = time.time() - t0
        tstride = eng.bound(BLOCK)
        d_tc = torch.empty(n * tstride, dtype=torch.uint8, device=dev)

        def run():
            eng.compress_batch(d_src, BLOCK, None, BLOCK, d_tc, tstride, d_clen, d_st, n, stream=stream)
        run()
        torch.cuda.synchronize()
        assert int(d_st.abs().sum()) == 0, "tree: status"
        tree_bytes = int(d_clen.sum())
        d_back = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
        d_blen = torch.zeros(n, dtype=torch.int32, device=dev)
        d_bst = torch.zeros(n, dtype=torch.int32, device=dev)
        eng.decompress_batch(d_tc, None, tstride, d_clen, d_back, BLOCK, BLOCK, d_blen, d_bst, n,
                             flags=hz.F_VERIFY_ADLER, stream=stream)
        torch.cuda.synchronize()
        assert int(d_bst.abs().sum()) == 0 and torch.equal(d_back, d_src.view(-1)), "tree: round trip"
        ms, best = _time_launches(run, steps)
        lit, dist = eng.tree
        out[name] = {"compress_gbps": n * BLOCK / (ms * 1e-3) / 1e9, "ms": ms, "ratio_tree": tree_bytes / (n * BLOCK),
                     "ratio_fixed": fixed_bytes / (n * BLOCK), "train_seconds": t_train,
                     "max_code_bits": [int(lit.max()), int(dist.max())], "round_trip_verified": True}
        del d_comp, d_tc, d_back
        eng.set_tree()
    out["what"] = ("hdlz_train_tree on the first 65536 blocks, then hdlz_compress_batch with the trained code (one "
                   "BTYPE = 10 block per stream, the reference's parse) against the fixed code; n = %d blocks" % n)
    return out


def run_long_stream(eng, hz, d_plain, dev, stream, steps, nbytes=4 << 20):
    """ONE compress stream (the reference's own use: one stream at a time): the first 4 MiB of the blocks as a
    single stream through hdlz_compress_batch with n = 1 — spread over the whole grid (k_compress<.., kLong>) and,
    for comparison, on one warp (HDLZ_NO_LONG=1).  The two outputs are compared, and the stream is inflated again
    by the engine (Adler-32 verified) and compared with the input."""
    import torch
    src = d_plain.view(-1)[:nbytes].contiguous()
    cap = hz.compress_bound(nbytes)
    outs, res = [], {}
    for mode in ("grid", "one_warp"):
        if mode == "one_warp":
            os.environ["HDLZ_NO_LONG"] = "1"
        d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        d_len = torch.zeros(1, dtype=torch.int32, device=dev)
        d_st = torch.zeros(1, dtype=torch.int32, device=dev)

        def run():
            eng.compress_batch(src, nbytes, None, nbytes, d_out, cap, d_len, d_st, 1, stream=stream)
        try:
            ms, best = _time_launches(run, steps if mode == "grid" else 2, warmup=1)
        finally:
            os.environ.pop("HDLZ_NO_LONG", None)
        assert int(d_st.item()) == 0, "long stream: status"
        outs.append(d_out[:int(d_len.item())].clone())
        res[mode + "_ms"] = ms
        res[mode + "_gbps"] = nbytes / (ms * 1e-3) / 1e9
    assert torch.equal(outs[0], outs[1]), "long stream: grid and one-warp outputs differ"
    n = outs[0].numel()
    d_in = torch.zeros((n + 31) & ~15, dtype=torch.uint8, device=dev)
    d_in[:n] = outs[0]
    d_back = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    d_blen = torch.zeros(1, dtype=torch.int32, device=dev)
    d_bst = torch.zeros(1, dtype=torch.int32, device=dev)
    d_clen = torch.tensor([n], dtype=torch.int32, device=dev)
    eng.decompress_batch(d_in, None, d_in.numel(), d_clen, d_back, nbytes, nbytes, d_blen, d_bst, 1,
                         flags=hz.F_VERIFY_ADLER, stream=stream)
    torch.cuda.synchronize()
    assert int(d_bst.item()) == 0 and torch.equal(d_back, src), "long stream: round trip"
    res.update({"bytes": nbytes, "compressed_bytes": n, "round_trip_verified": True,
                "what": "one %d-byte stream, hdlz_compress_batch with n = 1: tiles of the stream over the whole grid "
                        "(parse carry, bit cursor and Adler sums across tile borders by look-back) against one warp; "
                        "same bytes" % nbytes})
    return res


def config4_plain(nd, L=32768, seed=4):
    """Plain side of configs[3]: Zipf-like bytes over 64 symbols with repeats at distances up to 32 KiB
    (SURVEY 8(d)); compressible enough that zlib level 6 emits dynamic blocks."""
    import numpy as np
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, 65) ** 1.1
    p /= p.sum()
    plain = rng.choice(64, size=(nd, L), p=p).astype(np.uint8) + 32
    for _ in range(6):
        src = rng.integers(0, L - 4096, nd)
        dst = rng.integers(0, L - 4096, nd)
        ln = rng.integers(64, 4096, nd)
        for i in range(nd):
            plain[i, dst[i]:dst[i] + ln[i]] = plain[i, src[i]:src[i] + ln[i]].copy()
    return plain


def run_config4(eng, hz, n, nd, dev, stream, steps, threads, peak):
    """configs[3]: n x 32 KiB host-zlib level-6 (dynamic-tree) streams, drawn cyclically from nd distinct
    ones (~200 MB compressed, beyond the L2), fixed-stride slots, OBSIZE = 32768, decompress-only."""
    import numpy as np
    import torch
    L = 32768
    t0 = time.time()
    plain = config4_plain(nd)
    packed, off, lens = _zlib_many(plain, 6, 0, threads)
    t_fix = time.time() - t0
    assert (((packed[off + 2] >> 1) & 3) == 2).all(), "config4 streams must start with a dynamic block"
    stride = (int(lens.max()) + 15) & ~15
    idx = np.arange(stride)[None, :]
    comp = np.zeros((nd, stride), dtype=np.uint8)
    mask = idx < lens[:, None]
    comp[mask] = packed[(off[:, None] + idx)[mask]]
    sel = torch.arange(n, device=dev) % nd
    d_in = torch.from_numpy(comp).to(dev)[sel].contiguous()
    d_len = torch.from_numpy(lens.astype(np.int32)).to(dev)[sel].contiguous()
    d_plain = torch.from_numpy(plain).to(dev)
    d_out = torch.empty(n * L, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)

    def run(flags=0):
        eng.decompress_batch(d_in, None, stride, d_len, d_out, L, L, d_olen, d_st, n, flags=flags, stream=stream)
    run(hz.F_VERIFY_ADLER)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0 and bool((d_olen == L).all()), "config4: status / length"
    assert torch.equal(d_out.view(n, L), d_plain[sel]), "config4: output differs from the original"
    ms, best = _time_launches(run, steps)
    ms_v, _ = _time_launches(lambda: run(hz.F_VERIFY_ADLER), steps)
    cbytes = int(d_len.sum(dtype=torch.int64))
    alg = cbytes + n * L
    return {"workload": "BASELINE configs[3]: %d x 32 KiB host-zlib level-6 dynamic-tree streams (drawn cyclically from "
                        "%d distinct), OBSIZE = 32768, decompress-only" % (n, nd),
            "decompress_gbps": n * L / (ms * 1e-3) / 1e9, "decompress_gbps_verified": n * L / (ms_v * 1e-3) / 1e9,
            "ms": ms, "ms_best": best, "compressed_ratio": cbytes / (n * L), "byte_exact_vs_original": True,
            "roofline": {"bound": "hbm", "kernel": "dynamic-Huffman route", "achieved": alg / (ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg, "traffic": ncu_traffic("config4", n)},
            "fixture_seconds": t_fix}


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=1 << 20, help="2 KiB blocks per GPU")
    ap.add_argument("--e2e-blocks", type=int, default=1 << 20, help="blocks per step of the host-buffer (e2e) leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-blocks", type=int, default=1 << 16, help="blocks per step of the CPU arms")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[2] / [3] (decompress-only legs)")
    ap.add_argument("--c4-streams", type=int, default=100000)
    ap.add_argument("--c4-distinct", type=int, default=16384)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    import hdl_deflate_b200 as hz

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eng = hz.Engine(local)

    n = args.blocks
    ostride = hz.compress_bound(BLOCK)
    d_in = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    d_comp = torch.empty(n * ostride, dtype=torch.uint8, device=dev)
    d_clen = torch.zeros(n, dtype=torch.int32, device=dev)
    d_cst = torch.zeros(n, dtype=torch.int32, device=dev)
    d_back = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    d_blen = torch.zeros(n, dtype=torch.int32, device=dev)
    d_bst = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    eng.generate_blocks(d_in, BLOCK, BLOCK, n, first_block=rank * n, stream=stream)   # contiguous shard per rank
    torch.cuda.synchronize()

    def step(ev=None):
        if ev:
            ev[0].record()
        eng.compress_batch(d_in, BLOCK, None, BLOCK, d_comp, ostride, d_clen, d_cst, n, stream=stream)
        if ev:
            ev[1].record()
        eng.decompress_batch(d_comp, None, ostride, d_clen, d_back, BLOCK, BLOCK, d_blen, d_bst, n, flags=0,
                             stream=stream)
        if ev:
            ev[2].record()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # correctness of what is being timed: statuses clean, round trip identical, sample bit-exact
    assert int(d_cst.abs().sum()) == 0 and int(d_bst.abs().sum()) == 0
    assert torch.equal(d_back, d_in), "round trip differs"
    comp_bytes = int(d_clen.sum(dtype=torch.int64))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = eng.launch_count
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.steps):
        step(evs[k])
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches = eng.launch_count - launches0
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    t_total = evs[0][0].elapsed_time(evs[-1][2])                       # ms, first start .. last end
    t_c = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_d = sum(e[1].elapsed_time(e[2]) for e in evs)
    times = torch.tensor([t_total, t_c, t_d], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        tot = torch.tensor([comp_bytes], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        comp_bytes_all = int(tot.item())
    else:
        comp_bytes_all = comp_bytes
    t_total, t_c, t_d = [float(x) for x in times.tolist()]

    # the same decompress pass with the container checksum verified (zlib's contract; the reference itself
    # never checks it, deflate.py:1535), timed separately so both numbers stand side by side
    def dec_verified():
        eng.decompress_batch(d_comp, None, ostride, d_clen, d_back, BLOCK, BLOCK, d_blen, d_bst, n,
                             flags=hz.F_VERIFY_ADLER, stream=stream)
    t_dv, _ = _time_launches(dec_verified, max(3, args.steps // 2))
    assert int(d_bst.abs().sum()) == 0
    tdv = torch.tensor([t_dv], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tdv, op=dist.ReduceOp.MAX)
    t_dv = float(tdv.item())

    gather = None
    if dist:
        from hdl_deflate_b200 import sharding
        # configs[4] "NCCL used only to broadcast inputs and gather outputs" (none of it is in `value`):
        # (a) all_gather of the per-block stream lengths (every rank derives the packed offsets),
        # (b) the streams themselves, PACKED on each rank (hdlz_pack_batch) and gathered on rank 0 — only real
        #     stream bytes cross NVLink,
        # (c) the inputs once: rank 0 generates all N shards and sends every rank its own.
        def timed_ms(fn):
            fn()                                     # first call: NCCL's lazy channel set-up
            torch.cuda.synchronize()
            dist.barrier()
            g0 = torch.cuda.Event(enable_timing=True)
            g1 = torch.cuda.Event(enable_timing=True)
            g0.record()
            r = fn()
            g1.record()
            torch.cuda.synchronize()
            gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(gt, op=dist.ReduceOp.MAX)
            return float(gt.item()), r
        outs = [torch.empty_like(d_clen) for _ in range(world)]
        len_ms, _ = timed_ms(lambda: dist.all_gather(outs, d_clen))
        d_packed = torch.empty(n * ostride, dtype=torch.uint8, device=dev)
        d_poff = torch.zeros(n, dtype=torch.int64, device=dev)
        d_ptot = torch.zeros(1, dtype=torch.int64, device=dev)
        eng.pack_batch(d_comp, ostride, d_clen, d_packed, d_poff, d_ptot, n, stream=stream)
        torch.cuda.synchronize()
        my_total = int(d_ptot.item())
        out_ms, res = timed_ms(lambda: sharding.gather_streams(d_packed[:my_total], d_clen, n * world, dst=0))
        g_buf, g_off, g_len = res
        gathered_bytes = int(((g_len.to(torch.int64) + 3) & ~3).sum())
        ok = True
        if rank == 0:
            # rank 0's own shard must sit at the start of the gathered buffer, every stream where the offsets say
            ok = bool(torch.equal(g_buf[:my_total], d_packed[:my_total])) and int(g_off[n - 1]) == int(d_poff[n - 1])
            # and the gathered streams of ANOTHER rank must inflate to that rank's blocks (checked on a sample)
            m = min(4096, n)
            first = n * (world - 1)
            d_chk = torch.empty(m * BLOCK, dtype=torch.uint8, device=dev)
            d_cl = torch.zeros(m, dtype=torch.int32, device=dev)
            d_cs = torch.zeros(m, dtype=torch.int32, device=dev)
            eng.decompress_batch(g_buf, g_off[first:first + m].contiguous(), 0, g_len[first:first + m].contiguous(), d_chk,
                                 BLOCK, BLOCK, d_cl, d_cs, m, flags=hz.F_VERIFY_ADLER, stream=stream)
            d_ref = torch.empty(m * BLOCK, dtype=torch.uint8, device=dev)
            eng.generate_blocks(d_ref, BLOCK, BLOCK, m, first_block=first, stream=stream)
            torch.cuda.synchronize()
            ok = ok and int(d_cs.abs().sum()) == 0 and bool(torch.equal(d_chk, d_ref))
        del g_buf, d_packed
        d_all = None
        if rank == 0:
            d_all = torch.empty((n * world, BLOCK), dtype=torch.uint8, device=dev)
            eng.generate_blocks(d_all, BLOCK, BLOCK, n * world, first_block=0, stream=stream)
            torch.cuda.synchronize()
        d_recv = torch.empty((n, BLOCK), dtype=torch.uint8, device=dev)
        in_ms, _ = timed_ms(lambda: sharding.scatter_blocks(d_all, n * world, BLOCK, src=0, out=d_recv))
        same = torch.tensor([int(torch.equal(d_recv.view(-1), d_in))], dtype=torch.int32, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        del d_all, d_recv
        torch.cuda.empty_cache()
        gather = {"what": "NCCL: all_gather of out_len (4 B per block); packed streams of every rank gathered on rank 0 "
                          "(hdlz_pack_batch + point-to-point, only stream bytes cross NVLink); inputs scattered from rank 0 once",
                  "ms": len_ms, "outputs_ms": out_ms, "outputs_bytes": gathered_bytes,
                  "outputs_gbps": gathered_bytes / (out_ms * 1e-3) / 1e9,
                  "outputs_verified": bool(ok), "inputs_ms": in_ms, "inputs_bytes": n * world * BLOCK,
                  "inputs_gbps": n * (world - 1) * BLOCK / (in_ms * 1e-3) / 1e9,
                  "inputs_equal_local_generation": bool(int(same.item()))}

    # ---- e2e: the same step through the host-buffer C ABI (pinned host memory, copies timed), on
    # every rank at the same time (the ranks share the host's PCIe / memory system), max over ranks.
    # Two legs: (i) the two calls back to back on one context — each is bound by its larger PCIe direction
    # while the other direction idles; (ii) the calls on two contexts from two host threads, a step's
    # decompress working on the previous step's streams while the next batch is compressed: both directions
    # of the link stay busy.  (ii) is what a host that pipelines its batches gets and is the reported value.
    e2e = None
    if not args.no_e2e:
        from concurrent.futures import ThreadPoolExecutor
        import ctypes
        ne = min(args.e2e_blocks, n)
        h_in = torch.empty((ne, BLOCK), dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in.view(n, BLOCK)[:ne])
        h_comp = [torch.empty(ne * ostride, dtype=torch.uint8, pin_memory=True) for _ in range(2)]   # packed streams
        h_off = [torch.zeros(ne, dtype=torch.int64, pin_memory=True) for _ in range(2)]
        h_clen = [torch.zeros(ne, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        h_back = torch.empty((ne, BLOCK), dtype=torch.uint8, pin_memory=True)
        h_blen = torch.zeros(ne, dtype=torch.int32, pin_memory=True)
        h_st = [torch.zeros(ne, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        eng2 = hz.Engine(local)                                  # the decompressing context
        lib = eng._lib
        total = [ctypes.c_uint64(0), ctypes.c_uint64(0)]

        def comp(k):
            rc = lib.hdlz_compress_host_packed(eng._ctx, h_in.data_ptr(), BLOCK, None, BLOCK, h_comp[k].data_ptr(),
                                               ne * ostride, h_off[k].data_ptr(), h_clen[k].data_ptr(),
                                               h_st[0].data_ptr(), ne, ctypes.byref(total[k]))
            assert rc == 0, lib.hdlz_last_error()

        def decomp(k, ctx):
            rc = lib.hdlz_decompress_host(ctx, h_comp[k].data_ptr(), h_off[k].data_ptr(), 0, h_clen[k].data_ptr(),
                                          h_back.data_ptr(), BLOCK, BLOCK, h_blen.data_ptr(), h_st[1].data_ptr(), ne, 0)
            assert rc == 0, lib.hdlz_last_error()

        def sync_ranks():
            if dist:
                dist.barrier()

        def max_ranks(t):
            if not dist:
                return t
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        # (i) back to back
        comp(0)
        decomp(0, eng._ctx)
        assert torch.equal(h_back, h_in)
        sync_ranks()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            comp(0)
            decomp(0, eng._ctx)
        te_seq = max_ranks((time.perf_counter() - t0) / args.e2e_steps)
        # (ii) overlapped: step k compresses into buffer k & 1 while buffer (k - 1) & 1 is decompressed
        comp(1)
        h_back.zero_()
        pool = ThreadPoolExecutor(2)

        def ostep(k):
            fa = pool.submit(comp, k & 1)
            fb = pool.submit(decomp, (k - 1) & 1, eng2._ctx)
            fa.result()
            fb.result()
        ostep(0)
        ostep(1)
        assert torch.equal(h_back, h_in)
        sync_ranks()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            ostep(k)
        te = max_ranks((time.perf_counter() - t0) / args.e2e_steps)
        pool.shutdown()
        assert torch.equal(h_back, h_in)
        pk = int(total[0].value)
        # the ceiling of this leg: raw pinned copies of the same payload bytes in both directions at once, all
        # ranks at the same time (what the host's PCIe / memory system gives with no kernel and no pipeline)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        d_a = torch.empty(ne * BLOCK, dtype=torch.uint8, device=dev)
        d_b = torch.empty(max(pk, 16), dtype=torch.uint8, device=dev)
        hin, hback = h_in.view(-1), h_back.view(-1)
        cb = 48 << 20

        def raw_step():
            with torch.cuda.stream(s_up):
                for o in range(0, ne * BLOCK, cb):
                    d_a[o:o + cb].copy_(hin[o:o + cb], non_blocking=True)
                for o in range(0, pk, cb):
                    d_b[o:min(o + cb, pk)].copy_(h_comp[0][o:min(o + cb, pk)], non_blocking=True)
            with torch.cuda.stream(s_dn):
                for o in range(0, ne * BLOCK, cb):
                    hback[o:o + cb].copy_(d_a[o:o + cb], non_blocking=True)
                for o in range(0, pk, cb):
                    h_comp[1][o:min(o + cb, pk)].copy_(d_b[o:min(o + cb, pk)], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
        raw_step()
        sync_ranks()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            raw_step()
        t_raw = max_ranks((time.perf_counter() - t0) / args.e2e_steps)
        del d_a, d_b
        h2d = ne * BLOCK + pk + (8 + 4) * ne                 # blocks; packed streams + offsets + lengths
        d2h = pk + (8 + 4 + 4) * ne + ne * BLOCK + (4 + 4) * ne
        e2e = {"value": 2 * ne * BLOCK * world / te / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "blocks_per_step_per_gpu": ne, "ms_per_step": te * 1e3,
               "api": "hdlz_compress_host_packed on one context and hdlz_decompress_host on a second one, from two host "
                      "threads: a step compresses its batch while the previous step's streams are decompressed (pinned "
                      "host buffers, packed streams, chunked 3-stream pipelines); every step moves one full round trip",
               "back_to_back": {"value": 2 * ne * BLOCK * world / te_seq / 1e9, "ms_per_step": te_seq * 1e3,
                                "api": "the same two calls one after the other on one context"},
               "raw_copy_ceiling": {"value": 2 * ne * BLOCK * world / t_raw / 1e9, "ms_per_step": t_raw * 1e3,
                                    "what": "cudaMemcpyAsync of the step's payload (blocks + packed streams up, packed "
                                            "streams + blocks down) on two streams, 48 MiB pieces, no kernels: the same "
                                            "metric if the host link were the only cost"},
               "note": "all %d ranks concurrently, max over ranks" % world}
        eng2.close()
        del h_in, h_comp, h_back

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    K = args.steps
    unc = n * BLOCK * world                       # uncompressed bytes per pass, all ranks
    value = 2 * unc * K / (t_total * 1e-3) / 1e9
    c_gbps = unc * K / (t_c * 1e-3) / 1e9
    d_gbps = unc * K / (t_d * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel (per launch = per rank)
    alg = n * BLOCK + comp_bytes                  # compress: L read + C written; decompress: C read + L written
    if t_c >= t_d:
        kname, kt = "k_compress", t_c / K
    else:
        kname, kt = "k_inflate_lanes", t_d / K
    achieved = alg / (kt * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kname, n), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "launch_ms": kt,
                "other_kernel": {"kernel": "k_inflate_lanes" if kname == "k_compress" else "k_compress",
                                 "achieved": alg / ((t_d if kname == "k_compress" else t_c) / K * 1e-3) / 1e9}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": t_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(n), "blocks_per_gpu": n, "block_bytes": BLOCK,
                   "out_stride": ostride, "compressed_ratio": comp_bytes_all / unc,
                   "l2": "each pass streams > 4 GB per GPU, far beyond the 126 MB L2: no flush needed",
                   "parallelism": "independent blocks, contiguous shard per GPU, no data-path collective"},
        "compress_gbps": c_gbps, "decompress_gbps": d_gbps,
        "roofline": roofline, "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
    if gather:
        # kernel-only (`value`) and including the gather of the lengths
        gather["value_incl_gather"] = 2 * unc * K / ((t_total + K * gather["ms"]) * 1e-3) / 1e9
        gather["value_incl_output_gather"] = 2 * unc * K / ((t_total + K * (gather["ms"] + gather["outputs_ms"])) * 1e-3) / 1e9
        line["gather"] = gather
    # the north star's read-only variant of the compress roofline: input bytes / time against the HBM rate
    line["roofline"]["compress_input_read_frac"] = (n * BLOCK / (t_c / K * 1e-3) / 1e9) / peak
    sim = reference_sim_rate()
    if sim:
        line["reference_sim"] = sim

    if e2e:
        line["e2e"] = e2e
    line["decompress_gbps_verified"] = unc / (t_dv * 1e-3) / 1e9

    # ---- cpu_baseline: the oracle port on the host cores, bounded sample, N = 1 only ----
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ns = min(args.ref_blocks, n)
        sample = d_in.view(n, BLOCK)[:ns].cpu().numpy()
        r = cpu_roundtrip(sample, cores)
        rep = max(1, int(20.0 / max(r["seconds"], 1e-3)))
        if rep > 1:
            r = cpu_roundtrip(sample, cores, repeat=min(rep, 128))
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                                "sample": "first %d blocks of the workload, %.1f s of CPU work, %d threads"
                                          % (ns, r["seconds"], cores),
                                "compress_gbps": r["compress_gbps"], "decompress_gbps": r["decompress_gbps"],
                                "port_inflate_gbps": r["port_inflate_gbps"], "zlib_inflate_gbps": r["zlib_inflate_gbps"],
                                "what": "tuned C port of deflate.py's compressor (bit-identical output) + the faster of "
                                        "port / host-zlib inflate"}
        try:
            line["cpu_zlib"] = dict(cpu_zlib(sample, cores), cores=cores)
        except Exception as e:      # informational only
            line["cpu_zlib"] = {"error": repr(e)}

    if world == 1 and not args.no_configs:
        del d_comp, d_back
        torch.cuda.empty_cache()
        threads = os.cpu_count() or 1
        cfgs = {}
        cfgs["config3"] = run_config3(eng, hz, d_in, n, dev, stream, max(3, args.steps // 2), threads, peak)
        cfgs["config4"] = run_config4(eng, hz, args.c4_streams, min(args.c4_distinct, args.c4_streams), dev, stream,
                                      max(3, args.steps // 2), threads, peak)
        cfgs["tree"] = run_tree(eng, hz, d_in, n, dev, stream, max(3, args.steps // 2))
        cfgs["long_stream"] = run_long_stream(eng, hz, d_in, dev, stream, max(3, args.steps // 2))
        line["configs"] = cfgs

    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
