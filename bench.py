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
def cpu_roundtrip(sample, nthreads, repeat=1):
    """Time the port on `sample` (uint8 [n, 2048]) with `nthreads` threads.
    -> (roundtrip GB/s as defined for `value`, compress GB/s, decompress GB/s, seconds)."""
    import numpy as np
    from oracle import hdlz_oracle as O
    from hdl_deflate_b200 import compress_bound
    n = sample.shape[0]
    ostride = compress_bound(BLOCK)
    comp = np.empty((n, ostride), dtype=np.uint8)
    back = np.empty((n, BLOCK), dtype=np.uint8)
    ioff = np.arange(n, dtype=np.uint64) * BLOCK
    ooff = np.arange(n, dtype=np.uint64) * ostride
    lens = np.full(n, BLOCK, dtype=np.uint32)
    tc = td = 0.0
    for _ in range(repeat):
        t0 = time.perf_counter()
        clen, st = O.batch(O.KIND_PORT_COMPRESS, sample, ioff, lens, comp, ooff, ostride, nthreads)
        t1 = time.perf_counter()
        blen, st2 = O.batch(O.KIND_PORT_INFLATE, comp, ooff, clen, back, ioff, BLOCK, nthreads)
        t2 = time.perf_counter()
        tc += t1 - t0
        td += t2 - t1
    assert not st.any() and not st2.any() and np.array_equal(back, sample)
    byt = n * BLOCK * repeat
    return 2 * byt / (tc + td) / 1e9, byt / tc / 1e9, byt / td / 1e9, tc + td


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
    import numpy as np
    from hdl_deflate_b200 import workload
    return np.frombuffer(b"".join(workload.blocks(0, n, BLOCK)), dtype=np.uint8).reshape(n, BLOCK)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path.  deflate.py is a Python/MyHDL simulation that
    cannot travel to the GPU box (and runs at ~1.5 kB/s), so the arm times the oracle port of it
    (oracle/hdlz_oracle.c, checked byte-for-byte against the executing reference) on all host threads."""
    if rank != 0:
        return
    import __graft_entry__
    __graft_entry__.build()
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
    for _ in range(args.steps):
        v, c, d, secs = cpu_roundtrip(sample, cores)
        tot_c += n * BLOCK / c / 1e9
        tot_d += n * BLOCK / d / 1e9
    t = time.perf_counter() - t0
    byt = n * BLOCK * args.steps
    value = 2 * byt / (tot_c + tot_d) / 1e9
    sample_desc = "%d of the config's 2 KiB blocks per step (bounded sample), %d steps" % (n, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.blocks), "sample": sample_desc},
        "compress_gbps": byt / tot_c / 1e9, "decompress_gbps": byt / tot_d / 1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    gather = None
    if dist:
        # configs[4]: NCCL gather of the per-block stream lengths (the only exchange the path has); not in `value`
        outs = [torch.empty_like(d_clen) for _ in range(world)]
        torch.cuda.synchronize()
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather(outs, d_clen)
        g1.record()
        torch.cuda.synchronize()
        # second warm call timed (the first one includes NCCL's lazy channel set-up)
        g0.record()
        dist.all_gather(outs, d_clen)
        g1.record()
        torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather = {"what": "NCCL all_gather of out_len (4 B per block, every rank gets all lengths -> packed offsets)",
                  "ms": float(gt.item())}

    # ---- e2e: the same step through the host-buffer C ABI (pinned host memory, copies timed), on
    # every rank at the same time (the ranks share the host's PCIe / memory system), max over ranks
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_blocks, n)
        h_in = torch.empty((ne, BLOCK), dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in.view(n, BLOCK)[:ne])
        h_comp = torch.empty(ne * ostride, dtype=torch.uint8, pin_memory=True)       # packed streams land here
        h_off = torch.zeros(ne, dtype=torch.int64, pin_memory=True)
        h_back = torch.empty((ne, BLOCK), dtype=torch.uint8, pin_memory=True)
        h_clen = torch.zeros(ne, dtype=torch.int32, pin_memory=True)
        h_blen = torch.zeros(ne, dtype=torch.int32, pin_memory=True)
        h_st = torch.zeros(ne, dtype=torch.int32, pin_memory=True)
        lib, ctx = eng._lib, eng._ctx
        import ctypes
        total = ctypes.c_uint64(0)

        def e2e_step():
            rc = lib.hdlz_compress_host_packed(ctx, h_in.data_ptr(), BLOCK, None, BLOCK, h_comp.data_ptr(),
                                               ne * ostride, h_off.data_ptr(), h_clen.data_ptr(), h_st.data_ptr(), ne,
                                               ctypes.byref(total))
            assert rc == 0, lib.hdlz_last_error()
            rc = lib.hdlz_decompress_host(ctx, h_comp.data_ptr(), h_off.data_ptr(), 0, h_clen.data_ptr(),
                                          h_back.data_ptr(), BLOCK, BLOCK, h_blen.data_ptr(), h_st.data_ptr(), ne, 0)
            assert rc == 0, lib.hdlz_last_error()
        e2e_step()
        assert torch.equal(h_back, h_in)
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        te = (time.perf_counter() - t0) / args.e2e_steps
        if dist:
            tt = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
        pk = int(total.value)
        h2d = ne * BLOCK + pk + (8 + 4) * ne                 # blocks; packed streams + offsets + lengths
        d2h = pk + (8 + 4 + 4) * ne + ne * BLOCK + (4 + 4) * ne
        e2e = {"value": 2 * ne * BLOCK * world / te / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "blocks_per_step_per_gpu": ne, "ms_per_step": te * 1e3,
               "api": "hdlz_compress_host_packed + hdlz_decompress_host (pinned host buffers, packed streams, "
                      "chunked 3-stream pipeline)",
               "note": "all %d ranks concurrently, max over ranks" % world}
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
        line["gather"] = gather
    # the north star's read-only variant of the compress roofline: input bytes / time against the HBM rate
    line["roofline"]["compress_input_read_frac"] = (n * BLOCK / (t_c / K * 1e-3) / 1e9) / peak
    sim = reference_sim_rate()
    if sim:
        line["reference_sim"] = sim

    if e2e:
        line["e2e"] = e2e

    # ---- cpu_baseline: the oracle port on the host cores, bounded sample, N = 1 only ----
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ns = min(args.ref_blocks, n)
        sample = d_in.view(n, BLOCK)[:ns].cpu().numpy()
        v, c, d, secs = cpu_roundtrip(sample, cores)
        rep = max(1, int(20.0 / max(secs, 1e-3)))
        if rep > 1:
            v, c, d, secs = cpu_roundtrip(sample, cores, repeat=min(rep, 128))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first %d blocks of the workload, %.1f s of CPU work, %d threads"
                                          % (ns, secs, cores),
                                "compress_gbps": c, "decompress_gbps": d}
        try:
            line["cpu_zlib"] = dict(cpu_zlib(sample, cores), cores=cores)
        except Exception as e:      # informational only
            line["cpu_zlib"] = {"error": repr(e)}

    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
