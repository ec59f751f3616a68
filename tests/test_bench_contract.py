"""bench.py contract pieces that can be checked without a GPU: the reference arm (the oracle port on
the host cores) prints one JSON line with the agreed keys, and the GPU arm refuses to run without a
device instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True,
                          text=True, timeout=600, cwd=ROOT)


def test_reference_arm_line():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-blocks", "1024")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["metric"] == "deflate GB/s (compress+decompress)" and d["dtype"] == "u8" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "port+zlib") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_never_loads_the_product_library():
    """The CPU arm must be clean of libhdlz.so (round-1 review: the record listed it as loaded)."""
    code = ("import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', "
            "'--ref-blocks', '256']; import runpy\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit: pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libhdlz_oracle.so' in maps and 'libhdlz.so' not in maps, 'product library mapped'\n"
            "print('CLEAN')" % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0 and "CLEAN" in p.stdout, (p.stdout + p.stderr)[-2000:]


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run_bench("--steps", "1", "--no-cpu", "--no-e2e")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)


def test_reference_sim_rate_from_fixtures():
    sys.path.insert(0, ROOT)
    import bench
    r = bench.reference_sim_rate()
    assert r and 2.5 < r["cycles_per_byte"] < 4.5          # 3.0-3.9 cycles per input byte
