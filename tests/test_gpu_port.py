"""BASELINE config 1: the reference's port protocol on the CUDA engine."""
import os
import random
import subprocess
import sys
import zlib

import pytest

from port_driver import Port
from oracle import hdlz_oracle
from test_host_model import reference_style_data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_dir():
    """Where the unchanged test_deflate.py lies: $HDLZ_REFERENCE_DIR, the reference checkout (build container),
    or baseline/_ref/ (the copy __graft_entry__.build() places for the GPU box; git-ignored)."""
    for d in (os.environ.get("HDLZ_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.exists(os.path.join(d, "test_deflate.py")):
            return d
    return None



@pytest.mark.parametrize("mode", range(6))
def test_unit_test_flow_on_gpu(engine, mode):
    """Own transcription of TestDeflate.testMain (test_deflate.py:90-296) at the reference's sizes
    (tlen = 2500, slen = 10000), START jobs executed by the CUDA kernels."""
    rnd = random.Random(100 + mode)
    p = Port(engine)
    if mode == 0:
        p.pulse_reset()
    b_data = reference_style_data(mode, 2500, rnd)
    co = zlib.compressobj(wbits=p.m.LOBSIZE)
    zl_data = co.compress(b_data) + co.flush()
    out, _ = p.stream(p.m.STARTD, zl_data)
    assert out == b_data                                          # test_deflate.py:194
    slen = 10000
    src = bytes(b_data[i % len(b_data)] for i in range(slen)) if b_data else b""
    res, _ = p.stream(p.m.STARTC, src)
    rlen = min(len(b_data), slen)
    assert zlib.decompress(res)[:rlen] == b_data[:rlen]           # test_deflate.py:285
    if b_data:
        assert res == hdlz_oracle.compress(src)[1]                # and bit-exact with the reference format


def test_hw_bench_flow_on_gpu(engine):
    p = Port(engine)
    p.pulse_reset()
    data = " ".join("   Hello World! %d     " % i for i in range(100)).encode()[:2034]   # test_data(1, 100, IBSIZE)
    comp = p.preload(p.m.STARTC, data)
    assert comp == hdlz_oracle.compress(data)[1]
    assert p.preload(p.m.STARTD, comp) == data


@pytest.mark.skipif(_reference_dir() is None, reason="reference test bench not present (no baseline/_ref)")
def test_unchanged_reference_unittest_on_gpu():
    """BASELINE configs[0]: TestDeflate.testMain (test_deflate.py:90-321), unmodified, drives the CUDA engine."""
    env = dict(os.environ, HDLZ_REFERENCE_DIR=_reference_dir())
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_unittest.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    tail = (r.stdout + r.stderr)[-2000:]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "unchanged_unittest_gpu.log"), "w") as f:
        f.write(r.stdout[-6000:] + r.stderr[-6000:])
    assert r.returncode == 0 and "OK" in tail, tail
