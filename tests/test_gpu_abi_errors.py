"""Error behaviour of the C ABI on a live context: bad arguments come back as HDLZ_ERR_INVALID with a
message (never a crash or a silent success), per-stream problems as status words, and options round-trip."""
import ctypes

import numpy as np
import pytest

import hdl_deflate_b200 as hz
from hdl_deflate_b200 import workload

pytestmark = pytest.mark.gpu


def test_invalid_arguments(engine):
    lib, ctx = engine._lib, engine._ctx
    a = np.zeros(4096, dtype=np.uint8)
    o = np.zeros(8192, dtype=np.uint8)
    ln = np.zeros(4, dtype=np.uint32)
    st = np.zeros(4, dtype=np.uint32)
    # null buffers
    assert lib.hdlz_compress_host(ctx, None, 2048, None, 2048, o.ctypes.data, 2320, ln.ctypes.data, st.ctypes.data, 1) == -1
    assert b"null" in lib.hdlz_last_error()
    # strides that are not multiples of 16
    assert lib.hdlz_compress_host(ctx, a.ctypes.data, 2050, None, 2048, o.ctypes.data, 2320, ln.ctypes.data,
                                  st.ctypes.data, 1) == -1
    # out_cap larger than out_stride
    assert lib.hdlz_decompress_host(ctx, a.ctypes.data, None, 64, ln.ctypes.data, o.ctypes.data, 64, 128,
                                    ln.ctypes.data, st.ctypes.data, 1, 0) == -1
    # unknown container
    assert lib.hdlz_set_container(ctx, 7) == -1 and engine.container == hz.CONTAINER_ZLIB
    # zero streams is a no-op, not an error
    assert lib.hdlz_compress_host(ctx, a.ctypes.data, 2048, None, 2048, o.ctypes.data, 2320, ln.ctypes.data,
                                  st.ctypes.data, 0) == 0
    assert lib.hdlz_create(99, ctypes.byref(ctypes.c_void_p())) == -1


def test_slot_too_small_is_a_status_not_a_crash(engine):
    blocks = np.frombuffer(b"".join(workload.blocks(0, 4, 2048)), dtype=np.uint8).reshape(4, 2048)
    out, out_len, status = engine.compress_host(blocks, out_stride=1024)      # needs 2320
    assert status.tolist() == [6, 6, 6, 6] and not out_len.any()
    out, out_len, status = engine.compress_host(blocks)
    assert not status.any()
    back, back_len, bst = engine.decompress_host(out, out_len, 1000)          # output does not fit
    assert bst.tolist() == [6, 6, 6, 6]


def test_options_round_trip(engine):
    assert engine.match10 and engine.container == hz.CONTAINER_ZLIB
    engine.match10 = False
    engine.container = hz.CONTAINER_RAW
    assert not engine.match10 and engine.container == hz.CONTAINER_RAW
    engine.match10 = True
    engine.container = hz.CONTAINER_ZLIB
    assert hz.compress_bound(2048) == 2320 and hz.compress_bound(2048, hz.CONTAINER_RAW) == 2320 - 0
    assert hz.compress_bound(2048, hz.CONTAINER_GZIP) == 2336


def test_c_client_runs_a_job(tmp_path):
    """tests/c/abi_example.c (plain C against include/hdlz.h) compresses and inflates one stream, in one call and
    through hdlz_dstream_* in pieces."""
    import os
    import shutil
    import subprocess
    from hdl_deflate_b200 import _native
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_native.LIB_PATH)
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c", "abi_example.c"), "-o", exe, "-L", libdir, "-lhdlz",
                           "-Wl,-rpath," + libdir])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "compress 0 status 0" in p.stdout and "head 789c" in p.stdout and "same 1" in p.stdout
    assert "dstream 0 status 0" in p.stdout and "remaining 0 same 1" in p.stdout, p.stdout      # inflated in 7-byte pieces
