"""The C-ABI library: loads, exports every symbol include/hdlz.h declares, and fails loudly
without a GPU (no compute here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    import __graft_entry__
    __graft_entry__.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hdlz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hdlz_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    _built()
    from hdl_deflate_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libhdlz.so does not export %s" % s
        assert s in _native.SIGNATURES, "ctypes binding lacks %s" % s
    assert sorted(_native.SIGNATURES) == syms


def test_version_bound_and_names():
    _built()
    from hdl_deflate_b200 import _native, compress_bound
    lib = _native.load()
    assert lib.hdlz_version() == 0x000100
    assert compress_bound(2048) == 2320           # 2 + ceil((3+9*2048+7)/8) + 4 = 2312 -> 16-aligned
    assert compress_bound(5) % 16 == 0
    assert lib.hdlz_status_name(0) == b"OK" and lib.hdlz_status_name(4) == b"DIST_TOO_FAR"


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _built()
    import hdl_deflate_b200
    with pytest.raises(hdl_deflate_b200.HdlzError):
        hdl_deflate_b200.Engine(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or link it."""
    pkg = os.path.join(ROOT, "hdl-deflate_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|hdlz_oracle|oracle_engine", re.M)
    checked = 0
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                checked += 1
                assert not pat.search(open(os.path.join(dp, f)).read()), (dp, f)
    assert checked >= 8


def test_c_client_builds_and_runs(tmp_path):
    """include/hdlz.h is plain C: a C program compiles against it, links libhdlz.so and runs.  Without a
    GPU the library answers the device-independent calls and refuses to create a context."""
    import shutil
    import subprocess
    import torch
    _built()
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from hdl_deflate_b200 import _native
    libdir = os.path.dirname(_native.LIB_PATH)
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_example.c"), "-o", exe, "-L", libdir, "-lhdlz",
                           "-Wl,-rpath," + libdir])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    out = p.stdout
    assert p.returncode == 0, out + p.stderr
    assert "version 000100" in out and "bound2048 2320 2320 2336" in out and "status4 DIST_TOO_FAR" in out
    assert "tree 0 lenA" in out and "head 789c" in out            # device-independent helpers of the tree-coded mode
    if torch.cuda.is_available():
        assert "compress 0 status 0" in out and "head 789c" in out and "same 1" in out
        assert "dstream 0 status 0" in out and "remaining 0 same 1" in out
    else:
        assert "devices 0" in out and "create -" in out
