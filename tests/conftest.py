import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)["cases"]


def golden_input(case):
    """Input bytes of a compress golden case (inline hex or workload recipe)."""
    import hashlib
    from hdl_deflate_b200 import workload
    if "in_hex" in case:
        data = bytes.fromhex(case["in_hex"])
    else:
        r = case["recipe"]
        data = workload.block(r["index"], r["length"], r["seed"])
    assert hashlib.sha256(data).hexdigest() == case["in_sha256"], case["name"]
    return data


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine on cuda:0.  GPU tests fail (not skip) if it cannot be created."""
    import hdl_deflate_b200
    return hdl_deflate_b200.Engine(0)
