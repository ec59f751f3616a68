"""Run the UNCHANGED reference test bench (/root/reference/test_deflate.py) against the
drop-in `deflate` module of this repo.

    python tests/run_reference_unittest.py            # real engine (needs a B200)
    python tests/run_reference_unittest.py --oracle   # host-logic check: START jobs are answered
                                                      # by the CPU oracle (test double, tests only)
The reference file is executed from where it lies; nothing is copied.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DROPIN = os.path.join(ROOT, "hdl-deflate_b200", "dropin")
REF = os.environ.get("HDLZ_REFERENCE_DIR", "/root/reference")


def main():
    use_oracle = "--oracle" in sys.argv
    sys.argv = [os.path.join(REF, "test_deflate.py")]
    sys.path.insert(0, ROOT)
    sys.path.insert(0, DROPIN)
    import deflate
    assert deflate.__file__.startswith(DROPIN), deflate.__file__
    if use_oracle:
        sys.path.insert(0, HERE)
        from oracle_engine import OracleEngine
        deflate.set_backend(OracleEngine())
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
