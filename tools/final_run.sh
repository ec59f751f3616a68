set -x
R=r02i
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>> gpurun_out/${R}_bench.err
cat gpurun_out/${R}_pytest.log; head -c 600 gpurun_out/${R}_bench.json; echo; head -c 300 gpurun_out/${R}_bench_ref.json; tail -3 gpurun_out/${R}_bench.err
