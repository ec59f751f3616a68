# Round-end measurement set on one B200 (run through gpurun): GPU tests, the bench line, the reference arm,
# the other BASELINE configs, the ncu launch list of the bench command and the DRAM traffic of the two kernels.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1_pytest.log
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2>> gpurun_out/r1_bench.err
python tools/bench_configs.py > gpurun_out/r1_configs.json 2> gpurun_out/r1_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --no-cpu --no-e2e > gpurun_out/r1_launch_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_compress|k_inflate_lanes" -c 4 --csv --log-file gpurun_out/r1_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
cat gpurun_out/r1_pytest.log gpurun_out/r1_bench.json gpurun_out/r1_bench_ref.json gpurun_out/r1_configs.json
