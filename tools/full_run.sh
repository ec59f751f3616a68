# Round-end measurement set on one B200 (run through gpurun): GPU tests, the bench line (with BASELINE configs 3 / 4
# folded in), the reference arm, the ncu launch list of the bench command and the DRAM traffic of the kernels.
set -x
R=${R:-r2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>> gpurun_out/${R}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/${R}_launch_bench.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_compress|k_inflate_lanes|k_decode_tokens|k_resolve_tokens" -c 60 --csv --log-file gpurun_out/${R}_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
cat gpurun_out/${R}_pytest.log; head -c 1500 gpurun_out/${R}_bench.json; echo; cat gpurun_out/${R}_bench_ref.json | head -c 600; tail -3 gpurun_out/${R}_bench.err
