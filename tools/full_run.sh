# Round-end measurement set on one B200 (run through gpurun): GPU tests, the bench line (with BASELINE configs 3 / 4
# and the tree leg folded in), the reference arm, the ncu launch list of the bench command, the DRAM traffic of the
# kernels and one full ncu capture of the kernels of the config-4 route.  R names the output files.
set -x
R=${R:-r2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>> gpurun_out/${R}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/${R}_launch_bench.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_compress|k_inflate_lanes|k_decode_tokens|k_resolve_tokens" -c 80 --csv --log-file gpurun_out/${R}_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_tokens|k_resolve_tokens" --launch-skip 4 -c 2 -o gpurun_out/${R}_split python tools/bench_configs.py --only config4 --streams 60000 --distinct 8192 --steps 1 > /dev/null 2>&1
cat gpurun_out/${R}_pytest.log; head -c 1500 gpurun_out/${R}_bench.json; echo; cat gpurun_out/${R}_bench_ref.json | head -c 600; tail -3 gpurun_out/${R}_bench.err
