# compute-sanitizer passes over the GPU parity tests (run through gpurun; ~2 minutes of box time):
#   memcheck  - out-of-bounds / misaligned accesses of every kernel
#   racecheck - shared-memory hazards of the warp-synchronous code (the kernels order their shared-memory
#               phases with __syncwarp only)
set -x
SEL='not full_size and not config3 and not config4 and not batch_matches_oracle and not randomized'
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_containers.py tests/test_gpu_inflate.py tests/test_gpu_compress.py -m gpu -x -q -k "$SEL" 2>&1 | tail -5
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_containers.py tests/test_gpu_compress.py tests/test_gpu_inflate.py -m gpu -x -q -k "golden or ragged or same_body or raw_and_gzip or mixed_batch or multi_block" 2>&1 | tail -5
