#!/bin/bash
# round 2: first run of decode_step_fused2 (tiny + full-size parity, then the b64 bench line)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "bf16" > gpurun_out/r02_run4_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -5 gpurun_out/r02_run4_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_fullclip.py -m gpu -q -x -s -k "64_rows" > gpurun_out/r02_run4_pytest_b.log 2>&1; echo "pytest b rc=$?"; tail -5 gpurun_out/r02_run4_pytest_b.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r02_run4_bench.json 2> gpurun_out/r02_run4_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_run4_bench.json; tail -3 gpurun_out/r02_run4_bench.err
VAURA_FUSED2=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r02_run4_bench_old.json 2> /dev/null; echo "bench old rc=$?"; cut -c1-300 gpurun_out/r02_run4_bench_old.json
