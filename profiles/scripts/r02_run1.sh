#!/bin/bash
# round 2, first GPU call: whole -m gpu suite (new full-clip parity tests included), smoke(), default bench + reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r02_run1_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_run1_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run1_bench.json 2> gpurun_out/r02_run1_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_run1_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_run1_bench_ref.json 2>&1; echo "ref rc=$?"
