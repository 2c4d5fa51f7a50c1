#!/bin/bash
# state of the tree after parking decode_step_fused2: all GPU tests, the default bench line, smoke, prefill launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_run7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run7_pytest.log
tail -5 gpurun_out/r02_run7_pytest.log
python bench.py > gpurun_out/r02_run7_bench.json 2> gpurun_out/r02_run7_bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02_run7_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|attn|rmsnorm|embed|gemv|sample' -c 400 --csv --log-file gpurun_out/r02_prefill_launches_v3.csv python profiles/run_prefill.py > gpurun_out/r02_run7_ncu.log 2>&1; echo "ncu rc=$?"
