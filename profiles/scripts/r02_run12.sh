#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_driver.py tests/test_gpu_checkpoint.py -m gpu -q -x -s 2>&1 | tail -8
python bench.py --workload long_b1 --steps 3 --warmup 1 > gpurun_out/r02_run12_long_b1.json 2> gpurun_out/r02_run12.err; echo "long rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_run12_long_b1.json')); print(d['long_b1'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|attn|rmsnorm|embed|gemv|sample' -c 400 --csv --log-file gpurun_out/r02_prefill_launches_v4.csv python profiles/run_prefill.py > gpurun_out/r02_run12_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r02_prefill_launches_v4.csv | head -12
