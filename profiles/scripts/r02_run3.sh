#!/bin/bash
# round 2, third GPU call: tests touched by the prefill work, long_b1 bench, launch list of one prefill window
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_driver.py tests/test_gpu_checkpoint.py -m gpu -q -s > gpurun_out/r02_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run3_pytest.log
tail -6 gpurun_out/r02_run3_pytest.log
python bench.py --workload long_b1 --steps 3 --warmup 1 > gpurun_out/r02_run3_long_b1.json 2> gpurun_out/r02_run3_long_b1.err; echo "long rc=$?"; cat gpurun_out/r02_run3_long_b1.json | cut -c1-1100
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|attn|rmsnorm|embed|gemv|sample' -c 400 --csv --log-file gpurun_out/r02_prefill_launches_v2.csv python profiles/run_prefill.py > gpurun_out/r02_run3_ncu.log 2>&1; echo "ncu rc=$?"
