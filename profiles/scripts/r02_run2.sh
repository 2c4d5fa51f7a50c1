#!/bin/bash
# round 2, second GPU call: whole -m gpu suite (tensor-core prefill included), long_b1 bench, launch list of one prefill window
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_run2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run2_pytest.log
tail -8 gpurun_out/r02_run2_pytest.log
python bench.py --workload long_b1 --steps 3 --warmup 1 > gpurun_out/r02_run2_long_b1.json 2> gpurun_out/r02_run2_long_b1.err; echo "long rc=$?"; cat gpurun_out/r02_run2_long_b1.json | cut -c1-900
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/r02_prefill_launches.csv python profiles/run_prefill.py > gpurun_out/r02_run2_ncu.log 2>&1; echo "ncu rc=$?"
