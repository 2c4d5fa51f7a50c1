#!/bin/bash
# prompt prefill of a sampling call with bf16 operands (transformer_pass_tc1) vs the three-term prefill (VAURA_PREFILL_BF16=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_driver.py tests/test_gpu_parity.py -m gpu -q -x -s 2>&1 | grep -E "prefill|passed|failed|Error|error" | tail -15
for v in 1 0 1 0; do
VAURA_PREFILL_BF16=$v python bench.py --workload long_b1 --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/r02_run54_long_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PREFILL_BF16=$v long_b1', round(d['value'],2), round(d['long_b1']['ms_per_clip'],1), round(d['long_b1']['prefill_ms_per_window'],3))"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_run54_prefill_launches.csv python profiles/run_prefill.py > gpurun_out/r02_run54_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_run54_prefill_launches.csv 2>/dev/null | grep -v "at::" | head -14
