#!/bin/bash
# attn_prefill_kernel<2> (16 queries per CTA) check: GPU tests, long_b1 timing, prefill launch list
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02_run50_pytest.log
echo "pytest wall $(( $(date +%s) - S )) s"
python bench.py --workload long_b1 --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/r02_run50_long.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('long_b1', d['value'], d['long_b1']['ms_per_clip'], d['long_b1']['prefill_ms_per_window'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_run50_prefill_launches.csv python profiles/run_prefill.py > gpurun_out/r02_run50_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_run50_prefill_launches.csv 2>/dev/null | head -30
