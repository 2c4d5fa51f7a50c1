#!/bin/bash
# fp32 prefill attention: queries per warp 4 / 2 / 1 (query blocks issued last-first) on one 167-position prompt window
mkdir -p gpurun_out
for q in 4 2 1; do
VAURA_PREFILL_ATTN_QW=$q ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_prefill -c 48 --csv --log-file gpurun_out/r02_run56_attn_qw$q.csv python profiles/run_prefill.py > /dev/null 2>&1
echo "QW=$q"; python profiles/summarize_launches.py gpurun_out/r02_run56_attn_qw$q.csv | grep attn
done
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_driver.py tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x 2>&1 | tail -3
for v in 0 0; do
python bench.py --workload long_b1 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('long_b1', round(d['value'],2), round(d['long_b1']['ms_per_clip'],1), round(d['long_b1']['prefill_ms_per_window'],3))"
done
