#!/bin/bash
# decode_step_fused_bf16: device-wide barrier variants (VAURA_FUSED_BARRIER = 0 one poller + acquire loads, 1 relaxed polls + fence,
# 2/3/4 several staggered pollers + shared-memory mbarrier), alternated on one box
mkdir -p gpurun_out
VAURA_FUSED_BARRIER=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" 2>&1 | tail -3
VAURA_FUSED_BARRIER=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" 2>&1 | tail -3
for v in 0 1 2 3 4 0 1 2 3 4; do
VAURA_FUSED_BARRIER=$v python bench.py --workload b64 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>gpurun_out/r02_run52_b64_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('barrier $v b64', round(d['value'],1), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],4), d['decode_step']['p50_us'])"
done
