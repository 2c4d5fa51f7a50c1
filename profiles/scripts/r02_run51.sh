#!/bin/bash
# decode_step_fused_bf16: split-K partial slices added by the row CTAs (deterministic) vs float reductions (VAURA_FUSED_ATOMIC=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullclip.py tests/test_gpu_shapes.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_run51_pytest.log
for v in 0 1 0 1; do
echo "== VAURA_FUSED_ATOMIC=$v"
VAURA_FUSED_ATOMIC=$v python bench.py --workload b64 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>gpurun_out/r02_run51_b64_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('b64', d['value'], d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['decode_step']['p50_us'])"
done
VAURA_FUSED_ATOMIC=0 python bench.py --workload b64_cfg --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('b64_cfg part', d['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
VAURA_FUSED_ATOMIC=1 python bench.py --workload b64_cfg --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('b64_cfg atomic', d['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
