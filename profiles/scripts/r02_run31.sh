#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checkpoint.py -m gpu -q -x 2>&1 | tail -2
python profiles/host_glue_profile.py 2>&1 | grep "8-token"
python bench.py --workload b1 --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('b1', d['value'], d['e2e']['value'], d['roofline']['us_per_step_whole_call'], d['roofline']['us_per_launch'])"
