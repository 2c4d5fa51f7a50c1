#!/bin/bash
timeout 900 python -m pytest tests/test_driver.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
python bench.py --workload long_b1 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('long_b1', d['value'], d['long_b1']['ms_per_clip'], d['long_b1']['prefill_ms_per_window'])"
