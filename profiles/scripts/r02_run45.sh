#!/bin/bash
for v in 4096 8192 100000; do
echo "== VAURA_PREFILL_BN256_FROM=$v"
VAURA_PREFILL_BN256_FROM=$v python bench.py --workload long_b1 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('long_b1', d['value'], d['long_b1']['ms_per_clip'], d['long_b1']['prefill_ms_per_window'])"
done
