#!/bin/bash
for v in 1 5; do
echo "== VAURA_FUSED_L2_PREFETCH=$v"
VAURA_FUSED_L2_PREFETCH=$v python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['decode_step'])"
done
VAURA_FUSED_L2_PREFETCH=5 timeout 600 python -m pytest tests/test_gpu_fullclip.py -m gpu -q -x -k "64_rows" 2>&1 | tail -2
