#!/bin/bash
for v in 1 3; do
echo "== VAURA_FUSED_L2_PREFETCH=$v"
VAURA_FUSED_L2_PREFETCH=$v python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['decode_step'])"
done
