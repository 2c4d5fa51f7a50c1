#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1; do echo "== VAURA_BF16_STEP_FIRST=$v"; VAURA_BF16_STEP_FIRST=$v python bench.py --no-sub --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"; done
