#!/bin/bash
mkdir -p gpurun_out
for w in b1_cfg b64_cfg dataset long_b1; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02_run42_$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value'],2), d.get('e2e',{}).get('value'), d['ms_per_step'])" || tail -3 gpurun_out/r02_run42_$w.err
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['value'], [k for k in d if isinstance(d[k], dict) and 'error' in d[k]])"
