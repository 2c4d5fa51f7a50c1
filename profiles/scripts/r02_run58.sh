#!/bin/bash
# same-box A/B: the tree at cd71d29 (start of this session, _ab_old/) against the current tree, b64 and b1, alternated
mkdir -p gpurun_out
one() {  # dir label workload
  (cd $1 && python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],4), d['decode_step']['p50_us'], round(d.get('codec',{}).get('ms_per_batch',0),2))")
}
for i in 1 2 3; do
one _ab_old old b64
one . new b64
done
one _ab_old old b1
one . new b1
one _ab_old old b1
one . new b1
one _ab_old old b64_cfg
one . new b64_cfg
