#!/bin/bash
# same-box A/B after making the reproducible mode a separate instantiation: tree at cd71d29 (_ab_old/) vs current, alternated
mkdir -p gpurun_out
one() {  # dir label workload
  (cd $1 && python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), d['decode_step']['p50_us'], round(d.get('codec',{}).get('ms_per_batch',0),2), round(d['ms_per_step'],2))")
}
for i in 1 2 3; do
one _ab_old old b64
one . new b64
done
VAURA_DETERMINISTIC=1 one . new-det b64
one _ab_old old b64_cfg
one . new b64_cfg
python profiles/run_codec.py 16 2>&1 | tail -2
(cd _ab_old && python profiles/run_codec.py 16 2>&1 | tail -2)
