#!/bin/bash
for i in 1 2 3; do
python bench.py --workload b1 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('b1', round(d['value'],2), round(d['roofline']['us_per_launch'],2), round(d['roofline']['frac'],4), d['decode_step']['p50_us'], d['clocks'])"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
