#!/bin/bash
# fused step kernel: rolled MMA loop over the K blocks of a stage (smaller code) vs the fully unrolled one, same box
L=$PWD/vaura_b200/_lib
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k bf16 2>&1 | tail -2
one() {
  VAURA_B200_LIB=$1 python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],4), d['decode_step']['p50_us'])"
}
for i in 1 2; do
one $L/libvaura_b200_head.so head b64
one $L/libvaura_b200.so rolled b64
done
one $L/libvaura_b200_head.so head b64_cfg
one $L/libvaura_b200.so rolled b64_cfg
