#!/bin/bash
# decode_step_cluster with a provably warp-uniform warp index: same-box A/B at batch 1 (and batch 1 + CFG = 2 rows)
mkdir -p gpurun_out
L=$PWD/vaura_b200/_lib
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullclip.py -m gpu -q -x -k "not bf16" 2>&1 | tail -2
one() {
  VAURA_B200_LIB=$1 python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],2), round(d['roofline']['us_per_launch'],2), round(d['roofline']['frac'],4), d['decode_step']['p50_us'])"
}
for i in 1 2 3; do
one $L/libvaura_b200_head.so head b1
one $L/libvaura_b200.so new b1
done
one $L/libvaura_b200_head.so head b1_cfg
one $L/libvaura_b200.so new b1_cfg
