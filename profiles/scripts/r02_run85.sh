#!/bin/bash
# fused step kernel: wo without K split (96 tiles of 16 columns) with the FFN RMSNorm folded into its epilogue and into the w1|w3
# epilogue, against 24 x 6 split-K tiles + float reductions + a separate norm: parity tests, then same-box A/B
mkdir -p gpurun_out
L=$PWD/vaura_b200/_lib
timeout 1500 python -m pytest tests/test_gpu_fullclip.py tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|error|logit|agreement" | tail -12
one() {
  VAURA_B200_LIB=$1 python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],4), d['decode_step']['p50_us'])"
}
for i in 1 2; do
one $L/libvaura_b200_head.so head b64
one $L/libvaura_b200.so fold b64
done
one $L/libvaura_b200_head.so head b64_cfg
one $L/libvaura_b200.so fold b64_cfg
