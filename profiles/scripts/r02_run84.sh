#!/bin/bash
# attention phase at more items than warps (128 rows): second-round items cut into key ranges over the idle warps, merged in the
# CTA.  Parity (full clip, 128 rows with CFG) + same-box A/B against the build before (b64 must not move)
mkdir -p gpurun_out
L=$PWD/vaura_b200/_lib
timeout 1200 python -m pytest tests/test_gpu_fullclip.py tests/test_gpu_parity.py -m gpu -q -x -k "128 or 64_rows or bf16" 2>&1 | tail -3
one() {
  VAURA_B200_LIB=$1 python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],4), d['decode_step']['p50_us'])"
}
for i in 1 2; do
one $L/libvaura_b200_head.so head b64_cfg
one $L/libvaura_b200.so new b64_cfg
done
one $L/libvaura_b200_head.so head b64
one $L/libvaura_b200.so new b64
one $L/libvaura_b200_head.so head b64
one $L/libvaura_b200.so new b64
