#!/bin/bash
# decode_step_fused_bf16: K blocks per ring stage 4 (default: 3 stages of 64 KB) vs 2 (6 x 32 KB) vs 1 (12 x 16 KB), same box
mkdir -p gpurun_out
one() {  # lib label workload
  VAURA_B200_LIB=$1 python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), d['decode_step']['p50_us'])"
}
L=$PWD/vaura_b200/_lib
VAURA_B200_LIB=$L/libvaura_b200_ksub2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" 2>&1 | tail -2
VAURA_B200_LIB=$L/libvaura_b200_ksub1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" 2>&1 | tail -2
for i in 1 2; do
one $L/libvaura_b200.so ksub4 b64
one $L/libvaura_b200_ksub2.so ksub2 b64
one $L/libvaura_b200_ksub1.so ksub1 b64
done
one $L/libvaura_b200.so ksub4 b64_cfg
one $L/libvaura_b200_ksub2.so ksub2 b64_cfg
one $L/libvaura_b200_ksub1.so ksub1 b64_cfg
