#!/bin/bash
# decode_step_fused_bf16: (a) warp index / TMEM base / descriptor words broadcast with shfl (uniform registers) vs before;
# (b) timing-only variants that issue 1/2 and 1/4 of the MMAs of a stage (wrong results): is a GEMM phase MMA-rate-bound?
mkdir -p gpurun_out
one() {  # lib label workload
  VAURA_B200_LIB=$1 python bench.py --workload $3 --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2 $3', round(d['value'],1), round(d['roofline']['us_per_launch'],1), d['decode_step']['p50_us'])"
}
L=$PWD/vaura_b200/_lib
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" 2>&1 | tail -2
for i in 1 2; do
one $L/libvaura_b200_old.so old b64
one $L/libvaura_b200.so shfl b64
done
one $L/libvaura_b200_mmadiv2.so mma/2 b64
one $L/libvaura_b200_mmadiv4.so mma/4 b64
one $L/libvaura_b200_old.so old b64_cfg
one $L/libvaura_b200.so shfl b64_cfg
