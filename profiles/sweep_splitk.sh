for cfg in "64 6 64 8" "64 6 64 6" "32 3 32 4" "32 3 32 3" "64 4 64 6" "32 4 32 6" "128 12 128 12"; do
  set -- $cfg
  VAURA_WO_BN=$1 VAURA_WO_KSPLIT=$2 VAURA_W2_BN=$3 VAURA_W2_KSPLIT=$4 timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['decode_step']['mean_us'],1), round(d['value'],1))"
done
