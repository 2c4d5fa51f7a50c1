for cfg in "6 6" "4 6" "3 6" "6 4" "4 4" "3 4" "3 3"; do
  set -- $cfg
  r=$(VAURA_WO_KSPLIT=$1 VAURA_W2_KSPLIT=$2 python bench.py --workload b64 --no-cpu-baseline --steps 2 --warmup 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['value'],1), j['decode_step']['p50_us'])")
  echo "wo_ksplit $1 w2_ksplit $2 : $r"
done
