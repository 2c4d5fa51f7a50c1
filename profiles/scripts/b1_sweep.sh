for tu in 0 6 13 20 26; do
  echo "== tail_units $tu"
  VAURA_CLUSTER_TAIL_UNITS=$tu python bench.py --workload b1 --no-cpu-baseline --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print(j['value'], j['ms_per_step'], j.get('decode_step'), j['roofline']['frac'])
"
done
