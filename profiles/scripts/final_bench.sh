mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_b64.err | tail -1 > gpurun_out/bench_b64.json
for w in b64_cfg b1 b1_cfg; do python bench.py --workload $w --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.json; done
for f in gpurun_out/bench_*.json; do python - "$f" <<'PY'
import json, sys
j = json.load(open(sys.argv[1]))
print(sys.argv[1], j["value"], j["ms_per_step"], j.get("e2e", {}).get("value"), j.get("decode_step"), j["roofline"].get("frac"), j.get("clocks"))
PY
done
