#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r02_run10_bench.json 2> gpurun_out/r02_run10_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_run10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run10_bench.json'))
print(json.dumps(d.get('frames_b64'))[:1500])
print(d['value'], d['e2e']['value'], d['long_b1']['value'])
PY
