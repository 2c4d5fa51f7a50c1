#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r02_run22_bench.json 2> gpurun_out/r02_run22.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_run22.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run22_bench.json'))
print('b64', d['value'], d['e2e']['value'], d['ms_per_step'], d['codec'])
for k in ('b1','b64_cfg'):
    print(k, d[k]['e2e'], d[k]['roofline']['frac'])
print('long', d['long_b1']['value'], d['long_b1']['prefill_ms_per_window'])
print('frames', d['frames_b64']['e2e'], d['frames_b64']['avclip']['roofline']['frac'])
print('encode', d['codec_encode_b64'])
PY
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
