#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r02_run15_bench.json 2> gpurun_out/r02_run15.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_run15.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run15_bench.json'))
print('b64', d['value'], d['e2e']['value'], {k:d['roofline'][k] for k in ('frac','us_per_launch','launches_timed','us_per_step_whole_call','frac_whole_call')})
for k in ('b1','b64_cfg'):
    print(k, d[k]['e2e'], {kk:d[k]['roofline'][kk] for kk in ('frac','us_per_launch','launches_timed','us_per_step_whole_call','frac_whole_call')}, d[k]['decode_step']['p50_us'])
print('long', d['long_b1']['value'], d['long_b1']['prefill_ms_per_window'])
print('frames', d['frames_b64']['e2e'], d['frames_b64']['avclip']['roofline']['frac'])
PY
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
