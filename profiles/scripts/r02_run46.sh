#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py > gpurun_out/r02_run46_bench.json 2> gpurun_out/r02_run46.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run46_bench.json'))
print('b64', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['codec']['ms_per_batch'])
for k in ('b1','b64_cfg'):
    print(k, d[k]['e2e'], d[k]['roofline']['frac'], d[k]['decode_step']['p50_us'])
print('long', d['long_b1']['value'], d['long_b1']['ms_per_clip'], d['long_b1']['prefill_ms_per_window'])
print('frames', d['frames_b64']['e2e'], d['frames_b64']['ms_per_step'], d['frames_b64']['avclip']['ms_per_256_segments'])
print('encode', d['codec_encode_b64']['e2e_ms_per_batch'])
print('clocks', d['clocks'])
PY
