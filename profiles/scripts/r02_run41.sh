#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_run41_bench.json 2> gpurun_out/r02_run41.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_run41.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run41_bench.json'))
print('b64', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['codec']['ms_per_batch'], d['gpu_launches'])
for k in ('b1','b64_cfg'):
    print(k, d[k]['e2e'], d[k]['roofline']['frac'], d[k]['decode_step']['p50_us'])
print('long', d['long_b1']['value'], d['long_b1']['ms_per_clip'], d['long_b1']['prefill_ms_per_window'])
print('frames', d['frames_b64']['e2e'], d['frames_b64']['ms_per_step'], d['frames_b64']['avclip']['ms_per_256_segments'], d['frames_b64']['avclip']['roofline']['frac'])
print('encode', d['codec_encode_b64']['e2e_ms_per_batch'], d['codec_encode_b64']['roofline']['achieved'])
print('cpu', d['cpu_baseline'])
PY
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
python __graft_entry__.py --smoke 2>&1 | tail -6
