#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_avclip.py tests/test_gpu_codec_encode.py -m gpu -q -x 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/r02_run23_bench.json 2> gpurun_out/r02_run23.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_run23.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run23_bench.json'))
print('b64', d['value'], d['e2e']['value'])
print('frames', d['frames_b64']['e2e'], d['frames_b64']['ms_per_step'], d['frames_b64']['avclip']['ms_per_256_segments'])
PY
