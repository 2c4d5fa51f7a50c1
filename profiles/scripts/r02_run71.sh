#!/bin/bash
# final state of the round: all GPU tests, smoke(), default bench + reference arm
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02_run71_pytest.log
echo "pytest wall $(( $(date +%s) - S )) s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
S=$(date +%s)
python bench.py > gpurun_out/r02_run71_bench.json 2> gpurun_out/r02_run71_bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"
python bench.py --impl reference > gpurun_out/r02_run71_ref.json 2> gpurun_out/r02_run71_ref.err; echo "reference arm rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_run71_bench.json'))
print('b64', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['codec']['ms_per_batch'])
for k in ('b1','b64_cfg'):
    print(k, d[k]['e2e'], d[k]['roofline']['frac'], d[k]['decode_step']['p50_us'])
print('long', d['long_b1']['value'], d['long_b1']['ms_per_clip'], d['long_b1']['prefill_ms_per_window'])
print('frames', d['frames_b64']['e2e'], d['frames_b64']['ms_per_step'], d['frames_b64']['avclip']['ms_per_256_segments'])
print('encode', d['codec_encode_b64']['e2e_ms_per_batch'])
print('cpu', d.get('cpu_baseline',{}).get('value'))
print('clocks', d['clocks'])
PY
