#!/bin/bash
mkdir -p gpurun_out
for sk in 1 0; do
VAURA_CODEC_RU_SKEW=$sk ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_ru_fused --csv --log-file gpurun_out/r02_run76_ru_sk$sk.csv python profiles/run_codec.py 16 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_run76_ru_sk$sk.csv | head -4
python - <<PY
import csv
rows=[r for r in csv.DictReader([l for l in open('gpurun_out/r02_run76_ru_sk$sk.csv') if not l.startswith('==')]) if r.get('Metric Name')=='gpu__time_duration.sum']
print([ (r['Kernel Name'][:40], r['Metric Value']) for r in rows[-6:]])
PY
done
