#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_|conv|codes' -c 100 --csv --log-file gpurun_out/r02_codec_launches_b16_fused.csv python profiles/run_codec.py 16 > gpurun_out/r02_run21.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
txt=open('gpurun_out/r02_codec_launches_b16_fused.csv').read()
r=csv.DictReader(txt[txt.find('"ID"'):].splitlines())
rows=[(x['Kernel Name'][:70], float(x['Metric Value'].replace(',',''))/1000.0) for x in r if x.get('Metric Name')=='gpu__time_duration.sum']
n=len(rows)//2
print(n, sum(t for _,t in rows[n:]))
for i,(k,t) in enumerate(rows[n:]): print(i,k,round(t,1))
PY
