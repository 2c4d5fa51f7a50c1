#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|conv|codes|rvq|enc_' -c 200 --csv --log-file gpurun_out/r02_codec_launches_b16.csv python profiles/run_codec.py 16 > gpurun_out/r02_run19_a.log 2>&1; echo "rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|conv|codes|rvq|enc_' -c 200 --csv --log-file gpurun_out/r02_encode_launches_b8.csv python profiles/run_encode.py 8 > gpurun_out/r02_run19_b.log 2>&1; echo "rc=$?"
python profiles/summarize_launches.py gpurun_out/r02_encode_launches_b8.csv | head
python - <<'PY'
import csv
txt=open('gpurun_out/r02_codec_launches_b16.csv').read()
r=csv.DictReader(txt[txt.find('"ID"'):].splitlines())
rows=[(x['Kernel Name'][:58], float(x['Metric Value'].replace(',',''))/1000.0) for x in r if x.get('Metric Name')=='gpu__time_duration.sum']
print(len(rows), sum(t for _,t in rows))
for i,(k,t) in enumerate(rows): print(i,k,round(t,1))
PY
