#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/g1.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(1, 2).cuda()
kw = dict(max_new_tokens=8, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False)
for _ in range(2): m.generate(frames=feats, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.generate(frames=feats, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_b1_call_launches.csv python /tmp/g1.py > gpurun_out/r02_run30.log 2>&1; echo rc=$?
python - <<'PY'
import csv
txt=open('gpurun_out/r02_b1_call_launches.csv').read()
r=csv.DictReader(txt[txt.find('"ID"'):].splitlines())
rows=[(x['Kernel Name'][:90], float(x['Metric Value'].replace(',',''))/1000.0) for x in r if x.get('Metric Name')=='gpu__time_duration.sum']
print(len(rows), sum(t for _,t in rows))
for i,(k,t) in enumerate(rows): print(i,k,round(t,1))
PY
