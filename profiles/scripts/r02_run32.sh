#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/g64.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(64, 2).cuda()
ids = torch.arange(64, dtype=torch.int32)
kw = dict(max_new_tokens=220, use_sampling=True, top_k=128, prompt_is_encoded=True)
m.generate(frames=feats, clip_indices=ids, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.generate(frames=feats, clip_indices=ids, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
VAURA_FUSED2_NOCOOP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_b64_call_launches.csv python /tmp/g64.py > gpurun_out/r02_run32.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections
txt=open('gpurun_out/r02_b64_call_launches.csv').read()
r=csv.DictReader(txt[txt.find('"ID"'):].splitlines())
rows=[(x['Kernel Name'][:80], float(x['Metric Value'].replace(',',''))/1000.0) for x in r if x.get('Metric Name')=='gpu__time_duration.sum']
agg=collections.OrderedDict()
for k,t in rows:
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(t for _,t in rows)
print(len(rows), tot)
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]: print(f"{n:5d} {t:10.1f} {k}")
PY
