#!/bin/bash
mkdir -p gpurun_out
VAURA_AVCLIP_2CTA=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vit_|gemm_tc' -c 400 --csv --log-file gpurun_out/r02_avclip_launches_2cta.csv python profiles/run_avclip.py 32 1 > gpurun_out/r02_run39_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
txt=open('gpurun_out/r02_avclip_launches_2cta.csv').read()
r=csv.DictReader(txt[txt.find('"ID"'):].splitlines())
rows=[(x['Kernel Name'][:40], float(x['Metric Value'].replace(',',''))/1000.0) for x in r if x.get('Metric Name')=='gpu__time_duration.sum']
g=[t for k,t in rows if 'gemm_tc' in k]
per=g[77:154]
names=['patch']+['t_qkv','t_proj','s_qkv','s_proj','fc1','fc2']*12+['agg_qkv','agg_out','agg_l1','agg_l2']
agg=collections.defaultdict(list)
for n,t in zip(names,per): agg[n].append(t)
for n,v in agg.items(): print(f"{n:8s} n={len(v):2d} avg {sum(v)/len(v):7.1f} us")
print('gemm total per forward', sum(per), 'all kernels per forward', sum(t for _,t in rows)/2)
PY
