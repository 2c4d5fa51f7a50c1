#!/bin/bash
# Per-mode kernel time of the stand-alone sampling stage (sample_kernel) from an ncu launch list of sample_micro.py:
# 205 launches per mode, modes in the order argmax / no filter / top-k / top-p, first for 1 clip then for 64 clips.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sample_kernel --csv \
    --log-file gpurun_out/sample_micro_ncu.csv python profiles/sample_micro.py > gpurun_out/sample_micro.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/sample_micro_ncu.csv')) if len(r) > 10 and r[0].isdigit()]
d = [int(r[-1]) for r in rows]
names = ["argmax", "no filter", "top-k", "top-p"]
for i in range(0, len(d), 205):
    seg = sorted(d[i:i + 205])
    print(f"rows {(1, 64)[i // 205 // 4]:3d} {names[i // 205 % 4]:10s} median {seg[len(seg) // 2]} ns  min {seg[0]} ns  ({len(seg)} launches)")
PY
