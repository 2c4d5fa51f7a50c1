#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|attn|rmsnorm|embed|gemv|sample' -c 400 --csv --log-file gpurun_out/r02_prefill_launches_v5.csv python profiles/run_prefill.py > gpurun_out/r02_run44_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r02_prefill_launches_v5.csv | head -12
