#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_avclip.py -m gpu -q -x 2>&1 | tail -3
for c in 8 32 64; do CHUNK=$c timeout 300 python profiles/run_avclip.py 256 3; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vit_|gemm_tc' -c 400 --csv --log-file gpurun_out/r02_avclip_launches.csv python profiles/run_avclip.py 32 1 > gpurun_out/r02_run9_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r02_avclip_launches.csv | head -20
