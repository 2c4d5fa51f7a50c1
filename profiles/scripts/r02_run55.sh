#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_prefill_kernel' -s 12 -c 1 -f -o gpurun_out/r02_ncu_attn_prefill python profiles/run_prefill.py > /dev/null 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
