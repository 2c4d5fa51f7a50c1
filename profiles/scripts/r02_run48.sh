#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ru_fused -s 1 -c 1 -f -o gpurun_out/r02_ncu_ru_fused_192 python profiles/run_codec.py 16 2>&1 | tail -3; echo "rc=$?"
ls -la gpurun_out/r02_ncu_ru_fused_192.ncu-rep
