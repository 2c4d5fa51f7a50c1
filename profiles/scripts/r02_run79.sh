#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --kernel-name-base demangled -k regex:decode_step_fused --launch-skip 114 --launch-count 1 --set full --import-source on --clock-control none -f -o gpurun_out/r02_ncu_fused_b64_end python profiles/run_generate.py 64 120 1.0 nocodec > /dev/null 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_ncu_fused_b64_end.ncu-rep
