#!/bin/bash
mkdir -p gpurun_out
# headline kernel at a mid-clip position (launch 114 of a 120-token generate), the CTA-pair GEMM of the tower, a fused residual unit
timeout 900 ncu --kernel-name-base demangled -k regex:decode_step_fused --launch-skip 114 --launch-count 1 --set full --import-source on --clock-control none -f -o gpurun_out/r02_ncu_fused_b64 python profiles/run_generate.py 64 120 1.0 nocodec > /dev/null 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_persistent' -s 19 -c 1 -f -o gpurun_out/r02_ncu_avclip_gemm_2cta python profiles/run_avclip.py 32 1 > /dev/null 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_ru_fused_kernel<192' -s 1 -c 1 -f -o gpurun_out/r02_ncu_ru_fused_192 python profiles/run_codec.py 16 > /dev/null 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
