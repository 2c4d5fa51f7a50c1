#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -8
# ncu --set full of the tower's dominant kernels (32 segments): the q|k|v GEMM of a block and the space attention
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_persistent' -s 20 -c 1 -o gpurun_out/r02_ncu_avclip_gemm python profiles/run_avclip.py 32 1 > gpurun_out/r02_run14_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'space_attn_tc' -s 3 -c 1 -o gpurun_out/r02_ncu_avclip_space python profiles/run_avclip.py 32 1 > gpurun_out/r02_run14_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep
