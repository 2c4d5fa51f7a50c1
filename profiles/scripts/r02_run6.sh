#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16" 2>&1 | tail -3
for cta in 1 100; do
echo "== CTA $cta"; VAURA_TIMING_CTA=$cta timeout 300 python profiles/fused2_timing.py 64 120 2>&1 | tail -24
done > gpurun_out/r02_fused2_timeline_v2.txt 2>&1
cat gpurun_out/r02_fused2_timeline_v2.txt
