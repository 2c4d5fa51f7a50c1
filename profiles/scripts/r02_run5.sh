#!/bin/bash
mkdir -p gpurun_out
for cta in 1 100; do for fl in 0 1; do
echo "== CTA $cta flags $fl"; VAURA_TIMING_CTA=$cta VAURA_FUSED2_FLAGS=$fl timeout 300 python profiles/fused2_timing.py 64 120 2>&1 | tail -16
done; done > gpurun_out/r02_fused2_timeline_v1.txt 2>&1
cat gpurun_out/r02_fused2_timeline_v1.txt
