#!/bin/bash
# phase timeline of decode_step_fused_bf16 at the end of round 2 (CTA 0 and CTA 100, position 127 and 227)
mkdir -p gpurun_out
python profiles/fused_timing.py 64 120 2>&1 | tail -9
VAURA_TIMING_CTA=100 python profiles/fused_timing.py 64 120 2>&1 | tail -9
python profiles/fused_timing.py 64 220 2>&1 | tail -9
