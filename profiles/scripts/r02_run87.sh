#!/bin/bash
python profiles/fused_timing.py 64 120 2>&1 | tail -9
VAURA_TIMING_CTA=50 python profiles/fused_timing.py 64 120 2>&1 | tail -3
