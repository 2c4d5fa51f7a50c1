#!/bin/bash
# phase table of the 128-row instance (64 clips with CFG): python profiles/fused_timing.py does not set cfg -> use 128 clips without CFG (same 128 rows)
python profiles/fused_timing.py 128 120 2>&1 | tail -11
