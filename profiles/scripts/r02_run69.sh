#!/bin/bash
mkdir -p gpurun_out
python profiles/fused_arrivals.py 64 120 2>&1 | tail -26
