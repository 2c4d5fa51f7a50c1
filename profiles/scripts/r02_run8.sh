#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_avclip.py -m gpu -q -x -s 2>&1 | tail -40
