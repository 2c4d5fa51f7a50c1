#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x 2>&1 | tail -5
