#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "snr_bounds" 2>&1 | grep -v "^$" | tail -12
