#!/bin/bash
for v in 1 0; do echo "== VAURA_AVCLIP_EW8=$v"; VAURA_AVCLIP_EW8=$v timeout 300 python profiles/run_avclip.py 256 3; done
timeout 600 python -m pytest tests/test_gpu_avclip.py -m gpu -q -x 2>&1 | tail -2
