#!/bin/bash
for v in 1 0; do echo "== VAURA_AVCLIP_M_FASTEST=$v"; VAURA_AVCLIP_M_FASTEST=$v timeout 300 python profiles/run_avclip.py 256 3; done
timeout 300 python -m pytest tests/test_gpu_avclip.py -m gpu -q -x 2>&1 | tail -2
