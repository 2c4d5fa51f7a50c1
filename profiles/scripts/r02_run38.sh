#!/bin/bash
VAURA_AVCLIP_2CTA=1 timeout 180 python -m pytest tests/test_gpu_avclip.py -m gpu -q -x -s -k "golden or chunked" 2>&1 | grep -v "^$" | tail -12
echo "rc=$?"
for v in 0 1; do echo "== VAURA_AVCLIP_2CTA=$v"; VAURA_AVCLIP_2CTA=$v timeout 120 python profiles/run_avclip.py 256 3 2>&1 | tail -2; done
