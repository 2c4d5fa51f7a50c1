#!/bin/bash
for v in 0 1; do echo "== VAURA_CODEC_FUSED_RU=$v"; VAURA_CODEC_FUSED_RU=$v timeout 300 python profiles/codec_timing.py 16 2>&1 | grep "^tc"; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec_encode.py -m gpu -q -x -k "codec or encode" 2>&1 | tail -3
