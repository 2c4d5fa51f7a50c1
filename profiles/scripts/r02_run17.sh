#!/bin/bash
for v in 0 1; do echo "== VAURA_CONV_OCC2=$v"; VAURA_CONV_OCC2=$v timeout 300 python profiles/codec_timing.py 2>&1 | tail -4; done
for v in 1 0; do echo "== VAURA_AVCLIP_M_FASTEST=$v"; VAURA_AVCLIP_M_FASTEST=$v timeout 300 python profiles/run_avclip.py 256 3; done
timeout 600 python -m pytest tests/test_gpu_avclip.py tests/test_gpu_parity.py -m gpu -q -x -k "avclip or codec or features or chunked or passthrough or generate_from" 2>&1 | tail -2
