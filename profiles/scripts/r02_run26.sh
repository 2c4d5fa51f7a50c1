#!/bin/bash
timeout 300 python profiles/run_avclip.py 256 3
timeout 600 python -m pytest tests/test_gpu_avclip.py tests/test_gpu_parity.py -m gpu -q -x -s -k "avclip or features or chunked or passthrough or generate_from or pinned or snr_bounds" 2>&1 | tail -6
