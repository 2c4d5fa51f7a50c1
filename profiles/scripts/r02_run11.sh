#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_codec_encode.py -m gpu -q -x -s 2>&1 | tail -30
