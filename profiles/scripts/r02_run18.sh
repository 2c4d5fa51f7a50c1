#!/bin/bash
python -m pytest tests/test_gpu_fullclip.py -m gpu -q -x -s -k fused2 2>&1 | tail -30
