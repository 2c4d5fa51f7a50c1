#!/bin/bash
for v in 0 1; do echo "== VAURA_CONV_EW16=$v"; VAURA_CONV_EW16=$v timeout 300 python profiles/codec_timing.py 16 2>&1 | grep "^tc"; done
