#!/bin/bash
python profiles/fused_timing.py 64 120 2>&1 | tail -11
