#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python profiles/sanitize_small.py avclip 2>&1 | grep -v "^$" | tail -3
