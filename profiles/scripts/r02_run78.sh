#!/bin/bash
mkdir -p gpurun_out
for w in prefill_bf16 step; do
  echo "== $w"
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python profiles/sanitize_small.py $w 2>&1 | grep -v "^$" | tail -5
done >> gpurun_out/r02_sanitizer3.txt 2>&1
tail -14 gpurun_out/r02_sanitizer3.txt
