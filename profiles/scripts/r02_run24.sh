#!/bin/bash
mkdir -p gpurun_out
for w in avclip encode decode prefill; do
  echo "== $w"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python profiles/sanitize_small.py $w 2>&1 | grep -v "^$" | tail -6
done > gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
