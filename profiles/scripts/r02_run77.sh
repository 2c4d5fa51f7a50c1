#!/bin/bash
# memcheck of the kernels touched in the third session (skewed residual units in the encoder, bf16-operand prefill, prefill
# attention with two queries per warp, fused step kernel) on small inputs + the codec / encode tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_codec_encode.py tests/test_gpu_shapes.py -m gpu -q -x 2>&1 | tail -2
for w in encode decode prefill; do
  echo "== $w"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python profiles/sanitize_small.py $w 2>&1 | grep -v "^$" | tail -4
done > gpurun_out/r02_sanitizer3.txt 2>&1
cat gpurun_out/r02_sanitizer3.txt
