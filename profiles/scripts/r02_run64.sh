#!/bin/bash
# provably warp-uniform warp index / TMEM base in the codec, prefill and AVCLIP tcgen05 kernels as well: same-box A/B against the
# previous build (libvaura_b200_head.so)
mkdir -p gpurun_out
L=$PWD/vaura_b200/_lib
timeout 900 python -m pytest tests/test_gpu_avclip.py tests/test_gpu_codec_encode.py tests/test_gpu_parity.py tests/test_gpu_prefill.py -m gpu -q -x 2>&1 | tail -2
for lib in head new head new; do
  f=$L/libvaura_b200.so; [ $lib = head ] && f=$L/libvaura_b200_head.so
  VAURA_B200_LIB=$f python - <<PY
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
from vaura_b200.codec import DacModelWrapper
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict
m = DacModelWrapper(44100, dims=FULL_CODEC)
m.load_state_dict(make_codec_state_dict(FULL_CODEC, 100), device="cuda:0")
codes = torch.randint(0, 1024, (64, 9, 220)).cuda()
for _ in range(3): m.decode(codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m.decode(codes, validate=False)
e1.record(); torch.cuda.synchronize()
print("$lib codec decode 64 clips ms", round(e0.elapsed_time(e1) / 5, 3))
PY
done
for lib in head new; do
  f=$L/libvaura_b200.so; [ $lib = head ] && f=$L/libvaura_b200_head.so
  VAURA_B200_LIB=$f python bench.py --workload b64 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib frames', round(d['frames_b64']['avclip']['ms_per_256_segments'],2), 'encode', round(d['codec_encode_b64']['e2e_ms_per_batch'],2), 'prefill', round(d['long_b1']['prefill_ms_per_window'],3), 'b64', round(d['value'],1))"
done
