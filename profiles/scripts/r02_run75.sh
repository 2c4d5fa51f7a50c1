#!/bin/bash
# fused residual units of <= 128 channels skewed by one tile (two acc1 buffers, two h tiles): parity + A/B (VAURA_CODEC_RU_SKEW)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec_encode.py tests/test_gpu_shapes.py -m gpu -q -x -k "codec" 2>&1 | tail -3
for sk in 0 1 0 1; do
VAURA_CODEC_RU_SKEW=$sk python - <<PY
import torch, sys, os
sys.path.insert(0, os.getcwd())
from vaura_b200.codec import DacModelWrapper
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict
m = DacModelWrapper(44100, dims=FULL_CODEC)
m.load_state_dict(make_codec_state_dict(FULL_CODEC, 100, with_encoder=True), device="cuda:0")
codes = torch.randint(0, 1024, (64, 9, 220)).cuda()
for _ in range(3): m.decode(codes)
torch.cuda.synchronize()
def timed(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
dec = timed(lambda: m.decode(codes, validate=False))
wav = torch.randn(64, 1, 220 * 512, device="cuda") * 0.1
for _ in range(2): m.encode(wav)
torch.cuda.synchronize()
enc = timed(lambda: m.encode(wav))
print("SKEW=$sk codec decode 64 clips ms", round(dec, 3), "encode ms", round(enc, 3))
PY
done
