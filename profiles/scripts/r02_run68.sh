#!/bin/bash
# 96-channel fused residual unit: one ring stage per tap (three K blocks per tensor box) vs one per K block (VAURA_CODEC_RU_KSUB=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec_encode.py tests/test_gpu_shapes.py -m gpu -q -x -k "codec" 2>&1 | tail -2
for ks in 1 3 1 3; do
VAURA_CODEC_RU_KSUB=$ks python - <<PY
import torch, sys, os
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
print("KSUB=$ks codec decode 64 clips ms", round(e0.elapsed_time(e1) / 5, 3))
PY
done
for ks in 3 1; do
VAURA_CODEC_RU_KSUB=$ks ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_ru_fused --csv --log-file gpurun_out/r02_run68_ru_ks$ks.csv python profiles/run_codec.py 16 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_run68_ru_ks$ks.csv | head -4
done
