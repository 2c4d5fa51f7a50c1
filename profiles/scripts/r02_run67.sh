#!/bin/bash
# fused residual-unit kernel: 8 vs 16 epilogue warps (VAURA_CODEC_RU_EW), codec decode of 64 clips + encode, same box; parity tests
mkdir -p gpurun_out
VAURA_CODEC_RU_EW=16 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec_encode.py tests/test_gpu_shapes.py -m gpu -q -x -k "codec" 2>&1 | tail -2
for ew in 8 16 8 16; do
VAURA_CODEC_RU_EW=$ew python - <<PY
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
print("EW=$ew codec decode 64 clips ms", round(e0.elapsed_time(e1) / 5, 3))
PY
done
VAURA_CODEC_RU_EW=16 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_ru_fused --csv --log-file gpurun_out/r02_run67_ru16.csv python profiles/run_codec.py 16 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_run67_ru16.csv | head -5
VAURA_CODEC_RU_EW=8 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_ru_fused --csv --log-file gpurun_out/r02_run67_ru8.csv python profiles/run_codec.py 16 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_run67_ru8.csv | head -5
