"""DAC encode of `B` clips of 2.56 s for an ncu launch list:  python profiles/run_encode.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vaura_b200.codec import DacModelWrapper  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = DacModelWrapper(44100, dims=FULL_CODEC)
m.load_state_dict(make_codec_state_dict(FULL_CODEC, 100, with_encoder=True), device="cuda:0")
wav = 0.3 * torch.randn(B, 1, 220 * 512, device="cuda")
m.encode(wav)
torch.cuda.synchronize()
print("done")
