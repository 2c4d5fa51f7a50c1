"""Codec-only run for profiling: python profiles/run_codec.py <batch>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vaura_b200.codec import DacModelWrapper  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = DacModelWrapper(44100, dims=FULL_CODEC)
m.load_state_dict(make_codec_state_dict(FULL_CODEC, 100), device="cuda:0")
codes = torch.randint(0, 1024, (B, 9, 220)).cuda()
for _ in range(2):
    m.decode(codes)
torch.cuda.synchronize()
print("done")
