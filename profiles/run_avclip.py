"""Segment-AVCLIP tower: time `segments` segments (default 256 = 64 clips x 4) and print per-forward ms and TFLOP/s.
    python profiles/run_avclip.py [segments] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vaura_b200.features import MotionFormer  # noqa: E402
from vaura_b200.synthetic import FULL_AVCLIP, make_motionformer_state_dict  # noqa: E402
from vaura_b200.weights import avclip_flops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = MotionFormer(extract_features=True)
m.load_state_dict(make_motionformer_state_dict(7), device="cuda:0")
m.max_chunk_segments = int(os.environ.get("CHUNK", "32"))
frames = torch.randn(S // 4 if S >= 4 else 1, 4 if S >= 4 else S, 3, 16, 224, 224, device="cuda")
m(frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    m(frames)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nseg = frames.shape[0] * frames.shape[1]
print(f"{nseg} segments: {ms:.2f} ms per forward, {nseg / ms * 1e3:.0f} segments/s, "
      f"{avclip_flops(FULL_AVCLIP, nseg) / ms / 1e9:.0f} TFLOP/s (chunk {m.max_chunk_segments})")
