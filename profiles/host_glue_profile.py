"""Where the host time of a B = 1 generate() call goes (cProfile over 20 calls, token-only):  python profiles/host_glue_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features  # noqa: E402

m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(1, 2).cuda()
kw = dict(max_new_tokens=220, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False)
for _ in range(3):
    m.generate(frames=feats, **kw)
torch.cuda.synchronize()
# host time of a call when the GPU is not the limit: 8-token clips
kw_short = dict(kw, max_new_tokens=8)
for _ in range(3):
    m.generate(frames=feats, **kw_short)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    m.generate(frames=feats, **kw_short)
torch.cuda.synchronize()
print(f"8-token generate(): {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call (16 steps of 0.31 ms = 5 ms of GPU)")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    m.generate(frames=feats, **kw_short)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
