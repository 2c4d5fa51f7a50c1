"""Decode time per step at small row counts on the two precision paths: python profiles/scripts/rows_sweep.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vaura_b200.synthetic import build_model  # noqa: E402
from vaura_b200 import _cabi  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, make_avclip_features  # noqa: E402

m = build_model(FULL_SAMPLER, FULL_CODEC)
T = 60
for B in (1, 2, 3, 4, 6, 8, 12, 16, 32):
    feats = make_avclip_features(B, 2).cuda()
    for name, prec in (("fp32act", _cabi.PRECISION_FP32ACT), ("bf16", _cabi.PRECISION_BF16)):
        kw = dict(frames=feats, max_new_tokens=T, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False,
                  _precision=prec)
        m.generate(**kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.generate(**kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"rows {B:3d} {name:8s} {dt / (T + 8) * 1e6:8.1f} us per step")
