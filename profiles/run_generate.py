"""Short generate() run for profiling: python profiles/run_generate.py <batch> <tokens> [cfg_scale] [nocodec]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vaura_b200.synthetic import build_model  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, make_avclip_features  # noqa: E402

B, T = int(sys.argv[1]), int(sys.argv[2])
cfg = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(B, 2).cuda()
for _ in range(2):
    m.generate(frames=feats, max_new_tokens=T, use_sampling=True, top_k=128, prompt_is_encoded=True, cfg_scale=cfg,
               _decode_audio="nocodec" not in sys.argv)
torch.cuda.synchronize()
print("done")
