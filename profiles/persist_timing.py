"""Phase timeline of the persistent decode-step kernel (CTA 0), last step of a short generate.
VAURA_PERSIST_TIMING=1 python profiles/persist_timing.py [rows]"""
import os
import sys

os.environ["VAURA_PERSIST_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vaura_b200.synthetic import build_model  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, make_avclip_features  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(B, 2).cuda()
for _ in range(2):
    m.generate(frames=feats, max_new_tokens=T, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False)
torch.cuda.synchronize()
ws = m.sampler._buffers["ws"]
t = ws[256:256 + 16384].cpu().numpy().view(np.uint64).astype(np.int64)
names = ["stage1", "qkv", "bar1", "attn", "bar2", "comb", "wo", "bar3", "stage4", "w13", "bar4", "stage5", "w2", "bar5"]
L = FULL_SAMPLER.num_layers
per = np.zeros(len(names))
idx = 1
t0 = t[0]
for l in range(L):
    for i in range(len(names)):
        per[i] += t[idx] - t[idx - 1]
        idx += 1
tail = [t[idx + i] - t[idx + i - 1] for i in range(4)]
print(f"step total {(t[idx + 3] - t0) / 1e3:.1f} us (position {T + 7})")
for n, v in zip(names, per):
    print(f"  {n:8s} {v / L / 1e3:7.2f} us/layer")
print("  tail: final-norm %.2f heads %.2f bar %.2f sample %.2f us" % tuple(x / 1e3 for x in tail))


