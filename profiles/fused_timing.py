"""Phase timeline of decode_step_fused_bf16 (CTA 0), last step of a short generate: python profiles/fused_timing.py [B] [T]"""
import os
import sys

os.environ["VAURA_PERSIST_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests.test_gpu_parity import build_model  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, make_avclip_features  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(B, 2).cuda()
for _ in range(2):
    m.generate(frames=feats, max_new_tokens=T, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False)
torch.cuda.synchronize()
ws = m.sampler._buffers["ws"]
t = ws[256:256 + 16384].cpu().numpy().view(np.uint64).astype(np.int64)
L = FULL_SAMPLER.num_layers
# stamps: 2 per barrier (before, after); 7 barriers per layer
names = ["rms1", "qkv", "attn", "wo", "rms2", "w13", "w2"]
work = np.zeros(7)
bar = np.zeros(7)
prev = None
idx = 0
t0 = t[0]
for l in range(L):
    for i in range(7):
        before, after = t[idx], t[idx + 1]
        start = t[idx - 1] if idx > 0 else before
        work[i] += before - start
        bar[i] += after - before
        idx += 2
print(f"layers total {(t[idx - 1] - t0) / 1e3:.1f} us at position {T + 7}; first stamp = first barrier arrival")
for n, w, b in zip(names, work, bar):
    print(f"  {n:5s} work {w / L / 1e3:6.2f} us   barrier wait {b / L / 1e3:6.2f} us   (per layer, CTA 0)")
