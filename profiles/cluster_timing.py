"""Phase timeline of the cluster-persistent decode-step kernel (CTA 0), last step of a short generate.
python profiles/cluster_timing.py [rows] [tokens]"""
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
MODE = sys.argv[3] if len(sys.argv) > 3 else "topk"
SKW = {"topk": dict(use_sampling=True, top_k=128), "argmax": dict(use_sampling=False), "nofilter": dict(use_sampling=True, top_k=0),
       "topp": dict(use_sampling=True, top_k=0, top_p=0.9)}[MODE]
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(B, 2).cuda()
for _ in range(2):
    m.generate(frames=feats, max_new_tokens=T, prompt_is_encoded=True, **SKW, _decode_audio=False)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
m.generate(frames=feats, max_new_tokens=T, prompt_is_encoded=True, **SKW, _decode_audio=False)
ev[1].record()
torch.cuda.synchronize()
print(f"generate {T} tokens: {ev[0].elapsed_time(ev[1]):.2f} ms -> {ev[0].elapsed_time(ev[1]) / (T + 8) * 1e3:.1f} us per step (incl. first pass)")
ws = m.sampler._buffers["ws"]
t = ws[256:256 + 16384].cpu().numpy().view(np.uint64).astype(np.int64)
names = ["stage1", "qkv+xchg", "attn", "wo", "stage2", "w13+xchg", "w2"]
L = FULL_SAMPLER.num_layers
per = np.zeros(len(names))
idx = 1
t0 = t[0]
for l in range(L):
    for i in range(len(names)):
        per[i] += t[idx] - t[idx - 1]
        idx += 1
tail = [t[idx + i] - t[idx + i - 1] for i in range(3)]
print(f"step total {(t[idx + 2] - t0) / 1e3:.1f} us (position {T + 7}), layers {(t[idx - 1] - t0) / 1e3:.1f} us")
for n, v in zip(names, per):
    print(f"  {n:9s} {v / L / 1e3:7.2f} us/layer")
print("  tail: norm+heads %.2f bar %.2f sample %.2f us" % tuple(x / 1e3 for x in tail))

d = t[256:256 + 48]
labels = {0: "qkv u0 wait", 1: "qkv u0 got", 2: "u1 wait", 3: "u1 got", 4: "u2 wait", 5: "u2 got", 6: "u3 wait", 7: "u3 got",
          8: "u4 wait", 9: "u4 got", 10: "u5 wait", 11: "u5 got", 12: "halves written", 13: "cta sync", 14: "rs pushed",
          15: "xchg 1 done", 16: "rope/qkv written", 17: "cta sync", 18: "attn warp partials", 19: "attn cta pushed",
          20: "xchg 2 done", 21: "wo wait", 22: "wo got", 24: "w13 u0 wait", 25: "u0 got", 26: "u1 wait", 27: "u1 got",
          28: "u2 wait", 29: "u2 got", 30: "u3 wait", 31: "u3 got", 32: "part written", 33: "cta sync", 34: "h pushed",
          35: "xchg 3 done", 41: "wo adds issued, arrived", 42: "counter barrier done", 43: "x_mid words complete", 36: "w2 u0 wait", 37: "u0 got", 38: "w2 u1 wait", 39: "u1 got"}
base = d[0]
print("layer %d detail (us since first qkv wait):" % (L // 2))
prev = base
order = list(range(0, 23)) + [41, 42, 43] + list(range(24, 40))
for k in [o for o in order if o in labels]:
    print(f"  {labels[k]:22s} {(d[k] - base) / 1e3:7.2f}  (+{(d[k] - prev) / 1e3:.2f})")
    prev = d[k]
