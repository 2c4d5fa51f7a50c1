"""Phase timeline of decode_step_fused_bf16 (CTA 0), last step of a short generate: python profiles/fused_timing.py [B] [T]"""
import os
import sys

os.environ["VAURA_PERSIST_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vaura_b200.synthetic import build_model  # noqa: E402
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
# stamps: 2 per barrier (before, after); one barrier after the first norm, then 5 per layer.  "work" is thread 0's
# (the TMA producer's) time from the previous barrier to its arrival; the rest of the CTA's phase shows up as wait.
names = ["qkv", "attn", "wo+rms", "w13", "w2+rms"]
work = np.zeros(5)
bar = np.zeros(5)
idx = 2  # skip the barrier after the first norm
t0 = t[1]
for l in range(L):
    for i in range(5):
        before, after = t[idx], t[idx + 1]
        work[i] += before - t[idx - 1]
        bar[i] += after - before
        idx += 2
print(f"layers total {(t[idx - 1] - t0) / 1e3:.1f} us at position {T + 7}")
# outside the layers: kernel start (step_times[offset], stamped by CTA 0 at entry) -> first barrier open (embedding + first norm),
# last layer's barrier open -> heads tiles done -> barrier open -> next launch's start (sampling, teardown, launch gap)
st = t[1024:1024 + 300]
off = T + 8  # the last launch of the call sampled column `off`
if st[off] and st[off - 1]:
    print(f"step period {(st[off] - st[off - 1]) / 1e3:.1f} us; last launch: entry -> first barrier open {(t[1] - st[off]) / 1e3:.2f} us, "
          f"heads tiles {(t[idx] - t[idx - 1]) / 1e3:.2f} us, heads barrier {(t[idx + 1] - t[idx]) / 1e3:.2f} us")
    print(f"previous launch's layers end -> this launch's entry (heads + barrier + sampling + teardown + launch gap): "
          f"{(st[off] - st[off - 1]) / 1e3 - (t[idx - 1] - st[off]) / 1e3:.2f} us (period - (entry -> layers end))")
for n, w, b in zip(names, work, bar):
    print(f"  {n:7s} work {w / L / 1e3:6.2f} us   barrier wait {b / L / 1e3:6.2f} us   phase {(w + b) / L / 1e3:6.2f} us (per layer, CTA {os.environ.get('VAURA_TIMING_CTA', '0')})")
a = t[900:908]
if a[0] and a[7] > a[0]:
    lab = ["page table + first copies issued", "q loaded", "first K run landed", "scores done", "softmax done", "first V run landed", "P.V done"]
    print("attention sub-phases of warp 0, last layer (us): " + ", ".join(f"{n} {(a[i + 1] - a[i]) / 1e3:.2f}" for i, n in enumerate(lab)))
for name, base in (("wqkv", 910), ("w1|w3", 940)):
    d = t[base:base + 27]
    if d[26]:
        rel = lambda v: (v - d[26]) / 1e3
        print(f"{name} tile of the last layer, us after the producer's entry: loads issued " + " ".join(f"{rel(v):.2f}" for v in d[0:6]) +
              " | landed " + " ".join(f"{rel(v):.2f}" for v in d[8:14]) + " | MMAs issued " + " ".join(f"{rel(v):.2f}" for v in d[16:22]) +
              f" | accumulator complete {rel(d[24]):.2f} | epilogue done {rel(d[25]):.2f}")
