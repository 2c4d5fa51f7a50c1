"""Phase timeline of decode_step_fused2 (one CTA's barrier-polling warp), last step of a short generate:
    python profiles/fused2_timing.py [B] [T]        (VAURA_TIMING_CTA selects the CTA)
Stamps: timing[2 b] = warp 1 starts polling barrier b (its MMAs of the phase are issued), timing[2 b + 1] = barrier b passed.
Barrier 1 follows the embedding; per layer: q|k|v, attention, wo, w1|w3, w2; then the heads."""
import os
import sys

os.environ["VAURA_PERSIST_TIMING"] = "1"
os.environ.setdefault("VAURA_FUSED2", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(B, 2).cuda()
for _ in range(2):
    m.generate(frames=feats, max_new_tokens=T, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False)
torch.cuda.synchronize()
ws = m.sampler._buffers["ws"]
t = ws[256:256 + 8192].cpu().numpy().view(np.uint64).astype(np.int64)
L = FULL_SAMPLER.num_layers
names = ["qkv", "attn", "wo", "w13", "w2"]
phase = np.zeros(5)
poll = np.zeros(5)
for l in range(L):
    for i in range(5):
        b = 2 + 5 * l + i           # barrier that ends this phase
        phase[i] += t[2 * b + 1] - t[2 * (b - 1) + 1]
        poll[i] += t[2 * b + 1] - t[2 * b]
tot = t[2 * (5 * L + 1) + 1] - t[3]
print(f"layers total {tot / 1e3:.1f} us at position {T + 7}, CTA {os.environ.get('VAURA_TIMING_CTA', '0')}")
for n, ph, po in zip(names, phase, poll):
    print(f"  {n:5s} phase {ph / L / 1e3:6.2f} us per layer (of which the polling warp waited {po / L / 1e3:6.2f} us at the barrier)")
print(f"  embedding {(t[3] - t[2]) / 1e3:.2f} us poll; heads phase {(t[2 * (5 * L + 2) + 1] - t[2 * (5 * L + 1) + 1]) / 1e3:.2f} us")
# work-warp stamps of the last layer (warps 2 and 4, lane 0): 0 go seen, 1 before the accumulator wait, 2 accumulator complete,
# 3 exchange done, 6 before the work-warp barrier, 7 after the arrival
for wv, name in ((0, "warp 2"), (1, "warp 4")):
    for ph, pn in enumerate(names):
        d = t[600 + 200 * wv + 8 * ph: 600 + 200 * wv + 8 * ph + 8]
        if d[0] == 0:
            continue
        rel = lambda i: (d[i] - d[0]) / 1e3 if d[i] else float("nan")
        print(f"  {name} {pn:5s} last layer, us after go: rs/arm done {rel(1):5.2f}  acc complete {rel(2):5.2f}  exchange done {rel(3):5.2f}  "
              f"epilogue done {rel(6):5.2f}  arrived {rel(7):5.2f}")
# MMA-warp stamps of the last layer: 0 go, 1 activation loads issued, 2 first weight stage there, 3-5 activation boxes there, 8 MMAs issued
for ph, pn in enumerate(names):
    d = t[400 + 16 * ph: 400 + 16 * ph + 16]
    if d[0] == 0:
        continue
    rel = lambda i: f"{(d[i] - d[0]) / 1e3:5.2f}" if d[i] else "  -  "
    print(f"  MMA warp {pn:5s} last layer, us after go: loads issued {rel(1)}  weights there {rel(2)}  boxes there {rel(3)} {rel(4)} {rel(5)}  MMAs issued {rel(8)}")
