"""When every CTA of decode_step_fused_bf16 arrives at the five device-wide barriers of the last layer (last step of a short
generate): python profiles/fused_arrivals.py [B] [T].  Prints, per phase, the spread of the arrival times and the CTAs a phase
waits for.  Needs a build of the library with -DVAURA_FUSED_ARRIVALS (the stamps are compiled out by default), e.g.
    cd /tmp/x && for f in $REPO/vaura_b200/csrc/*.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 \
        -Xcompiler -fPIC -DVAURA_FUSED_ARRIVALS -c $f; done && nvcc -gencode arch=compute_100a,code=sm_100a -shared \
        -o $REPO/vaura_b200/_lib/libvaura_b200_arrivals.so *.o -lcudart
    VAURA_B200_LIB=$REPO/vaura_b200/_lib/libvaura_b200_arrivals.so python profiles/fused_arrivals.py"""
import os
import sys

os.environ["VAURA_PERSIST_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(B, 2).cuda()
G = torch.cuda.get_device_properties(0).multi_processor_count
late = np.zeros((5, G))
for rep in range(3):
    m.generate(frames=feats, max_new_tokens=T + rep, use_sampling=True, top_k=128, prompt_is_encoded=True, _decode_audio=False)
    torch.cuda.synchronize()
    ws = m.sampler._buffers["ws"]
    t = ws[256:256 + 16384].cpu().numpy().view(np.uint64).astype(np.int64)
    arr = t[1300:1300 + 5 * G].reshape(5, G)
    names = ["qkv", "attn", "wo+rms", "w13", "w2+rms"]
    print(f"--- position {T + rep + 7}, last layer: arrival of the CTAs at the barrier that ends each phase (us after the first arrival)")
    for i, n in enumerate(names):
        a = (arr[i] - arr[i].min()) / 1e3
        order = np.argsort(-a)
        late[i] += a
        q = np.percentile(a, [10, 50, 90, 99])
        print(f"  {n:7s} p10 {q[0]:5.2f} p50 {q[1]:5.2f} p90 {q[2]:5.2f} p99 {q[3]:5.2f} max {a.max():5.2f}   last CTAs: " +
              " ".join(f"{c}({a[c]:.2f})" for c in order[:8]))
print("mean lateness per CTA over the three runs, ten latest CTAs per phase:")
for i, n in enumerate(["qkv", "attn", "wo+rms", "w13", "w2+rms"]):
    order = np.argsort(-late[i])
    print(f"  {n:7s} " + " ".join(f"{c}({late[i][c] / 3:.2f})" for c in order[:10]))
