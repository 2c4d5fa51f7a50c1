"""Per-launch time of the sampling stage alone (vaura_sample_logits) for different modes; CUDA events over 200 launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vaura_b200.sampler import sample_logits  # noqa: E402

torch.manual_seed(0)
for rows in (1, 64):
    logits = torch.randn(rows, 9, 1024, device="cuda") * 2.0
    for name, kw in (("argmax", dict(use_sampling=False)), ("no filter", dict(use_sampling=True, top_k=0)),
                     ("top-k 256", dict(use_sampling=True, top_k=256)), ("top-p 0.9", dict(use_sampling=True, top_p=0.9))):
        for _ in range(5):
            sample_logits(logits, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            sample_logits(logits, **kw)
        e1.record()
        torch.cuda.synchronize()
        print(f"rows {rows:3d} {name:10s}: {e0.elapsed_time(e1) / 200 * 1e3:7.1f} us per call (launch + python overhead included)")
