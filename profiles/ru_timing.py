"""Per-tile timeline of gemm_ru_fused_kernel (CTA 0, tiles 8..15 of the last residual-unit launch of a codec decode = the
96-channel unit with dilation 9).  Needs a build with -DVAURA_RU_TIMING (stamps are compiled out by default):
    VAURA_B200_LIB=$REPO/vaura_b200/_lib/libvaura_b200_rutiming.so python profiles/ru_timing.py [clips]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vaura_b200 import _cabi  # noqa: E402
from vaura_b200.codec import DacModelWrapper  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = DacModelWrapper(44100, dims=FULL_CODEC)
m.load_state_dict(make_codec_state_dict(FULL_CODEC, 100), device="cuda:0")
codes = torch.randint(0, 1024, (B, 9, 220)).cuda()
for _ in range(2):
    m.decode(codes)
torch.cuda.synchronize()
lib = _cabi.load()
buf = (C.c_ulonglong * (8 * 3 * 8))()
assert lib.vaura_debug_ru_timing(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(8, 3, 8)
t0 = t[0, 0, 0]
rel = lambda v: (v - t0) / 1e3
print("us after the MMA warp started tile 8 (CTA 0); one line per tile")
print("tile | MMA warp: start, first stage landed, k7 issued (acc1 commit), acc2 free, h block 0 ready, k1 issued (acc2 commit) |"
      " epilogue warp 2: start, acc1 full, h stored, acc2 full, outputs stored | producer: start, k7 loads issued, W1 loads issued")
for i in range(8):
    mm, ep, pr = t[i, 0], t[i, 1], t[i, 2]
    print(f"{8 + i:4d} | " + " ".join(f"{rel(v):7.2f}" for v in mm[:6]) + " | " + " ".join(f"{rel(v):7.2f}" for v in ep[:5]) +
          " | " + " ".join(f"{rel(v):7.2f}" for v in pr[:3]))
d = np.diff(t[:, 0, 0]) / 1e3
print("tile period (us):", " ".join(f"{x:.2f}" for x in d))
