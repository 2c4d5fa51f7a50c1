"""TEST INFRASTRUCTURE ONLY — writes tests/golden/motionformer_full.npz by executing the UNMODIFIED reference
MotionFormer (models/modules/feature_extractors/avclip/motionformer.py) under the timm / omegaconf import stubs of
oracle/ref_stubs.py, in the shipped configuration, with the seeded synthetic weights and video segments of
vaura_b200/synthetic.py (so the fixture only stores the output features).

    python oracle/make_golden_motionformer.py

Needs /root/reference (build container only).  Also prints the distance of oracle/motionformer_oracle.py from it.
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import motionformer_oracle as mo  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from vaura_b200.synthetic import FULL_AVCLIP, make_motionformer_state_dict, make_video_segments  # noqa: E402

WEIGHT_SEED, FRAME_SEED, BATCH, SEGMENTS = 7, 11, 1, 2


def main():
    torch.manual_seed(0)
    ref = ref_stubs.build_reference_motionformer(WEIGHT_SEED)
    frames = make_video_segments(BATCH, FRAME_SEED, SEGMENTS)
    t0 = time.time()
    with torch.no_grad():
        feats, glob = ref(frames)
        feats_loop, _ = ref(frames, for_loop=True)
    print(f"reference forward: {time.time() - t0:.1f} s, features {tuple(feats.shape)}, global {glob}")
    assert glob is None and feats.shape == (BATCH, SEGMENTS, FULL_AVCLIP.temporal, FULL_AVCLIP.embed_dim)
    print("for_loop=True vs batched segments: max abs diff", float((feats - feats_loop).abs().max()))
    mine = mo.motionformer_features(frames, make_motionformer_state_dict(WEIGHT_SEED), FULL_AVCLIP)
    err = float((mine - feats).abs().max() / feats.abs().max())
    print(f"oracle vs reference: max abs err / max |feature| = {err:.3e}; feature std {float(feats.std()):.3f}, "
          f"max |feature| {float(feats.abs().max()):.3f}")
    assert err < 1e-5, err
    out = os.path.join(ROOT, "tests", "golden", "motionformer_full.npz")
    np.savez_compressed(out, features=feats.numpy().astype(np.float32), weight_seed=WEIGHT_SEED, frame_seed=FRAME_SEED,
                        batch=BATCH, segments=SEGMENTS)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
