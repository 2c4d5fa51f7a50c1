"""Segment-AVCLIP visual tower on the GPU against the reference's own MotionFormer output (golden made by
oracle/make_golden_motionformer.py) and against the CPU oracle on other inputs.  The GPU path multiplies in bf16 with fp32
accumulation, fp32 residual stream / LayerNorm / softmax: tolerance 1e-2 of max |feature| (stated here), measured ~3e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import motionformer_oracle as mo
from vaura_b200.synthetic import FULL_AVCLIP, make_motionformer_state_dict, make_video_segments

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-2


@pytest.fixture(scope="module")
def tower():
    from vaura_b200.features import MotionFormer

    m = MotionFormer(extract_features=True, factorize_space_time=True, agg_space_module="TransformerEncoderLayer",
                     agg_time_module="torch.nn.Identity", add_global_repr=False)
    m.load_state_dict(make_motionformer_state_dict(7), device="cuda:0")
    return m


def rel_err(a, ref):
    return float((a.float().cpu() - ref).abs().max() / ref.abs().max())


def test_features_match_reference_golden(tower):
    g = np.load(os.path.join(GOLD, "motionformer_full.npz"))
    frames = make_video_segments(int(g["batch"]), int(g["frame_seed"]), int(g["segments"]))
    feats, glob = tower(frames.cuda())
    assert glob is None and feats.shape == (1, 2, 8, 768) and feats.dtype == torch.float32
    ref = torch.from_numpy(g["features"])
    err = rel_err(feats, ref)
    cos = float(torch.nn.functional.cosine_similarity(feats.cpu().flatten(), ref.flatten(), dim=0))
    print(f"[avclip] vs reference golden: max err / max|feature| {err:.3e}, cosine {cos:.6f}")
    assert err < TOL, err
    assert cos > 0.9999, cos


def test_chunked_segments_match_oracle_and_are_independent(tower):
    """5 segments through a workspace that holds 2 at a time: ragged last chunk; segment results do not depend on their
    neighbours (motionformer.py:269-283: for_loop and batched evaluation agree in the reference too)."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    frames = make_video_segments(1, 23, 5)
    ref = mo.motionformer_features(frames, make_motionformer_state_dict(7), FULL_AVCLIP)
    tower.max_chunk_segments, tower._ws = 2, None
    feats, _ = tower(frames.cuda())
    tower.max_chunk_segments, tower._ws = 32, None
    err = rel_err(feats, ref)
    print(f"[avclip] 5 segments in chunks of 2 vs oracle: {err:.3e}")
    assert err < TOL, err
    alone, _ = tower(frames[:, 3:4].cuda())
    assert torch.equal(alone[0, 0], feats[0, 3]), "a segment's features must not depend on the batch it is in"


def test_passthrough_and_errors(tower):
    feats = torch.randn(2, 4, 8, 768)
    out, glob = tower(feats)
    assert out is feats and glob is None
    with pytest.raises(ValueError):
        tower(torch.zeros(1, 1, 3, 16, 112, 112).cuda())
    with pytest.raises(NotImplementedError):
        tower(torch.zeros(1, 1, 3, 16, 224, 224).cuda(), cont_mask=torch.ones(1))
    from vaura_b200.features import MotionFormer

    with pytest.raises(RuntimeError):
        MotionFormer(extract_features=True)(torch.zeros(1, 1, 3, 16, 224, 224).cuda())
    with pytest.raises(NotImplementedError):
        MotionFormer(extract_features=True, add_global_repr=True).load_state_dict({}, device="cuda:0")


def test_generate_from_raw_frames_matches_generate_from_features(tower):
    """BASELINE config 5 in miniature: frames -> tower -> AR decode, through VAURAModel.generate (vaura_model.py:194-214)."""
    from vaura_b200.synthetic import TINY_CODEC, TINY_SAMPLER, build_model

    model = build_model(TINY_SAMPLER, TINY_CODEC)
    model.visual_feature_extractor.load_state_dict(make_motionformer_state_dict(7), device="cuda:0")
    frames = make_video_segments(2, 31, 4).cuda()
    feats, _ = model.visual_feature_extractor(frames)
    kw = dict(max_new_tokens=12, use_sampling=False, prompt_is_encoded=True, return_sampled_indices=True)
    a = model.generate(frames=frames, **kw)
    b = model.generate(frames=feats, **kw)
    assert torch.equal(a["sampled_indices"], b["sampled_indices"])
    assert torch.equal(a["generated_audio"], b["generated_audio"])


def test_pinned_host_frames_are_copied_chunk_by_chunk_with_the_same_result(tower):
    """Frames handed over in pinned host memory take the overlapped copy path (two staging buffers, ragged last chunk)."""
    frames = make_video_segments(1, 57, 5)
    ref, _ = tower(frames.cuda())
    tower.max_chunk_segments, tower._ws = 2, None
    try:
        got, _ = tower(frames.pin_memory())
        again, _ = tower(frames.pin_memory())  # staging buffers re-used by a second call
    finally:
        tower.max_chunk_segments, tower._ws = 32, None
    assert got.device.type == "cuda" and torch.equal(got, ref) and torch.equal(again, ref)
