"""Pins the CPU oracle (oracle/) on fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import vaura_oracle as vo
from oracle.dac_oracle import DacDecodeOracle
from vaura_b200.synthetic import (FULL_CODEC, FULL_SAMPLER, TINY_CODEC, TINY_SAMPLER, make_avclip_features,
                                  make_codec_state_dict, make_sampler_state_dict)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.fixture(scope="module")
def tiny_oracle():
    return vo.SamplerOracle(make_sampler_state_dict(TINY_SAMPLER, 0), TINY_SAMPLER)


def snr_db(ref: torch.Tensor, x: torch.Tensor) -> float:
    ref, x = ref.double().flatten(), x.double().flatten()
    return float(10 * torch.log10(ref.pow(2).sum() / (ref - x).pow(2).sum().clamp_min(1e-30)))


def _check_greedy(g, oracle, cdims, logit_tol):
    B, T = int(g["B"]), int(g["T"])
    feats = make_avclip_features(B, int(g["feat_seed"])).reshape(B, 32, 768)
    prompt = torch.from_numpy(g["prompt"].astype(np.int64))
    codes, logits = vo.generate_tokens(oracle, feats, prompt=prompt if prompt.shape[-1] else None, max_new_tokens=T,
                                       cfg_scale=float(g["cfg_scale"]), collect_logits=True)
    # bit-exact tokens: the fixture was selected to have no top-2 gap below `min_gap`
    assert torch.equal(codes, torch.from_numpy(g["codes"].astype(np.int64)))
    start = prompt.shape[-1] + 1
    # logits[i] produced column start+i; golden stat arrays are indexed by (column - 1)
    keep = g["keep_steps"]
    for n, s in enumerate(keep):
        col = int(s) + 1
        if col < start:
            continue
        mine = logits[col - start]
        ref = torch.from_numpy(g["logits_keep"][n])
        scale = ref.abs().max()
        assert (mine - ref).abs().max() / scale < logit_tol, (s, (mine - ref).abs().max())
    lse = torch.logsumexp(logits.double(), -1).float()
    ref_lse = torch.from_numpy(g["stat_lse"])[start - 1:]
    assert torch.allclose(lse, ref_lse, atol=2e-5)
    # codec: golden waveform (reference decode path with transformers.DacModel arithmetic, fp32)
    wav = DacDecodeOracle(make_codec_state_dict(cdims, 100), cdims).decode(codes)
    ref_wav = torch.from_numpy(g["wav_fp16"]).float()
    assert wav.shape == ref_wav.shape
    assert snr_db(ref_wav, wav) > 60.0  # fixture is stored in fp16: ~66 dB quantisation floor


def test_tiny_greedy_matches_reference(tiny_oracle):
    _check_greedy(load("tiny_greedy"), tiny_oracle, TINY_CODEC, 1e-5)


def test_tiny_cfg_with_prompt_matches_reference(tiny_oracle):
    _check_greedy(load("tiny_cfg_prompt"), tiny_oracle, TINY_CODEC, 1e-5)


def test_teacher_forced_full_and_cached_match_reference(tiny_oracle):
    g = load("tiny_teacher_forced")
    seq = torch.from_numpy(g["seq"].astype(np.int64))
    feats = make_avclip_features(seq.shape[0], int(g["feat_seed"])).reshape(-1, 32, 768)
    logits = tiny_oracle.forward_full(seq, feats)  # (B,K,S,V)
    ref = torch.from_numpy(g["logits_keep"])
    keep = g["keep"].tolist()
    assert (logits[:, :, keep] - ref).abs().max() / ref.abs().max() < 1e-5
    assert torch.allclose(torch.logsumexp(logits.double(), -1).float(), torch.from_numpy(g["lse"]), atol=2e-5)
    # the KV-cached step restatement equals the full-prefix forward at every position (incl. >= 224)
    rows = tiny_oracle.cond_rows(feats)
    cache = tiny_oracle.new_cache(seq.shape[0])
    first = tiny_oracle.forward_cached(seq[..., :5], rows, cache)  # prefill of 5
    assert (first - logits[:, :, 4]).abs().max() < 2e-5
    for p in range(5, seq.shape[-1]):
        step = tiny_oracle.forward_cached(seq[..., p:p + 1], rows, cache)
        if p in keep or p % 37 == 0:
            assert (step - logits[:, :, p]).abs().max() < 2e-5, p


def test_sampling_filters_match_reference():
    g = load("sampling_filters")
    logits = torch.from_numpy(g["logits"])
    for key in g.files:
        if key.startswith("topk_"):
            temp, k = float(key.split("_t")[1].split("_k")[0]), int(key.split("_k")[1])
            mine = vo.filtered_probs(logits, temp, k, 0.0)
            assert torch.allclose(mine, torch.from_numpy(g[key]), atol=1e-7), key
        elif key.startswith("topp_sorted_"):
            temp, p = float(key.split("_t")[1].split("_p")[0]), float(key.split("_p")[-1])
            mine = vo.filtered_probs(logits, temp, 0, p)
            mine_sorted = torch.sort(mine, dim=-1, descending=True)[0]
            assert torch.allclose(mine_sorted, torch.from_numpy(g[key]), atol=1e-7), key
    # ties at the k-th value are all kept (utils/utils.py:172-175)
    tie = vo.filtered_probs(logits[0, 0][None], 1.0, 20, 0.0)[0]
    assert (tie[:41] > 0).sum() in (0, 41)


def test_philox_known_answers_and_draw():
    # Random123 known-answer vectors for Philox4x32-10
    assert vo.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert vo.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert vo.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    p = np.array([0.0, 0.25, 0.0, 0.5, 0.25, 0.0], dtype=np.float32)
    assert [vo.inverse_cdf_draw(p, u) for u in (0.0, 0.2499, 0.25, 0.74, 0.75, 0.999999)] == [1, 1, 3, 3, 4, 4]
    us = [vo.philox_uniform(1234, 3, s, 2) for s in range(2000)]
    assert 0.0 <= min(us) and max(us) < 1.0 and abs(np.mean(us) - 0.5) < 0.03


def test_full_size_greedy_matches_reference():
    """Config 1 of BASELINE.json: 24 layers, B=1, 2.56 s, greedy — the reference's own tokens."""
    g = load("full_greedy")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle = vo.SamplerOracle(make_sampler_state_dict(FULL_SAMPLER, 0), FULL_SAMPLER)
    _check_greedy(g, oracle, FULL_CODEC, 2e-5)


def test_motionformer_oracle_matches_reference_golden():
    """the restated Segment-AVCLIP tower against features the reference's own MotionFormer produced
    (oracle/make_golden_motionformer.py) on the same seeded weights and video segments."""
    from oracle import motionformer_oracle as mo
    from vaura_b200.synthetic import FULL_AVCLIP, make_motionformer_state_dict, make_video_segments

    g = np.load(os.path.join(GOLD, "motionformer_full.npz"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    frames = make_video_segments(int(g["batch"]), int(g["frame_seed"]), int(g["segments"]))
    mine = mo.motionformer_features(frames, make_motionformer_state_dict(int(g["weight_seed"])), FULL_AVCLIP)
    ref = torch.from_numpy(g["features"])
    assert mine.shape == ref.shape == (1, 2, 8, 768)
    err = float((mine - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


def test_dac_encode_oracle_matches_transformers():
    """the restated dac encoder + residual VQ against transformers.DacModel.encode (the independent statement of
    the same architecture; dac 1.0.0 itself is absent -> parity with it stays unpinned)."""
    from oracle import ref_stubs
    from oracle.dac_oracle import DacEncodeOracle
    from vaura_b200.synthetic import TINY_CODEC, make_codec_state_dict

    sd = make_codec_state_dict(TINY_CODEC, 100, with_encoder=True)
    hf = ref_stubs._hf_dac(TINY_CODEC, with_encoder=True)
    ref_stubs.load_dac_names_into_hf(hf, sd)
    ref_stubs.load_dac_encoder_names_into_hf(hf, sd)
    g = torch.Generator().manual_seed(3)
    wav = 0.3 * torch.randn(2, 1, 5000, generator=g)
    o = DacEncodeOracle(sd, TINY_CODEC)
    padded = o.preprocess(wav)
    assert padded.shape[-1] == 5120 and torch.equal(padded[..., :5000], wav) and float(padded[..., 5000:].abs().max()) == 0.0
    with torch.no_grad():
        out = hf.encode(padded)
        z_hf = hf.encoder(padded)
    assert torch.allclose(o.encode_latent(padded), z_hf, atol=1e-5)
    assert torch.equal(o.encode(wav), out.audio_codes)
