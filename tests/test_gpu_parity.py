"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every compute call goes through the C ABI
(libvaura_b200.so) via the host mirror; the CPU oracle is only the checker."""
import os

import numpy as np
import pytest
import torch

from oracle import vaura_oracle as vo
from oracle.dac_oracle import DacDecodeOracle
from vaura_b200 import VAURAModel, _cabi
from vaura_b200.synthetic import (FULL_CODEC, FULL_SAMPLER, TINY_CODEC, TINY_SAMPLER, build_model, make_avclip_features,
                                  make_codec_state_dict, make_sampler_state_dict)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def tiny_model():
    return build_model(TINY_SAMPLER, TINY_CODEC)


@pytest.fixture(scope="module")
def tiny_oracle():
    return vo.SamplerOracle(make_sampler_state_dict(TINY_SAMPLER, 0), TINY_SAMPLER)


def snr_db(ref, x):
    ref, x = ref.double().flatten().cpu(), x.double().flatten().cpu()
    return float(10 * torch.log10(ref.pow(2).sum() / (ref - x).pow(2).sum().clamp_min(1e-30)))


def rel_err(a, ref):
    return float((a.cpu().double() - ref.double()).abs().max() / ref.double().abs().max())


# fp32-activation path: logits tolerance of the north star for fp32 (1e-5, normalised by max |logit|,
# SURVEY §7 "hard parts"); measured error is summation-order noise.
FP32_LOGIT_TOL = 1e-5


def test_forward_teacher_forced_matches_oracle(tiny_model, tiny_oracle):
    g = torch.Generator().manual_seed(3)
    seq = torch.randint(0, 1025, (2, 9, 229), generator=g)
    feats = make_avclip_features(2, 5).reshape(2, 32, 768)
    logits, a, b = tiny_model.sampler(tgt=seq.cuda(), memory=feats.cuda())
    assert a is None and b is None and logits.shape == (2, 9, 229, 1024)
    ref = tiny_oracle.forward_full(seq, feats)
    assert rel_err(logits, ref) < FP32_LOGIT_TOL
    gold = np.load(os.path.join(GOLD, "tiny_teacher_forced.npz"))
    seq_g = torch.from_numpy(gold["seq"].astype(np.int64))
    lg, _, _ = tiny_model.sampler(tgt=seq_g.cuda(), memory=feats.cuda())
    assert rel_err(lg[:, :, gold["keep"].tolist()], torch.from_numpy(gold["logits_keep"])) < FP32_LOGIT_TOL


@pytest.mark.parametrize("name", ["tiny_greedy", "tiny_cfg_prompt"])
def test_generate_greedy_matches_reference_golden(tiny_model, tiny_oracle, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    B, T = int(g["B"]), int(g["T"])
    feats = make_avclip_features(B, int(g["feat_seed"]))
    prompt = torch.from_numpy(g["prompt"].astype(np.int64))
    out = tiny_model.generate(frames=feats.cuda(), audio=prompt.cuda() if prompt.shape[-1] else None, max_new_tokens=T,
                              use_sampling=False, prompt_is_encoded=True, return_sampled_indices=True,
                              cfg_scale=float(g["cfg_scale"]), check=True, _return_logits=True)
    codes = out["sampled_indices"].cpu()
    assert torch.equal(codes, torch.from_numpy(g["codes"].astype(np.int64)))  # bit-exact tokens
    # per-step post-CFG logits vs the reference's (kept steps) and vs the oracle (all steps)
    start = prompt.shape[-1] + 1
    for n, s in enumerate(g["keep_steps"]):
        if int(s) + 1 >= start:
            assert rel_err(out["_logits"][int(s) + 1], torch.from_numpy(g["logits_keep"][n])) < FP32_LOGIT_TOL
    _, ologits = vo.generate_tokens(tiny_oracle, feats.reshape(B, 32, 768), prompt=prompt if prompt.shape[-1] else None,
                                    max_new_tokens=T, cfg_scale=float(g["cfg_scale"]), collect_logits=True)
    assert rel_err(out["_logits"][start:], ologits) < FP32_LOGIT_TOL
    wav = out["generated_audio"]
    assert wav.dtype == torch.float16 and wav.shape == (B, 1, T * 512)
    assert snr_db(torch.from_numpy(g["wav_fp16"]).float(), wav.float()) > 40.0  # fp16 storage, fp32 accumulate


def test_generate_api_contract(tiny_model):
    feats = make_avclip_features(3, 9).cuda()
    out = tiny_model.generate(frames=feats, max_new_tokens=12, use_sampling=True, top_k=128, prompt_is_encoded=True)
    assert set(out) == {"generated_audio", "s_attn_weights", "mha_attn_weights", "sampled_indices"}
    assert out["sampled_indices"] is None and out["s_attn_weights"] is None and out["mha_attn_weights"] is None
    assert out["generated_audio"].shape == (3, 1, 12 * 512)
    out = tiny_model.generate(frames=feats, max_new_tokens=12, return_sampled_indices=True, prompt_is_encoded=True,
                              audio=torch.randint(0, 1024, (3, 9, 5)).cuda(), remove_prompts=True)
    assert out["sampled_indices"].shape == (3, 9, 7) and out["sampled_indices"].dtype == torch.long
    assert int(out["sampled_indices"].min()) >= 0 and int(out["sampled_indices"].max()) < 1024
    with pytest.raises(AssertionError):  # vaura_model.py:476-478
        tiny_model.generate(frames=feats, audio=torch.zeros(3, 9, 12, dtype=torch.long).cuda(), max_new_tokens=12,
                            prompt_is_encoded=True)
    with pytest.raises(RuntimeError):  # RoPE table has 256 rows (llama.py:364-368): 250+9 columns do not fit
        tiny_model.generate(frames=feats, max_new_tokens=250, prompt_is_encoded=True)
    # sampling is reproducible per (seed, clip id, column, codebook) and independent of batch composition as long as the
    # precision mode is the same (AUTO switches a sampling call to the bf16 path from 3 sequence rows, so it is pinned here)
    a = tiny_model.generate(frames=feats, max_new_tokens=12, top_k=64, return_sampled_indices=True,
                            clip_indices=torch.tensor([7, 8, 9]), _decode_audio=False,
                            _precision=_cabi.PRECISION_FP32ACT)["sampled_indices"]
    b = tiny_model.generate(frames=feats[1:2], max_new_tokens=12, top_k=64, return_sampled_indices=True,
                            clip_indices=torch.tensor([8]), _decode_audio=False,
                            _precision=_cabi.PRECISION_FP32ACT)["sampled_indices"]
    assert torch.equal(a[1:2], b)
    # AUTO: a sampling call with 3 rows takes the fused bf16 step kernel (include/vaura_b200.h: VAURA_PRECISION_AUTO)
    c = tiny_model.generate(frames=feats, max_new_tokens=12, top_k=64, return_sampled_indices=True,
                            clip_indices=torch.tensor([7, 8, 9]), _decode_audio=False)["sampled_indices"]
    assert c.shape == a.shape and int(c.min()) >= 0 and int(c.max()) < 1024


def test_codec_decode_matches_oracle(tiny_model):
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(0, 1024, (2, 9, 33), generator=g)  # ragged w.r.t. the 64-row tiles
    wav = tiny_model.audio_encoder.decode([(codes.cuda(), None)])
    ref = DacDecodeOracle(make_codec_state_dict(TINY_CODEC, 100), TINY_CODEC).decode(codes)
    assert wav.shape == ref.shape == (2, 1, 33 * 512)
    assert snr_db(ref, wav.float()) > 40.0
    with pytest.raises(IndexError):
        tiny_model.audio_encoder.decode(torch.full((1, 9, 4), 1024).cuda())  # special id is not a codec code


def test_codec_decode_full_size_both_snr_bounds():
    """two bounds for the full-size codec - against the fp32 oracle, and against the oracle with the weights the
    reference's `.half()` leaves (vaura_model.py:92: weight_g / weight_v rounded to fp16 before the weight-norm fold;
    activations kept in fp32 here, CPU fp16 convolutions being neither fast nor what a GPU accumulates).  The GPU path folds in
    fp32 and rounds the folded weight once, so the two differ by two independent roundings: the second bound is its own number,
    not a tighter one."""
    from vaura_b200.codec import DacModelWrapper
    from vaura_b200.synthetic import FULL_CODEC

    sd = make_codec_state_dict(FULL_CODEC, 100)
    m = DacModelWrapper(44100, dims=FULL_CODEC)
    m.load_state_dict(sd, device="cuda:0")
    codes = torch.randint(0, 1024, (1, 9, 40), generator=torch.Generator().manual_seed(2))
    wav = m.decode(codes.cuda()).float().cpu()
    ref32 = DacDecodeOracle(sd, FULL_CODEC).decode(codes)
    sd16 = {k: (v.half().float() if v.is_floating_point() else v) for k, v in sd.items()}
    ref16 = DacDecodeOracle(sd16, FULL_CODEC).decode(codes)
    s32, s16 = snr_db(ref32, wav), snr_db(ref16, wav)
    print(f"[codec full size] SNR vs fp32 oracle {s32:.1f} dB, vs fp16-weight oracle {s16:.1f} dB")
    assert s32 > 45.0 and s16 > 40.0


def test_full_size_greedy_matches_reference_golden():
    """BASELINE.json config 1 on the B200: 24 layers, B=1, 2.56 s clip, greedy; the tokens are the ones the
    unmodified reference produced (tests/golden/full_greedy.npz)."""
    g = np.load(os.path.join(GOLD, "full_greedy.npz"))
    m = build_model(FULL_SAMPLER, FULL_CODEC)
    feats = make_avclip_features(1, int(g["feat_seed"]))
    out = m.generate(frames=feats.cuda(), max_new_tokens=220, use_sampling=False, prompt_is_encoded=True,
                     return_sampled_indices=True, check=True, _return_logits=True)
    codes = out["sampled_indices"].cpu()
    ref = torch.from_numpy(g["codes"].astype(np.int64))
    # north star: bit-exact for the first 64 steps, >= 99 % agreement over the clip
    seq_m, _ = vo.build_pattern_sequence(codes, 1024)
    seq_r, _ = vo.build_pattern_sequence(ref, 1024)
    assert torch.equal(seq_m[..., :65], seq_r[..., :65])
    assert (codes == ref).float().mean().item() >= 0.99
    for n, s in enumerate(g["keep_steps"]):
        assert rel_err(out["_logits"][int(s) + 1], torch.from_numpy(g["logits_keep"][n])) < 2e-5
    wav = out["generated_audio"]
    assert wav.shape == (1, 1, 112640)
    if torch.equal(codes, ref):
        assert snr_db(torch.from_numpy(g["wav_fp16"]).float(), wav.float()) > 35.0
    # 4 clips + CFG at full size exercises the 8-row weight pass
    f4 = make_avclip_features(4, 77)
    o4 = m.generate(frames=f4.cuda(), max_new_tokens=16, use_sampling=False, prompt_is_encoded=True, cfg_scale=3.0,
                    return_sampled_indices=True, _decode_audio=False)["sampled_indices"].cpu()
    oracle = vo.SamplerOracle(make_sampler_state_dict(FULL_SAMPLER, 0), FULL_SAMPLER)
    r4, lg = vo.generate_tokens(oracle, f4.reshape(4, 32, 768), max_new_tokens=16, cfg_scale=3.0, collect_logits=True)
    gaps = torch.topk(lg, 2, dim=-1).values
    if float((gaps[..., 0] - gaps[..., 1]).min()) > 1e-3:
        assert torch.equal(o4, r4)
    else:
        assert (o4 == r4).float().mean().item() > 0.9
    # two sequence rows take the cluster decode kernel's NB = 2 instance: 1 clip + CFG, and 2 clips without
    for nclip, cfg in ((1, 3.0), (2, 1.0)):
        f = make_avclip_features(nclip, 91 + nclip)
        o = m.generate(frames=f.cuda(), max_new_tokens=24, use_sampling=False, prompt_is_encoded=True, cfg_scale=cfg,
                       return_sampled_indices=True, _decode_audio=False, _return_logits=True)
        r, lg = vo.generate_tokens(oracle, f.reshape(nclip, 32, 768), max_new_tokens=24, cfg_scale=cfg, collect_logits=True)
        assert rel_err(o["_logits"][1:].cpu(), lg) < 2e-5
        gaps = torch.topk(lg, 2, dim=-1).values
        if float((gaps[..., 0] - gaps[..., 1]).min()) > 1e-3:
            assert torch.equal(o["sampled_indices"].cpu(), r)
    # a prompt makes the first pass a multi-position prefill (multi-kernel path); the cluster kernel then continues on the
    # KV cache those kernels wrote, with classifier-free guidance on top (the reference's generate settings)
    g2 = torch.Generator().manual_seed(5)
    prompt = torch.randint(0, 1024, (1, 9, 10), generator=g2)
    f = make_avclip_features(1, 95)
    o = m.generate(frames=f.cuda(), audio=prompt.cuda(), max_new_tokens=26, use_sampling=False, prompt_is_encoded=True,
                   cfg_scale=6.0, return_sampled_indices=True, _decode_audio=False, _return_logits=True)
    r, lg = vo.generate_tokens(oracle, f.reshape(1, 32, 768), prompt=prompt, max_new_tokens=26, cfg_scale=6.0,
                               collect_logits=True)
    start = prompt.shape[-1] + 1
    # the guidance combine u + (c - u) * 6 amplifies the fp32 summation-order noise of both halves by up to 11; the prompt
    # prefill runs on the tensor cores (three bf16 terms per fp32 operand, fp32 accumulate in TMEM)
    assert rel_err(o["_logits"][start:].cpu(), lg) < 5e-5
    gaps = torch.topk(lg, 2, dim=-1).values
    if float((gaps[..., 0] - gaps[..., 1]).min()) > 1e-3:
        assert torch.equal(o["sampled_indices"].cpu(), r)


# ---- tensor-core (bf16 activation) path ------------------------------------------------------------------
# north star tolerance for bf16: per-step logits within 1e-2 relative error; here relative to max |logit| of the
# step (element-wise relative error is ill-posed near zero crossings).
BF16_LOGIT_TOL = 1e-2


def test_bf16_path_forward_logits(tiny_model, tiny_oracle):
    g = torch.Generator().manual_seed(4)
    seq = torch.randint(0, 1025, (2, 9, 229), generator=g)
    feats = make_avclip_features(2, 6).reshape(2, 32, 768)
    logits, _, _ = tiny_model.sampler(tgt=seq.cuda(), memory=feats.cuda(), precision=_cabi.PRECISION_BF16)
    ref = tiny_oracle.forward_full(seq, feats)
    assert rel_err(logits, ref) < BF16_LOGIT_TOL
    agree = (logits.cpu().argmax(-1) == ref.argmax(-1)).float().mean().item()
    assert agree >= 0.97, agree  # tiny random-init model: top-2 gaps are small, most flips are near-ties


def test_bf16_path_generate_batch16(tiny_model, tiny_oracle):
    B, T = 16, 24
    feats = make_avclip_features(B, 21)
    out = tiny_model.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                              return_sampled_indices=True, check=True, _return_logits=True, _decode_audio=False)
    codes = out["sampled_indices"].cpu()  # AUTO -> bf16 path at 16 rows
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    ref = tiny_oracle.forward_full(seq[..., :-1], feats.reshape(B, 32, 768))  # teacher-forced on OUR tokens
    mine = out["_logits"][1:].cpu().permute(1, 2, 0, 3)  # (S-1,B,K,V) -> (B,K,S-1,V)
    assert rel_err(mine, ref) < BF16_LOGIT_TOL
    gap = torch.topk(ref, 2, dim=-1).values
    clear = (gap[..., 0] - gap[..., 1]) > 0.05 * ref.abs().max()
    agree = (mine.argmax(-1) == ref.argmax(-1))
    assert agree[clear].float().mean().item() == 1.0
    assert agree.float().mean().item() >= 0.97
    # same call forced onto the fp32-activation path reproduces the oracle's free-running tokens
    o32 = tiny_model.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                              return_sampled_indices=True, _decode_audio=False, _precision=_cabi.PRECISION_FP32ACT)
    r32, lg = vo.generate_tokens(tiny_oracle, feats.reshape(B, 32, 768), max_new_tokens=T, collect_logits=True)
    g2 = torch.topk(lg, 2, dim=-1).values
    if float((g2[..., 0] - g2[..., 1]).min()) > 1e-4:
        assert torch.equal(o32["sampled_indices"].cpu(), r32)


@pytest.mark.parametrize("deterministic", [False, True])
def test_bf16_path_full_size_fused_step(deterministic, monkeypatch):
    """Full-size model, 24 sequence rows: the fused cooperative decode-step kernel (decode_step_fused_bf16) with every GEMM
    phase at its real tile count (144 / 128 tiles on 148 SMs), teacher-forced against the fp32 oracle on our own tokens.
    VAURA_DETERMINISTIC=1: split-K partial slices added in a fixed order -> a repeated call returns the same bits."""
    monkeypatch.setenv("VAURA_DETERMINISTIC", "1" if deterministic else "0")
    B, T = 24, 6
    m = build_model(FULL_SAMPLER, FULL_CODEC)
    oracle = vo.SamplerOracle(make_sampler_state_dict(FULL_SAMPLER, 0), FULL_SAMPLER)
    feats = make_avclip_features(B, 33)
    out = m.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                     return_sampled_indices=True, check=True, _return_logits=True, _decode_audio=False)
    codes = out["sampled_indices"].cpu()
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    ref = oracle.forward_full(seq[..., :-1], feats.reshape(B, 32, 768))
    mine = out["_logits"][1:].cpu().permute(1, 2, 0, 3)
    assert rel_err(mine, ref) < BF16_LOGIT_TOL
    if not deterministic:
        return
    # run-to-run reproducibility: the split-K sums of wo / w2 are added by one owner per row in a fixed order (no float
    # atomics in decode_step_fused_bf16), so a repeated call returns the same bits
    again = m.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                       return_sampled_indices=True, check=True, _return_logits=True, _decode_audio=False)
    assert torch.equal(again["_logits"][1:], out["_logits"][1:])  # without a prompt position 0 is a launch of the same kernel
    assert torch.equal(again["sampled_indices"], out["sampled_indices"])


def test_bf16_path_with_cfg_and_prompt_prefill(tiny_model, tiny_oracle):
    B, T, Tp = 8, 30, 11  # 16 rows with CFG; prompt -> prefill of 12 columns on the tensor-core path
    feats = make_avclip_features(B, 33)
    g = torch.Generator().manual_seed(8)
    prompt = torch.randint(0, 1024, (B, 9, Tp), generator=g)
    out = tiny_model.generate(frames=feats.cuda(), audio=prompt.cuda(), max_new_tokens=T, use_sampling=False,
                              prompt_is_encoded=True, return_sampled_indices=True, cfg_scale=2.5, check=True,
                              _return_logits=True, _decode_audio=False)
    codes = out["sampled_indices"].cpu()
    assert torch.equal(codes[..., :Tp], prompt)  # prompt preserved (vaura_model.py:540-544)
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    f = feats.reshape(B, 32, 768)
    fa = torch.cat([f, torch.zeros_like(f) + tiny_oracle.uncond], 0)
    lg = tiny_oracle.forward_full(seq[..., :-1].repeat(2, 1, 1), fa)
    ref = lg[B:] + (lg[:B] - lg[B:]) * 2.5
    mine = out["_logits"][Tp + 1:].cpu().permute(1, 2, 0, 3)
    assert rel_err(mine, ref[:, :, Tp:]) < 2 * BF16_LOGIT_TOL  # CFG amplifies the error by ~cfg_scale


@pytest.mark.parametrize("deterministic", [False, True])
def test_bf16_path_128_row_fused_step(tiny_model, tiny_oracle, deterministic, monkeypatch):
    """48 clips with CFG = 96 sequence rows: the fused decode-step kernel's 128-row instance (UMMA M = 128), in the default and
    in the reproducible mode (VAURA_DETERMINISTIC=1: a repeated call returns the same bits)."""
    monkeypatch.setenv("VAURA_DETERMINISTIC", "1" if deterministic else "0")
    B, T = 48, 12
    feats = make_avclip_features(B, 41)
    out = tiny_model.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                              return_sampled_indices=True, cfg_scale=2.0, check=True, _return_logits=True,
                              _decode_audio=False)
    codes = out["sampled_indices"].cpu()
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    f = feats.reshape(B, 32, 768)
    fa = torch.cat([f, torch.zeros_like(f) + tiny_oracle.uncond], 0)
    lg = tiny_oracle.forward_full(seq[..., :-1].repeat(2, 1, 1), fa)
    ref = lg[B:] + (lg[:B] - lg[B:]) * 2.0
    mine = out["_logits"][1:].cpu().permute(1, 2, 0, 3)
    assert rel_err(mine, ref) < 2 * BF16_LOGIT_TOL  # CFG amplifies the error by ~cfg_scale
    if deterministic:
        again = tiny_model.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                                    return_sampled_indices=True, cfg_scale=2.0, check=True, _return_logits=True,
                                    _decode_audio=False)
        assert torch.equal(again["_logits"][1:], out["_logits"][1:])
        assert torch.equal(again["sampled_indices"], out["sampled_indices"])


def test_bf16_fused_step_with_16_position_pages(tiny_model, tiny_oracle, monkeypatch):
    """K/V pages of 16 positions (the other page size the C ABI accepts): one staged attention run per page, a long
    enough clip for runs in several pages; teacher-forced against the fp32 oracle on our own tokens."""
    import vaura_b200.sampler as vs

    monkeypatch.setattr(vs, "PAGE_SIZE", 16)
    tiny_model.sampler._buffers.clear()  # page tables are cached per row count
    B, T = 20, 40
    feats = make_avclip_features(B, 57)
    try:
        out = tiny_model.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                                  return_sampled_indices=True, check=True, _return_logits=True, _decode_audio=False)
    finally:
        tiny_model.sampler._buffers.clear()
    codes = out["sampled_indices"].cpu()
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    ref = tiny_oracle.forward_full(seq[..., :-1], feats.reshape(B, 32, 768))
    mine = out["_logits"][1:].cpu().permute(1, 2, 0, 3)
    assert rel_err(mine, ref) < BF16_LOGIT_TOL


def test_standalone_ops_match_torch_fp32():
    """Op-level entry points of the C ABI (the kernels bench.py times for the roofline)."""
    import ctypes as C

    lib = _cabi.load()
    g = torch.Generator().manual_seed(1)
    st = torch.cuda.current_stream().cuda_stream
    for R, N, K in ((1, 4608, 1536), (4, 1536, 4096), (7, 64, 384)):
        W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
        x = torch.randn(R, K, generator=g).cuda()
        y = torch.empty(R, N, device="cuda")
        _cabi.check(lib.vaura_gemv_bf16w(W.data_ptr(), x.data_ptr(), y.data_ptr(), N, K, R, st), "gemv")
        ref = x.double() @ W.double().t()
        assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    for R, N, K, bn in ((64, 8192, 1536, 64), (200, 1536, 4096, 128), (16, 1152, 384, 32)):
        W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
        A = torch.randn(R, K, generator=g).to(torch.bfloat16).cuda()
        y = torch.empty(R, N, device="cuda")
        _cabi.check(lib.vaura_linear_bf16(A.data_ptr(), W.data_ptr(), y.data_ptr(), R, N, K, bn, st), "linear")
        ref = A.double() @ W.double().t()  # bf16 operands, fp32 accumulate: only summation-order error remains
        assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-5
