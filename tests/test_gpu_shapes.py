"""Shape sweep on the tiny model (same head width and vocabulary as the shipped one): batch sizes across every dispatch boundary
of `vaura_sampler_generate` (cluster kernel <= 2 rows, persistent 3-4, graph of GEMV kernels 5-15, fused bf16 step 16-64 and
65-128, multi-kernel bf16 above), with and without classifier-free guidance, with and without a prompt, odd clip lengths.
Teacher-forced comparison against the fp32 oracle: the oracle re-runs the full prefix on the tokens the GPU produced, so every
step is checked on a common history whatever near-ties do to the greedy sequence (models/vaura_model.py:502-547)."""
import os

import pytest
import torch

from oracle import vaura_oracle as vo
from vaura_b200.synthetic import TINY_CODEC, TINY_SAMPLER, build_model, make_avclip_features, make_sampler_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny_model():
    return build_model(TINY_SAMPLER, TINY_CODEC)


@pytest.fixture(scope="module")
def tiny_oracle():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    return vo.SamplerOracle(make_sampler_state_dict(TINY_SAMPLER, 0), TINY_SAMPLER)


CASES = [  # (batch, new tokens, prompt tokens, cfg_scale)
    (1, 1, 0, 1.0), (1, 7, 0, 6.0), (2, 13, 3, 1.0), (3, 9, 0, 1.0), (4, 5, 2, 1.0), (5, 11, 0, 6.0), (9, 6, 0, 1.0),
    (15, 4, 1, 1.0), (16, 10, 0, 1.0), (17, 3, 0, 1.0), (33, 12, 5, 6.0), (64, 6, 0, 1.0), (65, 5, 0, 1.0), (100, 4, 2, 1.0),
    (128, 3, 0, 1.0), (70, 4, 0, 6.0),
    # maximum length: T + prompt = 248 tokens -> 257 columns, the last step feeds position 255 = the last row of the RoPE table
    # (llama.py:364-368, :493-497) and the last slot of the eighth K/V page; with and without a long prompt prefill
    (1, 248, 0, 1.0), (16, 248, 0, 1.0), (3, 120, 128, 1.0), (40, 100, 148, 6.0),
]


@pytest.mark.parametrize("B,T,Tp,cfg", CASES)
def test_generate_shapes_against_teacher_forced_oracle(tiny_model, tiny_oracle, B, T, Tp, cfg):
    feats = make_avclip_features(B, 100 + B)
    total = T + Tp
    prompt = torch.randint(0, 1024, (B, 9, Tp), generator=torch.Generator().manual_seed(B)) if Tp else None
    out = tiny_model.generate(frames=feats.cuda(), audio=None if prompt is None else prompt.cuda(), max_new_tokens=total,
                              use_sampling=False, prompt_is_encoded=True, return_sampled_indices=True, cfg_scale=cfg,
                              _return_logits=True, _decode_audio=(B <= 4))
    codes = out["sampled_indices"].cpu()
    assert codes.shape == (B, 9, total) and int(codes.min()) >= 0 and int(codes.max()) < 1024
    if prompt is not None:
        assert torch.equal(codes[..., :Tp], prompt)                      # prompt-preserving write-back (vaura_model.py:536-544)
    if B <= 4:
        assert out["generated_audio"].shape == (B, 1, total * 512)
    rows = B * (2 if cfg > 1.0 else 1)
    bf16 = rows >= 16                                                     # greedy call: AUTO picks bf16 from 16 rows
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    f = feats.reshape(B, 32, 768)
    if cfg > 1.0:
        lg = tiny_oracle.forward_full(seq[..., :-1].repeat(2, 1, 1), torch.cat([f, torch.zeros_like(f) + tiny_oracle.uncond], 0))
        ref = lg[B:] + (lg[:B] - lg[B:]) * cfg
    else:
        ref = tiny_oracle.forward_full(seq[..., :-1], f)                  # (B, K, S-1, V)
    first = vo.first_step_with_timestep(Tp)                             # first sampled column (codebook_patterns.py:131-135)
    mine = out["_logits"][first:].cpu().permute(1, 2, 0, 3)              # logits of the steps that sampled columns first..S-1
    ref = ref[:, :, first - 1:]
    assert mine.shape == ref.shape, (mine.shape, ref.shape)
    err = float((mine - ref).abs().max() / ref.abs().max())
    # the CFG combine u + (c - u) * s amplifies the error of both halves by up to 2 s - 1 (tests/test_gpu_fullclip.py): the
    # bf16 bound is stated per model output, so it scales with the guidance; the fp32 path has the head-room to ignore it
    amp = (2 * cfg - 1) if cfg > 1.0 else 1.0
    assert err < (1e-2 * min(amp, 4.0) if bf16 else 2e-5 * amp), (err, rows)
    if not bf16:  # fp32-activation path: tokens follow the oracle's argmax wherever its top-2 gap is clear
        cols = torch.arange(first, total + 9)[None, :]
        tstep = cols - 1 - torch.arange(9)[:, None]
        mask = vo.pattern_mask(9, total)[:, first:] & (tstep >= Tp)     # sampled cells only (prompt cells are preserved)
        top2 = torch.topk(ref, 2, dim=-1).values
        clear = (top2[..., 0] - top2[..., 1]) > 1e-4
        agree = mine.argmax(-1) == ref.argmax(-1)
        assert bool(agree[clear & mask[None].expand_as(agree)].all())


@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (17, 5), (2, 129)])
def test_codec_decode_and_encode_shapes(tiny_model, B, T):
    """Codec at edge sizes (a single latent frame; a batch beyond the 16-clip workspace chunk; a length that is not a multiple of
    the 128-row tiles at any stage) against the fp32 oracles; decode -> encode keeps shapes (models/modules/dac/model.py:30-48)."""
    from oracle.dac_oracle import DacDecodeOracle
    from vaura_b200.synthetic import make_codec_state_dict

    codes = torch.randint(0, 1024, (B, 9, T), generator=torch.Generator().manual_seed(T))
    wav = tiny_model.audio_encoder.decode(codes.cuda())
    ref = DacDecodeOracle(make_codec_state_dict(TINY_CODEC, 100), TINY_CODEC).decode(codes)
    assert wav.shape == ref.shape == (B, 1, T * 512)
    r, x = ref.double().flatten(), wav.double().flatten().cpu()
    snr = float(10 * torch.log10(r.pow(2).sum() / (r - x).pow(2).sum().clamp_min(1e-30)))
    assert snr > 40.0, snr
    back = tiny_model.audio_encoder.encode(wav.float())
    assert back.shape == (B, 9, T) and int(back.min()) >= 0 and int(back.max()) < 1024
