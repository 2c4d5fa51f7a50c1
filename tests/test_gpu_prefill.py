"""Prompt prefill of a call that samples (csrc/cabi.cu: transformer_pass_tc1): GEMM operands rounded to bf16 once, residual
stream / q / K/V pages / attention in fp32, followed by the fp32-activation decode steps.  The oracle is teacher-forced on the
tokens the GPU produced (models/vaura_model.py:502-547 re-runs the whole prefix per step), so every step after the prompt is
compared on a common history.  Tolerance: the bf16 bound of the north star, 1e-2 of max |logit|; with VAURA_PREFILL_BF16=0 the
same call keeps the three-term (fp32-equivalent) prefill and meets the fp32 bound."""
import os

import pytest
import torch

from oracle import vaura_oracle as vo
from vaura_b200 import _cabi
from vaura_b200.synthetic import (FULL_CODEC, FULL_SAMPLER, TINY_CODEC, TINY_SAMPLER, build_model, make_avclip_features,
                                  make_sampler_state_dict)

pytestmark = pytest.mark.gpu
BF16_LOGIT_TOL, FP32_LOGIT_TOL = 1e-2, 2e-5


def rel_err(a, ref):
    return float((a.cpu().double() - ref.double()).abs().max() / ref.double().abs().max())


def run(model, oracle, B, Tp, T, seed, cfg=1.0):
    feats = make_avclip_features(B, seed)
    prompt = torch.randint(0, 1024, (B, 9, Tp), generator=torch.Generator().manual_seed(seed))
    out = model.generate(frames=feats.cuda(), audio=prompt.cuda(), max_new_tokens=Tp + T, use_sampling=True, top_k=64,
                         prompt_is_encoded=True, return_sampled_indices=True, cfg_scale=cfg, _return_logits=True,
                         _decode_audio=False, _precision=_cabi.PRECISION_FP32ACT)
    codes = out["sampled_indices"].cpu()
    assert torch.equal(codes[..., :Tp], prompt)  # prompt-preserving write-back (vaura_model.py:536-544)
    seq, _ = vo.build_pattern_sequence(codes, 1024)
    f = feats.reshape(B, 32, 768)
    if cfg > 1.0:
        lg = oracle.forward_full(seq[..., :-1].repeat(2, 1, 1), torch.cat([f, torch.zeros_like(f) + oracle.uncond], 0))
        ref = lg[B:] + (lg[:B] - lg[B:]) * cfg
    else:
        ref = oracle.forward_full(seq[..., :-1], f)
    first = vo.first_step_with_timestep(Tp)  # first sampled column (codebook_patterns.py:131-135)
    mine = out["_logits"][first:].cpu().permute(1, 2, 0, 3)
    return rel_err(mine, ref[:, :, first - 1:])


@pytest.mark.parametrize("B,Tp,cfg", [(1, 40, 1.0), (1, 23, 3.0), (2, 64, 1.0), (4, 17, 1.0)])
def test_sampling_prefill_bf16_operands_tiny(B, Tp, cfg, monkeypatch):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = build_model(TINY_SAMPLER, TINY_CODEC)
    oracle = vo.SamplerOracle(make_sampler_state_dict(TINY_SAMPLER, 0), TINY_SAMPLER)
    monkeypatch.setenv("VAURA_PREFILL_BF16", "1")
    e1 = run(model, oracle, B, Tp, 10, 300 + Tp, cfg)
    monkeypatch.setenv("VAURA_PREFILL_BF16", "0")
    e3 = run(model, oracle, B, Tp, 10, 300 + Tp, cfg)
    print(f"[prefill B={B} Tp={Tp} cfg={cfg}] logit error: bf16 operands {e1:.2e}, three terms {e3:.2e}")
    assert e1 < BF16_LOGIT_TOL * (cfg if cfg > 1.0 else 1.0)  # CFG amplifies the error by ~cfg_scale
    assert e3 < FP32_LOGIT_TOL * (cfg if cfg > 1.0 else 1.0)
    assert e1 > e3  # the knob does select another path


def test_sampling_prefill_bf16_operands_full_size():
    """One window of the chunked long clip (BASELINE config 3): a 166-token prompt -> 167-position prefill, then decode steps on
    the cluster kernel over the fp32 K/V pages the prefill wrote."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = build_model(FULL_SAMPLER, FULL_CODEC)
    oracle = vo.SamplerOracle(make_sampler_state_dict(FULL_SAMPLER, 0), FULL_SAMPLER)
    e1 = run(model, oracle, 1, 166, 8, 77)
    print(f"[prefill full size, 166-token prompt] logit error with bf16 operands {e1:.2e}")
    assert e1 < BF16_LOGIT_TOL
