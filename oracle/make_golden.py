"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the build container (needs /root/reference; the GPU box does not have it):

    python oracle/make_golden.py [--skip-full]

The reference has no tests or golden vectors for the generation path, so the oracle
(`oracle/vaura_oracle.py`, `oracle/dac_oracle.py`) is pinned on outputs of the reference's own
code: `VAURAModel.generate` (models/vaura_model.py:410-597), `Transformer.forward`
(models/modules/sampler/llama.py:520-539), `sample_top_k/top_p` (utils/utils.py:163-196) and the
`Pattern` class (models/modules/misc/codebook_patterns.py), executed under the import stubs in
`oracle/ref_stubs.py`.  Weights and features come from `vaura_b200/synthetic.py` (seeded, so the
fixtures only need to store outputs).  Codec goldens come from `transformers.DacModel` (the pip
package `descript-audio-codec` the reference uses is not available: parity with it is unpinned).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_stubs  # noqa: E402
from vaura_b200.synthetic import (FULL_CODEC, FULL_SAMPLER, TINY_CODEC, TINY_SAMPLER,  # noqa: E402
                                  make_avclip_features)

GOLD = os.path.join(ROOT, "tests", "golden")
MIN_GAP = 1e-3  # greedy goldens must not contain a near-tie (would make token parity ill-posed)


def step_stats(logits: torch.Tensor):
    """logits (steps,B,K,V) -> per-(step,b,k) summaries."""
    top2 = torch.topk(logits, 2, dim=-1).values
    return dict(
        argmax=logits.argmax(-1).to(torch.int16).numpy(),
        vmax=top2[..., 0].numpy(),
        gap=(top2[..., 0] - top2[..., 1]).numpy(),
        lse=torch.logsumexp(logits.double(), -1).float().numpy(),
    )


def teacher_forced_logits(model, codes, feats32, cfg_scale=1.0):
    """One reference forward over the built sequence gives the logits every decode step saw
    (SURVEY Appendix B corollary).  Returns (S-1, B, K, V) post-CFG logits for offsets 1..S-1."""
    from models.modules.misc.codebook_patterns import DelayedPatternProvider

    B, K, T = codes.shape
    pat = DelayedPatternProvider(K).get_pattern(T)
    seq, _, _ = pat.build_pattern_sequence(codes, model.special_token_id)
    S = seq.shape[-1]
    with torch.no_grad():
        cond = feats32
        inp = seq[..., : S - 1]
        if cfg_scale > 1.0:
            cond = torch.cat([cond, torch.zeros_like(cond) + model.sampler.cls_embeddings.uncond_embedding], 0)
            inp = inp.repeat(2, 1, 1)
        lg, _, _ = model.sampler(tgt=inp, memory=cond)
        if cfg_scale > 1.0:
            c, u = lg[:B], lg[B:]
            lg = u + (c - u) * cfg_scale
    return lg.permute(2, 0, 1, 3).contiguous(), seq  # (S-1,B,K,V)


def greedy_case(name, sdims, cdims, B, T, cfg_scale, prompt_len, feat_seed0, keep_steps, seed=0, min_gap_req=MIN_GAP):
    model = ref_stubs.build_reference_model(sdims, cdims, seed=seed)
    feat_seed = feat_seed0
    while True:
        feats = make_avclip_features(B, feat_seed)
        prompt = None
        if prompt_len:
            g = torch.Generator().manual_seed(feat_seed)
            prompt = torch.randint(0, sdims.d_codebook, (B, sdims.num_codebooks, prompt_len), generator=g)
        t0 = time.time()
        out = model.generate(frames=feats, audio=prompt, max_new_tokens=T, use_sampling=False,
                             prompt_is_encoded=True, return_sampled_indices=True, cfg_scale=cfg_scale,
                             check=True)
        dt = time.time() - t0
        codes = out["sampled_indices"]
        logits, seq = teacher_forced_logits(model, codes, feats.reshape(B, -1, feats.shape[-1]), cfg_scale)
        # logits[s-1] is what produced column s; only columns >= start and valid cells matter
        start = prompt_len + 1
        mask = torch.zeros(seq.shape[-1], sdims.num_codebooks, dtype=torch.bool)
        for k in range(sdims.num_codebooks):
            for s in range(start, seq.shape[-1]):
                mask[s, k] = (0 <= s - 1 - k < T) and (s - 1 - k >= prompt_len)
        st = step_stats(logits)
        gaps = torch.from_numpy(st["gap"])[start - 1:]  # (steps,B,K)
        m = mask[start:][:, None, :].expand(-1, B, -1)
        min_gap = gaps[m].min().item()
        # consistency: teacher-forced argmax reproduces the free-running tokens
        am = torch.from_numpy(st["argmax"]).long()[start - 1:]
        tgt = seq[..., start:].permute(2, 0, 1)
        assert torch.equal(am[m], tgt[m]), "reference generate != reference teacher-forced argmax"
        print(f"[{name}] feat_seed={feat_seed} generate {dt:.1f}s min top-2 gap {min_gap:.2e}")
        if min_gap >= min_gap_req:
            break
        feat_seed += 1000
    wav = out["generated_audio"].float()
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        codes=codes.to(torch.int16).numpy(),
        prompt=(prompt if prompt is not None else torch.zeros(B, sdims.num_codebooks, 0)).to(torch.int16).numpy(),
        feat_seed=feat_seed, weight_seed=seed, cfg_scale=cfg_scale, T=T, B=B, min_gap=min_gap,
        keep_steps=np.array(keep_steps), logits_keep=logits[keep_steps].numpy(),
        wav_fp16=wav.to(torch.float16).numpy(), ref_generate_seconds=dt,
        **{"stat_" + k: v for k, v in st.items()},
    )
    return model


def sampling_case():
    """Filtered / renormalised probabilities exactly as the reference computes them before
    torch.multinomial (utils/utils.py:163-196), captured by wrapping its `multinomial`."""
    import utils.utils as ru  # reference module

    g = torch.Generator().manual_seed(7)
    logits = torch.randn(6, 9, 1024, generator=g) * 1.5
    logits[0, 0, :40] = logits[0, 0, 40]  # exact ties across the top-k boundary region
    cap = {}
    orig = ru.multinomial

    def spy(inp, num_samples, **kw):
        cap["p"] = inp.clone()
        return orig(inp, num_samples, **kw)

    ru.multinomial = spy
    res = {"logits": logits.numpy()}
    try:
        for temp, k in ((1.0, 128), (0.7, 1), (1.3, 1024), (1.0, 256)):
            ru.sample_top_k(torch.softmax(logits / temp, -1), k)
            res[f"topk_t{temp}_k{k}"] = cap["p"].numpy()
        for temp, p in ((1.0, 0.9), (0.8, 0.5), (1.0, 1e-4)):
            ru.sample_top_p(torch.softmax(logits / temp, -1), p)
            res[f"topp_sorted_t{temp}_p{p}"] = cap["p"].numpy()
    finally:
        ru.multinomial = orig
    np.savez_compressed(os.path.join(GOLD, "sampling_filters.npz"), **res)


def teacher_case(sdims, cdims):
    """Reference forward on a random sequence that covers all 229 columns: special tokens,
    positions >= 224 (empty_video_emb rows) and the last RoPE rows."""
    model = ref_stubs.build_reference_model(sdims, cdims, seed=0)
    g = torch.Generator().manual_seed(11)
    B, K, S = 2, sdims.num_codebooks, 229
    seq = torch.randint(0, sdims.d_codebook + 1, (B, K, S), generator=g)
    feats = make_avclip_features(B, 5).reshape(B, 32, -1)
    with torch.no_grad():
        lg, _, _ = model.sampler(tgt=seq, memory=feats)
    keep = [0, 1, 6, 7, 8, 100, 223, 224, 228]
    np.savez_compressed(os.path.join(GOLD, "tiny_teacher_forced.npz"), seq=seq.to(torch.int16).numpy(),
                        feat_seed=5, keep=np.array(keep), logits_keep=lg[:, :, keep].numpy(),
                        lse=torch.logsumexp(lg.double(), -1).float().numpy())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    assert ref_stubs.reference_available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_float32_matmul_precision("highest")  # do NOT mirror main.py:34 (SURVEY §8c hygiene)

    greedy_case("tiny_greedy", TINY_SAMPLER, TINY_CODEC, B=2, T=20, cfg_scale=1.0, prompt_len=0,
                feat_seed0=1, keep_steps=[0, 1, 7, 8, 14, 27])
    greedy_case("tiny_cfg_prompt", TINY_SAMPLER, TINY_CODEC, B=2, T=24, cfg_scale=3.0, prompt_len=9,
                feat_seed0=2, keep_steps=[9, 10, 16, 31])
    teacher_case(TINY_SAMPLER, TINY_CODEC)
    sampling_case()
    if not args.skip_full:
        greedy_case("full_greedy", FULL_SAMPLER, FULL_CODEC, B=1, T=220, cfg_scale=1.0, prompt_len=0,
                    feat_seed0=1, keep_steps=[0, 1, 7, 63, 64, 128, 223, 224, 227], min_gap_req=1e-4)
    json.dump({"torch": torch.__version__, "made_by": "oracle/make_golden.py",
               "reference": "ilpoviertola/V-AURA at /root/reference (unmodified, run under oracle/ref_stubs.py)"},
              open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
