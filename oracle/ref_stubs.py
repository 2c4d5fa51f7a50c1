"""TEST INFRASTRUCTURE ONLY — not part of the product path.

Import stubs that let the *unmodified* reference (``/root/reference``) run in a container
that lacks pytorch_lightning / av / dac / omegaconf.  Only ``oracle/make_golden.py``
and ``tests/test_reference_crosscheck.py`` use this, and only when ``/root/reference`` exists
(it does not exist on the GPU box).

Nothing here re-implements reference arithmetic: the transformer, the delay pattern, the
samplers and ``VAURAModel.generate`` all execute from the reference's own files.  The only
arithmetic supplied from outside is the codec (the reference takes it from the pip package
``descript-audio-codec==1.0.0``, which is absent): ``transformers.DacModel`` (5.5.0,
architecture-equivalent) stands in for it.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("VAURA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "vaura_model.py"))


def _hf_dac(cdims):
    from transformers import DacConfig, DacModel

    cfg = DacConfig(
        sampling_rate=cdims.sample_rate,
        hidden_size=cdims.latent_dim,
        decoder_hidden_size=cdims.decoder_dim,
        upsampling_ratios=list(cdims.decoder_rates),
        downsampling_ratios=list(reversed(cdims.decoder_rates)),
        n_codebooks=cdims.n_codebooks,
        codebook_size=cdims.codebook_size,
        codebook_dim=cdims.codebook_dim,
        encoder_hidden_size=8,  # encoder is not on the path; keep it tiny
    )
    return DacModel(cfg).eval()


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| with the norm over all dims but 0."""
    dims = tuple(range(1, v.dim()))
    return g * v / v.pow(2).sum(dim=dims, keepdim=True).sqrt()


def load_dac_names_into_hf(hf_model, codec_sd):
    """Map dac-1.0.0 decoder/quantizer key names (weight-normed) onto transformers.DacModel
    (plain convs).  Used for the codec oracle."""
    def w(key):
        return fold_weight_norm(codec_sd[key + ".weight_g"], codec_sd[key + ".weight_v"])

    with torch.no_grad():
        for k, q in enumerate(hf_model.quantizer.quantizers):
            p = f"quantizer.quantizers.{k}"
            q.codebook.weight.copy_(codec_sd[f"{p}.codebook.weight"])
            q.out_proj.weight.copy_(w(f"{p}.out_proj"))
            q.out_proj.bias.copy_(codec_sd[f"{p}.out_proj.bias"])
        dec = hf_model.decoder
        dec.conv1.weight.copy_(w("decoder.model.0"))
        dec.conv1.bias.copy_(codec_sd["decoder.model.0.bias"])
        for i, blk in enumerate(dec.block):
            p = f"decoder.model.{i + 1}.block"
            blk.snake1.alpha.copy_(codec_sd[f"{p}.0.alpha"])
            blk.conv_t1.weight.copy_(w(f"{p}.1"))
            blk.conv_t1.bias.copy_(codec_sd[f"{p}.1.bias"])
            for j, ru in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
                q = f"{p}.{2 + j}.block"
                ru.snake1.alpha.copy_(codec_sd[f"{q}.0.alpha"])
                ru.conv1.weight.copy_(w(f"{q}.1"))
                ru.conv1.bias.copy_(codec_sd[f"{q}.1.bias"])
                ru.snake2.alpha.copy_(codec_sd[f"{q}.2.alpha"])
                ru.conv2.weight.copy_(w(f"{q}.3"))
                ru.conv2.bias.copy_(codec_sd[f"{q}.3.bias"])
        n = len(dec.block)
        dec.snake1.alpha.copy_(codec_sd[f"decoder.model.{n + 1}.alpha"])
        dec.conv2.weight.copy_(w(f"decoder.model.{n + 2}"))
        dec.conv2.bias.copy_(codec_sd[f"decoder.model.{n + 2}.bias"])
    return hf_model


def install_stubs(cdims, codec_sd, codec_dtype=torch.float32):
    """Pre-seed ``sys.modules`` so ``models.vaura_model`` imports."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl
    sys.modules["av"] = types.ModuleType("av")

    tu = types.ModuleType("utils.train_utils")

    def disabled_train(self, mode=True):  # utils/train_utils.py:198-201
        return self

    tu.disabled_train = disabled_train
    tu.generate_video_from_attn_weights = lambda *a, **k: None
    tu.combine_attn_weights_to_tensor = lambda *a, **k: None
    sys.modules["utils.train_utils"] = tu
    du = types.ModuleType("utils.data_utils")
    du.scale_tensor = lambda x, *a, **k: x
    sys.modules["utils.data_utils"] = du

    dac = types.ModuleType("dac")
    dac_model = types.ModuleType("dac.model")
    dac_model.DAC = object
    dac_nn = types.ModuleType("dac.nn")
    dac_layers = types.ModuleType("dac.nn.layers")
    dac_layers.WNConv1d = lambda *a, **k: torch.nn.utils.weight_norm(nn.Conv1d(*a, **k))
    dac.model, dac.nn, dac_nn.layers = dac_model, dac_nn, dac_layers
    sys.modules.update({"dac": dac, "dac.model": dac_model, "dac.nn": dac_nn, "dac.nn.layers": dac_layers})

    stubs = types.ModuleType("oracle_stub_modules")

    class DacModelWrapper(nn.Module):  # class name is checked at vaura_model.py:87
        def __init__(self, model_sr: int = 44100, ckpt_path=None):
            super().__init__()
            self.model_sr = model_sr
            self.model = load_dac_names_into_hf(_hf_dac(cdims), codec_sd)
            # the reference reads q.out_proj.in_channels etc. (llama.py:405-409): plain Conv1d has them

        @torch.no_grad()
        def decode(self, codes):
            # models/modules/dac/model.py:41-48
            if type(codes) == list:
                codes = codes[0][0]
            z, _, _ = self.model.quantizer.from_codes(codes)
            z = z.to(next(self.model.decoder.parameters()).dtype)
            return self.model.decoder(z)

    class MotionFormer(nn.Module):  # class name is checked at vaura_model.py:73-75
        """Pass-through: synthetic AVCLIP features go in as `frames` (B,S,t,D)."""

        def __init__(self, **kw):
            super().__init__()

        def forward(self, x):
            return x, None

    stubs.DacModelWrapper = DacModelWrapper
    stubs.MotionFormer = MotionFormer
    sys.modules["oracle_stub_modules"] = stubs
    return stubs


def build_reference_model(sdims, cdims, seed=0, codec_half=False):
    """The reference's own VAURAModel (models/vaura_model.py:27-120) with synthetic weights."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from vaura_b200.synthetic import make_codec_state_dict, make_sampler_state_dict

    codec_sd = make_codec_state_dict(cdims, seed + 100)
    install_stubs(cdims, codec_sd)
    from models.vaura_model import VAURAModel  # reference code

    sampler_cfg = {
        "target": "models.modules.sampler.llama.Transformer",
        "params": dict(num_layers=sdims.num_layers, d_model=sdims.d_model, d_codebook=sdims.d_codebook,
                       nhead=sdims.nhead, dim_feedforward=4096, dropout=0.1, activation="gelu",
                       layer_norm_eps=1e-5, batch_first=True, norm_first=True,
                       num_codebooks=sdims.num_codebooks, block_size_audio=sdims.block_size,
                       block_size_video=64, positional_embedder="learned",
                       cond_feature_channel_scaler=sdims.cond_feature_channel_scaler),
    }
    model = VAURAModel(
        use_visual_conditioning=True,
        feature_extractor_config={"target": "oracle_stub_modules.MotionFormer", "params": {}},
        audio_encoder_config={"target": "oracle_stub_modules.DacModelWrapper", "params": {"model_sr": 44100}},
        sampler_config=sampler_cfg,
        visual_bridge_config={"target": "torch.nn.Identity"},
        pattern_provider_config={"target": "models.modules.misc.codebook_patterns.DelayedPatternProvider",
                                 "params": {"n_q": sdims.num_codebooks}},
        flatten_vis_feats=True,
    )
    # vaura_model.py:92 halves the codec (weights rounded to fp16).  For the fp32 oracle run, restore
    # fp32 and reload the unrounded weights; with codec_half=True the reference's fp16 state is kept.
    if not codec_half:
        load_dac_names_into_hf(model.audio_encoder.model.float(), codec_sd)
    missing, unexpected = model.sampler.load_state_dict(make_sampler_state_dict(sdims, seed), strict=True)
    assert not missing and not unexpected
    model.eval()
    model.sampler.audio_tokens_per_video_frame = 7  # scripts/generate.py:216
    return model
