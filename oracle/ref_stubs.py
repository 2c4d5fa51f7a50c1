"""TEST INFRASTRUCTURE ONLY — not part of the product path.

Import stubs that let the *unmodified* reference (``/root/reference``) run in a container
that lacks pytorch_lightning / av / dac / omegaconf.  Only the golden-vector generators
(``oracle/make_golden.py``, ``oracle/make_golden_motionformer.py``) and the CPU tests that re-check
the oracle against the live reference (``tests/test_host.py``, ``tests/test_oracle_golden.py``)
use this, and only when ``/root/reference`` exists (it does not exist on the GPU box).

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


def _hf_dac(cdims, with_encoder=False):
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
        encoder_hidden_size=cdims.encoder_dim if with_encoder else 8,
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


def load_dac_encoder_names_into_hf(hf_model, codec_sd):
    """dac-1.0.0 encoder / in_proj key names (weight-normed) -> transformers.DacModel (plain convs)."""
    def w(key):
        return fold_weight_norm(codec_sd[key + ".weight_g"], codec_sd[key + ".weight_v"])

    with torch.no_grad():
        for k, q in enumerate(hf_model.quantizer.quantizers):
            p = f"quantizer.quantizers.{k}.in_proj"
            q.in_proj.weight.copy_(w(p))
            q.in_proj.bias.copy_(codec_sd[p + ".bias"])
        enc = hf_model.encoder
        enc.conv1.weight.copy_(w("encoder.block.0"))
        enc.conv1.bias.copy_(codec_sd["encoder.block.0.bias"])
        for i, blk in enumerate(enc.block):
            p = f"encoder.block.{i + 1}.block"
            for j, ru in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
                q = f"{p}.{j}.block"
                ru.snake1.alpha.copy_(codec_sd[f"{q}.0.alpha"])
                ru.conv1.weight.copy_(w(f"{q}.1"))
                ru.conv1.bias.copy_(codec_sd[f"{q}.1.bias"])
                ru.snake2.alpha.copy_(codec_sd[f"{q}.2.alpha"])
                ru.conv2.weight.copy_(w(f"{q}.3"))
                ru.conv2.bias.copy_(codec_sd[f"{q}.3.bias"])
            blk.snake1.alpha.copy_(codec_sd[f"{p}.3.alpha"])
            blk.conv1.weight.copy_(w(f"{p}.4"))
            blk.conv1.bias.copy_(codec_sd[f"{p}.4.bias"])
        n = len(enc.block)
        enc.snake1.alpha.copy_(codec_sd[f"encoder.block.{n + 1}.alpha"])
        enc.conv2.weight.copy_(w(f"encoder.block.{n + 2}"))
        enc.conv2.bias.copy_(codec_sd[f"encoder.block.{n + 2}.bias"])
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


# ------------------------------------------------------------------------------------------------
# Segment-AVCLIP / MotionFormer: stubs for timm + omegaconf so the reference's own
# models/modules/feature_extractors/avclip/motionformer.py imports and runs unmodified
# ------------------------------------------------------------------------------------------------
class _AttrDict(dict):
    """What the reference needs from an OmegaConf node: attribute read / write and .get()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_attr(o):
    if isinstance(o, dict):
        return _AttrDict({k: _to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_to_attr(v) for v in o]
    return o


def install_motionformer_stubs():
    import yaml

    avclip = os.path.join(REFERENCE_ROOT, "models", "modules", "feature_extractors", "avclip")
    for p in (REFERENCE_ROOT, avclip):  # motionformer.py imports `motionformer_src.*` as a top-level package
        if p not in sys.path:
            sys.path.insert(0, p)

    oc = types.ModuleType("omegaconf")

    class OmegaConf:
        @staticmethod
        def load(path):
            with open(path) as f:
                return _to_attr(yaml.safe_load(f))

    oc.OmegaConf = OmegaConf
    sys.modules["omegaconf"] = oc

    class DropPath(nn.Module):  # timm.models.layers.DropPath: stochastic depth, identity in eval mode
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training, "stub DropPath is eval-only"
            return x

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    timm = types.ModuleType("timm")
    t_models = types.ModuleType("timm.models")
    t_layers = types.ModuleType("timm.models.layers")
    t_layers.trunc_normal_ = trunc_normal_
    t_layers.DropPath = DropPath
    t_layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    t_data = types.ModuleType("timm.data")
    t_data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    t_data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    t_resnet = types.ModuleType("timm.models.resnet")
    t_resnet.resnet26d = t_resnet.resnet50d = None
    t_registry = types.ModuleType("timm.models.registry")
    t_registry.register_model = lambda f: f
    timm.models, timm.data = t_models, t_data
    t_models.layers, t_models.resnet, t_models.registry = t_layers, t_resnet, t_registry
    sys.modules.update({"timm": timm, "timm.models": t_models, "timm.models.layers": t_layers, "timm.data": t_data,
                        "timm.models.resnet": t_resnet, "timm.models.registry": t_registry})


def build_reference_motionformer(seed=7):
    """The reference's own MotionFormer in the shipped configuration
    (configs/modules/feature_extractors/avclip_vggsound.yaml: extract_features, factorize_space_time,
    agg_space_module TransformerEncoderLayer, agg_time_module Identity, no global representation) with the synthetic
    weights of vaura_b200.synthetic.make_motionformer_state_dict."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from vaura_b200.synthetic import make_motionformer_state_dict

    install_motionformer_stubs()
    from models.modules.feature_extractors.avclip.motionformer import MotionFormer  # reference code

    m = MotionFormer(extract_features=True, ckpt_path=None, factorize_space_time=True,
                     agg_space_module="TransformerEncoderLayer", agg_time_module="torch.nn.Identity",
                     add_global_repr=False)
    sd = make_motionformer_state_dict(seed)
    own = m.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    assert not unexpected, unexpected
    # keys the path never reads (2-D patch_embed, classification head leftovers) may stay at their init
    assert all(k.startswith(("patch_embed.", "pre_logits", "head")) for k in missing), missing
    m.load_state_dict(sd, strict=False)
    return m.eval()
