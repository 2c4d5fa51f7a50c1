"""Deterministic synthetic weights / inputs for the V-AURA generation path.

There is no network and no trained checkpoint in this environment, so every
test, golden fixture and benchmark runs on random-init weights of the named
architecture.  The tensors are produced here, *not* by the reference's own
initialisers, so the same bytes can be regenerated on a box that has no copy of
the reference: each tensor is drawn from a CPU ``torch.Generator`` seeded by
``(seed, crc32(key))`` and is therefore independent of creation order.

Key names follow the reference's Lightning checkpoint layout:
``sampler.*`` mirrors ``models/modules/sampler/llama.py`` module names,
``audio_encoder.model.*`` mirrors descript-audio-codec 1.0.0 (``dac.DAC``).

Deviations from the reference's own init, on purpose:
  * ``lm_heads`` are zero-initialised there (llama.py:383-385) -> all logits 0;
    here N(0, 0.02).
  * ``empty_video_emb`` is ``torch.empty`` there (llama.py:336-338); here N(0, 0.02).
  * every sampler tensor is rounded to a bf16-representable value so that the
    bf16 weight storage used on the GPU is lossless w.r.t. the fp32 oracle.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch


def find_multiple(n: int, k: int) -> int:
    # llama.py:24-27
    return n if n % k == 0 else n + k - (n % k)


@dataclass(frozen=True)
class SamplerDims:
    """Shape parameters of the AR transformer (llama.py:286-377)."""

    num_layers: int = 24
    d_model: int = 1536
    nhead: int = 16
    d_codebook: int = 1024          # vocab (special id == d_codebook)
    num_codebooks: int = 9
    block_size: int = 256           # RoPE table length (llama.py:317, :364-368)
    cond_in: int = 768              # AVCLIP feature width
    cond_tokens: int = 32           # AVCLIPEmbedder.token_num
    cond_feature_channel_scaler: int = 3
    codebook_dim: int = 8           # DAC factorised code dim
    norm_eps: float = 1e-5
    rope_base: int = 10000

    @property
    def head_dim(self) -> int:
        return self.d_model // self.nhead

    @property
    def cond_dim(self) -> int:
        return self.d_model // self.cond_feature_channel_scaler

    @property
    def tok_dim(self) -> int:
        # channel concat: d_model = cond_dim + tok_dim (llama.py:472)
        return self.d_model - self.cond_dim

    @property
    def ffn_dim(self) -> int:
        # llama.py:164-169 (the YAML's dim_feedforward is ignored there)
        return find_multiple(int(2 * (4 * self.d_model) / 3), 256)


@dataclass(frozen=True)
class CodecDims:
    """Shape parameters of the DAC decoder (dac 1.0.0 ``DAC.__init__``)."""

    latent_dim: int = 1024
    decoder_dim: int = 1536
    decoder_rates: Tuple[int, ...] = (8, 8, 4, 2)
    n_codebooks: int = 9
    codebook_size: int = 1024
    codebook_dim: int = 8
    sample_rate: int = 44100
    encoder_dim: int = 64  # dac 1.0.0 `encoder_dim`; channels double per EncoderBlock, strides = reversed decoder rates

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.decoder_rates))

    @property
    def encoder_rates(self) -> Tuple[int, ...]:
        return tuple(reversed(self.decoder_rates))


FULL_SAMPLER = SamplerDims()
FULL_CODEC = CodecDims()
# Small shapes for fast parity cases: same head_dim (96), same vocabulary.
TINY_SAMPLER = SamplerDims(num_layers=2, d_model=384, nhead=4)
TINY_CODEC = CodecDims(latent_dim=256, decoder_dim=512, encoder_dim=32)


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**63 - 1))
    return g


def _randn(seed: int, key: str, shape, std: float = 1.0) -> torch.Tensor:
    return torch.randn(*shape, generator=_gen(seed, key), dtype=torch.float32) * std


def _rand(seed: int, key: str, shape, lo: float, hi: float) -> torch.Tensor:
    return torch.rand(*shape, generator=_gen(seed, key), dtype=torch.float32) * (hi - lo) + lo


def _bf16r(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def make_sampler_state_dict(dims: SamplerDims = FULL_SAMPLER, seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict for ``llama.Transformer`` after ``initialize_embeddings`` (no prefix)."""
    sd: Dict[str, torch.Tensor] = {}
    d, F, C, V, K = dims.d_model, dims.ffn_dim, dims.cond_dim, dims.d_codebook, dims.num_codebooks
    std = 0.02

    def put(key, t):
        sd[key] = _bf16r(t)

    for k in range(K):
        p = f"tok_embeddings.{k}"
        put(f"{p}.emb.weight", _randn(seed, f"{p}.emb.weight", (V + 1, dims.codebook_dim)))
        v = _randn(seed, f"{p}.out_proj.weight_v", (dims.tok_dim, dims.codebook_dim, 1), 0.3)
        put(f"{p}.out_proj.weight_v", v)
        g = v.flatten(1).norm(dim=1).view(-1, 1, 1) * _rand(seed, f"{p}.out_proj.weight_g", (dims.tok_dim, 1, 1), 0.5, 1.5)
        put(f"{p}.out_proj.weight_g", g)
        put(f"{p}.out_proj.bias", _randn(seed, f"{p}.out_proj.bias", (dims.tok_dim,), std))
    put("cls_embeddings.projection.fc1.weight", _randn(seed, "fc1", (C, dims.cond_in), std))
    put("cls_embeddings.projection.fc2.weight", _randn(seed, "fc2", (C, C), std))
    put("cls_embeddings.uncond_embedding",
        _randn(seed, "uncond", (dims.cond_tokens, dims.cond_in)) / dims.cond_in ** 0.5)
    put("empty_video_emb", _randn(seed, "empty_video_emb", (1, 1, C), std))
    for i in range(dims.num_layers):
        p = f"layers.{i}"
        put(f"{p}.attention.wqkv.weight", _randn(seed, f"{p}.wqkv", (3 * d, d), std))
        put(f"{p}.attention.wo.weight", _randn(seed, f"{p}.wo", (d, d), std))
        put(f"{p}.feed_forward.w1.weight", _randn(seed, f"{p}.w1", (F, d), std))
        put(f"{p}.feed_forward.w3.weight", _randn(seed, f"{p}.w3", (F, d), std))
        put(f"{p}.feed_forward.w2.weight", _randn(seed, f"{p}.w2", (d, F), std))
        put(f"{p}.attention_norm.weight", 1.0 + _randn(seed, f"{p}.an", (d,), 0.1))
        put(f"{p}.ffn_norm.weight", 1.0 + _randn(seed, f"{p}.fn", (d,), 0.1))
    put("norm.weight", 1.0 + _randn(seed, "norm", (d,), 0.1))
    for k in range(K):
        put(f"lm_heads.{k}.weight", _randn(seed, f"lm_heads.{k}", (V, d), std))
    return sd


def _wn(sd, seed, key, shape, fan_in, norm_dims):
    """Old-style torch weight_norm parameters (dim=0): weight = g * v / ||v||."""
    v = _randn(seed, key + ".weight_v", shape, 1.0 / math.sqrt(fan_in))
    nrm = v.pow(2).sum(dim=norm_dims, keepdim=True).sqrt()
    g = nrm * _rand(seed, key + ".weight_g", tuple(nrm.shape), 0.7, 1.3)
    sd[key + ".weight_v"] = v
    sd[key + ".weight_g"] = g
    sd[key + ".bias"] = _randn(seed, key + ".bias", (shape[0],), 0.05)


def make_codec_state_dict(dims: CodecDims = FULL_CODEC, seed: int = 100, with_encoder: bool = False) -> Dict[str, torch.Tensor]:
    """State dict for the decode half of ``dac.DAC`` (dac 1.0.0 names, no prefix).

    ``quantizer.quantizers.{k}.codebook.weight``, ``...out_proj.{weight_g,weight_v,bias}``,
    ``decoder.model.0`` (conv k7), ``decoder.model.{1+i}.block.{0: Snake alpha, 1: ConvTranspose1d,
    2..4: ResidualUnit.block.{0: alpha, 1: conv k7 dilated, 2: alpha, 3: conv k1}}``,
    ``decoder.model.{n+1}.alpha``, ``decoder.model.{n+2}`` (conv k7 -> 1).  ``with_encoder`` adds the encode half
    (make_codec_encoder_state_dict).
    """
    sd: Dict[str, torch.Tensor] = {}
    for k in range(dims.n_codebooks):
        p = f"quantizer.quantizers.{k}"
        sd[f"{p}.codebook.weight"] = _randn(seed, f"{p}.codebook", (dims.codebook_size, dims.codebook_dim))
        _wn(sd, seed, f"{p}.out_proj", (dims.latent_dim, dims.codebook_dim, 1), dims.codebook_dim * dims.n_codebooks, (1, 2))
    ch = dims.decoder_dim
    _wn(sd, seed, "decoder.model.0", (ch, dims.latent_dim, 7), dims.latent_dim * 7, (1, 2))
    for i, s in enumerate(dims.decoder_rates):
        cin, cout = ch // 2 ** i, ch // 2 ** (i + 1)
        p = f"decoder.model.{i + 1}.block"
        sd[f"{p}.0.alpha"] = _rand(seed, f"{p}.0.alpha", (1, cin, 1), 0.5, 2.0)
        # ConvTranspose1d weight is (Cin, Cout, k); weight_norm dim=0 -> norm per *input* channel
        v = _randn(seed, f"{p}.1.weight_v", (cin, cout, 2 * s), 0.6 / math.sqrt(cin * 2))
        nrm = v.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
        sd[f"{p}.1.weight_v"] = v
        sd[f"{p}.1.weight_g"] = nrm * _rand(seed, f"{p}.1.weight_g", tuple(nrm.shape), 0.7, 1.3)
        sd[f"{p}.1.bias"] = _randn(seed, f"{p}.1.bias", (cout,), 0.05)
        for j in range(3):
            q = f"{p}.{2 + j}.block"
            sd[f"{q}.0.alpha"] = _rand(seed, f"{q}.0.alpha", (1, cout, 1), 0.5, 2.0)
            _wn(sd, seed, f"{q}.1", (cout, cout, 7), cout * 7 * 2, (1, 2))
            sd[f"{q}.2.alpha"] = _rand(seed, f"{q}.2.alpha", (1, cout, 1), 0.5, 2.0)
            _wn(sd, seed, f"{q}.3", (cout, cout, 1), cout * 8, (1, 2))
    n = len(dims.decoder_rates)
    cl = ch // 2 ** n
    sd[f"decoder.model.{n + 1}.alpha"] = _rand(seed, "final.alpha", (1, cl, 1), 0.5, 2.0)
    _wn(sd, seed, f"decoder.model.{n + 2}", (1, cl, 7), cl * 7 * 4, (1, 2))
    if with_encoder:
        sd.update(make_codec_encoder_state_dict(dims, seed))
    return sd


def make_codec_encoder_state_dict(dims: CodecDims = FULL_CODEC, seed: int = 100) -> Dict[str, torch.Tensor]:
    """The encode half (dac 1.0.0 names): ``encoder.block.0`` (conv k7, 1 -> encoder_dim), ``encoder.block.{1+i}.block.{0..2:
    ResidualUnit.block.{0: alpha, 1: conv k7 dilated, 2: alpha, 3: conv k1}, 3: Snake alpha, 4: conv k = 2 s, stride s}``,
    ``encoder.block.{n+1}.alpha``, ``encoder.block.{n+2}`` (conv k3 -> latent) and ``quantizer.quantizers.{k}.in_proj``."""
    sd: Dict[str, torch.Tensor] = {}
    c = dims.encoder_dim
    _wn(sd, seed, "encoder.block.0", (c, 1, 7), 7 * 0.25, (1, 2))
    for i, s in enumerate(dims.encoder_rates):
        p = f"encoder.block.{i + 1}.block"
        for j in range(3):
            q = f"{p}.{j}.block"
            sd[f"{q}.0.alpha"] = _rand(seed, f"{q}.0.alpha", (1, c, 1), 0.5, 2.0)
            _wn(sd, seed, f"{q}.1", (c, c, 7), c * 7 * 2, (1, 2))
            sd[f"{q}.2.alpha"] = _rand(seed, f"{q}.2.alpha", (1, c, 1), 0.5, 2.0)
            _wn(sd, seed, f"{q}.3", (c, c, 1), c * 8, (1, 2))
        sd[f"{p}.3.alpha"] = _rand(seed, f"{p}.3.alpha", (1, c, 1), 0.5, 2.0)
        _wn(sd, seed, f"{p}.4", (2 * c, c, 2 * s), c * 2 * s * 0.5, (1, 2))
        c *= 2
    n = len(dims.encoder_rates)
    sd[f"encoder.block.{n + 1}.alpha"] = _rand(seed, f"encoder.block.{n + 1}.alpha", (1, c, 1), 0.5, 2.0)
    _wn(sd, seed, f"encoder.block.{n + 2}", (dims.latent_dim, c, 3), c * 3 * 0.5, (1, 2))
    for k in range(dims.n_codebooks):
        _wn(sd, seed, f"quantizer.quantizers.{k}.in_proj", (dims.codebook_dim, dims.latent_dim, 1), dims.latent_dim * 0.05, (1, 2))
    return sd


def make_checkpoint_state_dict(sdims: SamplerDims = FULL_SAMPLER, cdims: CodecDims = FULL_CODEC,
                               seed: int = 0) -> Dict[str, torch.Tensor]:
    """Lightning-style flat state dict: ``sampler.*`` + ``audio_encoder.model.*``."""
    sd = {"sampler." + k: v for k, v in make_sampler_state_dict(sdims, seed).items()}
    # a real checkpoint carries both halves of the codec (dac.DAC: encoder + quantizer + decoder)
    sd.update({"audio_encoder.model." + k: v for k, v in make_codec_state_dict(cdims, seed + 100, with_encoder=True).items()})
    return sd


@dataclass(frozen=True)
class AvclipDims:
    """Segment-AVCLIP visual tower = MotionFormer `divided_224_16x4` (motionformer_src/divided_224_16x4.yaml: ViT-B/16,
    2-frame tubelets, divided space-time attention) + one spatial aggregation layer (motionformer.py:166-185)."""
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: int = 4
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    frames: int = 16          # frames per segment
    tubelet: int = 2          # PATCH_SIZE_TEMP

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def temporal(self) -> int:  # TEMPORAL_RESOLUTION
        return self.frames // self.tubelet

    @property
    def patches_per_frame(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:  # CLS + t * h * w
        return 1 + self.temporal * self.patches_per_frame

    @property
    def patch_k(self) -> int:
        return self.in_chans * self.tubelet * self.patch_size * self.patch_size


FULL_AVCLIP = AvclipDims()
TINY_AVCLIP = AvclipDims(embed_dim=128, depth=2, num_heads=2, img_size=64, frames=8)


def make_motionformer_state_dict(seed: int = 7, dims: AvclipDims = FULL_AVCLIP) -> Dict[str, torch.Tensor]:
    """State dict with the reference's parameter names (video_model_builder.py:44-123, vit_helper.py:80-171, :392-472,
    motionformer.py:166-185).  Matrices are bf16-representable (the GPU path stores them as bf16); the scales keep every
    sub-layer's output at O(1) so that attention is neither uniform nor one-hot."""
    sd: Dict[str, torch.Tensor] = {}
    D, H = dims.embed_dim, dims.mlp_ratio * dims.embed_dim

    def mat(key, shape, fan_in, gain=1.0):
        sd[key] = _bf16r(_randn(seed, key, shape, gain / fan_in ** 0.5))

    def vec(key, n, std=0.05, mean=0.0):
        sd[key] = mean + _randn(seed, key, (n,), std)

    sd["cls_token"] = _randn(seed, "cls_token", (1, 1, D), 0.5)
    sd["pos_embed"] = _randn(seed, "pos_embed", (1, dims.patches_per_frame + 1, D), 0.2)
    sd["temp_embed"] = _randn(seed, "temp_embed", (1, dims.temporal, D), 0.2)
    mat("patch_embed_3d.proj.weight", (D, dims.in_chans, dims.tubelet, dims.patch_size, dims.patch_size), dims.patch_k)
    vec("patch_embed_3d.proj.bias", D)
    for i in range(dims.depth):
        p = f"blocks.{i}"
        for n in ("norm1", "norm2", "norm3"):
            vec(f"{p}.{n}.weight", D, 0.1, 1.0)
            vec(f"{p}.{n}.bias", D)
        for a in ("attn", "timeattn"):
            mat(f"{p}.{a}.qkv.weight", (3 * D, D), D)
            vec(f"{p}.{a}.qkv.bias", 3 * D)
            mat(f"{p}.{a}.proj.weight", (D, D), D, 0.5)
            vec(f"{p}.{a}.proj.bias", D)
        mat(f"{p}.mlp.fc1.weight", (H, D), D)
        vec(f"{p}.mlp.fc1.bias", H)
        mat(f"{p}.mlp.fc2.weight", (D, H), H, 0.5)
        vec(f"{p}.mlp.fc2.bias", D)
    vec("norm.weight", D, 0.1, 1.0)
    vec("norm.bias", D)
    p = "spatial_attn_agg"
    sd[f"{p}.cls_token"] = _randn(seed, f"{p}.cls_token", (1, 1, D), 0.5)
    mat(f"{p}.self_attn.in_proj_weight", (3 * D, D), D)
    vec(f"{p}.self_attn.in_proj_bias", 3 * D)
    mat(f"{p}.self_attn.out_proj.weight", (D, D), D, 0.5)
    vec(f"{p}.self_attn.out_proj.bias", D)
    mat(f"{p}.linear1.weight", (H, D), D)
    vec(f"{p}.linear1.bias", H)
    mat(f"{p}.linear2.weight", (D, H), H, 0.5)
    vec(f"{p}.linear2.bias", D)
    for n in ("norm1", "norm2"):
        vec(f"{p}.{n}.weight", D, 0.1, 1.0)
        vec(f"{p}.{n}.bias", D)
    return sd


def make_video_segments(batch: int, seed: int, segments: int = 4, dims: AvclipDims = FULL_AVCLIP) -> torch.Tensor:
    """Synthetic normalised RGB segments ``(B, S, C, T, H, W)`` (the extractor's input, motionformer.py:252-264);
    clip ``b`` only depends on ``(seed, b)``."""
    return torch.stack([
        _randn(seed, f"frames.{b}", (segments, dims.in_chans, dims.frames, dims.img_size, dims.img_size))
        for b in range(batch)
    ])


def make_avclip_features(batch: int, seed: int, segments: int = 4, tokens_per_segment: int = 8,
                         width: int = 768) -> torch.Tensor:
    """Synthetic Segment-AVCLIP output ``(B, S, t, D)`` as MotionFormer returns it
    (motionformer.py:252-342); clip ``b`` only depends on ``(seed, b)``."""
    return torch.stack([
        _randn(seed, f"avclip.{b}", (segments, tokens_per_segment, width)) for b in range(batch)
    ])


def build_model(sdims: SamplerDims = FULL_SAMPLER, cdims: CodecDims = FULL_CODEC, seed: int = 0, device="cuda:0"):
    """``VAURAModel`` of the named shape with the synthetic weights above, built through the same constructor
    keywords an experiment's ``hparams.yaml`` carries (reference ``target:`` strings; vaura_model.py:28-48) and put
    in the state ``scripts/generate.py:212-216`` leaves it in (eval mode, 7 audio tokens per video frame)."""
    from .model import VAURAModel

    cfg = dict(
        use_visual_conditioning=True,
        feature_extractor_config={"target": "models.modules.feature_extractors.avclip.motionformer.MotionFormer",
                                  # configs/modules/feature_extractors/avclip_vggsound.yaml (ckpt comes with the model)
                                  "params": dict(extract_features=True, factorize_space_time=True,
                                                 agg_space_module="TransformerEncoderLayer",
                                                 agg_time_module="torch.nn.Identity", add_global_repr=False)},
        audio_encoder_config={"target": "models.modules.dac.model.DacModelWrapper",
                              "params": {"model_sr": 44100, "dims": cdims}},
        sampler_config={"target": "models.modules.sampler.llama.Transformer",
                        "params": dict(num_layers=sdims.num_layers, d_model=sdims.d_model, d_codebook=sdims.d_codebook,
                                       nhead=sdims.nhead, num_codebooks=sdims.num_codebooks,
                                       block_size_audio=sdims.block_size, block_size_video=64,
                                       cond_feature_channel_scaler=sdims.cond_feature_channel_scaler)},
        visual_bridge_config={"target": "torch.nn.Identity"},
        pattern_provider_config={"target": "models.modules.misc.codebook_patterns.DelayedPatternProvider",
                                 "params": {"n_q": sdims.num_codebooks}},
        flatten_vis_feats=True,
    )
    m = VAURAModel(**cfg)
    m.load_state_dict(make_checkpoint_state_dict(sdims, cdims, seed), device=device)
    m.eval()
    m.sampler.audio_tokens_per_video_frame = 7
    return m
