"""Drop-in for the generation half of ``VAURAModel`` (models/vaura_model.py).

Same constructor keywords (vaura_model.py:28-48; they are the keys of a Lightning ``hparams.yaml``),
same ``load_from_checkpoint(path, hparams_file=..., map_location=...)`` call shape
(scripts/generate.py:209-211), same ``generate(...)`` signature, defaults and return dict
(vaura_model.py:411-427, :577-597).  Training (`forward`, losses, optimisers, logging) is out of scope.

What changes underneath: the reference re-runs the whole prefix through 24 layers for every new column
(vaura_model.py:502-547); here the 228-step loop — embedding, KV-cached layers, heads, CFG, sampling,
mask-fix and write-back — runs on the device behind one C-ABI call.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import torch

from . import _cabi
from .config import instantiate_from_config, load_yaml
from .patterns import DelayedPatternProvider
from .weights import load_lightning_checkpoint, split_state_dict


class VAURAModel(torch.nn.Module):
    def __init__(self, learning_rate: float = 5e-6, lr_scheduler: dict = None, weight_decay: float = 0.01,
                 betas: tuple = (0.9, 0.95), batch_size: int = 1, use_visual_conditioning: bool = True,
                 feature_extractor_config: dict = None, audio_encoder_config: dict = None, sampler_config: dict = None,
                 visual_bridge_config: dict = None, pattern_provider_config: dict = None,
                 predict_at_val_start: bool = False, return_attention_weights: bool = False,
                 plot_distr_of_pred_indices: bool = False, freeze_feature_extractor: bool = False,
                 files_to_track_during_training: List[str] = None, flatten_vis_feats: bool = False,
                 apply_per_video_frame_mask: bool = False):
        super().__init__()
        self.hparams = dict(locals())
        self.hparams.pop("self"), self.hparams.pop("__class__", None)
        self.batch_size = batch_size
        self.use_visual_conditioning = use_visual_conditioning
        self.freeze_feature_extractor = freeze_feature_extractor
        self.visual_feature_extractor = (
            instantiate_from_config(feature_extractor_config) if use_visual_conditioning else None)
        self.using_avclip = self.visual_feature_extractor.__class__.__name__ == "MotionFormer"  # vaura_model.py:73-75
        self.flatten_vis_feats = self.using_avclip and flatten_vis_feats
        self.sampler = instantiate_from_config(self._update_sampler_config(sampler_config))
        self.visual_bridge = instantiate_from_config(visual_bridge_config) if use_visual_conditioning else None
        self.audio_encoder = instantiate_from_config(audio_encoder_config)
        if hasattr(self.sampler, "initialize_embeddings") and self.audio_encoder.__class__.__name__ == "DacModelWrapper":
            self.sampler.initialize_embeddings(self.audio_encoder.model)  # vaura_model.py:86-88
        self.audio_encoder.model.half()  # vaura_model.py:92
        self.num_codebooks = self.sampler.num_codebooks
        if pattern_provider_config is not None:
            self.pattern_provider = instantiate_from_config(pattern_provider_config)
        else:
            self.pattern_provider = DelayedPatternProvider(n_q=self.num_codebooks)
        if hasattr(self.sampler, "codebook_pattern"):
            self.sampler.codebook_pattern = self.pattern_provider.__class__.__name__
        self.apply_per_video_frame_mask = apply_per_video_frame_mask
        self.return_attention_weights = return_attention_weights
        # Philox key of the device sampler; per-draw counter = (clip id, column, codebook, stream id).  Calls without
        # clip_indices take consecutive stream ids (the reference's torch.multinomial advances a global generator, so
        # two calls never repeat their draws either); calls with clip_indices use stream 0 unless `_stream_id` is given,
        # which makes a clip's tokens a function of (seed, clip id) alone - independent of batching and of the GPU count.
        self.seed = 0
        self._auto_stream = 0

    # vaura_model.py:690-697
    def _update_sampler_config(self, sampler_config: dict) -> dict:
        sampler_config = dict(sampler_config)
        params = dict(sampler_config.get("params") or {})
        params["use_visual_conditioning"] = self.use_visual_conditioning
        sampler_config["params"] = params
        return sampler_config

    @property
    def special_token_id(self) -> int:
        return self.sampler.d_codebook

    @property
    def device(self):
        return self.sampler.device

    # ---- loading ----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True, device=None):
        parts = split_state_dict(state_dict)
        if not parts["sampler"] or not parts["codec"]:
            raise KeyError("state dict lacks `sampler.*` or `audio_encoder.model.*` entries")
        self.sampler.load_state_dict(parts["sampler"], device=device)
        self.audio_encoder.load_state_dict(parts["codec"], device=device)
        self._load_feature_extractor(parts["feature_extractor"], device)
        return self

    def _load_feature_extractor(self, sd, device):
        """`visual_feature_extractor.*` entries of the checkpoint (the Segment-AVCLIP tower, vaura_model.py:64-75).  A
        checkpoint without them leaves the extractor in pass-through mode (precomputed features only)."""
        if sd and self.using_avclip and any(k.startswith("blocks.") for k in sd):
            self.visual_feature_extractor.load_state_dict(sd, device=device)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, hparams_file=None, map_location=None, **overrides):
        """Lightning's call shape (scripts/generate.py:209-211) without Lightning: hyper-parameters come from
        ``hparams_file`` (YAML, keys = constructor kwargs) or the checkpoint's ``hyper_parameters``."""
        parts, hp = load_lightning_checkpoint(str(checkpoint_path), map_location="cpu")
        if hparams_file is not None:
            hp = load_yaml(str(hparams_file))
        if hp is None:
            raise ValueError("no hparams_file given and the checkpoint holds no hyper_parameters")
        hp = {k: v for k, v in dict(hp).items() if k in cls.__init__.__code__.co_varnames}
        hp.update(overrides)
        model = cls(**hp)
        device = None
        if map_location is not None and str(map_location) != "cpu":
            device = torch.device(map_location)
        model.sampler.load_state_dict(parts["sampler"], device=device)
        model.audio_encoder.load_state_dict(parts["codec"], device=device)
        model._load_feature_extractor(parts["feature_extractor"], device)
        return model

    # ---- generation --------------------------------------------------------------------------------
    def _handle_visual_conditioning(self, frames: torch.Tensor, clip_indices: torch.Tensor, B: int):
        # vaura_model.py:194-214
        if not self.use_visual_conditioning:
            return None
        assert frames is not None
        if self.using_avclip:
            vis_feats, _ = self.visual_feature_extractor(frames)
            if self.flatten_vis_feats:
                B, S, Tv, D = vis_feats.shape
                vis_feats = vis_feats.reshape(B, S * Tv, D)
        else:
            vis_feats = self.visual_feature_extractor(frames)
        return self.visual_bridge(vis_feats)

    @torch.no_grad()
    def generate(self, frames: Union[torch.Tensor, None] = None, audio: Union[torch.Tensor, None] = None,
                 clip_indices: Union[torch.Tensor, None] = None, max_new_tokens: int = 512,
                 return_attention_weights: bool = False, return_sampled_indices: bool = False, check: bool = False,
                 use_sampling: bool = True, temp: float = 1.0, top_k: int = 256, top_p: float = 0.0,
                 remove_prompts: bool = False, prompt_is_encoded: bool = False, cfg_scale: float = 1.0,
                 _return_logits: bool = False, _precision: int = _cabi.PRECISION_AUTO, _decode_audio: bool = True,
                 _stream_id: Optional[int] = None, _end_offset: Optional[int] = None) -> dict:
        assert not self.training, "do not use generation in training mode"  # vaura_model.py:437
        if return_attention_weights:
            # the reference sampler returns None weights and then indexes them (llama.py:539, vaura_model.py:528)
            raise NotImplementedError("attention weights are not produced by the Llama-style sampler")
        if not self.use_visual_conditioning or frames is None:
            raise Exception("Not implemented")  # memory=None -> llama.py:475-477
        dev = self.device
        num_samples = frames.shape[0]
        K = self.num_codebooks
        if audio is None:
            audio = torch.zeros((num_samples, K, 0), dtype=torch.long, device=dev)
        elif not prompt_is_encoded:
            # raw-audio prompt -> codes (B, K, Tp).  The reference's EnCodec-era unpacking of the encoder output
            # (vaura_model.py:464-469) does not type-check against DacModelWrapper.encode's tensor; the codes are used as they are
            audio = self.audio_encoder.encode(audio)
        B, K, T = audio.shape
        # raw video segments in pinned host memory stay there: the extractor overlaps their copy with its own compute
        host_frames = self.using_avclip and frames.dim() == 6 and frames.device.type == "cpu" and frames.is_pinned()
        vis_feats = self._handle_visual_conditioning(frames if host_frames else frames.to(dev), clip_indices, B)
        start_offset = T
        assert start_offset < max_new_tokens, "gt audio prompt can not be longer than max_new_tokens"
        pattern = self.pattern_provider.get_pattern(max_new_tokens)
        unknown_token = -1
        gen_codes = torch.full((B, K, max_new_tokens), unknown_token, dtype=torch.long, device=dev)
        gen_codes[..., :start_offset] = audio.to(dev)
        gen_sequence, _, mask = pattern.build_pattern_sequence(gen_codes, self.special_token_id)
        start_offset_sequence = pattern.get_first_step_with_timesteps(start_offset)
        assert start_offset_sequence is not None
        S = gen_sequence.shape[-1]

        use_cfg = cfg_scale > 1.0 and self.sampler.__class__.__name__ == "Transformer"  # vaura_model.py:786-788
        condition = vis_feats.to(dev, torch.float32)
        if use_cfg:
            cond_null = torch.zeros_like(condition) + self.sampler.cls_embeddings.uncond_embedding  # :789-793
            condition = torch.cat([condition, cond_null], dim=0)
        if self.sampler.audio_tokens_per_video_frame is None:
            self.sampler._set_audio_tokens_per_video_frame(start_offset_sequence, condition.shape[1])
        cond_rows = self.sampler.cond_rows(condition)

        seq32 = gen_sequence.to(torch.int32).contiguous()
        logits_out = None
        if _return_logits:
            logits_out = torch.zeros(S, B, K, self.sampler.d_codebook, dtype=torch.float32, device=dev)
        ids = None
        if clip_indices is not None:
            ids = torch.as_tensor(clip_indices).to(device=dev, dtype=torch.int32).contiguous()
        if _stream_id is None:
            _stream_id = 0
            if clip_indices is None:
                _stream_id = self._auto_stream
                self._auto_stream += 1
        self.sampler.generate_tokens(seq32, cond_rows, timesteps=max_new_tokens, start_offset=start_offset_sequence,
                                     use_cfg=use_cfg, cfg_scale=cfg_scale, use_sampling=use_sampling, temp=temp,
                                     top_k=top_k, top_p=top_p, seed=self.seed, clip_ids=ids, logits_out=logits_out,
                                     precision=_precision, stream_id=_stream_id, end_offset=_end_offset)
        gen_sequence = seq32.to(torch.long)
        if check:
            # vaura_model.py:550-558 (these force a device->host sync, as they do in the reference)
            assert not (gen_sequence == unknown_token).any()
            assert (gen_sequence == torch.where(mask[None].expand(B, -1, -1), gen_sequence, self.special_token_id)).all()
        out_codes, _, out_mask = pattern.revert_pattern_sequence(gen_sequence, special_token=unknown_token)
        out_start_offset = start_offset if remove_prompts else 0
        out_codes = out_codes[..., out_start_offset:max_new_tokens]
        generated_item: Dict[str, Any] = {}
        if _decode_audio:
            sampled_frames = [(out_codes[..., : self.num_codebooks, :], None)]
            # the tokens were written by the sampling stage (always < vocab; the special id only sits in the pattern cells
            # that revert_pattern_sequence drops): no host-synchronising range check needed
            generated_item["generated_audio"] = self.audio_encoder.decode(sampled_frames, validate=False)
        else:
            generated_item["generated_audio"] = None
        generated_item["s_attn_weights"] = None
        generated_item["mha_attn_weights"] = None
        generated_item["sampled_indices"] = out_codes if return_sampled_indices else None
        if _return_logits:
            generated_item["_logits"] = logits_out
        return generated_item
