"""The first thing a reference user does: VAURAModel.load_from_checkpoint(ckpt, hparams_file=..., map_location=device)
(scripts/generate.py:209-216).  A synthetic Lightning-style checkpoint + hparams.yaml go through that exact call and must
reproduce the tokens the unmodified reference produced from the same weights (tests/golden/tiny_greedy.npz)."""
import os

import numpy as np
import pytest
import torch
import yaml

from vaura_b200 import VAURAModel
from vaura_b200.config import get_file_with_best_val_loss
from vaura_b200.synthetic import TINY_CODEC, TINY_SAMPLER, make_avclip_features, make_checkpoint_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _write_experiment(tmp_path, with_codec_dims):
    sd = make_checkpoint_state_dict(TINY_SAMPLER, TINY_CODEC, 0)
    ckpt_dir = tmp_path / "checkpoints"
    ckpt_dir.mkdir()
    torch.save({"state_dict": sd, "epoch": 5, "global_step": 100}, ckpt_dir / "epoch=5-val_loss=3.21.ckpt")
    torch.save({"state_dict": {k: torch.zeros_like(v) for k, v in sd.items()}}, ckpt_dir / "epoch=2-val_loss=4.75.ckpt")
    s = TINY_SAMPLER
    codec_params = {"model_sr": 44100}
    if with_codec_dims:
        codec_params["dims"] = dict(latent_dim=TINY_CODEC.latent_dim, decoder_dim=TINY_CODEC.decoder_dim,
                                    decoder_rates=list(TINY_CODEC.decoder_rates), n_codebooks=9)
    hp = dict(
        learning_rate=5e-6, weight_decay=0.01, batch_size=2, use_visual_conditioning=True,
        feature_extractor_config={"target": "models.modules.feature_extractors.avclip.motionformer.MotionFormer",
                                  "params": {"ckpt_path": None, "extract_features": True}},
        audio_encoder_config={"target": "models.modules.dac.model.DacModelWrapper", "params": codec_params},
        sampler_config={"target": "models.modules.sampler.llama.Transformer",
                        "params": dict(num_layers=s.num_layers, d_model=s.d_model, d_codebook=1024, nhead=s.nhead,
                                       dim_feedforward=4096, num_codebooks=9, block_size_audio=256, block_size_video=64,
                                       layer_norm_eps=1e-5, cond_feature_channel_scaler=3)},
        visual_bridge_config={"target": "torch.nn.Identity"},
        pattern_provider_config={"target": "models.modules.misc.codebook_patterns.DelayedPatternProvider",
                                 "params": {"n_q": 9}},
        flatten_vis_feats=True, some_training_only_key=123,
    )
    hp_path = tmp_path / "hparams.yaml"
    hp_path.write_text(yaml.safe_dump(hp))
    return ckpt_dir, hp_path


@pytest.mark.parametrize("with_codec_dims", [False, True])
def test_load_from_checkpoint_reproduces_reference_tokens(tmp_path, with_codec_dims):
    ckpt_dir, hp_path = _write_experiment(tmp_path, with_codec_dims)
    ckpt = get_file_with_best_val_loss(ckpt_dir)  # scripts/generate.py resolve_ckpt -> utils/utils.py:30-45
    assert ckpt.name == "epoch=5-val_loss=3.21.ckpt"
    model = VAURAModel.load_from_checkpoint(ckpt, hparams_file=hp_path, map_location="cuda:0")
    model.eval()
    assert model.audio_encoder.__class__.__name__ == "DacModelWrapper" and model.audio_encoder.dims == TINY_CODEC
    model.sampler.audio_tokens_per_video_frame = 7  # scripts/generate.py:216
    g = np.load(os.path.join(GOLD, "tiny_greedy.npz"))
    B, T = int(g["B"]), int(g["T"])
    feats = make_avclip_features(B, int(g["feat_seed"]))
    out = model.generate(frames=feats.cuda(), max_new_tokens=T, use_sampling=False, prompt_is_encoded=True,
                         return_sampled_indices=True)
    assert torch.equal(out["sampled_indices"].cpu(), torch.from_numpy(g["codes"].astype(np.int64)))
    assert out["generated_audio"].shape == (B, 1, T * 512) and out["generated_audio"].dtype == torch.float16
    assert str(model.device) == "cuda:0"


def test_generate_without_explicit_audio_tokens_per_video_frame_raises(tmp_path):
    """Without scripts/generate.py:216 the sampler derives ceil((1 - 9) / 32) = 0 (llama.py:544-553) for a prompt-free call;
    the reference then fails on the host, and so does the mirror - the kernels never see a zero divisor."""
    ckpt_dir, hp_path = _write_experiment(tmp_path, True)
    model = VAURAModel.load_from_checkpoint(get_file_with_best_val_loss(ckpt_dir), hparams_file=hp_path,
                                            map_location="cuda:0").eval()
    feats = make_avclip_features(1, 3)
    with pytest.raises(ValueError):
        model.generate(frames=feats.cuda(), max_new_tokens=12, prompt_is_encoded=True)
    # a long enough prompt makes the derived value positive: ceil((64 + 1 - 9) / 32) = 2
    prompt = torch.randint(0, 1024, (1, 9, 64))
    out = model.generate(frames=feats.cuda(), audio=prompt.cuda(), max_new_tokens=80, prompt_is_encoded=True,
                         return_sampled_indices=True, _decode_audio=False)
    assert model.sampler.audio_tokens_per_video_frame == 2 and out["sampled_indices"].shape == (1, 9, 80)
