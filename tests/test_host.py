"""CPU tests of the host-side mirror: config loading, delay pattern, weight packing, C-ABI exports."""
import math
import os
import re

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_stubs, vaura_oracle as vo
from vaura_b200 import _cabi, config as vcfg
from vaura_b200.patterns import DelayedPatternProvider
from vaura_b200.synthetic import (FULL_SAMPLER, TINY_CODEC, TINY_SAMPLER, make_codec_state_dict,
                                  make_sampler_state_dict)
from vaura_b200.weights import convt_polyphase, fold_weight_norm, pack_codec, pack_sampler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = ref_stubs.reference_available()


# ---- config ------------------------------------------------------------------------------------------
def test_config_resolvers_and_targets():
    cfg = vcfg.load_config(os.path.join(ROOT, "tests/fixtures/configs/model_tiny.yaml"), base_dir=ROOT)
    assert cfg["model"]["batch_size"] == 2
    assert cfg["dataloader"]["partition_audio_to_clips"] is False  # negation of flatten_vis_feats
    sp = cfg["model"]["sampler_config"]
    assert sp["target"] == "models.modules.sampler.llama.Transformer"
    assert sp["params"]["layer_norm_eps"] == pytest.approx(1e-5)  # "1e-5" must parse as float like OmegaConf
    from vaura_b200.sampler import Transformer
    assert vcfg.get_obj_from_str(sp["target"]) is Transformer
    with pytest.raises(KeyError):
        vcfg.instantiate_from_config({"params": {}})
    ident = vcfg.instantiate_from_config({"target": "torch.nn.Identity"})
    assert isinstance(ident, torch.nn.Identity)


def test_config_dotlist_overrides():
    cfg = vcfg.load_config(os.path.join(ROOT, "tests/fixtures/configs/model_tiny.yaml"),
                           overrides=["dataloader.batch_size=7", "model.flatten_vis_feats=false"], base_dir=ROOT)
    assert cfg["model"]["batch_size"] == 7
    assert cfg["dataloader"]["partition_audio_to_clips"] is True


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_reference_yamls_load_unchanged():
    ref = ref_stubs.REFERENCE_ROOT
    cfg = vcfg.load_config(os.path.join(ref, "configs/vaura_defaults.yaml"), base_dir=ref)
    assert cfg["model"]["sampler_config"]["params"]["d_model"] == 1536
    assert cfg["model"]["pattern_provider_config"]["params"]["n_q"] == 9
    gen = vcfg.load_config(os.path.join(ref, "configs/generate_vgg.yaml"), base_dir=ref,
                           overrides=["experiment_path=/tmp/exp"])
    assert gen["checkpoint_path"] == "/tmp/exp/checkpoints/"
    assert (gen["top_k"], gen["cfg_scale"], gen["temperature"]) == (128, 6.0, 1.0)
    from vaura_b200 import VAURAModel
    m = VAURAModel(**{k: v for k, v in cfg["model"].items() if k != "name"})
    assert m.sampler.dims == FULL_SAMPLER and m.sampler.dims.ffn_dim == 4096 and m.num_codebooks == 9


def test_model_construction_mirrors_reference_attributes():
    from vaura_b200 import VAURAModel
    cfg = vcfg.load_config(os.path.join(ROOT, "tests/fixtures/configs/model_tiny.yaml"), base_dir=ROOT)
    m = VAURAModel(**cfg["model"])
    assert m.using_avclip and m.flatten_vis_feats
    assert m.special_token_id == 1024 and m.num_codebooks == 9
    assert m.sampler.config.block_size == 256  # scripts/generate.py:224 reads this
    assert m.sampler.codebook_pattern == "DelayedPatternProvider"
    assert m.sampler.audio_tokens_per_video_frame is None
    m.sampler.audio_tokens_per_video_frame = 7  # scripts/generate.py:216
    assert m.training  # like any nn.Module until .eval() is called
    with pytest.raises(AssertionError):
        m.generate(frames=torch.zeros(1, 4, 8, 768))  # vaura_model.py:437
    m.eval()
    with pytest.raises(NotImplementedError):
        m.generate(frames=torch.zeros(1, 4, 8, 768), return_attention_weights=True)


# ---- delay pattern -------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,T,Tp", [(9, 220, 0), (9, 24, 9), (2, 5, 1), (9, 1, 0)])
def test_pattern_matches_oracle_closed_form(K, T, Tp):
    g = torch.Generator().manual_seed(K * 1000 + T)
    codes = torch.randint(0, 1024, (3, K, T), generator=g)
    codes[..., Tp:] = -1
    pat = DelayedPatternProvider(K).get_pattern(T)
    seq, idx, mask = pat.build_pattern_sequence(codes, 1024)
    oseq, omask = vo.build_pattern_sequence(codes, 1024)
    assert torch.equal(seq, oseq) and torch.equal(mask, omask)
    assert seq.shape[-1] == T + K and (seq[..., 0] == 1024).all()
    assert pat.get_first_step_with_timesteps(Tp) == vo.first_step_with_timestep(Tp)
    full = torch.randint(0, 1024, (3, K, T), generator=g)
    s2, _, _ = pat.build_pattern_sequence(full, 1024)
    back, _, bmask = pat.revert_pattern_sequence(s2, special_token=-1)
    assert torch.equal(back, full) and bmask.all()
    assert torch.equal(back, vo.revert_pattern_sequence(s2, T))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
@pytest.mark.parametrize("K,T", [(9, 220), (9, 7), (4, 30)])
def test_pattern_matches_reference_class(K, T):
    ref_stubs.install_stubs(TINY_CODEC, make_codec_state_dict(TINY_CODEC, 100))
    from models.modules.misc.codebook_patterns import DelayedPatternProvider as RefProvider
    g = torch.Generator().manual_seed(T)
    codes = torch.randint(0, 1024, (2, K, T), generator=g)
    codes[..., T // 2:] = -1
    rp, mp = RefProvider(K).get_pattern(T), DelayedPatternProvider(K).get_pattern(T)
    rs, ri, rm = rp.build_pattern_sequence(codes, 1024)
    ms, mi, mm = mp.build_pattern_sequence(codes, 1024)
    assert torch.equal(rs, ms) and torch.equal(ri, mi) and torch.equal(rm, mm)
    rv, rvi, rvm = rp.revert_pattern_sequence(rs, special_token=-1)
    mv, mvi, mvm = mp.revert_pattern_sequence(ms, special_token=-1)
    assert torch.equal(rv, mv) and torch.equal(rvi, mvi) and torch.equal(rvm, mvm)
    for t in (0, 1, T - 1):
        assert rp.get_first_step_with_timesteps(t) == mp.get_first_step_with_timesteps(t)


# ---- weight packing ------------------------------------------------------------------------------------
def test_pack_sampler_layouts():
    d = TINY_SAMPLER
    sd = make_sampler_state_dict(d, 0)
    w = pack_sampler(sd, d, "cpu")
    o = vo.SamplerOracle(sd, d)
    assert torch.allclose(w["tok_tables"], torch.stack(o.tables), atol=1e-6)
    assert w["w13"].shape == (d.num_layers, 2 * d.ffn_dim, d.d_model)
    assert torch.equal(w["w13"][1, 0::2].float(), sd["layers.1.feed_forward.w1.weight"])  # bf16-exact weights
    assert torch.equal(w["w13"][1, 1::2].float(), sd["layers.1.feed_forward.w3.weight"])
    assert torch.equal(w["w_heads"][1024:2048].float(), sd["lm_heads.1.weight"])
    assert torch.allclose(w["rope"], o.rope)
    assert w["wqkv"].dtype == torch.bfloat16 and w["attn_norm"].dtype == torch.float32


@pytest.mark.parametrize("s", [2, 4, 8, 6])
def test_convtranspose_polyphase_equals_torch(s):
    g = torch.Generator().manual_seed(s)
    cin, cout, L = 6, 5, 11
    w = torch.randn(cin, cout, 2 * s, generator=g)
    x = torch.randn(2, cin, L, generator=g)
    ref = F.conv_transpose1d(x, w, stride=s, padding=math.ceil(s / 2))
    wp, offs = convt_polyphase(w, s)
    out = torch.zeros(2, cout, L * s)
    xp = F.pad(x, (1, 1))
    for r in range(s):
        for tap in range(2):
            o = offs[2 * r + tap]
            xs = xp[:, :, 1 + o: 1 + o + L]
            out[:, :, r::s] += torch.einsum("oc,bcl->bol", wp[r, tap], xs)
    assert ref.shape == out.shape and torch.allclose(ref, out, atol=1e-5)


def test_pack_codec_slots():
    c = TINY_CODEC
    sd = make_codec_state_dict(c, 100)
    blob, offs = pack_codec(sd, c, "cpu")
    n = len(c.decoder_rates)
    assert len(offs) == 3 + 21 * n + 3 + 1 and all(o % 256 == 0 for o in offs)
    w_in = fold_weight_norm(sd["decoder.model.0.weight_g"], sd["decoder.model.0.weight_v"]).permute(2, 0, 1)
    got = blob[offs[1]: offs[1] + w_in.numel() * 2].view(torch.float16).view(7, c.decoder_dim, c.latent_dim)
    assert torch.allclose(got.float(), w_in, atol=2e-3)
    taps = blob[offs[-1]:].view(torch.int32)
    assert taps[:7].tolist() == [-3, -2, -1, 0, 1, 2, 3] and taps[7:14].tolist() == [-9, -6, -3, 0, 3, 6, 9]
    assert taps[21].item() == 0 and taps[29:29 + 16].tolist() == [0, -1] * 4 + [0, 1] * 4  # stride 8, pad 4


# ---- C ABI ---------------------------------------------------------------------------------------------
def test_cabi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vaura_b200.h")).read()
    declared = set(re.findall(r"\b(vaura_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vaura_version() == 1 and lib.vaura_arch() == b"sm_100a"


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "vaura_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_environment_knobs_live_in_one_place():
    """csrc/knobs.h declares every environment variable the library reads; knobs.cu is the only file that calls getenv, and
    every variable it reads is documented in the header."""
    csrc = os.path.join(ROOT, "vaura_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh", ".h")) and f != "knobs.cu":
            assert "getenv" not in open(os.path.join(csrc, f)).read(), f
    read = set(re.findall(r'"(VAURA_[A-Z0-9_]+)"', open(os.path.join(csrc, "knobs.cu")).read()))
    documented = set(re.findall(r"\b(VAURA_[A-Z0-9_]+)", open(os.path.join(csrc, "knobs.h")).read()))
    assert read and read <= documented, read - documented


# ---- weight streams of the cluster-persistent decode kernel (weights.py: pack_cluster_stream) --------------------
STREAM_SHAPES = {"qkv": (4608, 1536), "wo": (1536, 1536), "w13": (8192, 1536), "w2": (1536, 4096), "heads": (9216, 1536)}


@pytest.mark.parametrize("phase", sorted(STREAM_SHAPES))
def test_cluster_stream_index_is_a_bijection(phase):
    """Every weight element appears exactly once in the streams (q|k|v once for the two clusters of a head)."""
    from vaura_b200.weights import CLUSTER_PHASE_SLOTS, CLUSTER_SLOT_ELEMS, cluster_stream_index

    rows, cols = STREAM_SHAPES[phase]
    idx = cluster_stream_index(phase)
    groups = 16 if phase in ("qkv", "wo") else 32
    assert idx.shape == (groups, 4, CLUSTER_PHASE_SLOTS[phase], 12, 2, 32, 8)
    assert idx.numel() == rows * cols and idx.numel() == groups * 4 * CLUSTER_PHASE_SLOTS[phase] * CLUSTER_SLOT_ELEMS
    counts = torch.bincount(idx.reshape(-1), minlength=rows * cols)
    assert int(counts.min()) == 1 and int(counts.max()) == 1


def test_cluster_stream_tile_is_an_mma_a_fragment():
    """A 512-byte tile holds a 16 x 16 block in mma.m16n8k16 A-fragment order: lane 4g+t owns rows g, g+8 and k 2t, 2t+1,
    2t+8, 2t+9; the warp/slot map is the one csrc/decode_cluster.cu consumes (wo: warp w = row tiles 2w, 2w+1; slots 0-2 hold
    k-tiles 0-2 of the head, slots 3-5 k-tiles 3-5)."""
    from vaura_b200.weights import cluster_stream_index

    idx = cluster_stream_index("wo")  # [head][rank][slot][warp][tile][lane][e]
    h, r, slot, w, t = 5, 2, 4, 7, 1
    tile = idx[h, r, slot, w, t]  # [32][8] flat indices into wo (1536 x 1536)
    jj = 2 * (slot % 3) + t
    row0 = 384 * r + 16 * (2 * w + jj // 3)
    col0 = 96 * h + 16 * (3 * (slot // 3) + jj % 3)
    for lane in (0, 5, 31):
        g, tq = lane // 4, lane % 4
        want = [(g, 2 * tq), (g, 2 * tq + 1), (g + 8, 2 * tq), (g + 8, 2 * tq + 1),
                (g, 2 * tq + 8), (g, 2 * tq + 9), (g + 8, 2 * tq + 8), (g + 8, 2 * tq + 9)]
        got = [(int(v) // 1536 - row0, int(v) % 1536 - col0) for v in tile[lane]]
        assert got == want


def test_precision_rule_and_environment_override(monkeypatch):
    """AUTO: fp32 activations below 16 sequence rows, bf16 from 16 (same rule as csrc/cabi.cu); VAURA_PRECISION overrides
    AUTO only, an explicit precision always wins."""
    from vaura_b200 import _cabi
    from vaura_b200.sampler import resolve_precision

    monkeypatch.delenv("VAURA_PRECISION", raising=False)
    assert resolve_precision(_cabi.PRECISION_AUTO, 1) == _cabi.PRECISION_FP32ACT
    assert resolve_precision(_cabi.PRECISION_AUTO, 15) == _cabi.PRECISION_FP32ACT
    assert resolve_precision(_cabi.PRECISION_AUTO, 16) == _cabi.PRECISION_BF16
    monkeypatch.setenv("VAURA_PRECISION", "bf16")
    assert resolve_precision(_cabi.PRECISION_AUTO, 4) == _cabi.PRECISION_BF16
    assert resolve_precision(_cabi.PRECISION_FP32ACT, 4) == _cabi.PRECISION_FP32ACT
    monkeypatch.setenv("VAURA_PRECISION", "fp32")
    assert resolve_precision(_cabi.PRECISION_AUTO, 64) == _cabi.PRECISION_FP32ACT


# ---- boundary: checkpoints, pickers, audio_tokens_per_video_frame --------------------------------------------------
def test_lightning_checkpoint_split_and_hparams(tmp_path):
    """A Lightning-style .ckpt (state_dict with the reference's prefixes + hyper_parameters) is split the way
    VAURAModel.load_from_checkpoint consumes it (scripts/generate.py:209-211)."""
    from vaura_b200.codec import DacModelWrapper
    from vaura_b200.synthetic import make_checkpoint_state_dict
    from vaura_b200.weights import load_lightning_checkpoint

    sd = make_checkpoint_state_dict(TINY_SAMPLER, TINY_CODEC, 0)
    sd["visual_feature_extractor.dummy"] = torch.zeros(1)
    sd["loss_weight"] = torch.ones(1)
    path = tmp_path / "epoch=3-val_loss=1.25.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": {"batch_size": 2}, "epoch": 3}, path)
    parts, hp = load_lightning_checkpoint(str(path))
    assert hp == {"batch_size": 2}
    assert set(parts) == {"sampler", "codec", "feature_extractor", "other"}
    assert "layers.0.attention.wqkv.weight" in parts["sampler"] and "decoder.model.0.weight_v" in parts["codec"]
    assert list(parts["feature_extractor"]) == ["dummy"] and list(parts["other"]) == ["loss_weight"]
    # codec shape parameters are read off the state dict when the YAML gives only model_sr
    assert DacModelWrapper.dims_from_state_dict(parts["codec"], 44100) == TINY_CODEC


def test_checkpoint_pickers(tmp_path):
    import time as _t

    for name in ("epoch=1-val_loss=2.50.ckpt", "epoch=7-val_loss=1.75.ckpt", "epoch=9-val_loss=1.80.ckpt"):
        (tmp_path / name).write_bytes(b"x")
        _t.sleep(0.01)
    assert vcfg.get_file_with_best_val_loss(tmp_path).name == "epoch=7-val_loss=1.75.ckpt"
    assert vcfg.get_latest_file(tmp_path, "*.ckpt").name == "epoch=9-val_loss=1.80.ckpt"
    only = tmp_path / "single"
    only.mkdir()
    (only / "last.ckpt").write_bytes(b"x")
    assert vcfg.get_file_with_best_val_loss(only).name == "last.ckpt"
    with pytest.raises(AssertionError):
        vcfg.get_file_with_best_val_loss(tmp_path / "single", "*.pt")


def test_audio_tokens_per_video_frame_must_be_positive():
    """llama.py:544-553 derives ceil((Ta - K) / Tv); without a prompt that is ceil(-8/32) = 0, which the kernels would
    divide by.  The host mirror raises instead of forwarding it (the reference fails on the host as well)."""
    from vaura_b200.sampler import Transformer

    t = Transformer(num_layers=2, d_model=384, nhead=4, num_codebooks=9, block_size_audio=256, cond_feature_channel_scaler=3)
    t.codebook_pattern = "DelayedPatternProvider"
    with pytest.raises(ValueError):
        t._set_audio_tokens_per_video_frame(1, 32)
    with pytest.raises(ValueError):
        t._set_audio_tokens_per_video_frame(5, 4)
    t._set_audio_tokens_per_video_frame(9 + 220, 32)
    assert t.audio_tokens_per_video_frame == 7
    t.weights = {"wqkv": torch.zeros(1)}
    t.audio_tokens_per_video_frame = 0
    with pytest.raises(ValueError):
        t.handle()


def test_precision_rule_takes_sampling_calls_from_three_rows(monkeypatch):
    from vaura_b200.sampler import resolve_precision

    monkeypatch.delenv("VAURA_PRECISION", raising=False)
    for rows, sampling, want in ((1, True, _cabi.PRECISION_FP32ACT), (2, True, _cabi.PRECISION_FP32ACT),
                                 (3, True, _cabi.PRECISION_BF16), (15, True, _cabi.PRECISION_BF16),
                                 (3, False, _cabi.PRECISION_FP32ACT), (15, False, _cabi.PRECISION_FP32ACT),
                                 (16, False, _cabi.PRECISION_BF16)):
        assert resolve_precision(_cabi.PRECISION_AUTO, rows, sampling) == want, (rows, sampling)


def test_generate_dataset_isolates_failures():
    """scripts/generate.py:386-389: one clip that raises does not end the run; it is reported and left silent."""
    from vaura_b200.driver import generate_dataset

    def gen(ids):
        if 5 in ids.tolist():
            raise RuntimeError("decode error")
        return ids.to(torch.float16).view(-1, 1, 1).expand(-1, 1, 4).contiguous()

    failed = []
    out = generate_dataset(gen, 9, 4, 0, 1, failed=failed)
    assert failed == [5] and out.shape == (9, 1, 4)
    assert out[:, 0, 0].tolist() == [0, 1, 2, 3, 4, 0, 6, 7, 8]


def test_strided_conv_as_three_frame_taps():
    """weights.strided_conv_frames: WNConv1d(k = 2 s, stride s, pad ceil(s / 2)) == a 3-tap conv over frames of s samples."""
    import torch.nn.functional as F

    from vaura_b200.weights import strided_conv_frames

    torch.manual_seed(0)
    for s in (2, 4, 8):
        cin, cout, T = 5, 7, 16 * s
        w, x = torch.randn(cout, cin, 2 * s), torch.randn(1, cin, T)
        ref = F.conv1d(x, w, stride=s, padding=(s + 1) // 2)
        wf = strided_conv_frames(w, s)                                  # [3][cout][s * cin]
        frames = x[0].t().reshape(T // s, s * cin)                      # channels-last [T][cin] -> [T / s][s * cin]
        fp = F.pad(frames, (0, 0, 1, 1))
        mine = sum(fp[f:f + T // s] @ wf[f].t() for f in range(3)).t()[None]
        assert torch.allclose(mine, ref, atol=1e-5), s


def test_avclip_and_codec_encoder_blobs_have_the_slot_counts_the_header_states():
    """include/vaura_b200.h: avclip blob = 4 + 18 * depth + 15 slots, codec encoder blob = 2 + 21 * n_blocks + 8 slots; the
    position table is pos_embed[1 + patch] + temp_embed[frame] (video_model_builder.py:238-245)."""
    from vaura_b200.synthetic import (TINY_AVCLIP, TINY_CODEC, make_codec_state_dict, make_motionformer_state_dict)
    from vaura_b200.weights import avclip_flops, codec_encoder_flops, pack_avclip, pack_codec_encoder

    sd = make_motionformer_state_dict(3, TINY_AVCLIP)
    blob, offs = pack_avclip(sd, TINY_AVCLIP, "cpu")
    assert len(offs) == 4 + 18 * TINY_AVCLIP.depth + 15 and all(o % 256 == 0 for o in offs) and offs == sorted(offs)
    t, n, D = TINY_AVCLIP.temporal, TINY_AVCLIP.patches_per_frame, TINY_AVCLIP.embed_dim
    pos = blob[offs[2]:offs[2] + t * n * D * 4].view(torch.float32).reshape(t, n, D)
    want = sd["pos_embed"][0, 1:][None] + sd["temp_embed"][0][:, None]
    assert torch.allclose(pos, want)
    cls = blob[offs[3]:offs[3] + D * 4].view(torch.float32)
    assert torch.allclose(cls, sd["cls_token"].reshape(-1) + sd["pos_embed"][0, 0])
    assert avclip_flops(TINY_AVCLIP, 2) == 2 * avclip_flops(TINY_AVCLIP, 1) > 0

    csd = make_codec_state_dict(TINY_CODEC, 100, with_encoder=True)
    blob, offs = pack_codec_encoder(csd, TINY_CODEC, "cpu")
    assert len(offs) == 2 + 21 * len(TINY_CODEC.decoder_rates) + 8 and all(o % 256 == 0 for o in offs)
    assert codec_encoder_flops(TINY_CODEC, 512 * 4) > 0


def test_motionformer_host_class_contract_without_a_gpu():
    """Class name / constructor keywords / pass-through / unsupported configurations (motionformer.py:47-75, :252-307)."""
    from vaura_b200.codec import DacModelWrapper
    from vaura_b200.features import MotionFormer

    m = MotionFormer(extract_features=True, ckpt_path="/path/to/vggsound/epoch_best.pt", factorize_space_time=True,
                     agg_space_module="TransformerEncoderLayer", agg_time_module="torch.nn.Identity", add_global_repr=False)
    assert m.__class__.__name__ == "MotionFormer" and m.embed_dim == 768 and not m.has_weights
    feats = torch.zeros(2, 4, 8, 768)
    out, glob = m(feats)
    assert out is feats and glob is None
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 3, 16, 224, 224))          # raw frames need the tower's weights
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 16, 224, 224))
    for bad in (dict(extract_features=False), dict(extract_features=True, factorize_space_time=False),
                dict(extract_features=True, agg_space_module="AveragePooling"),
                dict(extract_features=True, agg_time_module="TransformerEncoderLayer"),
                dict(extract_features=True, add_global_repr=True)):
        with pytest.raises(NotImplementedError):
            MotionFormer(**bad)._check_supported()
    c = DacModelWrapper(model_sr=44100)
    assert c.preprocess(torch.zeros(1, 1, 1000)).shape[-1] == 1024 and c.preprocess(torch.zeros(1, 1, 1024)).shape[-1] == 1024
    with pytest.raises(RuntimeError):
        c.encoder_handle()
