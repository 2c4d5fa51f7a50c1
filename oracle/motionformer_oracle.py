"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's Segment-AVCLIP visual tower.

Plain PyTorch fp32 from a state dict with the reference's parameter names.  Follows, function by function:

  * tubelet embedding + CLS + separate space / time position embeddings:
        motionformer_src/video_model_builder.py:174-249, motionformer_src/vit_helper.py:521-553 (PatchEmbed3D)
  * divided space-time block (time attention -> space attention -> MLP, pre-norm residuals, CLS attending everything):
        motionformer_src/vit_helper.py:392-472 (DividedSpaceTimeBlock), :80-171 (DividedAttention), :34-45 (qkv_attn),
        :475-499 (Mlp)
  * feature head: drop CLS, final LayerNorm, per-frame spatial aggregation by a pre-norm TransformerEncoderLayer whose own
    CLS token is the output: models/modules/feature_extractors/avclip/motionformer.py:309-342, :366-446, :449-476
  * segment batching: motionformer.py:252-307 (for_loop=False)

Only the shipped configuration is restated (configs/modules/feature_extractors/avclip_vggsound.yaml: divided attention,
separate position embeddings, extract_features, factorize_space_time, agg_space_module = TransformerEncoderLayer,
agg_time_module = Identity, add_global_repr = False, no content mask).  Pinned on outputs of the reference's own code:
tests/golden/motionformer_full.npz (written by oracle/make_golden_motionformer.py), checked by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _ln(x, sd, key, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def _heads(t, h):  # (b, n, h*d) -> (b*h, n, d)      vit_helper.py:112
    b, n, hd = t.shape
    return t.reshape(b, n, h, hd // h).permute(0, 2, 1, 3).reshape(b * h, n, hd // h)


def _qkv_attn(q, k, v):  # vit_helper.py:34-45
    return torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v


def divided_attention(x, sd, prefix, heads, mode, frames, patches):
    """vit_helper.py:100-171.  mode 'time': a patch attends the same location in every frame (+ CLS); 'space': a patch
    attends its own frame (+ CLS); CLS attends every token in both."""
    b, n, D = x.shape
    d = D // heads
    qkv = F.linear(x, sd[prefix + ".qkv.weight"], sd[prefix + ".qkv.bias"])
    q, k, v = (_heads(t, heads) for t in qkv.chunk(3, dim=-1))
    q = q * d ** -0.5
    cls_q, q_ = q[:, :1], q[:, 1:]
    cls_k, k_ = k[:, :1], k[:, 1:]
    cls_v, v_ = v[:, :1], v[:, 1:]
    cls_out = _qkv_attn(cls_q, k, v)
    bh = b * heads
    if mode == "time":  # "b (f n) d -> (b n) f d"
        def re(t):
            return t.reshape(bh, frames, patches, d).transpose(1, 2).reshape(bh * patches, frames, d)
        r = patches
    else:               # "b (f n) d -> (b f) n d"
        def re(t):
            return t.reshape(bh * frames, patches, d)
        r = frames
    q_, k_, v_ = re(q_), re(k_), re(v_)
    ck = cls_k.repeat_interleave(r, dim=0)
    cv = cls_v.repeat_interleave(r, dim=0)
    out = _qkv_attn(q_, torch.cat((ck, k_), dim=1), torch.cat((cv, v_), dim=1))
    if mode == "time":
        out = out.reshape(bh, patches, frames, d).transpose(1, 2).reshape(bh, frames * patches, d)
    else:
        out = out.reshape(bh, frames * patches, d)
    out = torch.cat((cls_out, out), dim=1)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, D)  # "(b h) n d -> b n (h d)"
    return F.linear(out, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])


def block(x, sd, p, heads, frames, patches):  # vit_helper.py:443-472 (DropPath is the identity in eval mode)
    x = x + divided_attention(_ln(x, sd, p + ".norm3"), sd, p + ".timeattn", heads, "time", frames, patches)
    x = x + divided_attention(_ln(x, sd, p + ".norm1"), sd, p + ".attn", heads, "space", frames, patches)
    h = F.linear(_ln(x, sd, p + ".norm2"), sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])
    h = F.gelu(h)  # nn.GELU(): exact erf form
    return x + F.linear(h, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])


def spatial_aggregate(tok, sd, heads, p="spatial_attn_agg"):
    """motionformer.py:366-446 with nn.TransformerEncoderLayer(norm_first=True, activation=GELU, eps 1e-6): the layer's own
    CLS token is prepended to the 196 patch tokens of a frame and its output row is the frame's feature.
    tok: (sequences, patches, D) -> (sequences, D)."""
    s, n, D = tok.shape
    d = D // heads
    x = torch.cat((sd[p + ".cls_token"].expand(s, -1, -1), tok), dim=1)
    y = _ln(x, sd, p + ".norm1")
    qkv = F.linear(y, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"])
    q, k, v = (_heads(t, heads) for t in qkv.chunk(3, dim=-1))
    a = _qkv_attn(q * d ** -0.5, k, v)
    a = a.reshape(s, heads, n + 1, d).permute(0, 2, 1, 3).reshape(s, n + 1, D)
    x = x + F.linear(a, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
    h = F.gelu(F.linear(_ln(x, sd, p + ".norm2"), sd[p + ".linear1.weight"], sd[p + ".linear1.bias"]))
    x = x + F.linear(h, sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return x[:, 0]


@torch.no_grad()
def motionformer_features(frames, sd, dims):
    """frames (B, S, C, T, H, W) fp32 -> (B, S, t, D): what MotionFormer.forward returns as its first output in the shipped
    configuration (motionformer.py:252-307)."""
    B, S = frames.shape[:2]
    x = frames.reshape(B * S, *frames.shape[2:]).float()
    sd = {k: v.float() for k, v in sd.items()}
    D, heads, t, n = dims.embed_dim, dims.num_heads, dims.temporal, dims.patches_per_frame
    # video_model_builder.py:185, vit_helper.py:547-552: Conv3d with stride = kernel, tokens ordered (t, h, w)
    x = F.conv3d(x, sd["patch_embed_3d.proj.weight"], sd["patch_embed_3d.proj.bias"],
                 stride=(dims.tubelet, dims.patch_size, dims.patch_size))
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd["cls_token"].expand(B * S, -1, -1), x), dim=1)             # :213-214
    pos = sd["pos_embed"]                                                        # :238-245 ("separate")
    total = torch.cat((pos[:, :1], pos[:, 1:].repeat(1, t, 1) + sd["temp_embed"].repeat_interleave(n, 1)), dim=1)
    x = x + total
    for i in range(dims.depth):
        x = block(x, sd, f"blocks.{i}", heads, t, n)
    x = _ln(x[:, 1:], sd, "norm")                                                # motionformer.py:316-319
    x = x.reshape(B * S * t, n, D)                                               # :320-342: one sequence per frame
    return spatial_aggregate(x, sd, heads).reshape(B, S, t, D)
