"""Checkpoint reading and weight packing (load-time plumbing; torch is used for memory only).

* ``load_lightning_checkpoint`` reads a reference ``.ckpt`` (``torch.load(...)["state_dict"]``) and
  splits it by the prefixes the reference uses: ``sampler.``, ``audio_encoder.model.``,
  ``visual_feature_extractor.``.
* ``pack_sampler`` folds the weight-normed token embeddings into tables (llama.py:60-73), stacks the
  per-layer matrices in the layouts ``include/vaura_b200.h`` documents (bf16), interleaves w1/w3 rows so
  the SwiGLU epilogue sees both halves in one warp, stacks the 9 heads, and precomputes the RoPE table
  (llama.py:593-603).
* ``pack_codec`` folds weight norm for every DAC conv, re-lays them out as per-tap [Cout][Cin] GEMM
  operands (polyphase for the transposed convs) in one fp16/fp32 blob.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .synthetic import CodecDims, SamplerDims


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm, dim=0: w = g * v / ||v||, norm over every dim but 0."""
    dims = tuple(range(1, v.dim()))
    return g * v / v.pow(2).sum(dim=dims, keepdim=True).sqrt()


def split_state_dict(sd: Dict[str, torch.Tensor]):
    out = {"sampler": {}, "codec": {}, "feature_extractor": {}, "other": {}}
    for k, v in sd.items():
        if k.startswith("sampler."):
            out["sampler"][k[len("sampler."):]] = v
        elif k.startswith("audio_encoder.model."):
            out["codec"][k[len("audio_encoder.model."):]] = v
        elif k.startswith("visual_feature_extractor."):
            out["feature_extractor"][k[len("visual_feature_extractor."):]] = v
        else:
            out["other"][k] = v
    return out


def load_lightning_checkpoint(path: str, map_location="cpu"):
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    return split_state_dict(sd), ckpt.get("hyper_parameters")


def rope_table(n_pos: int, head_dim: int, base: float = 10000.0) -> torch.Tensor:
    """llama.py:593-603 -> (n_pos, head_dim/2, 2) with (cos, sin)."""
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(n_pos).float(), freqs)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()


def pack_sampler(sd: Dict[str, torch.Tensor], dims: SamplerDims, device) -> Dict[str, torch.Tensor]:
    """-> dict of contiguous device tensors named like the fields of ``vaura_sampler_weights``."""
    L, d, F, K, V = dims.num_layers, dims.d_model, dims.ffn_dim, dims.num_codebooks, dims.d_codebook

    def f32(key):
        return sd[key].detach().to(torch.float32)

    def bf16_stack(keys):
        return torch.stack([sd[k].detach().to(device=device, dtype=torch.bfloat16) for k in keys]).contiguous()

    out = {}
    out["wqkv"] = bf16_stack([f"layers.{i}.attention.wqkv.weight" for i in range(L)])
    out["wo"] = bf16_stack([f"layers.{i}.attention.wo.weight" for i in range(L)])
    w13 = []
    for i in range(L):
        w1 = sd[f"layers.{i}.feed_forward.w1.weight"].detach().to(device=device, dtype=torch.bfloat16)
        w3 = sd[f"layers.{i}.feed_forward.w3.weight"].detach().to(device=device, dtype=torch.bfloat16)
        w13.append(torch.stack([w1, w3], dim=1).reshape(2 * F, d))
    out["w13"] = torch.stack(w13).contiguous()
    out["w2"] = bf16_stack([f"layers.{i}.feed_forward.w2.weight" for i in range(L)])
    out["w_heads"] = torch.cat([sd[f"lm_heads.{k}.weight"].detach().to(device=device, dtype=torch.bfloat16)
                                for k in range(K)]).contiguous()
    if cluster_stream_supported(dims):
        out["wstream"] = pack_cluster_stream(out, dims)
    out["attn_norm"] = torch.stack([f32(f"layers.{i}.attention_norm.weight") for i in range(L)]).to(device).contiguous()
    out["ffn_norm"] = torch.stack([f32(f"layers.{i}.ffn_norm.weight") for i in range(L)]).to(device).contiguous()
    out["final_norm"] = f32("norm.weight").to(device).contiguous()
    tables = []
    for k in range(K):
        p = f"tok_embeddings.{k}"
        if f"{p}.out_proj.weight_g" in sd:
            W = fold_weight_norm(f32(f"{p}.out_proj.weight_g"), f32(f"{p}.out_proj.weight_v"))[:, :, 0]
        else:  # new-style parametrisation or already folded
            W = f32(f"{p}.out_proj.weight")[:, :, 0]
        tables.append(f32(f"{p}.emb.weight") @ W.t() + f32(f"{p}.out_proj.bias"))
    out["tok_tables"] = torch.stack(tables).to(device).contiguous()
    assert out["tok_tables"].shape == (K, V + 1, dims.tok_dim), out["tok_tables"].shape
    out["rope"] = rope_table(dims.block_size, dims.head_dim, dims.rope_base).to(device)
    out["fc1"] = f32("cls_embeddings.projection.fc1.weight").to(device).contiguous()
    out["fc2"] = f32("cls_embeddings.projection.fc2.weight").to(device).contiguous()
    out["empty_video_emb"] = f32("empty_video_emb").reshape(-1).to(device).contiguous()
    out["uncond_embedding"] = f32("cls_embeddings.uncond_embedding").to(device).contiguous()
    return out


# ---- weight streams of the cluster-persistent decode kernel (csrc/decode_cluster.cu) ---------------------------
# 32 clusters (cl) x 4 CTAs (r); head h = cl // 2, s = cl % 2.  Bytes, in 12288-byte slots:
#   [16 heads][4 ranks][L][qkv 18 | wo 6]          shared streams, read by the CTAs (2h, r) and (2h+1, r)
#   [128 CTAs][L x (w13 16 | w2 8) | heads 18]       private streams
# slot = [warp w: 12][tile t: 2][lane: 32][e: 8] bf16.  A tile is 16 rows x 16 k in mma.m16n8k16 A-fragment order:
# lane = 4*gq + tq holds rows frow = gq + 8*((e>>1)&1), k fk = 2*tq + (e&1) + 8*(e>>2).  With j = 2*slot + t the tile
# (row-tile R, k-tile Kt) a warp owns and the source element of local (row lr = 16R + frow, k lk = 16Kt + fk) are:
#   qkv   R = 3*(w//2) + j//12, Kt = 12*(w%2) + j%12   row (lr//96)*d + 96h + lr%96        col 384r + lk
#   wo    slots 0-2: k-half 0, slots 3-5: k-half 1 of the head's 96 features (one sequence row: cluster s reads half s;
#         two rows: every cluster reads both halves for its own row); within a half jj = 2*(slot%3) + t:
#         R = 2w + jj//3, Kt = 3*(slot//3) + jj%3      row 384r + lr                       col 96h + lk
#   w13   R = j%4,              Kt = 8w + j//4         row 2*(128cl + 32r + 8R + frow%8) + frow//8 (interleaved w1|w3), col lk
#   w2    R = 2w + j//8,        Kt = j%8               row 384r + lr                       col 128cl + lk
#   heads R = 3*(w//2) + j//12, Kt = 12*(w%2) + j%12   row 288cl + lr                      col 384r + lk
CLUSTER_SLOT_ELEMS = 12 * 2 * 32 * 8
CLUSTER_PHASE_SLOTS = {"qkv": 18, "wo": 6, "w13": 16, "w2": 8, "heads": 18}


def cluster_stream_supported(dims: SamplerDims) -> bool:
    return (dims.d_model == 1536 and dims.nhead == 16 and dims.ffn_dim == 4096
            and dims.num_codebooks * dims.d_codebook == 9216)


def cluster_stream_index(phase: str, device="cpu") -> torch.Tensor:
    """Flat source index (into the row-major matrix of the phase) of every stream element:
    int64 [groups][4 ranks][slots][12 warps][2 tiles][32 lanes][8]; groups = 16 heads (qkv, wo) or 32 clusters."""
    G = CLUSTER_PHASE_SLOTS[phase]
    NG = 16 if phase in ("qkv", "wo") else 32
    ar = lambda n: torch.arange(n, device=device)
    c = ar(NG).view(NG, 1, 1, 1, 1, 1, 1)
    r = ar(4).view(1, 4, 1, 1, 1, 1, 1)
    g = ar(G).view(1, 1, G, 1, 1, 1, 1)
    w = ar(12).view(1, 1, 1, 12, 1, 1, 1)
    t = ar(2).view(1, 1, 1, 1, 2, 1, 1)
    lane = ar(32).view(1, 1, 1, 1, 1, 32, 1)
    e = ar(8).view(1, 1, 1, 1, 1, 1, 8)
    frow = lane // 4 + 8 * ((e >> 1) & 1)
    fk = 2 * (lane % 4) + (e & 1) + 8 * (e >> 2)
    j = 2 * g + t
    if phase in ("qkv", "heads"):
        R, Kt = 3 * (w // 2) + j // 12, 12 * (w % 2) + j % 12
    elif phase == "wo":
        jj = 2 * (g % 3) + t
        R, Kt = 2 * w + jj // 3, 3 * (g // 3) + jj % 3
    elif phase == "w2":
        R, Kt = 2 * w + j // 8, j % 8
    else:  # w13
        R, Kt = j % 4, 8 * w + j // 4
    lr, lk = 16 * R + frow, 16 * Kt + fk
    if phase == "qkv":
        row, col, ld = (lr // 96) * 1536 + 96 * c + lr % 96, 384 * r + lk, 1536
    elif phase == "wo":
        row, col, ld = 384 * r + lr, 96 * c + lk, 1536
    elif phase == "w13":
        row, col, ld = 2 * (128 * c + 32 * r + 8 * R + frow % 8) + frow // 8, lk, 1536
    elif phase == "w2":
        row, col, ld = 384 * r + lr, 128 * c + lk, 4096
    else:
        row, col, ld = 288 * c + lr, 384 * r + lk, 1536
    idx = row * ld + col
    return idx.expand(NG, 4, G, 12, 2, 32, 8).contiguous()


def pack_cluster_stream(packed: Dict[str, torch.Tensor], dims: SamplerDims) -> torch.Tensor:
    """-> flat uint8 tensor: the 64 shared streams followed by the 128 private streams (include/vaura_b200.h: wstream)."""
    assert cluster_stream_supported(dims)
    L = dims.num_layers
    dev = packed["wqkv"].device
    SE = CLUSTER_SLOT_ELEMS
    nsh = 64 * L * 24 * SE
    per_priv = (L * 24 + 18) * SE
    out = torch.empty(nsh + 128 * per_priv, dtype=torch.bfloat16, device=dev)
    sv = out[:nsh].view(64, L, 24 * SE)
    pv = out[nsh:].view(128, per_priv)
    iq = cluster_stream_index("qkv", dev).view(64, -1)
    io = cluster_stream_index("wo", dev).view(64, -1)
    for l in range(L):
        sv[:, l, :18 * SE] = packed["wqkv"][l].reshape(-1)[iq]
        sv[:, l, 18 * SE:] = packed["wo"][l].reshape(-1)[io]
    del iq, io
    off = 0
    for phase in ("w13", "w2"):
        idx = cluster_stream_index(phase, dev).view(128, -1)
        n = idx.shape[1]
        for l in range(L):
            pv[:, l * 24 * SE + off:l * 24 * SE + off + n] = packed[phase][l].reshape(-1)[idx]
        off += n
    assert off == 24 * SE
    idx = cluster_stream_index("heads", dev).view(128, -1)
    pv[:, L * 24 * SE:] = packed["w_heads"].reshape(-1)[idx]
    return out.view(torch.uint8)


def sampler_step_bytes(dims: SamplerDims) -> int:
    """Algorithmic weight bytes streamed by one decode step: layers + final norm + heads."""
    d, F = dims.d_model, dims.ffn_dim
    per_layer = (3 * d * d + d * d + 3 * d * F) * 2 + 2 * d * 4
    return dims.num_layers * per_layer + d * 4 + dims.num_codebooks * dims.d_codebook * d * 2


# ---- codec -------------------------------------------------------------------------------------------
def _conv_w(sd, key) -> torch.Tensor:
    if key + ".weight_g" in sd:
        return fold_weight_norm(sd[key + ".weight_g"].float(), sd[key + ".weight_v"].float())
    return sd[key + ".weight"].float()


def convt_polyphase(w: torch.Tensor, stride: int) -> Tuple[torch.Tensor, List[int]]:
    """ConvTranspose1d weight (Cin, Cout, 2s), padding ceil(s/2) -> ([s][2][Cout][Cin], offsets [s][2]).

    out[q*s + r] = sum_ci W[:, :, r+pad] x[q] + W[:, :, j1] x[q + o1]; (j1, o1) = (r+pad-s, +1) when
    r+pad >= s else (r+pad+s, -1).  Every output phase has exactly two taps."""
    if stride % 2 or w.shape[-1] != 2 * stride:
        raise ValueError("polyphase split needs an even stride and kernel = 2*stride (output length = L*stride)")
    pad = math.ceil(stride / 2)
    phases, offs = [], []
    for r in range(stride):
        j0 = r + pad
        j1, o1 = (j0 - stride, 1) if j0 >= stride else (j0 + stride, -1)
        phases.append(torch.stack([w[:, :, j0].t(), w[:, :, j1].t()]))
        offs += [0, o1]
    return torch.stack(phases).contiguous(), offs


def pack_codec(sd: Dict[str, torch.Tensor], dims: CodecDims, device):
    """-> (blob uint8 device tensor, offsets list[int] in bytes).  Slot order: see csrc/cabi.cu."""
    parts: List[torch.Tensor] = []

    def h(t):
        parts.append(t.to(torch.float16).contiguous())

    def f(t):
        parts.append(t.to(torch.float32).contiguous())

    def snake(alpha):
        """[C] alpha followed by [C] 1/(alpha + 1e-9): the epilogues multiply instead of dividing."""
        al = alpha.reshape(-1).to(torch.float32)
        parts.append(torch.cat([al, (al + 1e-9).reciprocal()]).contiguous())

    n = len(dims.decoder_rates)
    tables = []
    for k in range(dims.n_codebooks):
        p = f"quantizer.quantizers.{k}"
        W = _conv_w(sd, f"{p}.out_proj")[:, :, 0]  # (latent, 8)
        tables.append(sd[f"{p}.codebook.weight"].float() @ W.t() + sd[f"{p}.out_proj.bias"].float())
    h(torch.stack(tables))
    h(_conv_w(sd, "decoder.model.0").permute(2, 0, 1))  # (Cout,Cin,7) -> [7][Cout][Cin]
    f(sd["decoder.model.0.bias"])
    tap_tables: List[int] = []
    for dil in (1, 3, 9):
        tap_tables += [j * dil - (6 * dil) // 2 for j in range(7)]
    tap_tables += [0]
    tap_tables += [j - 3 for j in range(7)]
    for i, s in enumerate(dims.decoder_rates):
        p = f"decoder.model.{i + 1}.block"
        snake(sd[f"{p}.0.alpha"])
        wt, offs = convt_polyphase(_conv_w(sd, f"{p}.1"), s)
        h(wt)
        tap_tables += offs
        f(sd[f"{p}.1.bias"])
        for j in range(3):
            q = f"{p}.{2 + j}.block"
            snake(sd[f"{q}.0.alpha"])
            h(_conv_w(sd, f"{q}.1").permute(2, 0, 1))
            f(sd[f"{q}.1.bias"])
            snake(sd[f"{q}.2.alpha"])
            h(_conv_w(sd, f"{q}.3").permute(2, 0, 1))
            f(sd[f"{q}.3.bias"])
    snake(sd[f"decoder.model.{n + 1}.alpha"])
    f(_conv_w(sd, f"decoder.model.{n + 2}")[0].t())  # (1,Cl,7) -> [7][Cl] fp32
    f(sd[f"decoder.model.{n + 2}.bias"])
    parts.append(torch.tensor(tap_tables, dtype=torch.int32))

    offsets, cur = [], 0
    for t in parts:
        offsets.append(cur)
        cur += (t.numel() * t.element_size() + 255) // 256 * 256
    blob = torch.zeros(cur, dtype=torch.uint8)
    for o, t in zip(offsets, parts):
        blob[o:o + t.numel() * t.element_size()] = t.view(-1).view(torch.uint8)
    return blob.to(device), offsets


def strided_conv_frames(w: torch.Tensor, stride: int) -> torch.Tensor:
    """WNConv1d(Cin -> Cout, k = 2 s, stride s, padding ceil(s / 2)) weight (Cout, Cin, 2 s) -> [3][Cout][s * Cin]: the same
    convolution as three taps (frame offsets -1, 0, +1) over the channels-last input seen as frames of s samples
    ([T][Cin] -> [T / s][s * Cin], a free reshape): out[q] = sum_f W_f . frame[q + f].  Sample u = j - pad of tap j lies in
    frame floor(u / s) at row u mod s; positions no tap reaches stay zero."""
    cout, cin, k = w.shape
    assert k == 2 * stride
    pad = (stride + 1) // 2
    out = torch.zeros(3, cout, stride * cin, dtype=w.dtype)
    for j in range(k):
        u = j - pad
        f, r = u // stride, u % stride
        out[f + 1, :, r * cin:(r + 1) * cin] = w[:, :, j]
    return out


def pack_codec_encoder(sd: Dict[str, torch.Tensor], dims: CodecDims, device):
    """Encode half of a dac state dict -> (blob, offsets).  Slot order: include/vaura_b200.h (vaura_codec_encoder_create)."""
    parts: List[torch.Tensor] = []

    def h(t):
        parts.append(t.to(torch.float16).contiguous())

    def f(t):
        parts.append(t.to(torch.float32).contiguous())

    def snake(alpha):
        al = alpha.reshape(-1).to(torch.float32)
        parts.append(torch.cat([al, (al + 1e-9).reciprocal()]).contiguous())

    rates = dims.encoder_rates
    n = len(rates)
    f(_conv_w(sd, "encoder.block.0")[:, 0, :])  # (C0, 1, 7) -> [C0][7]
    f(sd["encoder.block.0.bias"])
    for i, s in enumerate(rates):
        p = f"encoder.block.{i + 1}.block"
        for j in range(3):
            q = f"{p}.{j}.block"
            snake(sd[f"{q}.0.alpha"])
            h(_conv_w(sd, f"{q}.1").permute(2, 0, 1))
            f(sd[f"{q}.1.bias"])
            snake(sd[f"{q}.2.alpha"])
            h(_conv_w(sd, f"{q}.3").permute(2, 0, 1))
            f(sd[f"{q}.3.bias"])
        snake(sd[f"{p}.3.alpha"])
        h(strided_conv_frames(_conv_w(sd, f"{p}.4").float(), s))
        f(sd[f"{p}.4.bias"])
    snake(sd[f"encoder.block.{n + 1}.alpha"])
    h(_conv_w(sd, f"encoder.block.{n + 2}").permute(2, 0, 1))
    f(sd[f"encoder.block.{n + 2}.bias"])
    w_in, b_in, cbn, tables = [], [], [], []
    for k in range(dims.n_codebooks):
        p = f"quantizer.quantizers.{k}"
        w_in.append(_conv_w(sd, f"{p}.in_proj")[:, :, 0].float())          # (Dc, latent)
        b_in.append(sd[f"{p}.in_proj.bias"].float())
        cb = sd[f"{p}.codebook.weight"].float()
        cbn.append(torch.nn.functional.normalize(cb))
        W = _conv_w(sd, f"{p}.out_proj")[:, :, 0].float()                   # (latent, Dc)
        tables.append(cb @ W.t() + sd[f"{p}.out_proj.bias"].float())
    f(torch.stack(w_in)); f(torch.stack(b_in)); f(torch.stack(cbn)); f(torch.stack(tables))
    taps: List[int] = []
    for dil in (1, 3, 9):
        taps += [j * dil - 3 * dil for j in range(7)]
    taps += [0]
    taps += [-1, 0, 1]
    parts.append(torch.tensor(taps, dtype=torch.int32))
    offsets, cur = [], 0
    for t in parts:
        offsets.append(cur)
        cur += (t.numel() * t.element_size() + 255) // 256 * 256
    blob = torch.zeros(cur, dtype=torch.uint8)
    for o, t in zip(offsets, parts):
        blob[o:o + t.numel() * t.element_size()] = t.view(-1).view(torch.uint8)
    return blob.to(device), offsets


def codec_encoder_flops(dims: CodecDims, samples: int) -> float:
    """2 * Cin * Cout * k * Lout per convolution of the encoder (the strided ones counted at their real 2 s taps)."""
    c, t, fl = dims.encoder_dim, samples, 2.0 * dims.encoder_dim * 7 * samples
    for s in dims.encoder_rates:
        fl += 3 * (2.0 * c * c * 7 * t + 2.0 * c * c * t)
        t //= s
        fl += 2.0 * c * (2 * c) * (2 * s) * t
        c *= 2
    return fl + 2.0 * c * dims.latent_dim * 3 * t


def pack_avclip(sd: Dict[str, torch.Tensor], dims, device):
    """MotionFormer state dict (reference names: video_model_builder.py:44-123, motionformer.py:166-185) ->
    (blob uint8 device tensor, offsets).  Slot order: include/vaura_b200.h (vaura_avclip_weights).  Matrices are stored as
    bf16 [out][in]; the tubelet Conv3d weight (D, C, 2, 16, 16) is already the [D][C*2*16*16] GEMM operand."""
    parts: List[torch.Tensor] = []

    def m(key):
        parts.append(sd[key].reshape(sd[key].shape[0], -1).to(torch.bfloat16).contiguous())

    def f(t):
        parts.append(t.reshape(-1).to(torch.float32).contiguous())

    def lin(prefix, wname="weight", bname="bias"):
        m(f"{prefix}.{wname}" if wname == "weight" else f"{prefix}{wname}")
        f(sd[f"{prefix}.{bname}" if bname == "bias" else f"{prefix}{bname}"])

    def norm(prefix):
        f(sd[prefix + ".weight"])
        f(sd[prefix + ".bias"])

    t, n = dims.temporal, dims.patches_per_frame
    lin("patch_embed_3d.proj")
    pos = sd["pos_embed"].float()[0]  # (1 + n, D)
    # total_pos_embed of video_model_builder.py:238-245 ("separate"): patch position tiled over frames + frame embedding
    f(pos[1:].repeat(t, 1) + sd["temp_embed"].float()[0].repeat_interleave(n, 0))
    f(sd["cls_token"].float().reshape(-1) + pos[0])
    for i in range(dims.depth):
        p = f"blocks.{i}"
        norm(f"{p}.norm3"); lin(f"{p}.timeattn.qkv"); lin(f"{p}.timeattn.proj")
        norm(f"{p}.norm1"); lin(f"{p}.attn.qkv"); lin(f"{p}.attn.proj")
        norm(f"{p}.norm2"); lin(f"{p}.mlp.fc1"); lin(f"{p}.mlp.fc2")
    norm("norm")
    a = "spatial_attn_agg"
    f(sd[f"{a}.cls_token"])
    norm(f"{a}.norm1")
    lin(f"{a}.self_attn.in_proj", "_weight", "_bias")
    lin(f"{a}.self_attn.out_proj")
    norm(f"{a}.norm2")
    lin(f"{a}.linear1")
    lin(f"{a}.linear2")
    offsets, cur = [], 0
    for tns in parts:
        offsets.append(cur)
        cur += (tns.numel() * tns.element_size() + 255) // 256 * 256
    blob = torch.zeros(cur, dtype=torch.uint8)
    for o, tns in zip(offsets, parts):
        blob[o:o + tns.numel() * tns.element_size()] = tns.view(-1).view(torch.uint8)
    return blob.to(device), offsets


def avclip_flops(dims, segments: int = 1) -> float:
    """Dense-contraction FLOPs of the tower for `segments` segments (2 * M * N * K per linear layer, attention included)."""
    D, T, t, n, F = dims.embed_dim, dims.tokens, dims.temporal, dims.patches_per_frame, dims.mlp_ratio * dims.embed_dim
    per_block = 2 * T * D * (3 * D) * 2 + 2 * T * D * D * 2 + 2 * T * D * F * 2
    attn = 2 * 2 * D * (t * n * (t + 1) + t * n * (n + 1) + 2 * T)
    agg = 2 * t * (n + 1) * D * 3 * D + 2 * 2 * D * t * (n + 1) + 2 * t * (D * D + 2 * D * F)
    return float(segments) * (2 * t * n * dims.patch_k * D + dims.depth * (per_block + attn) + agg)


def codec_flops(dims: CodecDims, frames: int) -> float:
    """2*Cin*Cout*k*Lout per conv (ConvT counted over its input length), SURVEY Appendix A."""
    fl = 2.0 * dims.latent_dim * dims.decoder_dim * 7 * frames
    t = frames
    for i, s in enumerate(dims.decoder_rates):
        cin, cout = dims.decoder_dim >> i, dims.decoder_dim >> (i + 1)
        fl += 2.0 * cin * cout * 2 * s * t
        t *= s
        fl += 3 * (2.0 * cout * cout * 7 * t + 2.0 * cout * cout * t)
    fl += 2.0 * (dims.decoder_dim >> len(dims.decoder_rates)) * 7 * t
    return fl
