"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of V-AURA's generation hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the *checker* (or, for the
CPU baseline, as the thing timed on host cores).  The product path (``vaura_b200``) never
imports it and has no CPU fallback.

What is restated, in fp32 torch on CPU, each function citing the reference lines it follows
(paths relative to /root/reference):

  * the delay pattern in closed form            models/modules/misc/codebook_patterns.py
  * the LlamaGen-style transformer              models/modules/sampler/llama.py:445-517
    - full-prefix causal forward (what the reference executes every step), and
    - a KV-cached single-position step (mathematically identical at the last position)
  * the decode loop with CFG / mask-fix         models/vaura_model.py:410-597, :775-827
  * top-k / top-p filtering                     utils/utils.py:163-196
  * Philox4x32-10 + inverse-CDF draw (OUR sampler's RNG; the reference uses torch.multinomial,
    whose stream is implementation defined, so sampling parity with it is statistical only)

Pinning: the reference holds no golden vectors or tests for this path, so this
restatement is pinned against outputs of the reference itself, executed in the build container
under import stubs: ``oracle/make_golden.py`` writes ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks this module against them on every run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

UNKNOWN = -1  # vaura_model.py:482


# ----------------------------------------------------------------------------------------------
# delay pattern, closed form (checked against the reference Pattern in tests)
# ----------------------------------------------------------------------------------------------
def pattern_mask(K: int, T: int) -> torch.Tensor:
    """mask[k, s] = 0 <= s-1-k < T  (codebook_patterns.py:164-178 for DelayedPatternProvider:390-406)."""
    s = torch.arange(T + K)[None, :]
    k = torch.arange(K)[:, None]
    t = s - 1 - k
    return (t >= 0) & (t < T)


def build_pattern_sequence(codes: torch.Tensor, special: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """codes (B,K,T) -> seq (B,K,T+K): seq[b,k,s] = codes[b,k,s-1-k] if valid else special
    (codebook_patterns.py:180-207)."""
    B, K, T = codes.shape
    mask = pattern_mask(K, T)
    s = torch.arange(T + K)[None, :]
    k = torch.arange(K)[:, None]
    t = (s - 1 - k).clamp(0, max(T - 1, 0))
    if T == 0:
        return torch.full((B, K, K), special, dtype=codes.dtype), mask
    seq = torch.gather(codes, 2, t[None].expand(B, -1, -1))
    seq = torch.where(mask[None], seq, torch.full_like(seq, special))
    return seq, mask


def revert_pattern_sequence(seq: torch.Tensor, T: int) -> torch.Tensor:
    """seq (B,K,S) -> codes (B,K,T): out[b,k,t] = seq[b,k,t+k+1] (codebook_patterns.py:260-285)."""
    B, K, S = seq.shape
    idx = torch.arange(T)[None, :] + torch.arange(K)[:, None] + 1
    return torch.gather(seq, 2, idx[None].expand(B, -1, -1))


def first_step_with_timestep(t: int) -> int:
    """codebook_patterns.py:131-135 for the delayed pattern: timestep t of codebook 0 sits at column t+1."""
    return t + 1


# ----------------------------------------------------------------------------------------------
# transformer
# ----------------------------------------------------------------------------------------------
def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    dims = tuple(range(1, v.dim()))
    return g * v / v.pow(2).sum(dim=dims, keepdim=True).sqrt()


def rope_table(n_pos: int, head_dim: int, base: int = 10000) -> torch.Tensor:
    """llama.py:593-603 -> (n_pos, head_dim/2, 2) = (cos, sin)."""
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(n_pos).float(), freqs)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)


def apply_rope(x: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """llama.py:633-650; x (B,T,H,Dh), table (T,Dh/2,2); adjacent-pair rotation in fp32."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    c = table[None, :, None, :, 0]
    s = table[None, :, None, :, 1]
    out = torch.stack([xs[..., 0] * c - xs[..., 1] * s, xs[..., 1] * c + xs[..., 0] * s], dim=-1)
    return out.flatten(3)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """llama.py:147-158."""
    return x * torch.rsqrt(torch.mean(x * x, dim=-1, keepdim=True) + eps) * w


@dataclass
class OracleDims:
    num_layers: int
    d_model: int
    nhead: int
    vocab: int
    K: int
    block_size: int
    cond_dim: int
    eps: float = 1e-5

    @property
    def head_dim(self):
        return self.d_model // self.nhead


class SamplerOracle:
    """fp32 CPU restatement of llama.Transformer.inference for eval mode (dropouts are no-ops)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims, audio_tokens_per_video_frame: int = 7):
        self.dims = OracleDims(dims.num_layers, dims.d_model, dims.nhead, dims.d_codebook,
                               dims.num_codebooks, dims.block_size, dims.cond_dim, dims.norm_eps)
        self.sd = {k: v.float() for k, v in sd.items()}
        self.atpvf = audio_tokens_per_video_frame
        d = self.dims
        # llama.py:60-73 folded: E_k = emb_k @ W_k^T + b_k, W_k = g*v/||v||  
        self.tables: List[torch.Tensor] = []
        for k in range(d.K):
            p = f"tok_embeddings.{k}"
            W = fold_weight_norm(self.sd[f"{p}.out_proj.weight_g"], self.sd[f"{p}.out_proj.weight_v"])[:, :, 0]
            self.tables.append(self.sd[f"{p}.emb.weight"] @ W.t() + self.sd[f"{p}.out_proj.bias"])
        self.rope = rope_table(d.block_size, d.head_dim, dims.rope_base)
        self.uncond = self.sd["cls_embeddings.uncond_embedding"]
        self.w_heads = torch.cat([self.sd[f"lm_heads.{k}.weight"] for k in range(d.K)], 0)

    # -- conditioning -------------------------------------------------------------------------
    def cond_rows(self, feats: torch.Tensor) -> torch.Tensor:
        """feats (B,Tv,768) -> (B,Tv+1,C): MLP rows (llama.py:79-92) + empty_video_emb row
        (llama.py:336-338).  Position p reads row min(p // atpvf, Tv) (llama.py:555-586)."""
        h = F.linear(feats.float(), self.sd["cls_embeddings.projection.fc1.weight"])
        h = F.gelu(h, approximate="tanh")
        h = F.linear(h, self.sd["cls_embeddings.projection.fc2.weight"])
        e = self.sd["empty_video_emb"].expand(h.shape[0], 1, -1)
        return torch.cat([h, e], dim=1)

    def embed(self, tokens: torch.Tensor, cond_rows: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """tokens (B,K,P), positions (P,) -> h (B,P,d): cat(cond, sum_k E_k[tok]) (llama.py:455-472)."""
        tok = sum(self.tables[k][tokens[:, k]] for k in range(self.dims.K))
        Tv = cond_rows.shape[1] - 1
        rows = torch.clamp(positions // self.atpvf, max=Tv)
        return torch.cat([cond_rows[:, rows], tok], dim=-1)

    # -- full-prefix forward (what the reference runs each step) ----------------------------------
    def forward_full(self, seq: torch.Tensor, feats: torch.Tensor, return_hidden: bool = False):
        """seq (B,K,S) int64, feats (B,Tv,768) -> logits (B,K,S,V)  (llama.py:445-517)."""
        d = self.dims
        B, K, S = seq.shape
        pos = torch.arange(S)
        h = self.embed(seq, self.cond_rows(feats), pos)
        rope = self.rope[:S]
        for i in range(d.num_layers):
            p = f"layers.{i}"
            x = rmsnorm(h, self.sd[f"{p}.attention_norm.weight"], d.eps)
            qkv = F.linear(x, self.sd[f"{p}.attention.wqkv.weight"])
            q, k, v = qkv.split(d.d_model, dim=-1)
            q = apply_rope(q.view(B, S, d.nhead, d.head_dim), rope).transpose(1, 2)
            k = apply_rope(k.view(B, S, d.nhead, d.head_dim), rope).transpose(1, 2)
            v = v.view(B, S, d.nhead, d.head_dim).transpose(1, 2)
            a = F.scaled_dot_product_attention(q, k, v, is_causal=True)  # llama.py:246-255
            a = a.transpose(1, 2).reshape(B, S, d.d_model)
            h = h + F.linear(a, self.sd[f"{p}.attention.wo.weight"])
            x = rmsnorm(h, self.sd[f"{p}.ffn_norm.weight"], d.eps)
            ff = F.silu(F.linear(x, self.sd[f"{p}.feed_forward.w1.weight"])) * F.linear(
                x, self.sd[f"{p}.feed_forward.w3.weight"])
            h = h + F.linear(ff, self.sd[f"{p}.feed_forward.w2.weight"])
        hn = rmsnorm(h, self.sd["norm.weight"], d.eps)
        logits = F.linear(hn, self.w_heads).view(B, S, K, d.vocab).permute(0, 2, 1, 3)
        return (logits, h) if return_hidden else logits

    # -- KV-cached restatement ---------------------------------------------------------------------
    def new_cache(self, B: int):
        d = self.dims
        shape = (d.num_layers, B, d.nhead, d.block_size, d.head_dim)
        return {"k": torch.zeros(shape), "v": torch.zeros(shape), "len": 0}

    def forward_cached(self, tokens: torch.Tensor, cond_rows: torch.Tensor, cache) -> torch.Tensor:
        """Append P positions (tokens (B,K,P)) starting at cache['len']; return logits of the LAST
        appended position (B,K,V).  P>1 is a prefill."""
        d = self.dims
        B, K, P = tokens.shape
        p0 = cache["len"]
        assert p0 + P <= d.block_size, "RoPE table overflow (llama.py:364-368)"
        pos = torch.arange(p0, p0 + P)
        h = self.embed(tokens, cond_rows, pos)
        rope = self.rope[p0:p0 + P]
        scale = 1.0 / math.sqrt(d.head_dim)
        causal = torch.arange(p0 + P)[None, :] <= pos[:, None]  # (P, p0+P)
        for i in range(d.num_layers):
            p = f"layers.{i}"
            x = rmsnorm(h, self.sd[f"{p}.attention_norm.weight"], d.eps)
            qkv = F.linear(x, self.sd[f"{p}.attention.wqkv.weight"])
            q, k, v = qkv.split(d.d_model, dim=-1)
            q = apply_rope(q.view(B, P, d.nhead, d.head_dim), rope).transpose(1, 2)
            k = apply_rope(k.view(B, P, d.nhead, d.head_dim), rope).transpose(1, 2)
            v = v.view(B, P, d.nhead, d.head_dim).transpose(1, 2)
            cache["k"][i, :, :, p0:p0 + P] = k
            cache["v"][i, :, :, p0:p0 + P] = v
            kk = cache["k"][i, :, :, :p0 + P]
            vv = cache["v"][i, :, :, :p0 + P]
            s = (q @ kk.transpose(-1, -2)) * scale
            s = s.masked_fill(~causal[None, None], float("-inf"))
            a = torch.softmax(s, dim=-1) @ vv
            a = a.transpose(1, 2).reshape(B, P, d.d_model)
            h = h + F.linear(a, self.sd[f"{p}.attention.wo.weight"])
            x = rmsnorm(h, self.sd[f"{p}.ffn_norm.weight"], d.eps)
            ff = F.silu(F.linear(x, self.sd[f"{p}.feed_forward.w1.weight"])) * F.linear(
                x, self.sd[f"{p}.feed_forward.w3.weight"])
            h = h + F.linear(ff, self.sd[f"{p}.feed_forward.w2.weight"])
        cache["len"] = p0 + P
        hn = rmsnorm(h[:, -1], self.sd["norm.weight"], d.eps)
        return F.linear(hn, self.w_heads).view(B, K, d.vocab)


# ----------------------------------------------------------------------------------------------
# sampling
# ----------------------------------------------------------------------------------------------
def filtered_probs(logits: torch.Tensor, temp: float, top_k: int, top_p: float) -> torch.Tensor:
    """softmax(logits/temp) then top-p (if >0) else top-k (if >0) masking + renormalisation, in
    vocabulary order (vaura_model.py:816-823; utils/utils.py:163-177, :180-196)."""
    probs = torch.softmax(logits / temp, dim=-1)
    if top_p > 0.0:
        ps, idx = torch.sort(probs, dim=-1, descending=True)
        cs = torch.cumsum(ps, dim=-1)
        ps = ps * (~(cs - ps > top_p)).float()
        ps = ps / ps.sum(dim=-1, keepdim=True)
        return torch.zeros_like(probs).scatter_(-1, idx, ps)
    if top_k > 0:
        kth = torch.topk(probs, top_k, dim=-1)[0][..., [-1]]
        probs = probs * (probs >= kth).float()
        probs = probs / probs.sum(dim=-1, keepdim=True)
    return probs


_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: Tuple[int, int, int, int], key: Tuple[int, int]) -> Tuple[int, int, int, int]:
    """Philox4x32-10 (Salmon et al. 2011), the generator our sampling kernel implements."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, \
                         ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_uniform(seed: int, clip_id: int, step: int, codebook: int, stream_id: int = 0) -> float:
    """u in [0,1): counter = (clip_id, step, codebook, stream_id), key = (seed lo, seed hi); first word,
    top 24 bits."""
    r = philox4x32_10((clip_id & 0xFFFFFFFF, step & 0xFFFFFFFF, codebook & 0xFFFFFFFF, stream_id & 0xFFFFFFFF),
                      (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return (r[0] >> 8) * (1.0 / 16777216.0)


def inverse_cdf_draw(probs: np.ndarray, u: float) -> int:
    """Smallest index whose inclusive prefix sum exceeds u*total (such an index always has
    probs > 0); if rounding leaves none, the last index with probs > 0."""
    c = np.cumsum(probs.astype(np.float64))
    idx = int(np.searchsorted(c, u * c[-1], side="right"))
    if idx >= len(probs):
        idx = int(np.nonzero(probs > 0)[0][-1])
    return idx


# ----------------------------------------------------------------------------------------------
# decode loop
# ----------------------------------------------------------------------------------------------
def generate_tokens(
    oracle: SamplerOracle,
    feats: torch.Tensor,
    prompt: Optional[torch.Tensor] = None,
    max_new_tokens: int = 220,
    use_sampling: bool = False,
    temp: float = 1.0,
    top_k: int = 256,
    top_p: float = 0.0,
    cfg_scale: float = 1.0,
    seed: int = 0,
    clip_ids: Optional[List[int]] = None,
    collect_logits: bool = False,
    stream_id: int = 0,
):
    """KV-cached restatement of VAURAModel.generate up to out_codes (vaura_model.py:455-572).

    feats (B,Tv,768); prompt (B,K,Tp) int64 or None.  Returns codes (B,K,max_new_tokens) and,
    optionally, the per-step post-CFG logits (steps,B,K,V)."""
    d = oracle.dims
    B = feats.shape[0]
    K, T = d.K, max_new_tokens
    special = d.vocab
    if prompt is None:
        prompt = torch.zeros((B, K, 0), dtype=torch.long)
    Tp = prompt.shape[-1]
    assert Tp < T, "gt audio prompt can not be longer than max_new_tokens"  # vaura_model.py:476-478
    codes = torch.full((B, K, T), UNKNOWN, dtype=torch.long)
    codes[..., :Tp] = prompt
    # build_pattern_sequence keeps -1 for not-yet-generated valid cells (vaura_model.py:485-493)
    seq, mask = build_pattern_sequence(codes, special)
    S = seq.shape[-1]
    start = first_step_with_timestep(Tp)  # vaura_model.py:496
    use_cfg = cfg_scale > 1.0  # vaura_model.py:786-788
    if use_cfg:
        # vaura_model.py:789-795: [cond; uncond] on the batch axis
        feats_all = torch.cat([feats, torch.zeros_like(feats) + oracle.uncond], dim=0)
    else:
        feats_all = feats
    rows = oracle.cond_rows(feats_all)
    cache = oracle.new_cache(feats_all.shape[0])
    clip_ids = list(range(B)) if clip_ids is None else clip_ids
    all_logits = []
    for offset in range(start, S):
        new = seq[..., cache["len"]:offset]  # columns not yet in the cache (prefill when >1)
        assert not (new == UNKNOWN).any()
        inp = new.repeat(2, 1, 1) if use_cfg else new
        logits = oracle.forward_cached(inp, rows, cache)
        if use_cfg:  # vaura_model.py:810-813
            c, u = logits[:B], logits[B:]
            logits = u + (c - u) * cfg_scale
        if collect_logits:
            all_logits.append(logits.clone())
        if use_sampling and temp > 0.0:  # vaura_model.py:816-823
            probs = filtered_probs(logits, temp, top_k, top_p).numpy()
            nxt = torch.empty((B, K), dtype=torch.long)
            for b in range(B):
                for k in range(K):
                    u01 = philox_uniform(seed, clip_ids[b], offset, k, stream_id)
                    nxt[b, k] = inverse_cdf_draw(probs[b, k], u01)
        else:
            nxt = torch.argmax(logits, dim=-1)  # vaura_model.py:825
        nxt[:, ~mask[:, offset]] = special  # vaura_model.py:536-537
        cur = seq[..., offset]
        seq[..., offset] = torch.where(cur == UNKNOWN, nxt, cur)  # vaura_model.py:540-544
    assert not (seq == UNKNOWN).any()  # vaura_model.py:550
    out = revert_pattern_sequence(seq, T)  # vaura_model.py:560-569
    assert (out >= 0).all() and (out <= special).all()  # vaura_model.py:572
    if collect_logits:
        return out, torch.stack(all_logits)
    return out
