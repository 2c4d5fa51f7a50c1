"""Host mirror of the reference AR transformer ("sampler") over the C ABI.

Same constructor keywords as ``models.modules.sampler.llama.Transformer`` (llama.py:286-306, fed from
configs/modules/samplers/llama_9cbs.yaml), same attributes callers poke and the same
``forward(tgt, memory, ...) -> (logits (B,K,T,V), None, None)`` contract (llama.py:520-539).  All
arithmetic happens in ``libvaura_b200.so``; torch only owns device buffers and the stream.
"""
from __future__ import annotations

import ctypes as C
import os
from math import ceil
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _cabi
from .synthetic import SamplerDims, find_multiple
from .weights import pack_sampler

PAGE_SIZE = int(os.environ.get("VAURA_PAGE_SIZE", "32"))


def resolve_precision(precision: int, rows: int, sampling: bool = False) -> int:
    """Same rule as csrc/cabi.cu: AUTO -> tcgen05/bf16 path from 16 sequence rows, and from 3 rows when the call samples
    (top-k / top-p / temperature draws have no bit-exactness contract); fp32-activation path otherwise.

    At 3...15 rows the bf16 path (one fused kernel per step, ~1 ms) is 1.2-3x faster than the fp32-activation paths
    (1.2-2.9 ms, profiles/scripts/rows_sweep.py) at bf16 tolerance instead of bit-exact greedy tokens; at 1-2 rows the
    fp32-activation cluster kernel is both exact and 3x faster.  ``VAURA_PRECISION=bf16|fp32`` overrides AUTO."""
    if precision != _cabi.PRECISION_AUTO:
        return precision
    env = os.environ.get("VAURA_PRECISION", "").lower()
    if env in ("bf16", "fp32"):
        return _cabi.PRECISION_BF16 if env == "bf16" else _cabi.PRECISION_FP32ACT
    return _cabi.PRECISION_BF16 if (rows >= 16 or (rows >= 3 and sampling)) else _cabi.PRECISION_FP32ACT


class _CondEmbedder:
    """Stand-in for ``AVCLIPEmbedder`` exposing what callers read (vaura_model.py:790-793)."""

    def __init__(self, token_num: int = 32, in_channels: int = 768):
        self.token_num = token_num
        self.in_channels = in_channels
        self.uncond_embedding: Optional[torch.Tensor] = None


class Transformer(torch.nn.Module):
    def __init__(self, num_layers: int = 12, d_model: int = 512, d_codebook: int = 1024, block_size_audio: int = 512,
                 block_size_video: int = 64, nhead: int = 8, dim_feedforward: int = 2048, dropout: float = 0.1,
                 activation: str = "relu", layer_norm_eps: float = 1e-5, batch_first: bool = False,
                 norm_first: bool = False, num_codebooks: int = 2, positional_embedder: str = "sinusoidal",
                 use_visual_conditioning: bool = True, use_delay_strategy: bool = False,
                 cond_feature_channel_scaler: int = 2):
        super().__init__()
        # dim_feedforward is ignored by the reference as well: the SwiGLU width comes from llama.py:164-169
        self.dims = SamplerDims(num_layers=num_layers, d_model=d_model, nhead=nhead, d_codebook=d_codebook,
                                num_codebooks=num_codebooks, block_size=max(block_size_audio, block_size_video),
                                cond_feature_channel_scaler=cond_feature_channel_scaler, norm_eps=layer_norm_eps)
        self.config = SimpleNamespace(dim=d_model, n_layer=num_layers, n_head=nhead, norm_eps=layer_norm_eps,
                                      block_size=self.dims.block_size, vocab_size=d_codebook, rope_base=10000,
                                      initializer_range=0.02)
        self.vocab_size = d_codebook
        self.n_layer = num_layers
        self.block_size = self.dims.block_size
        self.num_codebooks = num_codebooks
        self.d_codebook = d_codebook
        self.use_visual_conditioning = use_visual_conditioning
        self.audio_tokens_per_video_frame: Optional[int] = None
        self.codebook_pattern: Optional[str] = None
        self.cls_embeddings = _CondEmbedder(self.dims.cond_tokens, self.dims.cond_in)
        self.weights: Optional[Dict[str, torch.Tensor]] = None
        self._handle = None
        self._handle_atpvf = None
        self._buffers: Dict[str, torch.Tensor] = {}

    # ---- reference API surface ------------------------------------------------------------------
    def initialize_embeddings(self, dac_model=None):
        """llama.py:387-412 swaps the embedding tables for DAC-shaped ones; here the folded tables are
        built from the checkpoint's ``tok_embeddings.{k}.emb / out_proj`` entries in load_state_dict."""
        return None

    def _set_audio_tokens_per_video_frame(self, Ta: int, Tv: int):
        # llama.py:544-553
        pat = (self.codebook_pattern or "delayed").lower()
        Ta = Ta - self.num_codebooks if "delayed" in pat else Ta - 1
        atpvf = ceil(Ta / Tv)
        if atpvf < 1:
            # the reference goes on to `range(0, Ta, 0)` / a negative stride and raises on the host (llama.py:555-586);
            # the kernels divide the position by this value
            raise ValueError(f"audio_tokens_per_video_frame = ceil({Ta} / {Tv}) = {atpvf} must be >= 1: set "
                             "sampler.audio_tokens_per_video_frame explicitly (scripts/generate.py:216 sets 7)")
        self.audio_tokens_per_video_frame = atpvf

    def load_state_dict(self, state_dict, strict: bool = True, device=None):
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.weights = pack_sampler(state_dict, self.dims, device)
        self.cls_embeddings.uncond_embedding = self.weights["uncond_embedding"]
        self._destroy()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @property
    def device(self):
        if self.weights is None:
            raise RuntimeError("sampler weights are not loaded")
        return self.weights["wqkv"].device

    # ---- handle / buffers -----------------------------------------------------------------------
    def _destroy(self):
        if self._handle is not None:
            _cabi.load().vaura_sampler_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def last_loop_ms(self):
        """(ms, steps) of the decode-step launches of the last generate call: CUDA events recorded inside
        ``vaura_sampler_generate`` around the step launches alone (include/vaura_b200.h: vaura_sampler_last_loop_ms)."""
        ms, steps = C.c_float(), C.c_int32()
        _cabi.check(_cabi.load().vaura_sampler_last_loop_ms(self.handle(), C.byref(ms), C.byref(steps)), "vaura_sampler_last_loop_ms")
        return float(ms.value), int(steps.value)

    def handle(self):
        if self.weights is None:
            raise RuntimeError("sampler weights are not loaded (load_state_dict / load_from_checkpoint first)")
        if self.audio_tokens_per_video_frame is None:
            raise RuntimeError("sampler.audio_tokens_per_video_frame is not set (scripts/generate.py:216 sets 7)")
        if int(self.audio_tokens_per_video_frame) < 1:
            raise ValueError(f"sampler.audio_tokens_per_video_frame = {self.audio_tokens_per_video_frame} must be >= 1")
        if self._handle is not None and self._handle_atpvf == self.audio_tokens_per_video_frame:
            return self._handle
        self._destroy()
        lib = _cabi.load()
        d = self.dims
        dc = _cabi.SamplerDimsC(d.num_layers, d.d_model, d.nhead, d.ffn_dim, d.d_codebook, d.num_codebooks, d.block_size,
                                d.cond_dim, d.cond_in, d.cond_tokens, int(self.audio_tokens_per_video_frame), d.norm_eps)
        w = self.weights
        wc = _cabi.SamplerWeightsC(*[w[n].data_ptr() if n in w else None for n in (
            "wqkv", "wo", "w13", "w2", "w_heads", "attn_norm", "ffn_norm", "final_norm", "tok_tables", "rope", "fc1",
            "fc2", "empty_video_emb", "wstream")])
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(lib.vaura_sampler_create(C.byref(dc), C.byref(wc), C.byref(h)), "vaura_sampler_create")
        self._handle, self._handle_atpvf = h, self.audio_tokens_per_video_frame
        return h

    def _buffer(self, name: str, nbytes: int) -> torch.Tensor:
        buf = self._buffers.get(name)
        if buf is None or buf.numel() < nbytes or buf.device != self.device:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._buffers[name] = buf
        return buf

    def _kv(self, rows: int, dtype_code: int = _cabi.KV_F32):
        d = self.dims
        pages_per_seq = (d.block_size + PAGE_SIZE - 1) // PAGE_SIZE
        num_pages = rows * pages_per_seq
        esize = 4 if dtype_code == _cabi.KV_F32 else 2
        nbytes = d.num_layers * 2 * num_pages * d.nhead * PAGE_SIZE * d.head_dim * esize
        pages = self._buffer(f"kv{dtype_code}", nbytes)
        key = f"pt{rows}"
        pt = self._buffers.get(key)
        if pt is None or pt.device != self.device:
            # identity allocation: sequence row r owns pages [r*pps, (r+1)*pps).  The indirection is what the
            # kernels consume, so a caller may hand in any other assignment.
            pt = torch.arange(num_pages, dtype=torch.int32, device=self.device).reshape(rows, pages_per_seq).contiguous()
            self._buffers[key] = pt
        kv = _cabi.KvCacheC(pages.data_ptr(), pt.data_ptr(), num_pages, PAGE_SIZE, pages_per_seq, dtype_code)
        return kv

    # ---- compute entry points -------------------------------------------------------------------
    def cond_rows(self, memory: torch.Tensor) -> torch.Tensor:
        """memory (rows, Tv, 768) -> (rows, Tv+1, cond_dim): MLP rows + the empty_video_emb row."""
        lib = _cabi.load()
        memory = memory.to(device=self.device, dtype=torch.float32).contiguous()
        rows, tv, cin = memory.shape
        if cin != self.dims.cond_in:
            raise ValueError(f"conditioning width {cin} != {self.dims.cond_in}")
        if tv > self.dims.cond_tokens:
            raise ValueError(f"{tv} visual tokens > {self.dims.cond_tokens}: the conditioning table is built for at most "
                             f"{self.dims.cond_tokens} rows plus the empty_video_emb row")
        out = torch.empty(rows, tv + 1, self.dims.cond_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(lib.vaura_sampler_cond_project(self.handle(), memory.data_ptr(), rows, tv, out.data_ptr(), st),
                        "vaura_sampler_cond_project")
        if tv < self.dims.cond_tokens:
            # fewer visual tokens than table rows (short last window): positions whose frame index is >= Tv read
            # empty_video_emb (llama.py:569-572), so rows [Tv, cond_tokens] all hold that vector
            pad = out[:, tv:].expand(rows, self.dims.cond_tokens + 1 - tv, self.dims.cond_dim)
            out = torch.cat([out[:, :tv], pad], dim=1).contiguous()
        return out

    def forward(self, tgt: torch.Tensor, memory: torch.Tensor, use_conditioning: bool = True, tgt_mask=None,
                memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, tgt_is_causal: bool = False,
                memory_is_causal: bool = False, return_attention_weights: bool = False,
                apply_per_video_frame_mask: bool = False, precision: int = _cabi.PRECISION_AUTO):
        if memory is None or tgt is None:
            raise Exception("Not implemented")  # llama.py:475-477
        if tgt_mask is not None:
            raise NotImplementedError("explicit attention masks are not on the generation path (causal only)")
        lib = _cabi.load()
        B, K, S = tgt.shape
        if self.audio_tokens_per_video_frame is None:
            self._set_audio_tokens_per_video_frame(S, memory.shape[1])
        rows = self.cond_rows(memory)
        seq = tgt.to(device=self.device, dtype=torch.int32).contiguous()
        logits = torch.empty(B, K, S, self.dims.d_codebook, dtype=torch.float32, device=self.device)
        h = self.handle()
        precision = resolve_precision(precision, B)
        nbytes = lib.vaura_sampler_workspace_bytes(h, B, S, precision)
        ws = self._buffer("ws", nbytes)
        kv = self._kv(B, _cabi.KV_BF16 if precision == _cabi.PRECISION_BF16 else _cabi.KV_F32)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(lib.vaura_sampler_forward(h, seq.data_ptr(), rows.data_ptr(), B, S, logits.data_ptr(), C.byref(kv),
                                                  precision, ws.data_ptr(), ws.numel(), st), "vaura_sampler_forward")
        return logits, None, None

    def generate_tokens(self, sequence: torch.Tensor, cond_rows: torch.Tensor, timesteps: int, start_offset: int,
                        use_cfg: bool, cfg_scale: float, use_sampling: bool, temp: float, top_k: int, top_p: float,
                        seed: int = 0, clip_ids: Optional[torch.Tensor] = None, logits_out: Optional[torch.Tensor] = None,
                        end_offset: Optional[int] = None, precision: int = _cabi.PRECISION_AUTO,
                        stream_id: int = 0) -> torch.Tensor:
        """Run the fused decode loop in place on ``sequence`` (B,K,S) int32 (-1 = to be generated)."""
        lib = _cabi.load()
        assert sequence.dtype == torch.int32 and sequence.is_contiguous() and sequence.device == self.device
        B, K, S = sequence.shape
        rows = B * (2 if use_cfg else 1)
        assert cond_rows.shape[0] == rows and cond_rows.is_contiguous()
        h = self.handle()
        precision = resolve_precision(precision, rows, bool(use_sampling) and float(temp) > 0.0)
        nbytes = lib.vaura_sampler_workspace_bytes(h, rows, max(start_offset, 1), precision)
        ws = self._buffer("ws", nbytes)
        kv = self._kv(rows, _cabi.KV_BF16 if precision == _cabi.PRECISION_BF16 else _cabi.KV_F32)
        p = _cabi.GenerateParamsC(
            batch=B, use_cfg=int(use_cfg), timesteps=timesteps, start_offset=start_offset,
            end_offset=S if end_offset is None else end_offset, use_sampling=int(use_sampling), temp=float(temp),
            top_k=int(top_k), top_p=float(top_p), cfg_scale=float(cfg_scale), seed=int(seed) & (2 ** 64 - 1),
            clip_ids=clip_ids.data_ptr() if clip_ids is not None else None, sequence=sequence.data_ptr(),
            cond_rows=cond_rows.data_ptr(), logits_out=logits_out.data_ptr() if logits_out is not None else None,
            precision=precision, stream_id=int(stream_id) & 0xFFFFFFFF)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(lib.vaura_sampler_generate(h, C.byref(p), C.byref(kv), ws.data_ptr(), ws.numel(), st),
                        "vaura_sampler_generate")
        return sequence


def sample_logits(logits: torch.Tensor, use_cfg: bool = False, cfg_scale: float = 1.0, use_sampling: bool = True,
                  temp: float = 1.0, top_k: int = 0, top_p: float = 0.0, seed: int = 0, offset: int = 0,
                  clip_ids: Optional[torch.Tensor] = None, return_probs: bool = False):
    """Sampling stage alone (utils/utils.py:139-196 + vaura_model.py:810-825).  logits (rows_eff,K,V) f32 cuda."""
    lib = _cabi.load()
    logits = logits.contiguous().float()
    rows_eff, K, V = logits.shape
    rows = rows_eff // 2 if use_cfg else rows_eff
    tokens = torch.empty(rows, K, dtype=torch.int32, device=logits.device)
    probs = torch.empty(rows, K, V, dtype=torch.float32, device=logits.device) if return_probs else None
    with torch.cuda.device(logits.device):
        st = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.vaura_sample_logits(logits.data_ptr(), rows, K, V, int(use_cfg), float(cfg_scale), int(use_sampling),
                                            float(temp), int(top_k), float(top_p), int(seed) & (2 ** 64 - 1),
                                            clip_ids.data_ptr() if clip_ids is not None else None, int(offset),
                                            tokens.data_ptr(), probs.data_ptr() if probs is not None else None, st),
                    "vaura_sample_logits")
    return (tokens, probs) if return_probs else tokens
