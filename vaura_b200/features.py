"""Host mirror of the reference's Segment-AVCLIP visual feature extractor
(models/modules/feature_extractors/avclip/motionformer.py:47-342) for the shipped configuration
(configs/modules/feature_extractors/avclip_vggsound.yaml): MotionFormer ``divided_224_16x4`` with
``extract_features=True, factorize_space_time=True, agg_space_module="TransformerEncoderLayer",
agg_time_module="torch.nn.Identity", add_global_repr=False``.

The class keeps the reference's name — ``VAURAModel`` checks ``__class__.__name__ == "MotionFormer"``
(models/vaura_model.py:73-75) — its constructor keywords and its output contract ``forward(x) -> (feats, None)`` with
``x (B, S, C, T, H, W)`` normalised RGB segments and ``feats (B, S, t, D)`` (motionformer.py:252-307).  The arithmetic runs
in the sm_100a library (``vaura_avclip_forward``: csrc/avclip.cu + the tcgen05 linears of csrc/gemm_tcgen05.cu); there is no
PyTorch fallback.  Weights come from a state dict with the reference's parameter names (a Stage-I AVCLIP checkpoint's
``v_encoder.*`` entries or the Lightning checkpoint's ``visual_feature_extractor.*``); the reference's checkpoint/config
download (motionformer.py:28-45, :76-131) needs the network and is not mirrored.

Precomputed features ``(B, S, t, D)`` pass through unchanged (the synthetic-feature workloads of BASELINE configs 1-4).
Configurations the reference supports but the shipped model does not use (joint / trajectory attention, average-pooling
aggregation, temporal or global aggregation layers, content masks) raise ``NotImplementedError``.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import typing as tp

import torch

from . import _cabi
from .synthetic import AvclipDims
from .weights import pack_avclip


class MotionFormer(torch.nn.Module):
    def __init__(self, extract_features: bool = False, ckpt_path: tp.Optional[str] = None, factorize_space_time: bool = True,
                 agg_space_module: str = "TransformerEncoderLayer", agg_time_module: str = "torch.nn.Identity",
                 add_global_repr: bool = False, agg_segments_module: tp.Optional[str] = None,
                 max_segments: tp.Optional[int] = None, dims: tp.Optional[AvclipDims] = None, max_chunk_segments: int = 32,
                 **kwargs):
        super().__init__()
        self.config = dict(extract_features=extract_features, ckpt_path=ckpt_path, factorize_space_time=factorize_space_time,
                           agg_space_module=agg_space_module, agg_time_module=agg_time_module,
                           add_global_repr=add_global_repr, agg_segments_module=agg_segments_module,
                           max_segments=max_segments, **kwargs)
        if isinstance(dims, dict):
            dims = AvclipDims(**dims)
        self.dims = dims or AvclipDims()
        self.embed_dim = self.dims.embed_dim
        self.max_chunk_segments = max_chunk_segments
        self._blob = None
        self._offsets = None
        self._handle = None
        self._ws = None
        if ckpt_path is not None and not os.path.exists(str(ckpt_path)):
            # the reference would try to download here (motionformer.py:28-45); the full model's Lightning checkpoint
            # carries `visual_feature_extractor.*` and overrides Stage-I weights anyway (scripts/generate.py:209-211)
            logging.getLogger(__name__).warning(
                "MotionFormer ckpt_path %s does not exist: weights are expected from the model checkpoint", ckpt_path)
        elif ckpt_path is not None:
            ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            sd = ckpt.get("state_dict", ckpt.get("model_state", ckpt))
            # a Stage-I AVCLIP checkpoint holds both towers (motionformer.py:218-227)
            vis = {k.replace("module.", "").replace("v_encoder.", "", 1): v for k, v in sd.items()
                   if k.replace("module.", "").startswith("v_encoder.")}
            self.load_state_dict(vis or sd)

    # ---- weights ----------------------------------------------------------------------------------------------------
    def _check_supported(self):
        c = self.config
        if not c["extract_features"] or not c["factorize_space_time"]:
            raise NotImplementedError("only extract_features=True, factorize_space_time=True is built (the shipped configuration)")
        if c["agg_space_module"] != "TransformerEncoderLayer" or "Identity" not in str(c["agg_time_module"]):
            raise NotImplementedError("only agg_space_module='TransformerEncoderLayer' with agg_time_module=Identity is built")
        if c["add_global_repr"]:
            raise NotImplementedError("add_global_repr=True (global aggregation over segments) is not built")

    def load_state_dict(self, state_dict, strict: bool = True, device=None):
        self._check_supported()
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._blob, self._offsets = pack_avclip(state_dict, self.dims, device)
        self._destroy()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @property
    def has_weights(self) -> bool:
        return self._blob is not None

    def _destroy(self):
        if self._handle is not None:
            _cabi.load().vaura_avclip_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def handle(self):
        if self._blob is None:
            raise RuntimeError("MotionFormer weights are not loaded: raw frames need the visual tower's state dict "
                               "(precomputed AVCLIP features (B, S, t, D) pass through without it)")
        if self._handle is None:
            d = self.dims
            dc = _cabi.AvclipDimsC(d.embed_dim, d.depth, d.num_heads, d.mlp_ratio, d.img_size, d.patch_size, d.in_chans,
                                   d.frames, d.tubelet)
            offs = (C.c_int64 * len(self._offsets))(*self._offsets)
            wc = _cabi.AvclipWeightsC(self._blob.data_ptr(), offs, len(self._offsets))
            h = C.c_void_p()
            _cabi.check(_cabi.load().vaura_avclip_create(C.byref(dc), C.byref(wc), C.byref(h)), "vaura_avclip_create")
            self._handle = h
        return self._handle

    # ---- forward ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, for_loop: bool = False, cont_mask: tp.Optional[torch.Tensor] = None):
        """x (B, S, C, T, H, W) -> ((B, S, t, D), None); segments are independent, so ``for_loop`` (the reference's memory
        knob, motionformer.py:269-283) does not change the result: the library walks them in chunks either way."""
        if x.dim() == 4:  # already AVCLIP features (B, S, t, D)
            return x, None
        if cont_mask is not None:
            raise NotImplementedError("content masks (cont_mask) are not built")
        if x.dim() != 6:
            raise ValueError(f"expected (B, S, C, T, H, W) video segments or (B, S, t, D) features, got {tuple(x.shape)}")
        d = self.dims
        B, S, Cc, T, H, W = x.shape
        if (Cc, T, H, W) != (d.in_chans, d.frames, d.img_size, d.img_size):
            raise ValueError(f"segment shape {(Cc, T, H, W)} does not match the tower's "
                             f"{(d.in_chans, d.frames, d.img_size, d.img_size)}")
        h = self.handle()
        lib = _cabi.load()
        dev = self._blob.device
        out = torch.empty(B, S, d.temporal, d.embed_dim, dtype=torch.float32, device=dev)
        chunk = min(B * S, self.max_chunk_segments)
        nbytes = lib.vaura_avclip_workspace_bytes(h, chunk)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if x.device.type == "cpu" and x.dtype == torch.float32 and x.is_pinned() and x.is_contiguous() and B * S > chunk:
            return self._forward_from_pinned_host(x.view(B * S, Cc, T, H, W), out, chunk, h, lib, dev), None
        frames = x.to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(lib.vaura_avclip_forward(h, frames.data_ptr(), B * S, out.data_ptr(), self._ws.data_ptr(),
                                                 self._ws.numel(), st), "vaura_avclip_forward")
        return out, None

    def _forward_from_pinned_host(self, segs, out, chunk, h, lib, dev):
        """Frames in pinned host memory (2.4 MB of fp32 per segment, 2.5 GB for 64 clips): the host->device copy of chunk k + 1
        runs on a copy stream under the tower's pass over chunk k (two staging buffers, events both ways), so the PCIe time
        is hidden instead of preceding the first kernel."""
        n = segs.shape[0]
        with torch.cuda.device(dev):
            if getattr(self, "_stage", None) is None or self._stage[0].shape[0] < chunk or self._stage[0].device != dev:
                self._stage = [torch.empty(chunk, *segs.shape[1:], dtype=torch.float32, device=dev) for _ in range(2)]
                self._copy_stream = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream()
            flat = out.view(n, out.shape[2], out.shape[3])
            copied, used = [None, None], [None, None]
            self._copy_stream.wait_stream(main)  # staging buffers may still be read by an earlier call
            for k, s0 in enumerate(range(0, n, chunk)):
                ns, b = min(chunk, n - s0), k & 1
                with torch.cuda.stream(self._copy_stream):
                    if used[b] is not None:
                        self._copy_stream.wait_event(used[b])
                    self._stage[b][:ns].copy_(segs[s0:s0 + ns], non_blocking=True)
                    copied[b] = torch.cuda.Event()
                    copied[b].record(self._copy_stream)
                main.wait_event(copied[b])
                _cabi.check(lib.vaura_avclip_forward(h, self._stage[b].data_ptr(), ns, flat[s0:s0 + ns].data_ptr(),
                                                     self._ws.data_ptr(), self._ws.numel(), main.cuda_stream), "vaura_avclip_forward")
                used[b] = torch.cuda.Event()
                used[b].record(main)
        return out
