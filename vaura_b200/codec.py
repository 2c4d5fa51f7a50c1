"""Host mirror of ``DacModelWrapper`` (models/modules/dac/model.py:11-61) for the decode direction.

The class keeps the reference's name (checked at models/vaura_model.py:87) and call shapes:
``decode(codes | [(codes, None)]) -> (B, 1, hop*T) float16`` (fp16 because the reference halves the
codec, vaura_model.py:92).  Weights come from the Lightning checkpoint's ``audio_encoder.model.*``
entries (dac 1.0.0 names); the reference's ``dac.utils.download`` needs the network and is not mirrored.
``encode`` (wav -> codes, SURVEY §8f row 3: raw-audio prompts and ``compress_original_audio``, scripts/generate.py:286-301)
runs the DAC encoder + residual vector quantisation when the checkpoint carries the encode half.
"""
from __future__ import annotations

import ctypes as C
import typing as tp

import torch

from . import _cabi
from .synthetic import CodecDims
from .weights import pack_codec, pack_codec_encoder

MODEL_SR = [16000, 24000, 44000, 44100]


class DacModelWrapper(torch.nn.Module):
    def __init__(self, model_sr: int = 24000, ckpt_path: tp.Optional[str] = None, dims: tp.Optional[CodecDims] = None):
        super().__init__()
        assert model_sr in MODEL_SR, "Invalid model samplerate"
        self.model_sr = model_sr
        if isinstance(dims, dict):  # YAML-borne dims (hparams.yaml cannot carry a dataclass)
            dims = CodecDims(**{k: tuple(v) if isinstance(v, list) else v for k, v in dims.items()})
        self._dims_given = dims is not None
        self.dims = dims or CodecDims(sample_rate=model_sr)
        self._blob = None
        self._offsets = None
        self._handle = None
        self._ws = None
        self._enc_blob = None
        self._enc_offsets = None
        self._enc_handle = None
        self._enc_ws = None
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            self.load_state_dict(sd.get("state_dict", sd))

    @property
    def model(self):
        """Callers reach for ``.model.half()`` (vaura_model.py:92); the wrapper and the codec are one object here."""
        return self

    def half(self):  # the compute path already stores fp16 (vaura_model.py:92)
        return self

    @staticmethod
    def dims_from_state_dict(sd, sample_rate: int) -> CodecDims:
        """Shape parameters read off a dac state dict (the reference gets them from the downloaded model's metadata,
        models/modules/dac/model.py:23-25): conv_in (decoder_dim, latent, 7), ConvTranspose1d kernels (Cin, Cout, 2*rate),
        one quantizer entry per codebook."""
        def shape(key):
            for suffix in (".weight_v", ".weight"):
                if key + suffix in sd:
                    return tuple(sd[key + suffix].shape)
            raise KeyError(key)
        decoder_dim, latent, _ = shape("decoder.model.0")
        rates, i = [], 1
        while any(k.startswith(f"decoder.model.{i}.block.1.") for k in sd):
            rates.append(shape(f"decoder.model.{i}.block.1")[2] // 2)
            i += 1
        n_q = 0
        while f"quantizer.quantizers.{n_q}.codebook.weight" in sd:
            n_q += 1
        size, cdim = sd["quantizer.quantizers.0.codebook.weight"].shape
        extra = {}
        if any(k.startswith("encoder.block.0.") for k in sd):  # encode half present: its width is dac's `encoder_dim`
            extra["encoder_dim"] = int(shape("encoder.block.0")[0])
        return CodecDims(latent_dim=latent, decoder_dim=decoder_dim, decoder_rates=tuple(rates), n_codebooks=n_q,
                         codebook_size=size, codebook_dim=cdim, sample_rate=sample_rate, **extra)

    def load_state_dict(self, state_dict, strict: bool = True, device=None):
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if not self._dims_given:
            self.dims = self.dims_from_state_dict(state_dict, self.model_sr)
        self._blob, self._offsets = pack_codec(state_dict, self.dims, device)
        if "encoder.block.0.weight_v" in state_dict or "encoder.block.0.weight" in state_dict:
            # the encoder width is read off its first convolution whatever the YAML said (dac `encoder_dim`)
            enc_dim = self.dims_from_state_dict(state_dict, self.model_sr).encoder_dim
            if enc_dim != self.dims.encoder_dim:
                self.dims = CodecDims(**{**self.dims.__dict__, "encoder_dim": enc_dim})
            self._enc_blob, self._enc_offsets = pack_codec_encoder(state_dict, self.dims, device)
        self._destroy()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def _destroy(self):
        if self._handle is not None:
            _cabi.load().vaura_codec_destroy(self._handle)
            self._handle = None
        if self._enc_handle is not None:
            _cabi.load().vaura_codec_encoder_destroy(self._enc_handle)
            self._enc_handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def handle(self):
        if self._blob is None:
            raise RuntimeError("codec weights are not loaded")
        if self._handle is None:
            lib = _cabi.load()
            d = self.dims
            rates = (C.c_int32 * 8)(*d.decoder_rates)
            dc = _cabi.CodecDimsC(d.latent_dim, d.decoder_dim, len(d.decoder_rates), rates, d.n_codebooks, d.codebook_size)
            offs = (C.c_int64 * len(self._offsets))(*self._offsets)
            wc = _cabi.CodecWeightsC(self._blob.data_ptr(), offs, len(self._offsets))
            h = C.c_void_p()
            _cabi.check(lib.vaura_codec_create(C.byref(dc), C.byref(wc), C.byref(h)), "vaura_codec_create")
            self._handle = h
        return self._handle

    def forward(self, wav: torch.Tensor):
        return self.encode(wav)

    def preprocess(self, wav: torch.Tensor, sample_rate: tp.Optional[int] = None) -> torch.Tensor:
        """dac 1.0.0 ``DAC.preprocess``: zero-pad on the right to a multiple of the hop length."""
        assert sample_rate is None or sample_rate == self.dims.sample_rate or sample_rate == self.model_sr
        hop = self.dims.hop_length
        right = -wav.shape[-1] % hop
        return torch.nn.functional.pad(wav, (0, right)) if right else wav

    def encoder_handle(self):
        if self._enc_blob is None:
            raise RuntimeError("the checkpoint held no `encoder.*` / `quantizer.*.in_proj` entries: encode() needs the "
                               "encode half of the DAC weights")
        if self._enc_handle is None:
            lib = _cabi.load()
            d = self.dims
            rates = (C.c_int32 * 8)(*d.decoder_rates)
            dc = _cabi.CodecDimsC(d.latent_dim, d.decoder_dim, len(d.decoder_rates), rates, d.n_codebooks, d.codebook_size)
            offs = (C.c_int64 * len(self._enc_offsets))(*self._enc_offsets)
            wc = _cabi.CodecWeightsC(self._enc_blob.data_ptr(), offs, len(self._enc_offsets))
            h = C.c_void_p()
            _cabi.check(lib.vaura_codec_encoder_create(C.byref(dc), d.encoder_dim, d.codebook_dim, C.byref(wc), C.byref(h)),
                        "vaura_codec_encoder_create")
            self._enc_handle = h
        return self._enc_handle

    @torch.no_grad()
    def encode(self, wav: torch.Tensor, max_batch: int = 8, _return_latent: bool = False):
        """wav (L,) | (C, L) | (B, 1, L) -> codes (B, n_codebooks, ceil(L / hop)) int64
        (models/modules/dac/model.py:30-39: unsqueeze to 3 dims, ``model.preprocess``, codes of ``model.encode``)."""
        if wav.ndim < 2:
            wav = wav.unsqueeze(0)
        if wav.ndim < 3:
            wav = wav.unsqueeze(0)
        if wav.shape[1] != 1:
            raise ValueError(f"expected mono audio (B, 1, L), got {tuple(wav.shape)}")
        h = self.encoder_handle()
        lib = _cabi.load()
        dev = self._enc_blob.device
        x = self.preprocess(wav.to(device=dev, dtype=torch.float32), self.model_sr)[:, 0].contiguous()
        B, L = x.shape
        T = L // self.dims.hop_length
        codes = torch.empty(B, self.dims.n_codebooks, T, dtype=torch.int32, device=dev)
        latent = torch.empty(B, T, self.dims.latent_dim, dtype=torch.float16, device=dev) if _return_latent else None
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            for b0 in range(0, B, max_batch):
                nb = min(max_batch, B - b0)
                nbytes = lib.vaura_codec_encoder_workspace_bytes(h, nb, L)
                if self._enc_ws is None or self._enc_ws.numel() < nbytes:
                    self._enc_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _cabi.check(lib.vaura_codec_encode(h, x[b0:b0 + nb].data_ptr(), nb, L, codes[b0:b0 + nb].data_ptr(),
                                                   latent[b0:b0 + nb].data_ptr() if latent is not None else None,
                                                   self._enc_ws.data_ptr(), self._enc_ws.numel(), st), "vaura_codec_encode")
        codes = codes.to(torch.int64)
        return (codes, latent) if _return_latent else codes

    @torch.no_grad()
    def decode(self, codes: tp.Union[torch.Tensor, tp.List[tp.Tuple[torch.Tensor, tp.Any]]], max_batch: int = 16,
               validate: bool = True):
        """``validate``: range-check the codes on the host first (F.embedding raises IndexError in the reference); it
        costs a device->host synchronisation, so ``VAURAModel.generate`` - whose codes come from the sampling stage and
        are in range by construction - turns it off.  The kernel clamps indices either way (no out-of-bounds read)."""
        if type(codes) == list:  # EnCodec-style frames (models/modules/dac/model.py:43-44)
            codes = codes[0][0]
        lib = _cabi.load()
        dev = self._blob.device
        codes = codes.to(device=dev, dtype=torch.int32).contiguous()
        B, Kc, T = codes.shape
        if Kc != self.dims.n_codebooks:
            raise ValueError(f"expected {self.dims.n_codebooks} codebooks, got {Kc}")
        if validate and codes.numel() and (int(codes.min()) < 0 or int(codes.max()) >= self.dims.codebook_size):
            raise IndexError("codec code out of range")  # F.embedding would raise in the reference
        hop = self.dims.hop_length
        wav = torch.empty(B, 1, T * hop, dtype=torch.float16, device=dev)
        h = self.handle()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            for b0 in range(0, B, max_batch):  # bound the activation workspace
                nb = min(max_batch, B - b0)
                nbytes = lib.vaura_codec_workspace_bytes(h, nb, T)
                if self._ws is None or self._ws.numel() < nbytes:
                    self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _cabi.check(lib.vaura_codec_decode(h, codes[b0:b0 + nb].data_ptr(), nb, T, wav[b0:b0 + nb].data_ptr(),
                                                   self._ws.data_ptr(), self._ws.numel(), st), "vaura_codec_decode")
        return wav

    @property
    def sample_rate(self):
        return self.dims.sample_rate

    @property
    def channels(self):
        return 1

    @property
    def frame_rate(self):
        return None
