"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the DAC token->waveform decode.

The reference calls the pip package ``descript-audio-codec==1.0.0`` (conda_env_cuda12.1.yaml:298)
at models/modules/dac/model.py:41-48: ``z = quantizer.from_codes(codes)``; ``model.decode(z)``.
That package is not vendored in /root/reference and not installed here, so this file restates its
published algorithm (dac/model/dac.py ``Decoder``/``DecoderBlock``/``ResidualUnit``,
dac/nn/layers.py ``Snake1d``, dac/nn/quantize.py ``ResidualVectorQuantize.from_codes``) and is
cross-checked in tests against the architecture-equivalent ``transformers.DacModel`` (5.5.0,
modeling_dac.py:85-99, :173-207, :234-262, :345-369, :405-439).  The reference itself holds no
test or golden vector at this boundary -> parity with dac 1.0.0 proper is UNPINNED; what is pinned
is agreement with transformers' independent statement of the same architecture.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """old-style torch weight_norm, dim=0 (also for ConvTranspose1d, where dim 0 is Cin)."""
    dims = tuple(range(1, v.dim()))
    return g * v / v.pow(2).sum(dim=dims, keepdim=True).sqrt()


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """dac/nn/layers.py snake(): x + (alpha + 1e-9)^-1 * sin(alpha x)^2, alpha (1,C,1)."""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


class DacDecodeOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], cdims, dtype=torch.float32):
        self.c = cdims
        self.dtype = dtype
        self.sd = {k: v.to(dtype) for k, v in sd.items()}

    def _w(self, key):
        return fold_weight_norm(self.sd[key + ".weight_g"].float(), self.sd[key + ".weight_v"].float()).to(self.dtype)

    def from_codes(self, codes: torch.Tensor) -> torch.Tensor:
        """codes (B,Kc,T) int64 -> z (B,latent,T): sum_k out_proj_k(codebook_k[codes_k])."""
        z = 0.0
        for k in range(codes.shape[1]):
            p = f"quantizer.quantizers.{k}"
            e = F.embedding(codes[:, k], self.sd[f"{p}.codebook.weight"]).transpose(1, 2)
            z = z + F.conv1d(e, self._w(f"{p}.out_proj"), self.sd[f"{p}.out_proj.bias"])
        return z

    def decode_latent(self, z: torch.Tensor) -> torch.Tensor:
        sd, c = self.sd, self.c
        x = F.conv1d(z, self._w("decoder.model.0"), sd["decoder.model.0.bias"], padding=3)
        for i, s in enumerate(c.decoder_rates):
            p = f"decoder.model.{i + 1}.block"
            x = snake(x, sd[f"{p}.0.alpha"])
            x = F.conv_transpose1d(x, self._w(f"{p}.1"), sd[f"{p}.1.bias"], stride=s, padding=math.ceil(s / 2))
            for j, dil in enumerate((1, 3, 9)):
                q = f"{p}.{2 + j}.block"
                y = snake(x, sd[f"{q}.0.alpha"])
                y = F.conv1d(y, self._w(f"{q}.1"), sd[f"{q}.1.bias"], dilation=dil, padding=((7 - 1) * dil) // 2)
                y = snake(y, sd[f"{q}.2.alpha"])
                y = F.conv1d(y, self._w(f"{q}.3"), sd[f"{q}.3.bias"])
                x = x + y
        n = len(c.decoder_rates)
        x = snake(x, sd[f"decoder.model.{n + 1}.alpha"])
        x = F.conv1d(x, self._w(f"decoder.model.{n + 2}"), sd[f"decoder.model.{n + 2}.bias"], padding=3)
        return torch.tanh(x)

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes (B,Kc,T) -> waveform (B,1,hop*T)  (models/modules/dac/model.py:41-48)."""
        return self.decode_latent(self.from_codes(codes).to(self.dtype))


class DacEncodeOracle:
    """CPU restatement of the wav -> codes direction the reference reaches through ``DacModelWrapper.encode``
    (models/modules/dac/model.py:30-39: ``model.preprocess(wav, sr)``; ``_, codes, _, _, _ = model.encode(wav)``), i.e. of
    dac 1.0.0's ``DAC.preprocess`` (right-pad to a multiple of the hop length), ``Encoder`` / ``EncoderBlock`` /
    ``ResidualUnit`` (dac/model/dac.py) and ``ResidualVectorQuantize.forward`` / ``VectorQuantize.decode_latents``
    (dac/nn/quantize.py: factorised, l2-normalised nearest-neighbour look-up; the residual passed on is
    ``residual - out_proj(codebook[idx])``).  Cross-checked against ``transformers.DacModel.encode`` (modeling_dac.py:102-171,
    :210-232, :265-343, :442-473) in tests/test_oracle_golden.py; parity with dac 1.0.0 proper is UNPINNED (see the header)."""

    def __init__(self, sd: Dict[str, torch.Tensor], cdims, dtype=torch.float32):
        self.c = cdims
        self.dtype = dtype
        self.sd = {k: v.to(dtype) for k, v in sd.items()}

    def _w(self, key):
        return fold_weight_norm(self.sd[key + ".weight_g"].float(), self.sd[key + ".weight_v"].float()).to(self.dtype)

    def preprocess(self, wav: torch.Tensor) -> torch.Tensor:
        hop = self.c.hop_length
        length = wav.shape[-1]
        right = math.ceil(length / hop) * hop - length
        return F.pad(wav, (0, right))

    def encode_latent(self, wav: torch.Tensor) -> torch.Tensor:
        """wav (B,1,L) -> z (B,latent,L/hop)."""
        sd, c = self.sd, self.c
        x = F.conv1d(wav, self._w("encoder.block.0"), sd["encoder.block.0.bias"], padding=3)
        for i, s in enumerate(c.encoder_rates):
            p = f"encoder.block.{i + 1}.block"
            for j, dil in enumerate((1, 3, 9)):
                q = f"{p}.{j}.block"
                y = snake(x, sd[f"{q}.0.alpha"])
                y = F.conv1d(y, self._w(f"{q}.1"), sd[f"{q}.1.bias"], dilation=dil, padding=((7 - 1) * dil) // 2)
                y = snake(y, sd[f"{q}.2.alpha"])
                y = F.conv1d(y, self._w(f"{q}.3"), sd[f"{q}.3.bias"])
                x = x + y
            x = snake(x, sd[f"{p}.3.alpha"])
            x = F.conv1d(x, self._w(f"{p}.4"), sd[f"{p}.4.bias"], stride=s, padding=math.ceil(s / 2))
        n = len(c.encoder_rates)
        x = snake(x, sd[f"encoder.block.{n + 1}.alpha"])
        return F.conv1d(x, self._w(f"encoder.block.{n + 2}"), sd[f"encoder.block.{n + 2}.bias"], padding=1)

    def quantize(self, z: torch.Tensor, return_margins: bool = False):
        """z (B,latent,T) -> codes (B,Kc,T) int64 (+ per-cell top-2 similarity gaps when asked)."""
        sd = self.sd
        residual = z
        codes, margins = [], []
        for k in range(self.c.n_codebooks):
            p = f"quantizer.quantizers.{k}"
            e = F.conv1d(residual, self._w(f"{p}.in_proj"), sd[f"{p}.in_proj.bias"])          # (B, 8, T)
            B, Dc, T = e.shape
            enc = F.normalize(e.permute(0, 2, 1).reshape(B * T, Dc))
            cb = F.normalize(sd[f"{p}.codebook.weight"])
            dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cb.t() + cb.pow(2).sum(1, keepdim=True).t()
            idx = (-dist).max(1)[1].reshape(B, T)
            if return_margins:
                top2 = torch.topk(-dist, 2, dim=1).values
                margins.append((top2[:, 0] - top2[:, 1]).reshape(B, T))
            zq = F.embedding(idx, sd[f"{p}.codebook.weight"]).transpose(1, 2)
            zq = F.conv1d(zq, self._w(f"{p}.out_proj"), sd[f"{p}.out_proj.bias"])
            residual = residual - zq
            codes.append(idx)
        codes = torch.stack(codes, dim=1)
        return (codes, torch.stack(margins, dim=1)) if return_margins else codes

    @torch.no_grad()
    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        """wav (L,) | (C,L) | (B,1,L) -> codes (B,Kc,ceil(L/hop))  (models/modules/dac/model.py:30-39)."""
        if wav.ndim < 2:
            wav = wav.unsqueeze(0)
        if wav.ndim < 3:
            wav = wav.unsqueeze(0)
        return self.quantize(self.encode_latent(self.preprocess(wav.to(self.dtype))))
