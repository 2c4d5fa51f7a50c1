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
