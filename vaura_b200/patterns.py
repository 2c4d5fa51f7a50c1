"""Delay-pattern codebook interleaving in closed form.

Drop-in for the subset of the reference's ``Pattern`` / ``DelayedPatternProvider`` that the
generation path uses (models/modules/misc/codebook_patterns.py:131-135, :180-207, :260-285,
:350-406).  For delays = [0..n_q-1] the layout is: column 0 = special (BOS); codebook k at
column s holds timestep s-1-k; valid iff 0 <= s-1-k < T.  No index tables are needed, so the
device code and this host mirror share the same two formulas.
Other providers (parallel, unrolled, VALL-E, MusicLM) are not used by the shipped configs and
are out of scope.
"""
from __future__ import annotations

from functools import lru_cache
from typing import List, Optional, Tuple

import torch


class Pattern:
    def __init__(self, n_q: int, timesteps: int, delays: List[int]):
        self.n_q = n_q
        self.timesteps = timesteps
        self.delays = list(delays)

    @property
    def num_sequence_steps(self) -> int:
        return self.timesteps + max(self.delays)

    @property
    def max_delay(self) -> int:
        return max(self.delays)

    def _t_index(self, S: int, device) -> torch.Tensor:
        s = torch.arange(S, device=device)[None, :]
        d = torch.tensor(self.delays, device=device)[:, None]
        return s - 1 - d  # (K,S) timestep held by column s of codebook k

    def mask(self, timesteps: Optional[int] = None, device="cpu") -> torch.Tensor:
        T = self.timesteps if timesteps is None else timesteps
        t = self._t_index(self.num_sequence_steps + 1, device)
        return (t >= 0) & (t < T)

    def get_first_step_with_timesteps(self, t: int, q: Optional[int] = None) -> Optional[int]:
        # codebook_patterns.py:131-135: first column holding timestep t (of codebook q, or any)
        if t >= self.timesteps or t < 0:
            return None
        d = self.delays[q] if q is not None else min(self.delays)
        return t + 1 + d

    def build_pattern_sequence(self, z: torch.Tensor, special_token: int, keep_only_valid_steps: bool = False):
        """z (B,K,T) -> (values (B,K,S), indexes (K,S), mask (K,S)); codebook_patterns.py:180-207."""
        assert not keep_only_valid_steps, "only the generation-path variant is implemented"
        B, K, T = z.shape
        assert K == self.n_q and T <= self.timesteps
        S = self.num_sequence_steps + 1
        t = self._t_index(S, z.device)
        mask = (t >= 0) & (t < T)
        if T == 0:
            values = torch.full((B, K, S), special_token, dtype=z.dtype, device=z.device)
        else:
            values = torch.gather(z, 2, t.clamp(0, T - 1)[None].expand(B, -1, -1))
            values = torch.where(mask[None], values, torch.full_like(values, special_token))
        k = torch.arange(K, device=z.device)[:, None]
        indexes = torch.where(mask, t + k * T, torch.full_like(t, K * T))
        return values, indexes, mask

    def revert_pattern_sequence(self, s: torch.Tensor, special_token: int, keep_only_valid_steps: bool = False):
        """s (B,K,S) -> (values (B,K,T), indexes (K,T), mask (K,T)); codebook_patterns.py:260-285."""
        assert not keep_only_valid_steps
        B, K, S = s.shape
        T = self.timesteps
        d = torch.tensor(self.delays, device=s.device)[:, None]
        col = torch.arange(T, device=s.device)[None, :] + 1 + d  # (K,T)
        mask = col < S
        values = torch.gather(s, 2, col.clamp(max=S - 1)[None].expand(B, -1, -1))
        values = torch.where(mask[None], values, torch.full_like(values, special_token))
        k = torch.arange(K, device=s.device)[:, None]
        indexes = torch.where(mask, col + k * S, torch.full_like(col, K * S))
        return values, indexes, mask


class DelayedPatternProvider:
    """codebook_patterns.py:350-406 with the arguments the shipped config uses (n_q only)."""

    def __init__(self, n_q: int, delays: Optional[List[int]] = None, flatten_first: int = 0, empty_initial: int = 0):
        assert n_q > 0
        if flatten_first or empty_initial:
            raise NotImplementedError("flatten_first / empty_initial are not used by the shipped configs")
        self.n_q = n_q
        self.delays = list(range(n_q)) if delays is None else list(delays)
        assert len(self.delays) == n_q and sorted(self.delays) == self.delays
        self.flatten_first = flatten_first
        self.empty_initial = empty_initial
        self.get_pattern = lru_cache(100)(self.get_pattern)  # type: ignore

    def get_pattern(self, timesteps: int) -> Pattern:
        return Pattern(self.n_q, timesteps, self.delays)
