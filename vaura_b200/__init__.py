"""vaura_b200 — B200-native (sm_100a) implementation of V-AURA's generation hot path.

The directory is named ``vaura_b200`` (importable) rather than ``v-aura_b200``.  Public surface mirrors
the reference: ``VAURAModel.generate`` / ``load_from_checkpoint``, ``sampler.Transformer``,
``codec.DacModelWrapper``, ``patterns.DelayedPatternProvider`` and ``config.instantiate_from_config``.
All arithmetic runs in ``_lib/libvaura_b200.so`` (see include/vaura_b200.h); there is no CPU fallback.
"""
from .model import VAURAModel  # noqa: F401
from .synthetic import FULL_CODEC, FULL_SAMPLER, TINY_CODEC, TINY_SAMPLER  # noqa: F401

__all__ = ["VAURAModel", "FULL_CODEC", "FULL_SAMPLER", "TINY_CODEC", "TINY_SAMPLER"]
