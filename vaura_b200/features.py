"""Visual feature extractor boundary.

The north star feeds *synthetic Segment-AVCLIP features*; the MotionFormer backbone itself
(models/modules/feature_extractors/avclip/motionformer.py:252-342) is a "next" row and is
not built yet.  This class keeps the reference's class name — ``VAURAModel`` checks
``__class__.__name__ == "MotionFormer"`` (models/vaura_model.py:73-75) — and its output contract
``(feats (B,S,t,D), None)``, accepting precomputed features ``(B,S,t,768)`` and failing loudly on raw
frames ``(B,S,3,16,224,224)``.
"""
from __future__ import annotations

import torch


class MotionFormer(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.config = dict(kwargs)

    def forward(self, x: torch.Tensor):
        if x.dim() == 4:  # already AVCLIP features (B, S, t, D)
            return x, None
        raise NotImplementedError(
            "raw-frame Segment-AVCLIP extraction is outside the built hot path (SURVEY §8f row 2); "
            "pass precomputed AVCLIP features of shape (B, segments, 8, 768)")
