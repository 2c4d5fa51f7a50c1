"""TEST INFRASTRUCTURE ONLY — writes tests/golden/postprocess.npz by executing the reference's OWN normalisation functions
(utils/data_utils.py:346-466: normalize_loudness, _clip_wav, normalize_audio) on seeded waveforms.  The module they live in
cannot be imported here (it imports PyAV and torchvision.io.read_video), so the three function definitions are cut out of
the unmodified source file with `ast` and executed as they stand.

    python oracle/make_golden_postprocess.py        (needs /root/reference: build container only)
"""
import ast
import os
import sys
from typing import Optional  # noqa: F401  (used by the extracted source)

import numpy as np
import torch
import torchaudio  # noqa: F401  (used by the extracted source)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/utils/data_utils.py"
CASES = [("peak", 6.0), ("clip", 6.0), ("clip", 1.0), ("rms", 6.0), ("loudness", 6.0)]


def reference_functions():
    src = open(SRC).read()
    tree = ast.parse(src)
    want = {"normalize_loudness", "_clip_wav", "normalize_audio"}
    code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want)
    ns = {"torch": torch, "torchaudio": torchaudio, "sys": sys, "Optional": Optional}
    exec(compile(code, SRC, "exec"), ns)
    return ns


def waveforms():
    g = torch.Generator().manual_seed(21)
    t = torch.arange(44100 * 2) / 44100.0
    tone = 0.4 * torch.sin(2 * torch.pi * 440 * t) + 0.2 * torch.sin(2 * torch.pi * 1870 * t)
    return {
        "loud": (tone * 3.0 + 0.3 * torch.randn(t.numel(), generator=g))[None],
        "quiet": (tone * 0.05 + 0.01 * torch.randn(t.numel(), generator=g))[None],
        "silent": (1e-4 * torch.randn(t.numel(), generator=g))[None],
    }


def main():
    ns = reference_functions()
    out = {}
    for name, wav in waveforms().items():
        for strategy, db in CASES:
            y = ns["normalize_audio"](wav.clone(), strategy=strategy, sample_rate=44100, peak_clip_headroom_db=db)
            out[f"{name}|{strategy}|{db}"] = y.numpy().astype(np.float32)[:, ::37]  # every 37th sample keeps the file small
            out[f"{name}|{strategy}|{db}|stats"] = np.array([float(y.abs().max()), float(y.pow(2).mean().sqrt()), float(y.sum())],
                                                            dtype=np.float64)
    p = os.path.join(ROOT, "tests", "golden", "postprocess.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
