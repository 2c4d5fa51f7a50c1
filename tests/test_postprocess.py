"""Output tail: vaura_b200.postprocess against the reference's own normalize_audio / normalize_loudness /
_clip_wav outputs (tests/golden/postprocess.npz, made by oracle/make_golden_postprocess.py from utils/data_utils.py:346-466)
and the wav writer's round trip.  CPU-only."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_postprocess import CASES, waveforms
from vaura_b200 import postprocess as pp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postprocess.npz")


def test_normalize_audio_matches_reference_golden():
    g = np.load(GOLD)
    for name, wav in waveforms().items():
        for strategy, db in CASES:
            y = pp.normalize_audio(wav.clone(), strategy=strategy, sample_rate=44100, peak_clip_headroom_db=db)
            ref = torch.from_numpy(g[f"{name}|{strategy}|{db}"])
            assert torch.allclose(y[:, ::37], ref, rtol=0, atol=1e-6), (name, strategy, db)
            stats = g[f"{name}|{strategy}|{db}|stats"]
            assert abs(float(y.abs().max()) - stats[0]) < 1e-6
            assert abs(float(y.pow(2).mean().sqrt()) - stats[1]) < 1e-6


def test_scale_audio_shape_dtype_and_batch_form():
    wavs = torch.stack([w for w in waveforms().values()]).to(torch.float16)  # (3, 1, N) like generate()'s output
    for strategy in ("clip", "peak", "rms", "loudness"):
        batch = pp.normalize_batch(wavs, strategy, 44100, 6.0)
        assert batch.shape == wavs.shape and batch.dtype == torch.float32
        for i in range(wavs.shape[0]):
            one = pp.scale_audio(wavs[i], strategy, 44100, 6.0)
            assert one.shape == (1, wavs.shape[-1]) and one.dtype == torch.float32 and one.device.type == "cpu"
            assert torch.allclose(batch[i].reshape(1, -1), one, atol=1e-6), strategy
    with pytest.raises(AssertionError):
        pp.normalize_audio(torch.ones(1, 8) * 2, strategy="none")
    with pytest.raises(AssertionError):
        pp.normalize_audio(torch.zeros(1, 8), strategy="bogus")


def test_save_results_writes_float_wav(tmp_path):
    wav = waveforms()["loud"].to(torch.float16)[None]  # (1, 1, N)
    written = pp.save_results(wav[0], None, tmp_path / "out", "clip_0001.mp4", generated_a_fps=44100, audio_norm_strategy="clip")
    assert written["wav"].endswith("clip_0001.wav") and written["mp4"] is None
    back, sr = pp.read_wav_f32(written["wav"])
    assert sr == 44100 and back.shape == (1, wav.shape[-1])
    assert torch.equal(back, pp.scale_audio(wav[0], "clip", 44100))
    assert float(back.abs().max()) <= 10 ** (-6 / 20) + 1e-7
    # frames given but no PyAV in this image: the wav is still written, the mp4 is reported as skipped
    frames = torch.zeros(4, 3, 32, 32, dtype=torch.uint8)
    written = pp.save_results(wav[0], frames, tmp_path / "out", "clip_0002", audio_norm_strategy="peak")
    assert os.path.exists(written["wav"])
    try:
        import av
        has_av = hasattr(av, "open")
    except ImportError:
        has_av = False
    assert (written["mp4"] is not None) == has_av


def test_driver_save_waveforms(tmp_path):
    from vaura_b200.driver import save_waveforms

    wavs = torch.stack([w for w in waveforms().values()]).to(torch.float16)
    for strategy in ("clip", "loudness"):
        res = save_waveforms(wavs, ["a.mp4", "b", "c.wav"], tmp_path / strategy, strategy)
        assert [os.path.basename(r["wav"]) for r in res] == ["a.wav", "b.wav", "c.wav"]
        for i, r in enumerate(res):
            back, sr = pp.read_wav_f32(r["wav"])
            assert sr == 44100 and torch.allclose(back, pp.scale_audio(wavs[i], strategy, 44100), atol=1e-6)
