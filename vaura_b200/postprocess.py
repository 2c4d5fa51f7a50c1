"""Output tail of the generation driver: loudness / peak / RMS / clip normalisation of the generated
waveforms and writing them out, mirroring scripts/generate.py:392-461 (`save_results`, `scale_audio`) and
utils/data_utils.py:346-466 (`normalize_loudness`, `_clip_wav`, `normalize_audio`; audiocraft's functions, MIT).

This part of the reference is CPU-side I/O; here the arithmetic is plain torch and therefore also runs on the device the
waveforms already live on (`normalize_batch` normalises a whole batch before the device->host copy).  The `loudness`
strategy needs the ITU-R BS.1770 meter, which both the reference and this file take from `torchaudio.transforms.Loudness`.
Files: 32-bit float WAV through a dependency-free RIFF writer (what `torchaudio.save` writes for a float32 tensor); the
mp4 mux of the reference's `write_video` needs PyAV, which this image does not have - `save_results` writes the wav and
skips the mp4 with a warning unless PyAV is importable.
"""
from __future__ import annotations

import logging
import struct
import sys
import typing as tp
from pathlib import Path

import torch

logger = logging.getLogger(__name__)


def normalize_loudness(wav: torch.Tensor, sample_rate: int, loudness_headroom_db: float = 14,
                       loudness_compressor: bool = False, energy_floor: float = 2e-3) -> torch.Tensor:
    """utils/data_utils.py:346-387: scale to -loudness_headroom_db LUFS; signals below the energy floor pass through."""
    import torchaudio

    energy = wav.pow(2).mean().sqrt().item()
    if energy < energy_floor:
        return wav
    try:
        input_loudness_db = torchaudio.transforms.Loudness(sample_rate).to(wav.device)(wav).item()
    except Exception as e:  # the reference logs and returns the input unchanged (data_utils.py:378-386)
        print(f"Error in normalize_loudness: {e}", file=sys.stderr)
        return wav
    gain = 10.0 ** ((-loudness_headroom_db - input_loudness_db) / 20.0)
    output = gain * wav
    if loudness_compressor:
        output = torch.tanh(output)
    if not output.isfinite().all():
        return wav
    return output


def _clip_wav(wav: torch.Tensor, log_clipping: bool = False, stem_name: tp.Optional[str] = None) -> None:
    """utils/data_utils.py:390-404 (in place)."""
    max_scale = wav.abs().max()
    if log_clipping and max_scale > 1:
        clamp_prob = (wav.abs() > 1).float().mean().item()
        print(f"CLIPPING {stem_name or ''} happening with proba (a bit of clipping is okay):", clamp_prob,
              "maximum scale: ", max_scale.item(), file=sys.stderr)
    wav.clamp_(-1, 1)


def normalize_audio(wav: torch.Tensor, normalize: bool = True, strategy: str = "peak", peak_clip_headroom_db: float = 6,
                    rms_headroom_db: float = 18, loudness_headroom_db: float = 12, loudness_compressor: bool = False,
                    log_clipping: bool = False, sample_rate: tp.Optional[int] = None,
                    stem_name: tp.Optional[str] = None) -> torch.Tensor:
    """utils/data_utils.py:407-466, same strategies and defaults: 'peak', 'clip', 'rms', 'loudness', '' / 'none'."""
    scale_peak = 10 ** (-peak_clip_headroom_db / 20)
    scale_rms = 10 ** (-rms_headroom_db / 20)
    if strategy == "peak":
        rescaling = scale_peak / wav.abs().max()
        if normalize or rescaling < 1:
            wav = wav * rescaling
    elif strategy == "clip":
        wav = wav.clamp(-scale_peak, scale_peak)
    elif strategy == "rms":
        mono = wav.mean(dim=0)
        rescaling = scale_rms / mono.pow(2).mean().sqrt()
        if normalize or rescaling < 1:
            wav = wav * rescaling
        _clip_wav(wav, log_clipping=log_clipping, stem_name=stem_name)
    elif strategy == "loudness":
        assert sample_rate is not None, "Loudness normalization requires sample rate."
        wav = normalize_loudness(wav, sample_rate, loudness_headroom_db, loudness_compressor)
        _clip_wav(wav, log_clipping=log_clipping, stem_name=stem_name)
    else:
        assert wav.abs().max() < 1
        assert strategy == "" or strategy == "none", f"Unexpected strategy: '{strategy}'"
    return wav


def scale_audio(audio: torch.Tensor, strategy: str = "loudness", sample_rate: int = 24000, db: float = 6.0) -> torch.Tensor:
    """scripts/generate.py:443-461: to float32, normalise, flatten to (1, samples) on the host."""
    if audio.dtype not in [torch.float32, torch.int32, torch.int16, torch.uint8]:
        audio = audio.to(torch.float32)
    audio = normalize_audio(audio, strategy=strategy, sample_rate=sample_rate, peak_clip_headroom_db=db)
    return audio.reshape(1, -1).to("cpu")


def normalize_batch(wavs: torch.Tensor, strategy: str = "loudness", sample_rate: int = 44100, db: float = 6.0) -> torch.Tensor:
    """The same per-clip arithmetic for a whole batch (B, 1, samples) on whatever device it lives on (the generated fp16
    waveforms are still in HBM after `generate`): 'clip', 'peak' and 'rms' are batched tensor ops, 'loudness' walks the
    clips because the BS.1770 meter is per signal.  Returns float32 (B, 1, samples)."""
    x = wavs.to(torch.float32)
    if x.dim() == 2:
        x = x[:, None]
    scale_peak = 10 ** (-db / 20)
    if strategy == "clip":
        return x.clamp(-scale_peak, scale_peak)
    if strategy == "peak":
        return x * (scale_peak / x.abs().amax(dim=(1, 2), keepdim=True))
    if strategy == "rms":
        scale_rms = 10 ** (-18 / 20)
        mono = x.mean(dim=1)
        return (x * (scale_rms / mono.pow(2).mean(dim=-1).sqrt())[:, None, None]).clamp_(-1, 1)
    return torch.stack([normalize_audio(c, strategy=strategy, sample_rate=sample_rate, peak_clip_headroom_db=db) for c in x])


def write_wav_f32(path: tp.Union[str, Path], audio: torch.Tensor, sample_rate: int) -> None:
    """(channels, samples) float -> RIFF/WAVE, format tag 3 (IEEE float), 32 bits: the file `torchaudio.save(path, audio,
    sr)` produces for a float32 tensor (scripts/generate.py:423)."""
    a = audio.detach().to("cpu", torch.float32)
    if a.dim() == 1:
        a = a[None]
    ch, n = a.shape
    data = a.t().contiguous().numpy().astype("<f4").tobytes()
    fmt = struct.pack("<HHIIHH", 3, ch, sample_rate, sample_rate * ch * 4, ch * 4, 32)
    fact = struct.pack("<I", n)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + \
        b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def read_wav_f32(path: tp.Union[str, Path]) -> tp.Tuple[torch.Tensor, int]:
    """Inverse of `write_wav_f32` (tests, round trips)."""
    import numpy as np

    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE", "not a RIFF/WAVE file"
    pos, ch, sr, data = 12, 1, 0, b""
    while pos + 8 <= len(raw):
        tag, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        chunk = raw[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            code, ch, sr, _, _, bits = struct.unpack("<HHIIHH", chunk[:16])
            assert code == 3 and bits == 32, "only 32-bit float files"
        elif tag == b"data":
            data = chunk
        pos += 8 + size + (size & 1)
    a = torch.from_numpy(np.frombuffer(data, dtype="<f4").copy()).reshape(-1, ch).t().contiguous()
    return a, sr


def write_video(filename: str, video_array, fps: float, audio_array: tp.Optional[torch.Tensor], audio_fps: int,
                video_codec: str = "h264", audio_codec: str = "aac", options: tp.Optional[dict] = None) -> bool:
    """The mux of utils/utils.py `write_video` (PyAV).  Returns False (after a warning) when PyAV is not installed."""
    try:
        import av
        if not hasattr(av, "open"):
            raise ImportError("av without av.open")
    except Exception:
        logger.warning("PyAV is not installed: %s not written (the wav next to it is)", filename)
        return False
    import numpy as np

    with av.open(filename, mode="w") as container:
        stream = container.add_stream(video_codec, rate=int(round(fps)))
        arr = np.asarray(video_array, dtype=np.uint8)
        stream.width, stream.height = arr.shape[2], arr.shape[1]
        stream.pix_fmt = "yuv420p" if video_codec != "libx264rgb" else "rgb24"
        stream.options = options or {}
        a_stream = None
        if audio_array is not None:
            a_stream = container.add_stream(audio_codec, rate=audio_fps)
            frame = av.AudioFrame.from_ndarray(audio_array.numpy().astype("<f4"), format="fltp",
                                               layout="mono" if audio_array.shape[0] == 1 else "stereo")
            frame.sample_rate = audio_fps
            for packet in a_stream.encode(frame):
                container.mux(packet)
            for packet in a_stream.encode():
                container.mux(packet)
        for img in arr:
            frame = av.VideoFrame.from_ndarray(img, format="rgb24")
            for packet in stream.encode(frame):
                container.mux(packet)
        for packet in stream.encode():
            container.mux(packet)
    return True


def save_results(audio: torch.Tensor, frames: tp.Optional[torch.Tensor], output_dir_path: Path, fn: str, v_fps: float = 25,
                 generated_a_fps: int = 44100, original_a_fps: int = 24000, original_audio: tp.Optional[torch.Tensor] = None,
                 audio_norm_strategy: str = "clip", save_original: bool = False,
                 pre_normalized: bool = False) -> tp.Dict[str, tp.Optional[str]]:
    """scripts/generate.py:392-440: normalise, write `<fn>.wav` (+ `<fn>.mp4` with the frames and, on request,
    `<fn>_original.mp4` with the source audio).  Returns the paths written.  ``pre_normalized``: the waveform already went
    through `normalize_batch` on the device (driver.save_waveforms)."""
    output_dir_path = Path(output_dir_path)
    output_dir_path.mkdir(parents=True, exist_ok=True)
    if pre_normalized:
        audio = audio.to(torch.float32).reshape(1, -1).to("cpu")
    else:
        audio = scale_audio(audio, audio_norm_strategy, generated_a_fps)
    if fn.endswith(".mp4") or fn.endswith(".wav"):
        fn = fn[:-4]
    audio_path = output_dir_path / f"{fn}.wav"
    video_path = output_dir_path / f"{fn}.mp4"
    written: tp.Dict[str, tp.Optional[str]] = {"wav": audio_path.as_posix(), "mp4": None, "original_mp4": None}
    if frames is not None:
        if video_path.exists():
            logger.warning("File %s already exists. Overwriting...", video_path.as_posix())
        if write_video(video_path.as_posix(), frames.permute(0, 2, 3, 1).cpu().numpy(), v_fps, audio, generated_a_fps,
                       options={"crf": "10", "pix_fmt": "yuv420p"}):
            written["mp4"] = video_path.as_posix()
    write_wav_f32(audio_path, audio, generated_a_fps)
    if original_audio is not None and save_original and frames is not None:
        original_audio = scale_audio(original_audio, audio_norm_strategy, original_a_fps)
        p = output_dir_path / f"{fn}_original.mp4"
        if write_video(p.as_posix(), frames.permute(0, 2, 3, 1).cpu().numpy(), v_fps, original_audio.reshape(1, -1).to("cpu"),
                       original_a_fps, options={"crf": "10", "pix_fmt": "yuv420p"}):
            written["original_mp4"] = p.as_posix()
    return written
