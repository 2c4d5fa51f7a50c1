"""Drivers around ``VAURAModel.generate``: long clips by overlapping windows, and data-parallel generation.

* ``generate_long`` mirrors the chunked branch of the reference driver (scripts/generate.py:327-370, same
  arithmetic as demo.ipynb cell 8): windows of ``model_max_duration`` (2.56 s) advanced by ``stride`` (0.64 s),
  the tail of the previous window fed back as an encoded token prompt, per-window visual segments picked by
  ``positions = arange(ceil(t*vfps)//16, (ceil(t*vfps)+ceil(dur*vfps))//16) % n_segments``; every window restarts
  positions at 0, i.e. a prefill of the prompt columns followed by decode steps; one codec decode at the end.
* ``shard_range`` / ``generate_dataset`` / ``gather_waveforms``: clips are independent, so rank r of R takes a
  contiguous block of clip indices; nothing is communicated inside the hot loop.  The only exchange
  step is the final all-gather of the fp16 waveforms (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from math import ceil
from typing import Callable, List, Optional, Tuple

import torch

COMPRESSION_MODEL_FRAME_RATE = 86  # DAC 44.1 kHz: 44100 / 512 frames per second (scripts/generate.py)


def chunk_schedule(duration: float, stride: float = 0.64, model_max_duration: float = 2.56, vfps: int = 25,
                   frame_rate: int = COMPRESSION_MODEL_FRAME_RATE) -> List[dict]:
    """The (prompt_len, max_gen_len, segment positions) sequence of scripts/generate.py:327-365."""
    total_gen_len = int(duration * frame_rate)
    stride_tokens = int(frame_rate * stride)
    out, current, prompt_length = [], 0, 0
    while current + prompt_length < total_gen_len:
        time_offset = current / frame_rate
        chunk_duration = min(duration - time_offset, model_max_duration)
        max_gen_len = ceil(chunk_duration * frame_rate)
        initial_position = ceil(time_offset * vfps)
        video_target_length = ceil(chunk_duration * vfps)
        positions = list(range(initial_position // 16, (initial_position + video_target_length) // 16))
        out.append(dict(prompt_len=prompt_length, max_gen_len=max_gen_len, positions=positions))
        prompt_length = max_gen_len - stride_tokens
        current += stride_tokens
    return out


@torch.no_grad()
def generate_long(model, frames: torch.Tensor, duration: float, stride: float = 0.64, model_max_duration: float = 2.56,
                  vfps: int = 25, decode_audio: bool = True, **gen_kwargs) -> dict:
    """frames: per-segment visual input (B, n_segments, ...), e.g. AVCLIP features (B, S, 8, 768)."""
    all_tokens, prompt_tokens = [], None
    stride_tokens = int(COMPRESSION_MODEL_FRAME_RATE * stride)
    stream0 = int(gen_kwargs.pop("_stream_id", 0) or 0)
    for wi, ch in enumerate(chunk_schedule(duration, stride, model_max_duration, vfps)):
        pos = torch.tensor(ch["positions"], device=frames.device) % frames.shape[1]
        selected = frames[:, pos]
        # every window restarts at column 0 with the same clip ids: the window index goes into the Philox counter so
        # that overlapping windows do not draw the same uniforms for the same columns
        item = model.generate(frames=selected, audio=prompt_tokens, max_new_tokens=ch["max_gen_len"],
                              return_sampled_indices=True, remove_prompts=False, prompt_is_encoded=True,
                              _decode_audio=False, _stream_id=stream0 + wi, **gen_kwargs)
        gen_tokens = item["sampled_indices"]
        all_tokens.append(gen_tokens if prompt_tokens is None else gen_tokens[:, :, prompt_tokens.shape[-1]:])
        prompt_tokens = gen_tokens[:, :, stride_tokens:]
    gen_tokens = torch.cat(all_tokens, dim=-1)
    audio = model.audio_encoder.decode([(gen_tokens[..., : model.num_codebooks, :], None)]) if decode_audio else None
    return {"generated_audio": audio, "sampled_indices": gen_tokens}


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rank ``rank``: [r*ceil(n/R), min(n, (r+1)*ceil(n/R)))."""
    per = (n_items + world - 1) // world
    return min(n_items, rank * per), min(n_items, (rank + 1) * per)


def gather_waveforms(local: torch.Tensor, n_items: int, rank: int, world: int, group=None) -> Optional[torch.Tensor]:
    """All-gather per-rank waveform blocks ``(n_local, 1, L)`` (padded to the common block size) into
    ``(n_items, 1, L)`` in clip order on every rank."""
    import torch.distributed as dist

    if world == 1:
        return local
    per = (n_items + world - 1) // world
    L = local.shape[-1]
    padded = torch.zeros(per, 1, L, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * per, 1, L, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_items]


@torch.no_grad()
def generate_dataset(generate_batch: Callable[[torch.Tensor], torch.Tensor], n_items: int, batch_size: int, rank: int = 0,
                     world: int = 1, gather: bool = True, group=None, failed: Optional[List[int]] = None):
    """Data-parallel generation of ``n_items`` clips.  ``generate_batch(clip_ids int32 (b,)) -> waveforms (b,1,L)``
    must depend only on the clip ids (features and the Philox counters are keyed by clip id), which makes the
    result independent of ``world`` and ``batch_size``.  Returns all waveforms in clip order (every rank) when
    ``gather`` is set, else this rank's block.

    Failure isolation (scripts/generate.py:386-389 wraps every clip in try/except, prints and goes on): a batch that
    raises is retried clip by clip; a clip that still raises is reported in ``failed`` (its global index is appended
    when a list is passed) and its waveform is left zero, so one bad input does not end a 15k-clip run."""
    lo, hi = shard_range(n_items, rank, world)
    blocks: List[Optional[torch.Tensor]] = []
    missing: List[Tuple[int, int]] = []  # (block position, clip id) of clips that failed
    for b0 in range(lo, hi, batch_size):
        ids = torch.arange(b0, min(hi, b0 + batch_size), dtype=torch.int32)
        try:
            blocks.append(generate_batch(ids))
            continue
        except Exception as e:  # noqa: BLE001 - the reference driver catches everything per clip as well
            print(f"[generate_dataset] rank {rank}: batch {b0}..{int(ids[-1])} failed ({type(e).__name__}: {e}); "
                  f"retrying clip by clip", flush=True)
        for cid in ids.tolist():
            try:
                blocks.append(generate_batch(torch.tensor([cid], dtype=torch.int32)))
            except Exception as e:  # noqa: BLE001
                print(f"[generate_dataset] rank {rank}: clip {cid} failed ({type(e).__name__}: {e})", flush=True)
                missing.append((len(blocks), cid))
                blocks.append(None)
                if failed is not None:
                    failed.append(cid)
    good = [b for b in blocks if b is not None]
    if good:
        for pos, _ in missing:
            blocks[pos] = torch.zeros_like(good[0][:1])
        local = torch.cat(blocks, dim=0)
    elif blocks:  # every clip of the shard failed: the waveform shape is learnt from the peers below
        local = None
    else:  # rank beyond the data
        local = None
    n_failed_here = len(missing)
    if not gather or world == 1:
        if local is None and blocks:
            raise RuntimeError(f"all {len(blocks)} clips of rank {rank} failed")
        return local
    import torch.distributed as dist

    # ranks with no clips need the waveform length / dtype / device of the others
    meta = torch.tensor([local.shape[-1] if local is not None else 0], dtype=torch.int64,
                        device=local.device if local is not None else _default_device())
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)
    if local is None:
        local = torch.zeros(n_failed_here, 1, int(meta.item()), dtype=torch.float16, device=meta.device)
    return gather_waveforms(local, n_items, rank, world, group)


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def save_waveforms(waveforms: torch.Tensor, names, output_dir, audio_norm_strategy: str = "clip", sample_rate: int = 44100,
                   frames=None, v_fps: float = 25.0) -> list:
    """The output loop of scripts/generate.py:372-384 for a batch (or the gathered set) of generated clips: normalise on
    the device the waveforms live on, then one `<name>.wav` (+ mp4 when frames and PyAV are there) per clip.
    Returns the per-clip dicts of paths written (postprocess.save_results)."""
    from . import postprocess as pp

    norm = pp.normalize_batch(waveforms, audio_norm_strategy, sample_rate).cpu()
    out = []
    for i, name in enumerate(names):
        out.append(pp.save_results(norm[i], None if frames is None else frames[i], output_dir, str(name), v_fps, sample_rate,
                                   audio_norm_strategy=audio_norm_strategy, pre_normalized=True))
    return out
