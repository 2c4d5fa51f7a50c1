#!/usr/bin/env python
"""Benchmark of the V-AURA generation hot path (BASELINE.json metric: generated audio-sec/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload b64|b64_cfg|b1|b1_cfg|long_b1|dataset]
                    [--impl reference] [--no-sub] [--no-cpu-baseline]

One "step" = one pass of the hot path over one batch of synthetic clips: VAURAModel.generate(...)
(228 device-side decode steps + sampling + codec decode) on `batch` clips of 2.56 s.  Under torchrun
every rank runs the same per-GPU workload on its own clips (data parallel, weak scaling); the one exchange
step of the path - the NCCL all-gather of the fp16 waveforms - is inside the timed region of every
step at N > 1.

Printed JSON (rank 0, one line): value = whole-job audio-seconds per second with inputs resident in HBM;
e2e = the same through the public API with pinned host inputs and the waveform read back to the host;
roofline = the dominant kernel (one launch = one decode step) timed with CUDA events against MEASURED_PEAKS.json;
cpu_baseline = the oracle port of the reference's algorithm (no KV cache, fp32) run for one whole clip on this box's
host cores.  At N = 1 the default workload also carries sub-records for the other halves of BASELINE's metric:
`b1` (one clip, batch 1), `b64_cfg` (128 sequence rows), `long_b1` (10.24 s clip by overlapping windows, config 3) and
`frames_b64` (config 5 on one GPU: raw video frames -> Segment-AVCLIP tower -> decode -> codec, frames copied from pinned host
memory inside the timed region, with the tower's own tensor-roofline record).

`--workload dataset` is BASELINE config 4 in miniature: `driver.generate_dataset` over --clips-per-gpu x N clips
sharded by clip index, ending in the waveform all-gather, all inside the timed region.

`--impl reference` times the reference's own algorithm on the host cores: ONE real 2.56 s clip end to end (228
full-prefix forwards, llama.py:445-517 re-run per step as vaura_model.py:502-547 does, then the CPU codec decode),
cut into K consecutive slices that are reported as the K steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

AUDIO_SEC_PER_TOKEN = 512.0 / 44100.0
WORKLOADS = {
    # BASELINE.json configs[1]: batch 64, 2.56 s clips, top-k sampling (generate_vgg.yaml: temp 1.0, top_k 128)
    "b64": dict(batch=64, T=220, use_sampling=True, top_k=128, temp=1.0, cfg_scale=1.0,
                name="V-AURA 9cb LlamaGen decoder, random-init, 64 x 2.56 s clips, top-k 128 sampling, cfg off"),
    "b64_cfg": dict(batch=64, T=220, use_sampling=True, top_k=128, temp=1.0, cfg_scale=6.0,
                    name="same, classifier-free guidance 6.0 (128 sequence rows)"),
    # BASELINE.json configs[0]/[2]: batch-1 low latency
    "b1": dict(batch=1, T=220, use_sampling=True, top_k=128, temp=1.0, cfg_scale=1.0,
               name="one 2.56 s clip, batch 1, top-k 128 sampling"),
    # the reference's own generate settings (configs/generate_vgg.yaml: cfg_scale 6.0) at batch 1: two sequence rows
    "b1_cfg": dict(batch=1, T=220, use_sampling=True, top_k=128, temp=1.0, cfg_scale=6.0,
                   name="one 2.56 s clip, batch 1, top-k 128 sampling, classifier-free guidance 6.0 (2 sequence rows)"),
    # BASELINE.json configs[2]: one 10.24 s clip, batch 1, overlapping 2.56 s windows (scripts/generate.py:327-370)
    "long_b1": dict(batch=1, T=220, use_sampling=True, top_k=128, temp=1.0, cfg_scale=1.0, duration=10.24,
                    name="one 10.24 s clip, batch 1, 13 overlapping 2.56 s windows (stride 0.64 s), top-k 128 sampling"),
    # BASELINE.json configs[3] in miniature: clips sharded over the ranks by index + waveform all-gather
    "dataset": dict(batch=64, T=220, use_sampling=True, top_k=128, temp=1.0, cfg_scale=1.0,
                    name="synthetic clip set sharded data-parallel by clip index, batch 64 per GPU, NCCL waveform all-gather"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()] or [0])
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port; /root/reference does not exist on the GPU box)
# ---------------------------------------------------------------------------------------------------------------------
class CpuReferenceClip:
    """One real 2.56 s clip, B = 1, greedy, exactly as the reference executes it: no KV cache, every step re-runs the
    whole prefix through the 24 layers in fp32 (models/vaura_model.py:502-547 -> llama.py:445-517, restated by
    oracle/vaura_oracle.py: forward_full), the mask-fix and write-back of :536-544, then the codec decode on the CPU
    (models/modules/dac/model.py:41-48, restated by oracle/dac_oracle.py).  `run_steps` advances the loop so that the
    caller can cut the clip into timed slices."""

    def __init__(self, threads: int):
        from oracle import vaura_oracle as vo
        from oracle.dac_oracle import DacDecodeOracle
        from vaura_b200.synthetic import (FULL_CODEC, FULL_SAMPLER, make_avclip_features, make_codec_state_dict,
                                          make_sampler_state_dict)

        torch.set_num_threads(threads)
        self.vo = vo
        self.oracle = vo.SamplerOracle(make_sampler_state_dict(FULL_SAMPLER, 0), FULL_SAMPLER)
        self.codec = DacDecodeOracle(make_codec_state_dict(FULL_CODEC, 100), FULL_CODEC)
        self.feats = make_avclip_features(1, 1).reshape(1, 32, 768)
        self.T = 220
        codes = torch.full((1, 9, self.T), vo.UNKNOWN, dtype=torch.long)
        self.seq, self.mask = vo.build_pattern_sequence(codes, 1024)
        self.S = self.seq.shape[-1]
        self.offset = 1
        self.total_steps = self.S - 1

    def warm(self, n: int):
        with torch.no_grad():
            for _ in range(max(1, n)):
                self.oracle.forward_full(torch.full((1, 9, 8), 1024, dtype=torch.long), self.feats)

    def run_steps(self, n: int) -> int:
        done = 0
        with torch.no_grad():
            while done < n and self.offset < self.S:
                o = self.offset
                logits = self.oracle.forward_full(self.seq[..., :o], self.feats)[:, :, -1]  # last position only is used
                nxt = torch.argmax(logits, dim=-1)
                nxt[:, ~self.mask[:, o]] = 1024
                cur = self.seq[..., o]
                self.seq[..., o] = torch.where(cur == self.vo.UNKNOWN, nxt, cur)
                self.offset += 1
                done += 1
        return done

    def decode_audio(self):
        with torch.no_grad():
            return self.codec.decode(self.vo.revert_pattern_sequence(self.seq, self.T))


def cpu_reference_clip(threads: int, slices: int, warmup: int):
    """-> (audio-s/s, total seconds, per-slice seconds, codec seconds)."""
    clip = CpuReferenceClip(threads)
    clip.warm(warmup)
    slices = max(1, min(slices, clip.total_steps))
    per, t_slices = [], []
    base, extra = divmod(clip.total_steps, slices)
    t_all = time.perf_counter()
    codec_s = 0.0
    for i in range(slices):
        t0 = time.perf_counter()
        clip.run_steps(base + (1 if i < extra else 0))
        if i == slices - 1:
            tc = time.perf_counter()
            wav = clip.decode_audio()
            codec_s = time.perf_counter() - tc
            assert wav.shape == (1, 1, 220 * 512)
        t_slices.append(time.perf_counter() - t0)
    total = time.perf_counter() - t_all
    return 220 * AUDIO_SEC_PER_TOKEN / total, total, t_slices, codec_s


CPU_SAMPLE = ("oracle port of the reference's own loop, one whole 2.56 s clip, B=1, greedy, fp32: 228 full-prefix forwards "
              "(no KV cache, vaura_model.py:502-547) + CPU codec decode; measured, not extrapolated")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    steps = max(1, args.steps)
    value, total, t_slices, codec_s = cpu_reference_clip(threads, steps, args.warmup)
    line = {
        "impl": "reference", "metric": "generated audio-sec/sec", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"],
                   "note": "CPU reference algorithm: per-clip cost is batch-independent on the host cores, so one clip is the "
                           "bounded sample; the K steps are K consecutive slices of its 228 decode iterations (the last one "
                           "includes the codec decode)",
                   "clip_seconds": total, "codec_seconds": codec_s, "slices": len(t_slices)},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": CPU_SAMPLE},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="b64", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the b1 / b64_cfg / long_b1 / frames_b64 sub-records")
    ap.add_argument("--clips-per-gpu", type=int, default=128, help="dataset workload: clips per rank and step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch.distributed as dist

    from vaura_b200 import _cabi
    from vaura_b200.driver import gather_waveforms, generate_dataset, generate_long
    from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features
    from vaura_b200.weights import codec_flops, sampler_step_bytes

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    wl = WORKLOADS[args.workload]
    B, T = wl["batch"], wl["T"]
    d = FULL_SAMPLER
    model = build_model(FULL_SAMPLER, FULL_CODEC, device=str(dev))
    model.seed = 1234
    peak, peak_src = load_peaks()
    step_bytes = sampler_step_bytes(d)
    traffic_tab = {}
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        traffic_tab = json.load(open(prof))

    def gen_kw(w):
        return dict(max_new_tokens=w["T"], use_sampling=w["use_sampling"], temp=w["temp"], top_k=w["top_k"], top_p=0.0,
                    cfg_scale=w["cfg_scale"], prompt_is_encoded=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collective=True):
        """CUDA-event time of `steps` calls, barrier + synchronize on both sides, max over ranks."""
        if collective:
            barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if collective:
            barrier()
        else:
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1 and collective:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_latency(S):
        """p50 / p99 of the step-to-step latency from the %globaltimer stamp every step kernel (decode_step_cluster,
        decode_step_fused_bf16) leaves at the start of the launch that samples column `offset` (csrc/cabi.cu)."""
        try:
            ws_buf = model.sampler._buffers["ws"]
            st_ns = ws_buf[256 + 8 * 1024:256 + 8 * (1024 + S)].cpu().numpy().view(np.uint64).astype(np.int64)
            d_ns = np.diff(st_ns[2:S])
            d_ns = d_ns[(d_ns > 0) & (d_ns < 10**8)]
            if d_ns.size >= 50:
                return float(np.median(d_ns)) / 1e3, float(np.percentile(d_ns, 99)) / 1e3
        except Exception:  # paths without a persistent step kernel leave no stamps
            pass
        return None, None

    def decode_roofline(w, feats, ids):
        """The dominant kernel of a 2.56 s workload is the decode step (one launch = one generated column over all
        sequence rows).  Timed through the token-only generate: CUDA events around 228 back-to-back launches (the first
        one is the first-pass kernels), on torch's current stream, which is the stream the C ABI launches on."""
        rows = w["batch"] * (2 if w["cfg_scale"] > 1.0 else 1)
        kw = gen_kw(w)
        sampling = bool(w["use_sampling"]) and w["temp"] > 0

        def tokens_only():
            model.generate(frames=feats, clip_indices=ids, _decode_audio=False, **kw)
        tokens_only()
        ms_tok = timed(tokens_only, 3, collective=False) / 3
        nsteps = w["T"] + 8
        # the step kernel's average launch duration: CUDA events recorded inside vaura_sampler_generate right around the step
        # launches of the last call (the whole-call time ms_tok / nsteps also carries the first pass and the host glue of
        # generate(): ~1 ms per call; it is reported next to it as us_per_step_whole_call)
        loop_ms, loop_steps = model.sampler.last_loop_ms()
        k_ms = loop_ms / loop_steps if loop_steps > 0 else ms_tok / nsteps
        bf16 = rows >= 16 or (rows >= 3 and sampling)
        kv_el = 2 if bf16 else 4
        kv_bytes = d.num_layers * 2 * d.d_model * kv_el * rows * nsteps / 2  # K/V read at the mean context (S/2 positions)
        alg = step_bytes + kv_bytes
        if bf16:
            tm = 64 if rows <= 64 else 128
            kname = (f"decode_step_fused_bf16<{tm}> ({rows} sequence rows; one launch = embedding + 24 layers + heads + "
                     f"CFG/sampling/write-back of one column)")
            key = f"decode_step_fused_bf16_rows{rows}"
        elif rows <= 2:
            kname = f"decode_step_cluster<{rows}> (one launch = the whole decode step: 24 layers + heads + sampling)"
            key = f"decode_step_cluster_rows{rows}"
        else:
            kname = f"fp32-activation decode step at {rows} rows (decode_step_persistent / graph of GEMV kernels)"
            key = f"decode_step_persistent_rows{rows}"
        p50, p99 = step_latency(w["T"] + 9)
        ach = alg / (k_ms * 1e-3) / 1e9
        roof = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic_tab.get(key), "peak_source": peak_src, "us_per_launch": k_ms * 1e3,
                "launches_timed": loop_steps, "us_per_step_whole_call": ms_tok / nsteps * 1e3,
                "frac_whole_call": alg / (ms_tok / nsteps * 1e-3) / 1e9 / peak,
                "algorithmic_bytes_per_launch": alg,
                "algorithmic_bytes": f"weights {step_bytes} + K/V read {int(kv_bytes)} (24 layers x 2 x 1536 x {kv_el} B x "
                                     f"{rows} rows x mean context {nsteps // 2})"}
        step = {"p50_us": p50, "p99_us": p99, "mean_us": k_ms * 1e3, "weight_bytes": step_bytes,
                "weights_only_hbm_frac_of_measured": (step_bytes / (k_ms * 1e-3) / 1e9) / peak}
        return roof, step, ms_tok

    clocks = ClockSampler(local_rank)
    line = None

    # ---- dataset workload (BASELINE config 4 in miniature) ----------------------------------------------------------
    if args.workload == "dataset":
        per_gpu = args.clips_per_gpu
        n_items = per_gpu * world
        kw = gen_kw(wl)
        failed = []
        feat_cache = {}

        def gen_batch(ids):
            key = (int(ids[0]), int(ids[-1]))
            if key not in feat_cache:  # features depend only on the clip index (seed = clip index)
                feat_cache[key] = torch.stack([make_avclip_features(1, 100000 + int(c))[0] for c in ids]).to(dev)
            return model.generate(frames=feat_cache[key], clip_indices=ids, **kw)["generated_audio"]

        def step():
            return generate_dataset(gen_batch, n_items, B, rank, world, gather=True, failed=failed)

        for _ in range(max(args.warmup, 1)):
            out = step()
        assert out.shape == (n_items, 1, T * 512), out.shape
        if rank == 0:
            clocks.start()
        l0 = lib.vaura_launch_count()
        ms = timed(step, args.steps)
        launches = lib.vaura_launch_count() - l0
        clk = clocks.stop() if rank == 0 else None
        # the exchange step alone
        local = out[rank * per_gpu:(rank + 1) * per_gpu].contiguous()
        gather_ms = timed(lambda: gather_waveforms(local, n_items, rank, world), 5) / 5 if world > 1 else 0.0
        audio = n_items * T * AUDIO_SEC_PER_TOKEN
        value = audio * args.steps / (ms / 1000.0)
        line = {
            "metric": "generated audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 weights+activations, fp32 accumulate (tcgen05); codec fp16, fp32 accumulate",
            "data": "synthetic",
            "config": {"workload": wl["name"], "clips": n_items, "clips_per_gpu": per_gpu, "per_gpu_batch": B,
                       "parallelism": f"dp{world}: contiguous clip-index shards, full weight replica per GPU, one NCCL "
                                      "all_gather_into_tensor of the fp16 waveforms per step (inside the timed region)",
                       "l2": "weights 1.39 GB per decode step >> 126 MB L2; no flush needed"},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": per_gpu * 32 * 768 * 4,
                    "d2h_bytes_per_step": 0, "note": "features are made on the host per clip index and copied once (cached "
                                                      "across steps); the gathered waveforms stay on the device"},
            "gpu_launches": int(launches),
            "gather": {"ms": gather_ms, "bytes_per_rank": int(local.numel() * 2), "bytes_total": int(out.numel() * 2),
                       "algbw_GBps": (out.numel() * 2 / 1e9) / (gather_ms / 1e3) if gather_ms > 0 else None},
            "failed_clips": failed, "clocks": clk,
        }
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- long clip workload (BASELINE config 3) -------------------------------------------------------------------------
    def long_clip_record(steps):
        w = WORKLOADS["long_b1"]
        duration = w["duration"]
        feats = make_avclip_features(1, 3, segments=16).to(dev)  # 16 segments of AVCLIP features (10.24 s of video)
        kw = dict(use_sampling=True, temp=w["temp"], top_k=w["top_k"], top_p=0.0, cfg_scale=w["cfg_scale"])

        def run():
            return generate_long(model, feats, duration, **kw)
        out = run()
        n_tok = out["sampled_indices"].shape[-1]
        ms = timed(run, steps, collective=False) / steps
        ms_tok = timed(lambda: generate_long(model, feats, duration, decode_audio=False, **kw), steps, collective=False) / steps
        # one window's prefill alone: prompt of 165 tokens -> first pass over 166 columns + the first sampled column
        from vaura_b200.driver import chunk_schedule
        sched = chunk_schedule(duration)
        pl = sched[1]["prompt_len"]
        prompt = out["sampled_indices"][:, :, :pl].contiguous()
        sel = feats[:, torch.tensor(sched[1]["positions"], device=dev) % feats.shape[1]]

        def prefill_only():  # stops after the first sampled column (_end_offset): first pass over the prompt columns only
            model.generate(frames=sel, audio=prompt, max_new_tokens=sched[1]["max_gen_len"], prompt_is_encoded=True,
                           _decode_audio=False, _end_offset=pl + 2, **kw)
        prefill_only()
        ms_pre = timed(prefill_only, 5, collective=False) / 5
        audio = n_tok * AUDIO_SEC_PER_TOKEN
        return {"workload": w["name"], "value": audio / (ms / 1e3), "unit": "audio-s/s", "ms_per_clip": ms,
                "ms_tokens_only": ms_tok, "tokens": int(n_tok), "windows": len(sched),
                "prefill_positions": pl + 1, "prefill_ms_per_window": ms_pre,
                "prefill_note": "generate() with the window's prompt, stopped after the first sampled column: host glue + "
                                "the first pass over the prompt columns + one sampling launch"}

    # ---- BASELINE config 5 on one GPU: video frames -> Segment-AVCLIP features -> AR decode -> codec ------------------------
    def frames_record(steps):
        from vaura_b200.synthetic import FULL_AVCLIP, make_motionformer_state_dict
        from vaura_b200.weights import avclip_flops
        w = WORKLOADS["b64"]
        nb, segs = w["batch"], 4
        fx = model.visual_feature_extractor
        if not fx.has_weights:
            fx.load_state_dict(make_motionformer_state_dict(7), device=str(dev))
        gsrc = torch.Generator().manual_seed(5)
        fr_host = torch.empty(nb, segs, 3, FULL_AVCLIP.frames, FULL_AVCLIP.img_size, FULL_AVCLIP.img_size).pin_memory()
        fr_host.normal_(generator=gsrc)
        ids = torch.arange(nb, dtype=torch.int32)
        wh = torch.empty(nb, 1, w["T"] * 512, dtype=torch.float16).pin_memory()
        kws = gen_kw(w)

        def e2e_step():
            # the pinned host tensor goes straight into generate(): the extractor copies it chunk by chunk under its own compute
            wv = model.generate(frames=fr_host, clip_indices=ids, **kws)["generated_audio"]
            wh.copy_(wv, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_step()
        ms_s = timed(e2e_step, steps, collective=False) / steps
        fr_dev = fr_host.to(dev)
        fx(fr_dev)
        ms_fx = timed(lambda: fx(fr_dev), steps, collective=False) / steps
        tf = avclip_flops(FULL_AVCLIP, nb * segs) / 1e12
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1434.4) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1434.4
        del fr_dev
        return {"workload": "64 x 2.56 s clips from raw frames (64 x 4 segments of 16 x 3 x 224 x 224 fp32): Segment-AVCLIP tower "
                            "-> AR decode (top-k 128) -> codec",
                "e2e": nb * w["T"] * AUDIO_SEC_PER_TOKEN / (ms_s / 1e3), "unit": "audio-s/s", "ms_per_step": ms_s,
                "h2d_bytes_per_step": fr_host.numel() * 4, "d2h_bytes_per_step": wh.numel() * 2,
                "avclip": {"ms_per_256_segments": ms_fx, "segments_per_s": nb * segs / (ms_fx / 1e3),
                           "roofline": {"kernel": "gemm_tc_persistent_kernel<256,64,4,bf16,EpiVit> (70 % of the tower's time) + "
                                                  "attention / LayerNorm kernels: whole tower", "bound": "tensor",
                                        "achieved": tf / (ms_fx / 1e3), "peak": pk, "unit": "TFLOP/s",
                                        "frac": tf / (ms_fx / 1e3) / pk, "traffic": None,
                                        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)",
                                        "algorithmic_flops": tf * 1e12}}}

    # ---- DAC encode: 64 x 2.56 s of audio -> codes, waveforms from pinned host memory ----------------------------
    def encode_record(steps):
        from vaura_b200.synthetic import make_codec_state_dict
        from vaura_b200.weights import codec_encoder_flops
        enc = model.audio_encoder
        if enc._enc_blob is None:
            enc.load_state_dict(make_codec_state_dict(FULL_CODEC, 100, with_encoder=True), device=str(dev))
        nb, L = 64, 220 * FULL_CODEC.hop_length
        wav_h = (0.3 * torch.randn(nb, 1, L, generator=torch.Generator().manual_seed(6))).pin_memory()
        codes_h = torch.empty(nb, FULL_CODEC.n_codebooks, 220, dtype=torch.int64).pin_memory()

        def e2e_step():
            c = enc.encode(wav_h.to(dev, non_blocking=True))
            codes_h.copy_(c, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_step()
        ms_s = timed(e2e_step, steps, collective=False) / steps
        wav_d = wav_h.to(dev)
        ms_k = timed(lambda: enc.encode(wav_d), steps, collective=False) / steps
        tf = codec_encoder_flops(FULL_CODEC, L) * nb / 1e12
        pk = 1434.4
        pf = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pf):
            pk = json.load(open(pf)).get("bf16_tflops_sustained", pk)
        return {"workload": "64 x 2.56 s of 44.1 kHz audio -> 9 x 220 codes (DAC encoder + residual VQ)",
                "e2e_ms_per_batch": ms_s, "audio_s_per_s": nb * 220 * AUDIO_SEC_PER_TOKEN / (ms_s / 1e3),
                "h2d_bytes_per_step": wav_h.numel() * 4, "d2h_bytes_per_step": codes_h.numel() * 8,
                "roofline": {"kernel": "gemm_tc_persistent_kernel<.., EpiConv> implicit-GEMM convolutions of the encoder + "
                                       "enc_conv_in_kernel + rvq_encode_kernel: whole encode", "bound": "tensor",
                             "achieved": tf / (ms_k / 1e3), "peak": pk, "unit": "TFLOP/s", "frac": tf / (ms_k / 1e3) / pk,
                             "traffic": None, "ms_resident": ms_k, "algorithmic_flops": tf * 1e12,
                             "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"}}

    if args.workload == "long_b1":
        for _ in range(max(args.warmup, 1)):
            pass
        if rank == 0:
            clocks.start()
        rec = long_clip_record(max(args.steps, 1))
        clk = clocks.stop() if rank == 0 else None
        line = {"metric": "generated audio-sec/sec", "value": rec["value"] * world, "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": rec["ms_per_clip"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16 weights, fp32 activations/accumulate; codec fp16",
                "data": "synthetic", "config": {"workload": rec["workload"], "parallelism": f"dp{world} replicas"},
                "long_b1": rec, "clocks": clk}
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- 2.56 s workloads (b64 default) -------------------------------------------------------------------------------
    kw = gen_kw(wl)
    # rank r owns clips [r*B, (r+1)*B) of every step; features depend only on the clip index
    feats_host = make_avclip_features(B, 2 + rank).pin_memory()
    feats_dev = feats_host.to(dev)
    clip_ids = torch.arange(rank * B, (rank + 1) * B, dtype=torch.int32)
    wav_host = torch.empty(B * world, 1, T * 512, dtype=torch.float16).pin_memory()
    n_items = B * world

    def step_resident():
        wav = model.generate(frames=feats_dev, clip_indices=clip_ids, **kw)["generated_audio"]
        return gather_waveforms(wav, n_items, rank, world)  # identity at N = 1; NCCL all-gather otherwise

    def step_e2e():
        f = feats_host.to(dev, non_blocking=True)
        wav = model.generate(frames=f, clip_indices=clip_ids, **kw)["generated_audio"]
        wav = gather_waveforms(wav, n_items, rank, world)
        wav_host.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return wav_host

    for _ in range(max(args.warmup, 3)):
        step_resident()
    if rank == 0:
        clocks.start()
    l0 = lib.vaura_launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.vaura_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    audio_per_step = B * T * AUDIO_SEC_PER_TOKEN * world
    value = audio_per_step * args.steps / (ms / 1000.0)
    e2e = audio_per_step * args.steps / (ms_e2e / 1000.0)
    gather_ms = None
    if world > 1:
        wav_local = model.generate(frames=feats_dev, clip_indices=clip_ids, **kw)["generated_audio"]
        gather_ms = timed(lambda: gather_waveforms(wav_local, n_items, rank, world), 5) / 5

    roof, step_rec, ms_tok = decode_roofline(wl, feats_dev, clip_ids)
    rows = B * (2 if wl["cfg_scale"] > 1.0 else 1)
    line = {
        "metric": "generated audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": ("bf16 weights+activations, fp32 accumulate (tcgen05)" if rows >= 3 else "bf16 weights, fp32 activations/accumulate")
                 + "; codec fp16, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": wl["name"], "per_gpu_batch": B, "tokens_per_clip": T, "decode_steps": T + 8,
                   "l2": "weights 1.39 GB per decode step >> 126 MB L2; no flush needed", "cfg_scale": wl["cfg_scale"],
                   "parallelism": f"dp{world} (clips sharded by index, no collective in the decode loop; NCCL all-gather of the "
                                  "fp16 waveforms at the end of every step)" if world > 1 else "dp1"},
        "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": feats_host.numel() * 4,
                "d2h_bytes_per_step": wav_host.numel() * 2, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roof,
        "decode_step": step_rec,
        "clocks": clk,
        "codec": {"ms_per_batch": ms / args.steps - ms_tok - (gather_ms or 0.0), "gflop_per_clip": codec_flops(FULL_CODEC, T) / 1e9},
    }
    if gather_ms is not None:
        line["gather"] = {"ms": gather_ms, "bytes_per_rank": int(B * T * 512 * 2), "bytes_total": int(n_items * T * 512 * 2)}

    # ---- the other halves of BASELINE's metric, N = 1 only (the driver runs the default line) --------------------------
    if world == 1 and not args.no_sub and args.workload == "b64":
        sub_steps = 3

        def small_record(name):
            w = WORKLOADS[name]
            fh = make_avclip_features(w["batch"], 2).pin_memory()
            ids = torch.arange(w["batch"], dtype=torch.int32)
            wh = torch.empty(w["batch"], 1, w["T"] * 512, dtype=torch.float16).pin_memory()
            kws = gen_kw(w)

            def e2e_step():
                f = fh.to(dev, non_blocking=True)
                wv = model.generate(frames=f, clip_indices=ids, **kws)["generated_audio"]
                wh.copy_(wv, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            for _ in range(3):
                e2e_step()
            ms_s = timed(e2e_step, sub_steps, collective=False) / sub_steps
            r, s, _ = decode_roofline(w, fh.to(dev), ids)
            return {"workload": w["name"], "e2e": w["batch"] * w["T"] * AUDIO_SEC_PER_TOKEN / (ms_s / 1e3),
                    "unit": "audio-s/s", "ms_per_step": ms_s, "roofline": r, "decode_step": s}

        # a sub-record that fails (e.g. no room for 2.5 GB of pinned frames on the host) is reported as such; it never takes
        # the headline line with it
        for name, fn in (("b1", lambda: small_record("b1")), ("b64_cfg", lambda: small_record("b64_cfg")),
                         ("long_b1", lambda: long_clip_record(2)), ("frames_b64", lambda: frames_record(2)),
                         ("codec_encode_b64", lambda: encode_record(3))):
            try:
                line[name] = fn()
            except Exception as e:  # noqa: BLE001
                line[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.synchronize()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        try:
            v, total, t_slices, codec_s = cpu_reference_clip(threads, 1, 1)
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                    "sample": CPU_SAMPLE + f"; {total:.1f} s per clip (codec {codec_s:.1f} s)"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300], "kind": "port", "cores": threads}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
