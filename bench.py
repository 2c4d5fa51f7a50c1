#!/usr/bin/env python
"""Benchmark of the V-AURA generation hot path (BASELINE.json metric: generated audio-sec/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload b64|b64_cfg|b1|b1_cfg] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic clips: VAURAModel.generate(...)
(228 device-side decode steps + sampling + codec decode) on `batch` clips of 2.56 s.  Under torchrun
every rank runs the same per-GPU workload on its own clips (data parallel, weak scaling, no collective
in the timed path; the final waveform gather is outside the hot loop).

Printed JSON (rank 0, one line): value = whole-job audio-seconds per second with inputs resident in HBM;
e2e = the same through the public API with pinned host inputs and the waveform read back to the host;
roofline = the dominant kernel timed alone with CUDA events against MEASURED_PEAKS.json;
cpu_baseline = the oracle port of the reference's algorithm timed on this box's host cores.
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


def cpu_reference_sample(threads: int):
    """Reference algorithm on host cores: the reference has no KV cache and re-runs the whole prefix every
    step (models/vaura_model.py:502-547).  Bounded sample: the oracle's full-prefix forward
    (oracle/vaura_oracle.py: forward_full, restating llama.py:445-517) at prefix lengths 1/76/152/228,
    integrated piecewise-linearly over the 228 steps of one 2.56 s clip (B=1, fp32)."""
    from oracle import vaura_oracle as vo
    from vaura_b200.synthetic import FULL_SAMPLER, make_avclip_features, make_sampler_state_dict

    torch.set_num_threads(threads)
    oracle = vo.SamplerOracle(make_sampler_state_dict(FULL_SAMPLER, 0), FULL_SAMPLER)
    feats = make_avclip_features(1, 1).reshape(1, 32, 768)
    g = torch.Generator().manual_seed(0)
    seq = torch.randint(0, 1024, (1, 9, 229), generator=g)
    with torch.no_grad():
        oracle.forward_full(seq[..., :8], feats)  # warm-up
        pts = []
        for n in (1, 76, 152, 228):
            t0 = time.perf_counter()
            oracle.forward_full(seq[..., :n], feats)
            pts.append((n, time.perf_counter() - t0))
    total = 0.0
    for (n0, t0), (n1, t1) in zip(pts[:-1], pts[1:]):
        for n in range(n0, n1):
            total += t0 + (t1 - t0) * (n - n0) / (n1 - n0)
    total += pts[-1][1]
    audio = 220 * AUDIO_SEC_PER_TOKEN
    return audio / total, total, pts


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    vals = []
    for _ in range(max(1, args.steps)):
        v, total, pts = cpu_reference_sample(threads)
        vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "generated audio-sec/sec", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * 220 * AUDIO_SEC_PER_TOKEN / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "note": "CPU reference algorithm, throughput per clip is batch-independent"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": "oracle port of the reference's no-KV-cache loop: full-prefix forwards at prefix "
                                   "1/76/152/228 integrated over 228 steps, B=1, fp32"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="b64", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    from tests.test_gpu_parity import build_model
    from vaura_b200 import _cabi
    from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, make_avclip_features
    from vaura_b200.weights import codec_flops, sampler_step_bytes

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    wl = WORKLOADS[args.workload]
    B, T = wl["batch"], wl["T"]
    model = build_model(FULL_SAMPLER, FULL_CODEC, device=str(dev))
    model.seed = 1234
    kw = dict(max_new_tokens=T, use_sampling=wl["use_sampling"], temp=wl["temp"], top_k=wl["top_k"], top_p=0.0,
              cfg_scale=wl["cfg_scale"], prompt_is_encoded=True)
    # rank r owns clips [r*B, (r+1)*B) of every step; features depend only on the clip index
    feats_host = make_avclip_features(B, 2 + rank).pin_memory()
    feats_dev = feats_host.to(dev)
    clip_ids = torch.arange(rank * B, (rank + 1) * B, dtype=torch.int32)
    wav_host = torch.empty(B, 1, T * 512, dtype=torch.float16).pin_memory()

    def step_resident():
        return model.generate(frames=feats_dev, clip_indices=clip_ids, **kw)["generated_audio"]

    def step_e2e():
        f = feats_host.to(dev, non_blocking=True)
        wav = model.generate(frames=f, clip_indices=clip_ids, **kw)["generated_audio"]
        wav_host.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return wav_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local_rank)
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

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peak, peak_src = load_peaks()
    rows = B * (2 if wl["cfg_scale"] > 1.0 else 1)
    d = FULL_SAMPLER
    st = torch.cuda.current_stream().cuda_stream
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    step_bytes = sampler_step_bytes(d)
    if rows <= 4:
        # rows <= 4: the whole decode step is ONE persistent kernel (rows <= 2: decode_step_cluster, 32 clusters x 4
        # CTAs; rows 3-4: decode_step_persistent); timed through the token-only generate (CUDA events around 227
        # back-to-back launches of that kernel + 1 first pass)
        def tokens_only_r():
            model.generate(frames=feats_dev, clip_indices=clip_ids, _decode_audio=False, **kw)
        tokens_only_r()
        k_ms = timed(tokens_only_r, 3) / 3 / (T + 8)
        kv_bytes = 24 * 2 * d.d_model * 4 * rows * (T + 8) / 2  # fp32 KV read, mean context (S/2 positions)
        alg_bytes = step_bytes + kv_bytes
        kern = "decode_step_cluster" if rows <= 2 else "decode_step_persistent"
        kname = f"{kern}<{rows}> (whole decode step: 24 layers + heads + sampling)"
        key = f"{kern}_rows{rows}"
    elif rows <= 64:
        # rows 16..64: the whole decode step is ONE cooperative kernel (decode_step_fused_bf16: rmsnorm, tcgen05 GEMM tiles,
        # attention, device-wide barriers between phases); timed through the token-only generate, which is 228 x (embed
        # kernel + that kernel + sampling kernel) back to back
        def tokens_only_r():
            model.generate(frames=feats_dev, clip_indices=clip_ids, _decode_audio=False, **kw)
        tokens_only_r()
        k_ms = timed(tokens_only_r, 3) / 3 / (T + 8)
        kv_bytes = 24 * 2 * d.d_model * 2 * rows * (T + 8) / 2  # bf16 KV read, mean context (S/2 positions)
        alg_bytes = step_bytes + kv_bytes
        kname = f"decode_step_fused_bf16 ({rows} rows: 24 layers + heads in one launch; embed + sampling kernels included in the time)"
        key = f"decode_step_fused_bf16_rows{rows}"
    else:
        # rows > 64: tcgen05 linear over w1|w3 (25.2 MB of weights per launch), 24 different matrices back to back
        w13 = model.sampler.weights["w13"]
        x = torch.randn(rows, d.d_model, device=dev).to(torch.bfloat16)
        y = torch.empty(rows, 2 * d.ffn_dim, device=dev)
        bn = 64 if rows <= 128 else 128

        def gemm_pass():
            for l in range(d.num_layers):  # 604 MB of distinct weights > L2: nothing is re-read from cache
                _cabi.check(lib.vaura_linear_bf16(x.data_ptr(), w13[l].data_ptr(), y.data_ptr(), rows, 2 * d.ffn_dim,
                                                  d.d_model, bn, st), "vaura_linear_bf16")

        for _ in range(3):
            gemm_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gemm_pass()
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / (5 * d.num_layers)
        alg_bytes = 2 * d.ffn_dim * d.d_model * 2 + rows * d.d_model * 2 + rows * 2 * d.ffn_dim * 4
        kname = f"gemm_tc_kernel<{bn},64,...,EpiLinear> w1|w3 [8192x1536] bf16 x {rows} rows (tcgen05)"
        key = f"gemm_tc_w13_rows{rows}"
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get(key)

    line = {
        "metric": "generated audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": ("bf16 weights, fp32 activations/accumulate" if rows < 16 else "bf16 weights+activations, fp32 accumulate (tcgen05)") + "; codec fp16, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": wl["name"], "per_gpu_batch": B, "tokens_per_clip": T, "decode_steps": T + 8,
                   "l2": "weights 1.39 GB per decode step >> 126 MB L2; no flush needed", "cfg_scale": wl["cfg_scale"],
                   "parallelism": f"dp{world} (clips sharded, no collective in the hot loop)"},
        "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": feats_host.numel() * 4,
                "d2h_bytes_per_step": wav_host.numel() * 2, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"kernel": kname, "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "us_per_launch": k_ms * 1e3, "algorithmic_bytes_per_launch": alg_bytes},
        "decode_step": {"p50_us": None, "weight_bytes": sampler_step_bytes(d)},
        "clocks": clk,
    }
    # decode-step latency: tokens only, no codec
    def tokens_only():
        model.generate(frames=feats_dev, clip_indices=clip_ids, _decode_audio=False, **kw)
    tokens_only()
    ms_tok = timed(tokens_only, 2) / 2
    line["decode_step"]["mean_us"] = ms_tok * 1e3 / (T + 8)
    # p50 of the step-to-step latency: the step kernels (decode_step_cluster, decode_step_fused_bf16) stamp %globaltimer at
    # the start of the launch that samples column `offset` into the workspace (csrc/cabi.cu: ws.timing + 1024)
    try:
        import numpy as np
        ws_buf = model.sampler._buffers["ws"]
        S = T + 9
        st_ns = ws_buf[256 + 8 * 1024:256 + 8 * (1024 + S)].cpu().numpy().view(np.uint64).astype(np.int64)
        d_ns = np.diff(st_ns[2:S])
        d_ns = d_ns[(d_ns > 0) & (d_ns < 10**8)]
        if d_ns.size >= 100:
            line["decode_step"]["p50_us"] = float(np.median(d_ns)) / 1e3
            line["decode_step"]["p99_us"] = float(np.percentile(d_ns, 99)) / 1e3
    except Exception:  # paths without a persistent step kernel leave no stamps
        pass
    line["decode_step"]["hbm_frac_of_measured"] = (sampler_step_bytes(d) / (ms_tok * 1e-3 / (T + 8)) / 1e9) / peak
    line["codec"] = {"ms_per_batch": ms / args.steps - ms_tok, "gflop_per_clip": codec_flops(FULL_CODEC, T) / 1e9}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, total, pts = cpu_reference_sample(threads)
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                "sample": "oracle port of the reference's no-KV-cache loop: full-prefix forwards at "
                                          "prefix 1/76/152/228 integrated over 228 steps, B=1, fp32; "
                                          f"{total:.1f} s per clip"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
