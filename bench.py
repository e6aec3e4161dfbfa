#!/usr/bin/env python
"""bench.py — audio-seconds tokenized per second on the TASTE speech-tokenization hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]                      # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     # the CPU arm (oracle port), rank 0 only

One "step" = one pass of the whole path (16 kHz waveform -> log-mel -> Whisper encoder -> aggregator -> word pooling
-> RVQ indices) over one batch of synthetic input.  Workload at every N: BASELINE.json configs[1] — batch 64 x 30 s
utterances (480 000 samples, 3000 mel frames), 64 transcript tokens each, random-init distil-large-v3 geometry, bf16
tensor-core GEMMs with fp32 accumulation / residual stream / RVQ.  Every rank owns its own batch (weak scaling, data
parallel by utterance); the only collective is an all-gather of per-rank token counts and indices after the last step.

Printed keys (one JSON line from rank 0): see the task contract; `value` = device-resident throughput, `e2e` = the
same metric through the drop-in modules with host (pinned) buffers and the H2D / D2H copies inside the timed region,
`roofline` = the dominant kernel measured live with CUDA events on the launching stream, `stages` = the same for every
kernel class, `cpu_baseline` = the CPU oracle timed on this box's host cores on one utterance of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "audio-sec tokenized/sec (30 s utts)"
UNIT = "audio-s/s"
UTT_SECONDS = 30.0
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {k: float(d[k]) for k in FALLBACK_PEAKS if k in d} | {"source": "measured"}
        except Exception:
            pass
    return dict(FALLBACK_PEAKS, source="fallback")


# ---------------------------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------------------------
def make_batch(seed: int, batch: int, tokens: int, device):
    """Speech-like waveforms generated on the device from a seed (8 AM sinusoids + noise), plus transcripts."""
    from taste_spokenlm_b200 import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = 480000
    t = torch.arange(n, device=device, dtype=torch.float32) / 16000.0
    wav = 0.01 * torch.randn(batch, n, device=device, generator=g)
    f = 80.0 + 7520.0 * torch.rand(batch, 8, device=device, generator=g)
    ph = 6.2831853 * torch.rand(batch, 8, device=device, generator=g)
    am = 0.5 + 4.5 * torch.rand(batch, 8, device=device, generator=g)
    amp = 0.02 + 0.1 * torch.rand(batch, 8, device=device, generator=g)
    for i in range(8):
        env = 0.5 * (1.0 + torch.sin(6.2831853 * am[:, i:i + 1] * t + ph[:, i:i + 1]))
        wav += amp[:, i:i + 1] * env * torch.sin(6.2831853 * f[:, i:i + 1] * t + ph[:, (i + 3) % 8:(i + 3) % 8 + 1])
    ids, wids = zip(*[synth.synth_transcript(seed * 104729 + b, tokens, tokens) for b in range(batch)])
    return {
        "wav": wav.contiguous(),
        "n_samples": torch.full((batch,), n, dtype=torch.int32, device=device),
        "ids": torch.stack(ids).to(device), "wid": torch.stack(wids).to(device),
        "lengths_host": np.full(batch, tokens, dtype=np.int64),
    }


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (one utterance of the same workload per step)
# ---------------------------------------------------------------------------------------------------------------
def cpu_step_factory(tokens: int, layers: int):
    from taste_spokenlm_b200 import synth
    from oracle import taste_oracle as O                       # checker / CPU baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = synth.FULL if layers == synth.FULL.enc_layers else synth.TowerConfig(enc_layers=layers)
    W = synth.random_weights(cfg, 1234)
    b = synth.synth_batch(11, [UTT_SECONDS], [tokens])

    def step():
        with torch.no_grad():
            feats, _ = O.log_mel(b["wav"])
            out = O.tower_forward(W, b["asr_token_ids"], b["asr_token_lengths"], feats, b["asr_word_ids"], cfg.heads,
                                  cfg.enc_layers, cfg.dec_layers, cfg.num_quantizers, cfg.target_hidden_layer)
        return out["quantized_indices"]
    return step


def run_reference_arm(args, rank):
    if rank != 0:
        return
    step = cpu_step_factory(args.tokens, args.layers)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * UTT_SECONDS / dt
    cores = os.cpu_count() or 1
    sample = f"1 utterance (30 s, {args.tokens} tokens) of the batch-64 workload per step, fp32, torch CPU, {cores} threads"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1, cpu=True),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def workload_config(args, world, cpu=False):
    return {
        "workload": f"configs[1]: batch {args.batch} x 30 s utterances (480000 samples, 3000 mel frames), "
                    f"{args.tokens} transcript tokens each, per GPU",
        "geometry": f"distil-large-v3 tower: d_model 1280, {args.layers} encoder layers, 2 aggregator layers, RVQ 4x512x256, random init",
        "batch_per_gpu": 1 if cpu else args.batch, "tokens_per_utt": args.tokens, "parallelism": f"dp{world} by utterance",
        "l2": "working set (>= 1.9 GB of activations per step) exceeds the 126 MB L2; no explicit flush",
    }


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _claim_stdout():
    """Everything native libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the single JSON line is
    written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--layers", type=int, default=32, help="encoder layers (32 = the named config; others are for debugging and flagged)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--encoder-mode", type=int, default=0, help="taste_encoder_set_mode (A/B runs): 0 default, 1 no LayerNorm folding, 2 fold both")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from taste_spokenlm_b200 import _lib, synth
    from taste_spokenlm_b200.tower import TasteAudioTowerB200
    from taste_spokenlm_b200.frontend import WhisperFrontendB200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.taste_encoder_set_mode(args.encoder_mode), "taste_encoder_set_mode")

    cfg = synth.FULL if args.layers == synth.FULL.enc_layers else synth.TowerConfig(enc_layers=args.layers)
    torch.set_grad_enabled(False)
    tower = TasteAudioTowerB200.from_config(cfg).eval()
    tower.load_state_dict(synth.random_weights(cfg, 1234), strict=True)
    tower = tower.to(dev)
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True).to(dev)
    eng = tower.engine()

    batch = make_batch(1000 + rank, args.batch, args.tokens, dev)
    B, T = args.batch, args.tokens

    def step_device():
        return eng.tokenize_device(batch["wav"], batch["n_samples"], batch["ids"], batch["wid"], batch["lengths_host"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        qz, idx = step_device()
    barrier()
    lib.taste_prof_reset()
    lib.taste_prof_enable(1)
    launches0 = lib.taste_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        qz, idx = step_device()
    if world > 1:
        # the path's only collective: per-rank token counts, then the packed indices
        counts = torch.tensor([int(batch["lengths_host"].sum())], device=dev, dtype=torch.int64)
        all_counts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_counts, counts)
        packed = idx.to(torch.int16).reshape(-1).view(torch.uint8)       # NCCL has no int16: move the raw bytes
        all_idx = torch.empty(world * packed.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(all_idx, packed)
    e1.record()
    barrier()
    t_wall1 = time.time()
    lib.taste_prof_enable(0)
    launches = int(lib.taste_launch_count() - launches0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    prof = _lib.prof_collect()
    lib.taste_prof_reset()

    # ---- end-to-end timing through the drop-in modules with host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        h_wav = batch["wav"].cpu().pin_memory()
        h_ids = batch["ids"].cpu().pin_memory()
        h_wid = batch["wid"].cpu().pin_memory()
        h_len = torch.full((B,), T, dtype=torch.int32).pin_memory()
        h_ns = batch["n_samples"].cpu().pin_memory()
        feat_len = torch.full((B,), 3000, dtype=torch.int32, device=dev)

        # Double-buffered ingest: the H2D copy of step i+1's inputs runs on a copy stream under
        # step i's kernels.  Every step's inputs still cross PCIe inside the timed region (K copies for K steps).
        copy_stream = torch.cuda.Stream(device=dev)
        host = (h_wav, h_ids, h_wid, h_len, h_ns)
        dev_bufs = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def h2d(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])            # the previous user of this slot has finished
                for d_t, h_t in zip(dev_bufs[slot], host):
                    d_t.copy_(h_t, non_blocking=True)
                ready[slot].record(copy_stream)

        def compute(slot):
            torch.cuda.current_stream(dev).wait_event(ready[slot])
            wav, ids, wid, lens, ns = dev_bufs[slot]
            _, feats = fe.forward_device(wav, ns, want_f32=False, want_bf16=True)         # WhisperFrontend (WF:87-113)
            out = tower(ids, lens, feats, feat_len, asr_word_ids=wid)                     # TasteAudioTower.forward
            consumed[slot].record(torch.cuda.current_stream(dev))
            return out["quantized_indices"].cpu(), out["audio_unit_lengths"].cpu()        # XV:39-40

        def run_e2e(n_steps):
            h2d(0)
            res = None
            for i in range(n_steps):
                if i + 1 < n_steps:
                    h2d((i + 1) & 1)
                res = compute(i & 1)
            return res

        for c in consumed:
            c.record(torch.cuda.current_stream(dev))
        r_idx, r_len = run_e2e(max(args.warmup, 3))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        r_idx, r_len = run_e2e(args.steps)
        f1.record()
        barrier()
        ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e_ms = float(ms2.item())
        assert torch.equal(r_idx.to(dev), idx), "e2e and device-resident paths disagree"
        h2d = sum(t.numel() * t.element_size() for t in (h_wav, h_ids, h_wid, h_len, h_ns))
        d2h = r_idx.numel() * r_idx.element_size() + r_len.numel() * r_len.element_size()
        e2e = {"value": world * B * UTT_SECONDS * args.steps / (e2e_ms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
               "api": "WhisperFrontendB200.forward_device + TasteAudioTowerB200.forward (pinned host in, host indices out; "
                      "H2D of step i+1 double-buffered under step i)"}

    # ---- single-utterance latency (BASELINE config 5's tokenizer call: B = 1, 30 s window, same T) ------------------
    lat_ms = None
    if rank == 0 and not args.no_e2e:
        one = {k: (v[:1].contiguous() if torch.is_tensor(v) else v[:1]) for k, v in batch.items()}
        for _ in range(3):
            eng.tokenize_device(one["wav"], one["n_samples"], one["ids"], one["wid"], one["lengths_host"])
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            eng.tokenize_device(one["wav"], one["n_samples"], one["ids"], one["wid"], one["lengths_host"])
        g1.record()
        torch.cuda.synchronize()
        lat_ms = g0.elapsed_time(g1) / 10

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = world * B * UTT_SECONDS * args.steps / (ms_total / 1e3)
    stages = []
    tot_kernel_ms = sum(p["total_ms"] for p in prof) or 1.0
    for p in sorted(prof, key=lambda r: -r["total_ms"]):
        tensor = p["name"].startswith("gemm") or p["name"].startswith("attention")
        sec = p["total_ms"] / 1e3
        if tensor:
            ach, peak, unit = p["flops"] / sec / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = p["bytes"] / sec / 1e9, peaks["hbm_gbs"], "GB/s"
        stages.append({"kernel": p["name"], "launches_per_step": p["launches"] / args.steps,
                       "ms_per_step": p["total_ms"] / args.steps, "share": p["total_ms"] / tot_kernel_ms,
                       "bound": "tensor" if tensor else "hbm", "achieved": ach, "peak": peak, "unit": unit,
                       "frac": ach / peak})
    top = stages[0] if stages else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if top and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(top["kernel"])
        except Exception:
            traffic = None
    roofline = None
    if top:
        roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
                    "unit": top["unit"], "frac": top["frac"], "traffic": traffic,
                    "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}), sustained bf16 (kernel timed inside a long step)"
                    if top["bound"] == "tensor" else f"MEASURED_PEAKS.json ({peaks['source']}) hbm_gbs",
                    "avg_launch_ms": top["ms_per_step"] / max(top["launches_per_step"], 1e-9)}
    flops_per_utt = sum(p["flops"] for p in prof) / (args.steps * B)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        step = cpu_step_factory(args.tokens, args.layers)
        step()
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            ref_idx = step()
        dt = (time.perf_counter() - t0) / reps
        cores = os.cpu_count() or 1
        cpu_baseline = {"value": UTT_SECONDS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{reps} x 1 utterance (30 s, {args.tokens} tokens) of the batch-64 workload after 1 warm-up, "
                                  f"fp32 oracle (torch CPU, {cores} threads), {dt:.2f} s per utterance"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
        "utt_per_s": value / UTT_SECONDS, "per_gpu": value / world,
        "tflops_per_gpu": flops_per_utt * (value / UTT_SECONDS / world) / 1e12,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "stages": stages,
        "cpu_baseline": cpu_baseline,
        "latency_b1_ms": lat_ms,
    }
    if args.layers != 32:
        line["INVALID"] = "debug run with a reduced layer count; not the named config"
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
