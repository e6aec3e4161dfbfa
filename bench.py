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
def cpu_step_factory(tokens: int, layers: int, rows=None):
    """One utterance per step through the CPU oracle.  `rows`: optional list of (wav [N] f32, ids [T], word ids [T]) host
    tensors — utterances of the batch the GPU arm timed, so the oracle's indices double as the bench's parity check."""
    from taste_spokenlm_b200 import synth
    from oracle import taste_oracle as O                       # checker / CPU baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = synth.FULL if layers == synth.FULL.enc_layers else synth.TowerConfig(enc_layers=layers)
    W = synth.random_weights(cfg, 1234)
    if rows is None:
        b = synth.synth_batch(11, [UTT_SECONDS], [tokens])
        rows = [(b["wav"][0], b["asr_token_ids"][0], b["asr_word_ids"][0])]

    def step(i: int = 0, with_aggregated: bool = False):
        wav, ids, wid = rows[i % len(rows)]
        with torch.no_grad():
            feats, _ = O.log_mel(wav[None])
            out = O.tower_forward(W, ids[None], torch.tensor([ids.shape[0]], dtype=torch.int32), feats, wid[None],
                                  cfg.heads, cfg.enc_layers, cfg.dec_layers, cfg.num_quantizers, cfg.target_hidden_layer,
                                  stages=with_aggregated)
        if with_aggregated:
            return out["quantized_indices"][0], out["_aggregated"][0]
        return out["quantized_indices"][0]
    step.weights = W
    return step


def itemise_misses(W, ref_idx, ref_agg, got_idx):
    """Every token whose indices differ from the oracle's: level of the first divergence and the fp64 margin between the
    two candidate codes on the ORACLE's residual, (d_got - d_ref) / d_ref.  A margin of 1e-4 or less is a near tie that
    the 16-bit operand rounding of the encoder (aggregator error 3.5e-3 bf16 / 4.4e-4 fp16) flips; a pooled multi-token
    word repeats one decision on each of its tokens.  ref_idx / got_idx [R, T, Q], ref_agg [R, T, D]."""
    Win = W["vq.rvq.project_in.weight"].double()
    b_in = W["vq.rvq.project_in.bias"].double()
    Q = ref_idx.shape[-1]
    code = [W[f"vq.rvq.layers.{q}._codebook.embed"][0].double() for q in range(Q)]
    out = []
    for u, t in torch.nonzero((ref_idx != got_idx).any(-1)).tolist():
        q = int(torch.nonzero(ref_idx[u, t] != got_idx[u, t])[0])
        r = ref_agg[u, t].double() @ Win.T + b_in
        for p in range(q):
            r = r - code[p][ref_idx[u, t, p]]
        d_ref = float((r - code[q][ref_idx[u, t, q]]).norm())
        d_got = float((r - code[q][got_idx[u, t, q]]).norm())
        out.append({"utt": u, "t": t, "level": q, "margin_rel": (d_got - d_ref) / d_ref})
    return out


def run_reference_arm(args, rank):
    if rank != 0:
        return
    step = cpu_step_factory(args.tokens, args.layers)
    for _ in range(args.warmup):
        step(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(0)
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
# BASELINE config 5: the tokenizer's share of an audio-conditioned inference_completion prompt (MT:1664-1704)
# ---------------------------------------------------------------------------------------------------------------
def measure_config5(tower, fe, batch, dev, args):
    """B = 1 latency of the tokenizer call as `inference_completion` makes it: host tensors in (PT:246-247 leaves the
    processor's outputs on the CPU, scripts/generate_audio.py:186-190 moves them), the drop-in `WhisperFrontend.forward`
    + the patched `TasteForCausalLM.extract_vq` (tower forward incl. its state-key check and the token-length sync,
    llm-token mapping), llm indices back on the host.  Wall clock around each call, median of 30 after 5 warm-ups.

    Beside it, the LM side of the same prompt measured in this run on the same GPU: the reference's spoken LM is a
    Llama-3.2-1B (`text_config` of configs/model/taslm.json) driven by `TasteSpokenLM.generate` (MT:1111-1117,
    1196-1199), which re-forwards the WHOLE growing `inputs_embeds` on every step with no KV cache.  It is timed here as
    HF transformers' stock `LlamaModel` (random init of that geometry, bf16, sdpa) + lm_head on a prompt of L llm
    tokens followed by `new_tokens` re-forwards of L + k tokens; bridge, sampler and speech decoder are not included, so
    the LM figure is a lower bound and the tokenizer share an upper bound."""
    from taste_spokenlm_b200 import tower as T

    class Model:                                   # stands in for TasteForCausalLM: extract_vq touches .audio_tower only
        extract_vq = T.extract_vq

    m = Model()
    m.audio_tower = tower
    Tn = args.tokens
    wav_h = batch["wav"][:1].cpu()
    ns_h = torch.tensor([480000], dtype=torch.int32)
    ids_h, wid_h = batch["ids"][:1].cpu(), batch["wid"][:1].cpu()
    len_h = torch.tensor([Tn], dtype=torch.int32)
    n_words = int(wid_h.max()) + 1
    g = torch.Generator().manual_seed(5)
    pieces = torch.randint(1, 3, (n_words,), generator=g)
    lwid_h = torch.repeat_interleave(torch.arange(n_words, dtype=torch.int32), pieces)[None]
    L = lwid_h.shape[1]
    llen_h = torch.tensor([L], dtype=torch.int32)
    lids_h = torch.randint(0, 128256, (1, L), generator=g)
    flen_h = torch.tensor([3000], dtype=torch.int32)

    def call():
        feats, _ = fe(wav_h.to(dev, non_blocking=True), ns_h)                       # WhisperFrontend.forward (WF:87-113)
        _, llm_idx = m.extract_vq(ids_h.to(dev), len_h.to(dev), wid_h.to(dev), lids_h.to(dev), llen_h.to(dev),
                                  lwid_h.to(dev), feats, flen_h.to(dev))            # MT:1695-1704
        return llm_idx.cpu()

    for _ in range(5):
        out = call()
    ts = []
    for _ in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = call()
        ts.append((time.perf_counter() - t0) * 1e3)
    tok_ms = statistics.median(ts)
    res = {"tokenizer_b1_e2e_ms": tok_ms, "tokenizer_b1_e2e_ms_min": min(ts), "asr_tokens": Tn, "llm_tokens": int(L),
           "api": "host tensors -> WhisperFrontendB200.forward -> patched TasteForCausalLM.extract_vq -> llm indices on host"}
    try:
        from transformers import LlamaConfig, LlamaForCausalLM
        cfg = LlamaConfig(hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32,
                          num_key_value_heads=8, head_dim=64, vocab_size=128256, rms_norm_eps=1e-5, rope_theta=500000.0,
                          max_position_embeddings=131072, tie_word_embeddings=True)
        with torch.device(dev):
            lm = LlamaForCausalLM(cfg).to(torch.bfloat16).eval()
        new_tokens = 48                                   # extra_words = 32 (MT:1670) at ~1.5 llm tokens per word
        emb = torch.randn(1, L + new_tokens, 2048, device=dev, dtype=torch.bfloat16) * 0.02

        def prefill():
            return lm.lm_head(lm.model(inputs_embeds=emb[:, :L], use_cache=False).last_hidden_state)

        def generate_like_reference():
            for k in range(new_tokens):                   # MT:1111-1117: whole sequence re-forwarded every step
                lm.lm_head(lm.model(inputs_embeds=emb[:, : L + k], use_cache=False).last_hidden_state)

        def timed(fn, n):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3 / n

        def generate_kv_cached():                     # taste_spokenlm_b200/generate.py: prompt once, then one position per step
            out = lm.model(inputs_embeds=emb[:, :L], use_cache=True)
            past = out.past_key_values
            lm.lm_head(out.last_hidden_state[:, -1:])
            for k in range(1, new_tokens):
                out = lm.model(inputs_embeds=emb[:, L + k - 1: L + k], past_key_values=past, use_cache=True)
                past = out.past_key_values
                lm.lm_head(out.last_hidden_state[:, -1:])

        from taste_spokenlm_b200.generate import _GraphedDecoder
        graphed = {}

        def generate_kv_cached_graph():                # + the one-position step replayed from a CUDA graph (static KV cache)
            dec = graphed.get("dec")
            if dec is None:
                dec = graphed["dec"] = _GraphedDecoder(lm.model, dev, torch.bfloat16, L + new_tokens + 8)
            out = dec.prefill(emb[:, :L])
            lm.lm_head(out.last_hidden_state[:, -1:])
            for k in range(1, new_tokens):
                out = dec.step(emb[:, L + k - 1: L + k])
                lm.lm_head(out.last_hidden_state[:, -1:])

        pre_ms = timed(prefill, 5)
        gen_ms = timed(generate_like_reference, 2)
        gen_kv_ms = timed(generate_kv_cached, 2)
        try:
            gen_graph_ms = timed(generate_kv_cached_graph, 3)
            res.update({"lm_generate_kv_cached_cuda_graph_ms": gen_graph_ms,
                        "tokenizer_share_of_completion_kv_cached_cuda_graph": tok_ms / (tok_ms + gen_graph_ms)})
        except Exception as e:  # noqa: BLE001
            res["lm_graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        res.update({"lm_prefill_ms": pre_ms, "lm_generate_ms": gen_ms, "lm_new_tokens": new_tokens,
                    "lm_generate_kv_cached_ms": gen_kv_ms,
                    "tokenizer_share_of_completion_kv_cached": tok_ms / (tok_ms + gen_kv_ms),
                    "lm": "HF transformers LlamaForCausalLM, random-init Llama-3.2-1B geometry (taslm.json text_config), bf16, "
                          "no KV cache as MT:1111-1117; measured in this run on the same GPU; bridge / sampler / speech "
                          "decoder excluded",
                    "tokenizer_share_of_prompt": tok_ms / (tok_ms + pre_ms),
                    "tokenizer_share_of_completion": tok_ms / (tok_ms + gen_ms)})
        del lm
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        res["lm_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    return res


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs 3 / 4: ragged corpus through the real corpus driver (strong scaling)
# ---------------------------------------------------------------------------------------------------------------
def run_corpus(args, rank, world, local_rank):
    import shutil
    import tempfile
    import torch.distributed as dist
    from taste_spokenlm_b200 import _lib, shard, synth
    from taste_spokenlm_b200.tower import TasteAudioTowerB200
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    cfg = synth.FULL if args.layers == synth.FULL.enc_layers else synth.TowerConfig(enc_layers=args.layers)
    torch.set_grad_enabled(False)
    tower = TasteAudioTowerB200.from_config(cfg).eval()
    tower.load_state_dict(synth.random_weights(cfg, 1234), strict=True)
    tower = tower.to(dev)
    eng = tower.engine()
    corpus = synth.SynthCorpus(args.utts, seed=4, pool=args.pool)
    out_dir = tempfile.mkdtemp(prefix=f"taste_corpus_r{rank}_")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: a small corpus of the same distribution through the same driver (allocations, first-launch costs)
    warm = synth.SynthCorpus(args.batch * 3 * world, seed=5, pool=2)
    shard.tokenize_corpus(eng, warm, world, rank, batch_size=args.batch, writer=shard.ShardWriter(out_dir + "/warm", rank),
                          gather=True)
    if world > 1:
        # ... and one gather of the real job's message size: NCCL sets up its large-message channels on first use
        n_u = args.utts // world + 1
        n_r = int(sum(corpus.token_counts)) // world + 1
        shard.gather_indices(torch.zeros(n_u, 2, dtype=torch.int32, device=dev),
                             torch.zeros(n_r, cfg.num_quantizers, dtype=torch.int16, device=dev), cfg.num_quantizers, dev)
    barrier()
    launches0 = lib.taste_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    tm = {}
    writer = shard.ShardWriter(out_dir, rank, flush_every=2048)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    got = shard.tokenize_corpus(eng, corpus, world, rank, batch_size=args.batch, writer=writer, gather=True, timings=tm)
    e1.record()
    barrier()
    t_wall1 = time.time()
    t_fin0 = time.perf_counter()
    writer.finalize()
    fin_ms = (time.perf_counter() - t_fin0) * 1e3
    launches = int(lib.taste_launch_count() - launches0)
    my_ms = e0.elapsed_time(e1)
    all_ms = torch.tensor([my_ms], device=dev)
    stats = torch.tensor([tm["fetch_ms"], tm["stall_ms"], tm["result_wait_ms"], tm["writer_ms"], tm["gather_ms"],
                          float(tm["batches"]), float(tm["h2d_bytes"]), float(tm["d2h_bytes"])], device=dev, dtype=torch.float64)
    if world > 1:
        gl = [torch.zeros_like(all_ms) for _ in range(world)]
        dist.all_gather(gl, all_ms)
        rank_ms = [float(t.item()) for t in gl]
        sl = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(sl, stats)
        stats_all = torch.stack(sl).cpu()
    else:
        rank_ms = [my_ms]
        stats_all = stats[None].cpu()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ok = len(got) == args.utts and all(int(t.shape[0]) == corpus.token_counts[u] for u, t in got[:: max(1, args.utts // 256)])
    shutil.rmtree(out_dir, ignore_errors=True)
    if rank == 0:
        ms = max(rank_ms)
        nb = stats_all[:, 5]
        per_batch = lambda col: float((stats_all[:, col] / nb.clamp_min(1)).max())      # noqa: E731
        line = {
            "metric": "audio-sec tokenized/sec (ragged 1-30 s corpus)", "value": corpus.audio_seconds / (ms / 1e3), "unit": UNIT,
            "windows_audio_s_per_s": args.utts * UTT_SECONDS / (ms / 1e3), "utt_per_s": args.utts / (ms / 1e3),
            "n_gpus": world, "steps": int(nb.max()), "warmup": 3, "ms_per_step": ms / max(float(nb.max()), 1.0),
            "ms_total": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"configs[2]/[3]: {args.utts} utterances, durations U[1,30] s (mean {corpus.audio_seconds / args.utts:.1f} s), "
                                   f"T = clip(round(2.7 dur + N(0,2)), 1, 443) (mean {np.mean(corpus.token_counts):.1f}), seed 4; "
                                   f"length-bucketed batches of {args.batch}; host-resident audio, double-buffered pinned H2D, "
                                   f"tokenize + llm-token mapping, ShardWriter (Arrow, load_from_disk layout), final all-gather",
                       "geometry": f"distil-large-v3 tower: d_model 1280, {args.layers} encoder layers, random init",
                       "parallelism": f"dp{world} by utterance (shard_indices round-robin inside length buckets)",
                       "note": "every utterance costs one full 30 s encoder window (SURVEY 0.3): windows/s is the rate the GPU "
                               "sees, real audio-s/s the rate the corpus sees"},
            "rank_ms": {"max": max(rank_ms), "min": min(rank_ms)},
            "host_ms_per_batch_max_over_ranks": {"fetch (worker thread, overlapped)": per_batch(0), "stall (compute thread waits for worker)": per_batch(1),
                                                  "result_wait": per_batch(2), "writer": per_batch(3)},
            # max over ranks = the fastest rank waiting for the slowest at the collective (GPUs of one box differ by ~2 %
            # under the power cap); min over ranks = what the collective + unpacking cost the rank that arrived last
            "gather_ms": float(stats_all[:, 4].max()), "gather_ms_min_over_ranks": float(stats_all[:, 4].min()),
            "writer_ms_total": float(stats_all[:, 3].max()), "finalize_ms": fin_ms,
            "e2e": {"value": corpus.audio_seconds / (ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": float((stats_all[:, 6] / nb.clamp_min(1)).mean()),
                    "d2h_bytes_per_step": float((stats_all[:, 7] / nb.clamp_min(1)).mean()),
                    "api": "shard.tokenize_corpus (host corpus in, Arrow shards + gathered indices out)"},
            "gpu_launches": launches, "clocks": clocks, "complete": bool(ok),
        }
        if args.layers != 32:
            line["INVALID"] = "debug run with a reduced layer count; not the named config"
        emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-ragged", action="store_true", help="skip the short config-3 (ragged corpus) measurement")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"],
                    help="library flavour of the headline arm (bf16 = BASELINE config 2's dtype)")
    ap.add_argument("--no-alt-precision", action="store_true", help="skip the second, other-flavour timing of the same step")
    ap.add_argument("--workload", default="fixed", choices=["fixed", "corpus"],
                    help="fixed = configs[1] (the headline); corpus = configs[2]/[3] through shard.tokenize_corpus (strong scaling)")
    ap.add_argument("--utts", type=int, default=16384, help="corpus workload: utterances in the whole job")
    ap.add_argument("--pool", type=int, default=64, help="corpus workload: distinct decoded waveforms held on the host")
    ap.add_argument("--encoder-mode", type=int, default=0, help="taste_encoder_set_mode (A/B runs): 0 default, 1 no LayerNorm folding, 2 fold both")
    args = ap.parse_args()

    if args.layers <= 6:
        raise SystemExit("--layers must exceed 6: the aggregator's values are the state entering encoder layer 6 (JES:192-193)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.workload == "corpus":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        run_corpus(args, rank, world, local_rank)
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
    lib = _lib.load(args.precision)
    _lib.check(lib.taste_encoder_set_mode(args.encoder_mode), "taste_encoder_set_mode", lib)

    cfg = synth.FULL if args.layers == synth.FULL.enc_layers else synth.TowerConfig(enc_layers=args.layers)
    torch.set_grad_enabled(False)
    weights = synth.random_weights(cfg, 1234)
    tower = TasteAudioTowerB200.from_config(cfg, precision=args.precision).eval()
    tower.load_state_dict(weights, strict=True)
    tower = tower.to(dev)
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True, precision=args.precision).to(dev)
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
    prof = _lib.prof_collect(lib)
    lib.taste_prof_reset()

    # ---- the same step in the other library flavour (same kernels, other 16-bit operand type) ----------------------
    alt = None
    alt_idx_head = None
    if not args.no_alt_precision:
        other = "fp16" if args.precision == "bf16" else "bf16"
        tower2 = TasteAudioTowerB200.from_config(cfg, precision=other).eval()
        tower2.load_state_dict(weights, strict=True)
        tower2 = tower2.to(dev)
        eng2 = tower2.engine()

        def step2():
            return eng2.tokenize_device(batch["wav"], batch["n_samples"], batch["ids"], batch["wid"], batch["lengths_host"])

        for _ in range(max(args.warmup, 3)):
            _, idx2 = step2()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            _, idx2 = step2()
        a1.record()
        barrier()
        ms_alt = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_alt, op=dist.ReduceOp.MAX)
        alt = {"dtype": other, "value": world * B * UTT_SECONDS * args.steps / (float(ms_alt.item()) / 1e3), "unit": UNIT,
               "ms_per_step": float(ms_alt.item()) / args.steps,
               "index_agreement_with_headline_arm": float((idx2 == idx).float().mean()),
               "note": "same workload and step, library flavour with the other 16-bit operand type "
                       "(libtaste_b200_f16.so = fp16 operands: the reference's own autocast dtype, JES:133)"}
        alt_idx_head = idx2[:8].cpu()
        del tower2, eng2
        torch.cuda.empty_cache()

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

    config5 = None
    if rank == 0 and not args.no_e2e and not args.no_config5:
        config5 = measure_config5(tower, fe, batch, dev, args)

    # ---- BASELINE config 3 in short: 512 ragged utterances through the corpus driver on this rank (full job: --workload corpus)
    ragged = None
    if rank == 0 and not args.no_e2e and not args.no_ragged and args.batch >= 8:
        try:
            from taste_spokenlm_b200 import shard
            rc = synth.SynthCorpus(8 * args.batch, seed=4, pool=8)
            shard.tokenize_corpus(eng, synth.SynthCorpus(2 * args.batch, seed=5, pool=2), 1, 0, batch_size=args.batch, gather=False)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            got = shard.tokenize_corpus(eng, rc, 1, 0, batch_size=args.batch, gather=False)
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1)
            ragged = {"utterances": rc.n, "mean_duration_s": rc.audio_seconds / rc.n, "mean_tokens": float(np.mean(rc.token_counts)),
                      "real_audio_s_per_s": rc.audio_seconds / (rms / 1e3), "window_audio_s_per_s": rc.n * UTT_SECONDS / (rms / 1e3),
                      "ms_per_batch": rms / (rc.n / args.batch), "complete": len(got) == rc.n,
                      "api": "shard.tokenize_corpus: host-resident ragged audio (durations U[1,30] s), pinned H2D pipeline, one D2H per "
                             "batch; every utterance still costs one 30 s encoder window (SURVEY 0.3)"}
        except Exception as e:  # noqa: BLE001
            ragged = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = world * B * UTT_SECONDS * args.steps / (ms_total / 1e3)
    stages = []
    tot_kernel_ms = sum(p["total_ms"] for p in prof) or 1.0
    for p in sorted(prof, key=lambda r: -r["total_ms"]):
        tensor = p["name"].startswith("gemm") or p["name"].startswith("attention") or p["name"].startswith("rvq_encode")
        sec = p["total_ms"] / 1e3
        if tensor:
            ach, peak, unit = p["flops"] / sec / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = p["bytes"] / sec / 1e9, peaks["hbm_gbs"], "GB/s"
        extra = {}
        if p["name"] in ("attention_encoder", "attention_tcgen05"):
            # head_dim 64: 16 ex2 per clock per SM against 32 scores per clock for the tensor pipe; with 4 of every 16
            # exponentials on the FMA pipe the MUFU caps the kernel at 2/3 of the tensor peak
            extra = {"ceiling": "mufu_ex2 (16/clk/SM, 12 of 16 exponentials)", "ceiling_frac_of_tensor_peak": 2.0 / 3.0,
                     "frac_of_ceiling": ach / (peak * 2.0 / 3.0)}
        stages.append({**extra, "kernel": p["name"], "launches_per_step": p["launches"] / args.steps,
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
                    "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from the "
                                      "committed ncu --set full capture (scripts/gpu_ncu.sh), not re-measured in this run",
                    "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}), sustained bf16 (kernel timed inside a long step)"
                    if top["bound"] == "tensor" else f"MEASURED_PEAKS.json ({peaks['source']}) hbm_gbs",
                    "avg_launch_ms": top["ms_per_step"] / max(top["launches_per_step"], 1e-9)}
    flops_per_utt = sum(p["flops"] for p in prof) / (args.steps * B)

    cpu_baseline = None
    parity_check = None
    if world == 1 and not args.no_cpu_baseline:
        # the CPU leg runs utterances 0 .. reps of the batch that was just timed: its indices are the parity check of
        # the timed output (the timed batch itself was never compared with the oracle)
        reps = min(8, B)
        rows = [(batch["wav"][i].cpu(), batch["ids"][i].cpu(), batch["wid"][i].cpu()) for i in range(reps)]
        step = cpu_step_factory(args.tokens, args.layers, rows)
        ref_rows = [step(0, True)]                             # warm-up, kept as a checked row
        t0 = time.perf_counter()
        for i in range(1, reps):
            ref_rows.append(step(i, True))
        if reps == 1:
            step(0, True)
        dt = (time.perf_counter() - t0) / max(reps - 1, 1)
        cores = os.cpu_count() or 1
        cpu_baseline = {"value": UTT_SECONDS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{max(reps - 1, 1)} x 1 utterance (30 s, {args.tokens} tokens) of the timed batch after 1 warm-up, "
                                  f"fp32 oracle (torch CPU, {cores} threads), {dt:.2f} s per utterance"}
        got = idx[:reps].cpu()
        ref = torch.stack([r[0] for r in ref_rows])
        ref_agg = torch.stack([r[1] for r in ref_rows])
        lvl = [float((got[..., q] == ref[..., q]).float().mean()) for q in range(ref.shape[-1])]
        if alt is not None and alt_idx_head is not None:
            alt["parity_check_index_agreement"] = float((alt_idx_head[:reps] == ref).float().mean())
            alt["parity_check_misses"] = itemise_misses(step.weights, ref, ref_agg, alt_idx_head[:reps])
            alt["parity_check_max_margin_rel"] = max([m["margin_rel"] for m in alt["parity_check_misses"]], default=0.0)
        parity_check = {"utterances": reps, "n_indices": int(ref.numel()), "index_agreement": float((got == ref).float().mean()),
                        "per_level": lvl, "misses": itemise_misses(step.weights, ref, ref_agg, got),
                        "misses_note": "differing tokens: level of the first divergence and the fp64 margin "
                                       "(d_got - d_ref) / d_ref between the two codes on the oracle's residual",
                        "against": f"fp32 CPU oracle on utterances 0..{reps - 1} of the timed batch (device-resident arm; "
                                   "the e2e arm is asserted bit-equal to it)"}
        parity_check["max_margin_rel"] = max([m["margin_rel"] for m in parity_check["misses"]], default=0.0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "config": workload_config(args, world),
        "utt_per_s": value / UTT_SECONDS, "per_gpu": value / world,
        "tflops_per_gpu": flops_per_utt * (value / UTT_SECONDS / world) / 1e12,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "stages": stages,
        "cpu_baseline": cpu_baseline, "parity_check": parity_check, "alt_precision": alt,
        "latency_b1_ms": lat_ms, "config5": config5, "ragged_config3": ragged,
    }
    if args.layers != 32:
        line["INVALID"] = "debug run with a reduced layer count; not the named config"
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
