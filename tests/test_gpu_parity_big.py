"""Statistically meaningful end-to-end parity at FULL geometry (-m gpu): >= 2000 tokens against outputs of the REAL
reference (tests/golden/tower_full_b34.npz: one reference forward over 34 utterances, 28 of config 2's shape + 6 ragged;
tests/golden/make_golden.py::tower_big_case).

Two batch paths of the product are measured against the same fixture:
  * batched  — all 34 utterances in one call: CTA-pair GEMMs, LayerNorm folded into the QKV GEMM, persistent attention walk;
  * solo     — every utterance alone (B = 1): the config-5 path.

For each path the report (gpurun_out/r2_parity_report.json -> profiles/r2_parity_report.json) holds per-level and
per-token agreement, the entropy of the indices, plain and centred relative L2 of the aggregator output, and an
ITEMISED list of every token whose codes differ: the level of the first divergence, the two candidate codes, the fp64
margin between them on the reference's own residual, the size of the product's aggregator error on that token, and
whether the flip is the one that error predicts (`explained`).  Levels after the first divergence quantise a different
residual and are consequences, not independent misses.

Both library flavours are measured (`precision=`): bf16 operands (BASELINE config 2) and fp16 operands (the reference's
own GPU dtype under autocast; same tensor-core rate, 8x smaller operand rounding).

Assertions: every miss must be explained by the measured aggregator error (a miss that is not would be an RVQ kernel
defect, not operand rounding upstream); aggregator error <= 1e-2 (north star); index agreement >= the floor below
(>= 99.5 % for fp16).
"""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from taste_spokenlm_b200 import synth
from taste_spokenlm_b200.frontend import WhisperFrontendB200
from taste_spokenlm_b200.tower import TasteAudioTowerB200

torch.set_grad_enabled(False)
REPORT_PATH = "gpurun_out/r2_parity_report.json"
# Element-wise index agreement floors = measured value minus a small margin (profiles/r2_parity_report.json).
#   bf16 operands (BASELINE config 2's dtype): 0.988 batched / 0.983 solo over 8272 indices.  Every one of the 60-75
#       missed tokens is a near tie (fp64 margin between the two codes <= 1.6e-4 of the distance) flipped by a 3e-3
#       aggregator error, i.e. this IS the bf16 operand-rounding floor (SURVEY section 7 hard part 3), itemised in the report.
#   fp16 operands (the reference's own autocast dtype; same tensor-core rate): the >= 99.5 % bar of the north star.
FLOOR = {"bf16": {"batched": 0.98, "solo": 0.975}, "fp16": {"batched": 0.995, "solo": 0.995}}
AGG_TOL = {"bf16": 1e-2, "fp16": 2e-3}


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _entropy_bits(codes: np.ndarray, k: int = 512) -> float:
    c = np.bincount(codes.astype(np.int64), minlength=k).astype(np.float64)
    p = c[c > 0] / c.sum()
    return float(-(p * np.log2(p)).sum())


def analyse(idx: np.ndarray, agg_packed: np.ndarray, z, W, lens) -> dict:
    """idx [B,Tm,Q] of the product, agg_packed [N,D] fp32 of the product (valid tokens, (b, t) order)."""
    ridx = z["quantized_indices"].astype(np.int64)
    B, Tm, Q = ridx.shape
    valid = np.arange(Tm)[None, :] < np.asarray(lens)[:, None]
    assert np.array_equal(idx < 0, ridx < 0), "padding pattern differs"
    gi, ri = idx[valid], ridx[valid]                                     # [N, Q]
    N = gi.shape[0]
    a_ref = torch.from_numpy(z["aggregated_packed"]).double()
    a_got = torch.from_numpy(agg_packed).double()
    res = torch.from_numpy(z["residuals_packed"]).double()               # [N, Q, dc]
    Win = W["vq.rvq.project_in.weight"].double()
    code = [W[f"vq.rvq.layers.{q}._codebook.embed"][0].double() for q in range(Q)]
    delta = (a_got - a_ref) @ Win.T                                      # error of the residual entering level 0 (and of
    #                                                                      every level up to the first divergence)
    starts = np.concatenate([[0], np.cumsum(lens)])
    cen = []
    for b in range(B):
        s, e = int(starts[b]), int(starts[b + 1])
        if e - s > 1:
            g, r = a_got[s:e], a_ref[s:e]
            cen.append(_rel(g - g.mean(0, keepdim=True), r - r.mean(0, keepdim=True)))
    misses = []
    tok_b = np.repeat(np.arange(B), lens)
    tok_t = np.concatenate([np.arange(l) for l in lens])
    first_div = np.zeros(Q, dtype=int)
    for n in np.nonzero((gi != ri).any(-1))[0]:
        q = int(np.argmax(gi[n] != ri[n]))
        first_div[q] += 1
        r = res[n, q]
        ea, eb = code[q][ri[n, q]], code[q][gi[n, q]]
        dA, dB = float((r - ea).norm()), float((r - eb).norm())
        # the product quantises r + delta: it prefers b exactly when |r-eb|^2 - |r-ea|^2 < 2 delta.(eb - ea)
        lhs = dB * dB - dA * dA
        rhs = float(2.0 * (delta[n] @ (eb - ea)))
        misses.append(dict(utt=int(tok_b[n]), t=int(tok_t[n]), level=q, ref_code=int(ri[n, q]), got_code=int(gi[n, q]),
                           margin_rel=(dB - dA) / dA, delta_rel=float(delta[n].norm() / r.norm()),
                           explained=bool(lhs < rhs + 1e-5 * dA * dA)))     # slack: fp32 rounding of the distances
    out = dict(
        n_utterances=B, n_tokens=int(N), n_indices=int(N * Q),
        index_agreement=float((gi == ri).mean()),
        per_level=[float((gi[:, q] == ri[:, q]).mean()) for q in range(Q)],
        token_agreement=float((gi == ri).all(-1).mean()),
        first_divergence_per_level=first_div.tolist(),
        entropy_bits_ref=[_entropy_bits(ri[:, q]) for q in range(Q)],
        entropy_bits_got=[_entropy_bits(gi[:, q]) for q in range(Q)],
        distinct_codes_ref=[int(len(np.unique(ri[:, q]))) for q in range(Q)],
        aggregator_rel=_rel(a_got, a_ref), aggregator_centred_rel_max=max(cen), aggregator_centred_rel_mean=float(np.mean(cen)),
        misses=misses, misses_unexplained=int(sum(not m["explained"] for m in misses)),
        max_margin_rel=max([m["margin_rel"] for m in misses], default=0.0),
    )
    return out


def _run(tower, fe, batch, rows, want_agg=True):
    """One product call on the utterances `rows` (un-padded to their own longest transcript, like a collated batch)."""
    lens = batch["asr_token_lengths"][rows]
    T = int(lens.max())
    wav = batch["wav"][rows].cuda()
    _, b16 = fe.forward_device(wav, batch["n_samples"][rows].cuda(), False, True)
    fl = torch.full((len(rows),), 3000, device="cuda")
    args = (batch["asr_token_ids"][rows][:, :T].cuda(), lens.cuda(), b16, fl)
    kw = dict(asr_word_ids=batch["asr_word_ids"][rows][:, :T].cuda())
    out = tower(*args, **kw)
    agg = tower(*args, **kw, skip_vq_in_audio_encoder=True)["audio_unit_embeds"] if want_agg else None
    return out["quantized_indices"].cpu(), (agg.cpu() if want_agg else None), out


def _save(report):
    os.makedirs("gpurun_out", exist_ok=True)
    old = {}
    if os.path.exists(REPORT_PATH):
        with open(REPORT_PATH) as f:
            old = json.load(f)
    old.update(report)
    with open(REPORT_PATH, "w") as f:
        json.dump(old, f, indent=1)


@pytest.fixture(scope="module", params=["bf16", "fp16"])
def big(request, built_lib, golden_dir):
    prec = request.param
    z = np.load(os.path.join(golden_dir, "tower_full_b34.npz"))
    meta = json.loads(str(z["meta"]))
    W = synth.random_weights(synth.FULL, meta["weight_seed"])
    tower = TasteAudioTowerB200.from_config(synth.FULL, precision=prec).eval()
    tower.load_state_dict(W, strict=True)
    tower = tower.to("cuda:0")
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True, precision=prec).to("cuda:0")
    batch = synth.synth_batch(meta["batch_seed"], meta["durations"], meta["tokens"], pad_wave_to=480000)
    yield z, meta, W, tower, fe, batch, prec
    del tower
    torch.cuda.empty_cache()


def _pack(agg, lens):
    return np.concatenate([agg[b, : int(l)].numpy() for b, l in enumerate(lens)])


def test_full_b34_batched_vs_reference(big):
    z, meta, W, tower, fe, batch, prec = big
    lens = meta["tokens"]
    B = len(lens)
    idx, agg, out = _run(tower, fe, batch, list(range(B)))
    assert np.array_equal(out["audio_unit_lengths"].cpu().numpy(), z["audio_unit_lengths"])
    # encoder states (sub-sampled in the fixture)
    _, b16 = fe.forward_device(batch["wav"].cuda(), batch["n_samples"].cuda(), False, True)
    h_last, h_t = tower.engine().encode(b16)
    r_last = _rel(h_last.float().cpu()[:, ::100, ::8], z["h_last_sub"])
    r_tgt = _rel(h_t.float().cpu()[:, ::100, ::8], z["h_target_sub"])
    assert r_last < 1e-2 and r_tgt < 1e-2, (r_last, r_tgt)
    rep = analyse(idx.numpy(), _pack(agg, lens), z, W, lens)
    rep.update(h_last_rel=r_last, h_target_rel=r_tgt, precision=prec,
               path="one call, B = 34: CTA-pair GEMMs + folded LayerNorm")
    _save({f"full_b34_batched_{prec}": rep})
    print(f"batched {prec}:", {k: v for k, v in rep.items() if k != "misses"})
    assert rep["aggregator_rel"] < AGG_TOL[prec]
    assert rep["misses_unexplained"] == 0, [m for m in rep["misses"] if not m["explained"]]
    assert rep["index_agreement"] >= FLOOR[prec]["batched"], rep["index_agreement"]


def test_full_b34_solo_vs_reference(big):
    """B = 1 (config 5's path: single-CTA GEMM tiles below 2048 rows) on every utterance of the same fixture."""
    z, meta, W, tower, fe, batch, prec = big
    lens = meta["tokens"]
    B = len(lens)
    Tm = max(lens)
    idx = torch.full((B, Tm, 4), -1, dtype=torch.int64)
    aggs = []
    for b in range(B):
        i, a, _ = _run(tower, fe, batch, [b])
        idx[b, : lens[b]] = i[0]
        aggs.append(a[0, : lens[b]].numpy())
    rep = analyse(idx.numpy(), np.concatenate(aggs), z, W, lens)
    rep.update(path="34 calls, B = 1", precision=prec)
    _save({f"full_b34_solo_{prec}": rep})
    print(f"solo {prec}:", {k: v for k, v in rep.items() if k != "misses"})
    assert rep["aggregator_rel"] < AGG_TOL[prec]
    assert rep["misses_unexplained"] == 0, [m for m in rep["misses"] if not m["explained"]]
    assert rep["index_agreement"] >= FLOOR[prec]["solo"], rep["index_agreement"]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_default_init_weight_set_is_reported(built_lib, golden_dir, prec):
    """SURVEY section 7 hard part 3: the HF default init (std 0.02) is ill-conditioned (a handful of codes in play, the
    token-varying part of the aggregator output is a few per cent of its norm).  Reported, not asserted - except the
    aggregator tolerance and that every miss is explained by the aggregator error."""
    z = np.load(os.path.join(golden_dir, "tower_full_default_init.npz"))
    meta = json.loads(str(z["meta"]))
    W = synth.default_init_weights(synth.FULL, meta["weight_seed"])
    tower = TasteAudioTowerB200.from_config(synth.FULL, precision=prec).eval()
    tower.load_state_dict(W, strict=True)
    tower = tower.to("cuda:0")
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True, precision=prec).to("cuda:0")
    batch = synth.synth_batch(meta["batch_seed"], meta["durations"], meta["tokens"], pad_wave_to=480000)
    lens = meta["tokens"]
    idx, agg, _ = _run(tower, fe, batch, list(range(len(lens))))
    rep = analyse(idx.numpy(), _pack(agg, lens), z, W, lens)
    rep.update(path="one call, B = 8, HF default init (reported only)", precision=prec)
    _save({f"full_default_init_{prec}": rep})
    print(f"default init {prec}:", {k: v for k, v in rep.items() if k != "misses"})
    assert rep["aggregator_rel"] < 1e-2
    assert rep["misses_unexplained"] == 0
