"""(f)2 ingest on the GPU (-m gpu): `taste_resample_mean_f32` through the C ABI against the CPU oracle and the
torchaudio / process_one_sample fixtures, then the whole arrow-row -> llm-aligned indices driver.

Tolerance: fp32 FIR with a different summation order than torchaudio's conv1d: <= 1e-6 relative L2, <= 4e-6 absolute
on signals of amplitude <= 1."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import taste_oracle as O                      # checker only
from taste_spokenlm_b200 import ingest, synth
from taste_spokenlm_b200.frontend import WhisperFrontendB200
from taste_spokenlm_b200.shard import ShardWriter
from taste_spokenlm_b200.tower import TasteAudioTowerB200

torch.set_grad_enabled(False)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def resampler(built_lib):
    return ingest.ResampleMeanB200("cuda:0")


def test_resample_vs_torchaudio_fixture(resampler, golden_dir):
    z = np.load(os.path.join(golden_dir, "ingest.npz"))
    meta = json.loads(str(z["meta"]))
    for nm, seed, sr, ch, n in meta["resample"]:
        x = synth.synth_pcm(seed, n, ch)
        wav, ns = resampler([x], sr)
        ref = z[nm].reshape(-1)
        assert int(ns[0]) == ref.shape[0], nm
        got = wav[0, : ref.shape[0]].cpu().numpy()
        assert np.abs(got - ref).max() <= 4e-6, nm
        if n > 100:
            assert _rel(got, ref) < 1e-6, nm


@pytest.mark.parametrize("sr", [24000, 44100, 22050, 8000, 48000, 16000, 11025])
def test_resample_ragged_batch_vs_oracle(built_lib, sr):
    # small stride so that one row is longer than the window (trimmed, WF:98-99) and guard bands catch stray stores
    stride = 20000
    rs = ingest.ResampleMeanB200("cuda:0", wav_stride=stride)
    specs = [(1, 1), (2, 7), (1, 2049), (2, 12345), (3, 30011), (1, 50000), (1, 3 * stride)]
    arrays = [synth.synth_pcm(900 + i, n, c) for i, (c, n) in enumerate(specs)]
    B = len(arrays)
    buf = torch.full((B + 2, stride), 7.0, device="cuda")             # guard rows before and after
    out = buf[1: B + 1]
    wav, ns = rs(arrays, sr, out=out)
    torch.cuda.synchronize()
    assert torch.all(buf[0] == 7.0) and torch.all(buf[B + 1] == 7.0)
    for b, x in enumerate(arrays):
        ref = O.resample_mean(x, sr, 16000)
        m = min(ref.shape[0], stride)
        assert int(ns[b]) == m
        got = wav[b].cpu().numpy()
        assert np.abs(got[:m] - ref[:m]).max() <= 4e-6, (sr, b)
        if m > 100:
            assert _rel(got[:m], ref[:m]) < 1e-6, (sr, b)
        assert np.all(got[m:] == 7.0)                                  # nothing written past the utterance


def test_resample_back_to_back_calls_do_not_share_live_staging(built_lib):
    """the H2D copy out of the pinned staging buffer is asynchronous; many large back-to-back calls (the
    mixed-sampling-rate loop of CorpusIngestB200.tokenize_batch) must not let the host overwrite a buffer a DMA is still
    reading.  Every call's output is checked against the same call made alone after a full synchronize."""
    rs = ingest.ResampleMeanB200("cuda:0")
    groups = [(24000, [synth.synth_pcm(1200 + i, 700000, 1) for i in range(6)]),
              (44100, [synth.synth_pcm(1300 + i, 1300000, 1) for i in range(6)]),
              (8000, [synth.synth_pcm(1400 + i, 240000, 2) for i in range(6)]),
              (48000, [synth.synth_pcm(1500 + i, 1400000, 1) for i in range(6)])]
    outs = []
    for rate, arrs in groups * 2:                                  # 8 calls with no synchronisation in between
        w, ns = rs(arrs, rate)
        outs.append((w.clone(), ns.clone()))
    torch.cuda.synchronize()
    for k, (rate, arrs) in enumerate(groups * 2):
        torch.cuda.synchronize()
        w, ns = rs(arrs, rate)
        torch.cuda.synchronize()
        assert torch.equal(ns, outs[k][1])
        for b in range(len(arrs)):                                 # rows are written up to n_samples only
            n = int(ns[b])
            assert torch.equal(w[b, :n], outs[k][0][b, :n]), (k, rate, b)


def test_resample_properties_full_size(resampler):
    """Batch 64 x 30 s at 24 kHz (the Emilia rate): linearity, channel-mean consistency, pass-band gain."""
    B, n = 64, 720000
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, n, device="cuda", generator=g) * 0.1
    y = torch.randn(B, n, device="cuda", generator=g) * 0.1
    off = np.arange(B + 1, dtype=np.int64) * n
    ones, nin = np.ones(B, np.int64), np.full(B, n, np.int64)
    rx, ns = resampler.run_device(x.reshape(-1), off, ones, nin, 24000)
    assert int(ns.min()) == 480000 and int(ns.max()) == 480000
    rx = rx.clone()
    ry = resampler.run_device(y.reshape(-1), off, ones, nin, 24000)[0].clone()
    rxy = resampler.run_device((2.0 * x - 0.5 * y).reshape(-1), off, ones, nin, 24000)[0]
    lin = (rxy - (2.0 * rx - 0.5 * ry)).abs().max()
    assert float(lin) < 5e-6
    # two channels (x, y) of every utterance == mean of the mono results
    xy = torch.stack([x, y], 1).reshape(-1)
    rmean = resampler.run_device(xy, off * 2, 2 * ones, nin, 24000)[0]
    assert float((rmean - 0.5 * (rx + ry)).abs().max()) < 2e-6
    # a 1 kHz tone (far inside the pass band) keeps its amplitude
    t = torch.arange(n, device="cuda", dtype=torch.float64)
    tone = torch.sin(2 * np.pi * 1000.0 / 24000.0 * t).float()[None].repeat(2, 1)
    rt = resampler.run_device(tone.reshape(-1), off[:3], ones[:2], nin[:2], 24000)[0]
    ref = torch.sin(2 * np.pi * 1000.0 / 16000.0 * torch.arange(480000, device="cuda", dtype=torch.float64)).float()
    assert float((rt[0, 100:-100] - ref[100:-100]).abs().max()) < 2e-3


def test_process_one_sample_features_vs_reference(resampler, golden_dir):
    """PCM -> resample + mean -> log-mel on the device == `audio_features` of the reference's process_one_sample."""
    z = np.load(os.path.join(golden_dir, "ingest.npz"))
    meta = json.loads(str(z["meta"]))
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True).to("cuda:0")
    for j, (pseed, tseed, sr, ch, n, nwords) in enumerate(meta["samples"]):
        wav, ns = resampler([synth.synth_pcm(pseed, n, ch)], sr)
        f32, _ = fe.forward_device(wav, ns, True, False)
        assert _rel(f32[0, ::20, :].cpu().numpy(), z[f"s{j}_feats_sub"]) < 1e-4
        np.testing.assert_allclose(float(f32.double().sum()), float(z[f"s{j}_feats_sum"]), rtol=1e-4)


def test_corpus_ingest_rows_vs_oracle(built_lib, tmp_path):
    """Arrow-schema rows (mixed rates / channels) -> llm-aligned indices in the reference's output columns; resumable."""
    cfg = synth.TINY
    W = synth.random_weights(cfg, 1234)
    tower = TasteAudioTowerB200.from_config(cfg).eval()
    tower.load_state_dict(W, strict=True)
    tower = tower.to("cuda:0")
    asr_tok, llm_tok = synth.StubTokenizer(50257, 3, 1), synth.StubTokenizer(128256, 4, 2)
    rows = []
    for i, (sr, ch, dur, nw) in enumerate([(24000, 1, 3.0, 6), (24000, 2, 1.2, 2), (44100, 1, 2.5, 9), (16000, 1, 4.0, 4),
                                           (24000, 1, 0.7, 1), (8000, 2, 2.0, 5), (24000, 1, 5.0, 11)]):
        rows.append({"mp3": {"array": synth.synth_pcm(40 + i, int(sr * dur), ch), "sampling_rate": sr},
                     "json": {"text": synth.synth_text(80 + i, nw)}})
    ing = ingest.CorpusIngestB200(tower, asr_tok, llm_tok, batch_size=3)
    w = ShardWriter(str(tmp_path), rank=0, flush_every=2)
    assert ing.run(rows[:4], w) == 4
    # restart: a fresh writer skips what is on disk
    w2 = ShardWriter(str(tmp_path), rank=0, flush_every=2)
    assert ing.run(rows, w2) == 3
    got = {r["utt_id"]: r for r in w2.read_all()}
    assert sorted(got) == list(range(len(rows)))
    agree = total = 0
    for i, row in enumerate(rows):
        a_ids, a_wid, l_ids, l_wid = O.split_transcript(row["json"]["text"], asr_tok.encode, llm_tok.encode)
        wav = O.resample_mean(row["mp3"]["array"], row["mp3"]["sampling_rate"], 16000)
        feats, _ = O.log_mel(torch.from_numpy(wav)[None], [wav.shape[0]])
        T = len(a_ids)
        out = O.tower_forward(W, torch.tensor([a_ids]), torch.tensor([T], dtype=torch.int32), feats,
                              torch.tensor([a_wid], dtype=torch.int32), cfg.heads, cfg.enc_layers)
        ref = O.map_indices_to_llm_tokens(out["quantized_indices"], torch.tensor([T]), torch.tensor([a_wid]),
                                          torch.tensor([len(l_ids)]), torch.tensor([l_wid]))[0].numpy()
        r = got[i]
        assert r["llm_token_ids"] == [l_ids] and r["llm_word_ids"] == [l_wid] and r["llm_token_lengths"] == [len(l_ids)]
        mine = np.asarray(r["llm_indices"], dtype=np.int64)[0]                 # rows carry the reference's [1, L, Q]
        assert mine.shape == ref.shape
        np.testing.assert_array_equal(mine < 0, ref < 0)             # the word-start pattern is exact
        agree += int((mine == ref).sum())
        total += ref.size
    assert agree / total >= 0.97, agree / total                      # bf16 encoder vs fp32 oracle on random weights
