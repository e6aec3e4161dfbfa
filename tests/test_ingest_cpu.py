"""(f)2 ingest on the CPU: the oracle's restatement of torchaudio's Resample + channel mean and of the transcript
split, against fixtures produced by the reference's own `process_one_sample` (DS:37-113) and torchaudio
(tests/golden/make_golden.py ingest); and the product's host-side tap tables / transcript split against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import taste_oracle as O
from taste_spokenlm_b200 import ingest, synth

torch.set_grad_enabled(False)


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "ingest.npz"))
    return z, json.loads(str(z["meta"]))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_oracle_resample_matches_torchaudio_fixture(golden_dir):
    z, meta = _golden(golden_dir)
    for nm, seed, sr, ch, n in meta["resample"]:
        x = synth.synth_pcm(seed, n, ch)
        y = O.resample_mean(x, sr, 16000)
        ref = z[nm].reshape(-1)                             # DS:60 `.squeeze(0)` turns a 1-sample result into a scalar
        assert y.shape == ref.shape, nm                     # ceil(new * n / orig), incl. the 1- and 2-sample inputs
        assert np.abs(y - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), nm
        if n > 100:
            assert _rel(y, ref) < 1e-6, nm


def test_oracle_transcript_split_and_features_match_process_one_sample(golden_dir):
    z, meta = _golden(golden_dir)
    asr = synth.StubTokenizer(50257, 3, 1)
    llm = synth.StubTokenizer(128256, 4, 2)
    for j, (pseed, tseed, sr, ch, n, nwords) in enumerate(meta["samples"]):
        text = "  " + synth.synth_text(tseed, nwords) + " "
        a_ids, a_wid, l_ids, l_wid = O.split_transcript(text, asr.encode, llm.encode)
        np.testing.assert_array_equal(a_ids, z[f"s{j}_asr_token_ids"])
        np.testing.assert_array_equal(a_wid, z[f"s{j}_asr_word_ids"])
        np.testing.assert_array_equal(l_ids, z[f"s{j}_llm_token_ids"])
        np.testing.assert_array_equal(l_wid, z[f"s{j}_llm_word_ids"])
        # product host code: same lists
        assert ingest.split_transcript(text, asr, llm) == (a_ids, a_wid, l_ids, l_wid)
        # waveform -> features: resample + mean (oracle) then the log-mel oracle == the reference's audio_features
        wav = O.resample_mean(synth.synth_pcm(pseed, n, ch), sr, 16000)
        feats, lens = O.log_mel(torch.from_numpy(wav)[None], [wav.shape[0]])
        assert _rel(feats[0, ::20, :].numpy(), z[f"s{j}_feats_sub"]) < 5e-6
        np.testing.assert_allclose(float(feats.double().sum()), float(z[f"s{j}_feats_sum"]), rtol=1e-5)
        assert int(lens[0]) == int(z[f"s{j}_feat_len"][0])


@pytest.mark.parametrize("sr", [24000, 44100, 22050, 8000, 48000, 11025, 32000])
def test_sparse_tap_tables_reproduce_the_full_kernel(sr):
    full, width, orig, new = O.sinc_resample_kernel(sr, 16000)
    tb = ingest.polyphase_taps(sr, 16000)
    assert (tb["orig"], tb["new"], tb["width"]) == (orig, new, width)
    assert tb["knz_ld"] % 2 == 1 and tb["knz_ld"] >= tb["knz"]
    dense = np.zeros_like(full)
    for p in range(new):
        k0 = int(tb["kstart"][p])
        assert 0 <= k0 and k0 + tb["knz"] <= full.shape[1]
        dense[p, k0: k0 + tb["knz"]] = tb["taps"][p, : tb["knz"]]
    # identical where kept; what is dropped is the clamped window's tail
    kept = dense != 0
    np.testing.assert_array_equal(dense[kept], full[kept])
    assert np.abs(full[~kept]).max(initial=0.0) <= 1e-20
    # the dense run is far shorter than torchaudio's 2*width+orig taps when orig is large
    if orig > 100:
        assert tb["knz"] * 8 < full.shape[1]


def test_identity_rate_tables():
    tb = ingest.polyphase_taps(16000, 16000)
    assert (tb["orig"], tb["new"], tb["width"], tb["knz"]) == (1, 1, 0, 1)
    x = synth.synth_pcm(3, 1000, 2)
    np.testing.assert_array_equal(O.resample_mean(x, 16000), x.mean(0, dtype=np.float32))


def test_oracle_resample_live_against_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for sr, ch, n in [(24000, 2, 30001), (44100, 1, 12345), (8000, 1, 999)]:
        x = synth.synth_pcm(11, n, ch)
        pt = torch.from_numpy(np.atleast_2d(x))
        ref = ta.transforms.Resample(orig_freq=sr, new_freq=16000)(pt).mean(0).numpy()
        assert _rel(O.resample_mean(x, sr), ref) < 1e-6
        k, w, _, _ = O.sinc_resample_kernel(sr, 16000)
        r = ta.transforms.Resample(orig_freq=sr, new_freq=16000)
        np.testing.assert_array_equal(k, r.kernel[:, 0].numpy())
        assert w == r.width


def test_ingest_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(ingest.TasteError):
        ingest.ResampleMeanB200("cpu")
