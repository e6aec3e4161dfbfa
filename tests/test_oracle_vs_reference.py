"""Live pin of the CPU oracle against the REAL reference (only where /root/reference exists, i.e. the build container;
skipped on the GPU box).  The reference modules are executed where they lie through oracle/ref_shim.py."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle import taste_oracle as O
from taste_spokenlm_b200 import synth

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")
torch.set_grad_enabled(False)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ref_frontend():
    return ref_shim.build_reference_frontend()


def test_frontend_live(ref_frontend):
    for seed, n in ((3, 16000), (4, 200123), (5, 500000)):
        wav = synth.synth_waveform(seed, n)[None]
        f_ref, l_ref = ref_frontend(wav, torch.tensor([n]))                     # WF:87-113
        f, l = O.log_mel(wav, [n])
        assert f.shape == f_ref.shape == (1, 3000, 128)
        assert _rel(f, f_ref) < 2e-6
        assert int(l[0]) == int(l_ref[0])


@pytest.mark.parametrize("durs,toks,seed", [([4.0, 11.0], [9, 23], 0), ([2.0], [1], 1), ([7.5, 1.0, 30.0], [5, 2, 40], 2)])
def test_tower_live_tiny(ref_frontend, durs, toks, seed):
    cfg = synth.TINY
    tower = ref_shim.build_reference_tower(d_model=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
                                           heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab)
    W = synth.random_weights(cfg, 100 + seed)
    tower.load_state_dict(W, strict=True)                                        # same keys as the reference
    batch = synth.synth_batch(seed, durs, toks)
    feats = torch.cat([ref_frontend(batch["wav"][b:b + 1, : int(batch["n_samples"][b])],
                                    torch.tensor([int(batch["n_samples"][b])]))[0] for b in range(len(durs))])
    fl = torch.tensor([3000] * len(durs))
    ref = tower(batch["asr_token_ids"], batch["asr_token_lengths"], feats, fl, asr_word_ids=batch["asr_word_ids"])
    out = O.tower_forward(W, batch["asr_token_ids"], batch["asr_token_lengths"], feats, batch["asr_word_ids"],
                          cfg.heads, cfg.enc_layers)
    assert torch.equal(out["quantized_indices"] < 0, ref["quantized_indices"] < 0)
    agree = (out["quantized_indices"] == ref["quantized_indices"]).float().mean().item()
    assert agree >= 0.99, agree
    assert np.array_equal(out["audio_unit_lengths"].numpy(), ref["audio_unit_lengths"].numpy())
    same = (out["quantized_indices"] == ref["quantized_indices"]).all(-1)
    assert _rel(out["audio_unit_embeds"][same], ref["audio_unit_embeds"][same]) < 1e-4
    agg_ref = tower(batch["asr_token_ids"], batch["asr_token_lengths"], feats, fl, asr_word_ids=batch["asr_word_ids"],
                    skip_vq_in_audio_encoder=True)["audio_unit_embeds"]
    agg = O.tower_forward(W, batch["asr_token_ids"], batch["asr_token_lengths"], feats, batch["asr_word_ids"],
                          cfg.heads, cfg.enc_layers, skip_vq=True)["audio_unit_embeds"]
    for b, t in enumerate(toks):
        assert _rel(agg[b, :t], agg_ref[b, :t]) < 5e-5


def test_state_dict_keys_match_reference():
    cfg = synth.TINY
    tower = ref_shim.build_reference_tower(d_model=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
                                           heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab)
    ref_sd = tower.state_dict()
    spec = synth.state_dict_spec(cfg)
    assert set(ref_sd.keys()) == set(spec.keys())
    for k, shape in spec.items():
        assert tuple(ref_sd[k].shape) == tuple(shape), k
    from taste_spokenlm_b200.tower import TasteAudioTowerB200
    ours = TasteAudioTowerB200.from_config(cfg).state_dict()
    assert set(ours.keys()) == set(ref_sd.keys())
    assert all(ours[k].shape == ref_sd[k].shape and ours[k].dtype == ref_sd[k].dtype for k in ref_sd)
