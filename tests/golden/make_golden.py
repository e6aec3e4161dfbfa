"""Generate the committed golden fixtures from the REAL reference (run in the build container only).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

Imports the reference modules where they lie under /root/reference through `oracle/ref_shim.py`, loads the
seeded synthetic weights of `taste_spokenlm_b200.synth` into the reference's own `TasteAudioTower`
(`load_state_dict(strict=True)`), runs the reference's own `WhisperFrontend.forward` (WF:87-113) and
`TasteAudioTower.forward` (MT:108-211) on seeded synthetic inputs, and stores small outputs / sub-sampled
intermediates as .npz.  Inputs and weights are NOT stored: they are regenerated from the seeds recorded here.
"""
import os
import sys
import json
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim                      # noqa: E402
from taste_spokenlm_b200 import synth            # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)

CASES = {
    # name: (config name, weight seed, batch seed, durations [s], token counts)
    "tower_tiny": ("TINY", 1234, 0, [10.0, 3.7, 30.0], [30, 11, 64]),
    "tower_tiny_single_word": ("TINY", 1234, 5, [2.0, 5.0], [1, 7]),
    "tower_small": ("SMALL", 4321, 1, [12.5, 30.0], [40, 80]),
    "tower_full": ("FULL", 1234, 0, [10.0, 30.0], [30, 64]),
}


def ref_frontend(fe, wav, n_samples):
    feats = []
    for b in range(wav.shape[0]):
        n = int(n_samples[b])
        f, _ = fe(wav[b:b + 1, :n], torch.tensor([n]))
        feats.append(f)
    return torch.cat(feats)


def tower_case(name, cfg_name, wseed, bseed, durs, toks, fe):
    cfg = getattr(synth, cfg_name)
    tower = ref_shim.build_reference_tower(d_model=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
                                           heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab)
    W = synth.random_weights(cfg, wseed)
    tower.load_state_dict(W, strict=True)
    batch = synth.synth_batch(bseed, durs, toks)
    feats = ref_frontend(fe, batch["wav"], batch["n_samples"])
    B = feats.shape[0]
    fl = torch.tensor([3000] * B)
    out = tower(batch["asr_token_ids"], batch["asr_token_lengths"], feats, fl, asr_word_ids=batch["asr_word_ids"])
    agg = tower(batch["asr_token_ids"], batch["asr_token_lengths"], feats, fl, asr_word_ids=batch["asr_word_ids"],
                skip_vq_in_audio_encoder=True)
    enc = tower.audio_joint_encoder_segmenter.audio_encoder(feats, fl, output_hidden_states=6)["encoded_feats"]
    cs = 8 if cfg.d_model > 256 else 1
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=json.dumps(dict(config=cfg_name, weight_seed=wseed, batch_seed=bseed, durations=durs, tokens=toks,
                             chan_stride=cs)),
        feats_sub=feats[:, ::50, :].numpy(),
        feats_sum=feats.double().sum(dim=(1, 2)).numpy(),
        h_last_sub=enc["last_hidden"][:, ::100, ::cs].numpy(),
        h_target_sub=enc["6"][:, ::100, ::cs].numpy(),
        aggregated=agg["audio_unit_embeds"].numpy(),
        audio_unit_embeds=out["audio_unit_embeds"].numpy(),
        audio_unit_lengths=out["audio_unit_lengths"].numpy(),
        quantized_indices=out["quantized_indices"].numpy(),
    )
    print(name, "ok", tuple(out["quantized_indices"].shape))


# FULL-geometry parity fixtures with >= 2000 tokens: 28 fixed-length 30 s x 64-token utterances
# (config 2's shape) + 6 ragged ones.  ~10 min of CPU in the build container for the conditioned weight set.
BIG_RAGGED = [(5.0, 14), (12.0, 32), (20.0, 54), (27.3, 74), (8.1, 22), (29.9, 80)]
BIG_CASES = {
    # name: (weight init, weight seed, batch seed, durations, tokens)
    "tower_full_b34": ("conditioned", 1234, 11, [30.0] * 28 + [d for d, _ in BIG_RAGGED], [64] * 28 + [t for _, t in BIG_RAGGED]),
    # SURVEY §7 hard part 3: the default HF init (std 0.02) is REPORTED, not asserted (ill-conditioned: ~4 codes in play)
    "tower_full_default_init": ("hf_default", 1234, 12, [30.0] * 6 + [11.0, 21.5], [64] * 6 + [30, 58]),
}


def tower_big_case(name, init, wseed, bseed, durs, toks, fe):
    cfg = synth.FULL
    tower = ref_shim.build_reference_tower()
    W = synth.random_weights(cfg, wseed) if init == "conditioned" else synth.default_init_weights(cfg, wseed)
    tower.load_state_dict(W, strict=True)
    batch = synth.synth_batch(bseed, durs, toks)
    B = len(durs)
    residuals = [[] for _ in range(cfg.num_quantizers)]
    hooks = [layer.register_forward_pre_hook(lambda m, a, q=q: residuals[q].append(a[0].detach().clone()))
             for q, layer in enumerate(tower.vq.rvq.layers)]
    outs, aggs, hl, ht = [], [], [], []
    # ONE reference forward over the whole batch (~20 GB of eager attention scores at B = 34); the aggregator output and the encoder states are captured by hooks on the
    # reference's own sub-modules (input of `tower.vq`, MT:178-181; output of the encoder wrapper, JES:133-223)
    hooks.append(tower.vq.register_forward_pre_hook(lambda m, a: aggs.append(a[0].detach().clone())))

    def enc_hook(m, a, out):
        e = out["encoded_feats"]
        hl.append(e["last_hidden"][:, ::100, ::8].clone()); ht.append(e["6"][:, ::100, ::8].clone())
    hooks.append(tower.audio_joint_encoder_segmenter.audio_encoder.register_forward_hook(enc_hook))
    feats = ref_frontend(fe, batch["wav"], batch["n_samples"])
    outs.append(tower(batch["asr_token_ids"], batch["asr_token_lengths"], feats, torch.tensor([3000] * B),
                      asr_word_ids=batch["asr_word_ids"]))
    for h in hooks:
        h.remove()
    Tm = max(toks)
    def padT(x, fill=0):
        if x.shape[1] == Tm:
            return x
        pad = torch.full((x.shape[0], Tm - x.shape[1]) + tuple(x.shape[2:]), fill, dtype=x.dtype)
        return torch.cat([x, pad], 1)
    idx = torch.cat([padT(o["quantized_indices"], -1) for o in outs])
    agg = torch.cat([padT(a) for a in aggs])
    assert len(aggs) == len(outs) == len(hl)
    lens = torch.cat([o["audio_unit_lengths"] for o in outs])
    valid = torch.arange(Tm)[None] < lens[:, None]
    # residual entering each RVQ level, as the reference's own VectorQuantize layers received it
    res = torch.stack([torch.cat([padT(r) for r in residuals[q]]) for q in range(cfg.num_quantizers)], 2)   # [B,Tm,Q,dc]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=json.dumps(dict(config="FULL", init=init, weight_seed=wseed, batch_seed=bseed, durations=durs, tokens=toks,
                             chan_stride=8)),
        h_last_sub=torch.cat(hl).numpy(), h_target_sub=torch.cat(ht).numpy(),
        aggregated_packed=agg[valid].numpy(),                      # [sum T, 1280] fp32, valid tokens in (b, t) order
        residuals_packed=res[valid].numpy(),                       # [sum T, Q, 256] fp32
        audio_unit_lengths=lens.numpy(),
        quantized_indices=idx.numpy().astype(np.int16),
    )
    print(name, "ok", tuple(idx.shape), int(valid.sum()), "tokens")


def frontend_cases(fe):
    specs = [("1s", 16000), ("10s", 160000), ("29.99s", 479840), ("30s", 480000), ("33s_trimmed", 528000),
             ("odd", 123457)]
    d = {}
    names = []
    for i, (nm, n) in enumerate(specs):
        wav = synth.synth_waveform(100 + i, n)[None]
        f, l = fe(wav, torch.tensor([n]))
        d[f"{nm}_sub"] = f[0, ::25, :].numpy()
        d[f"{nm}_sum"] = f.double().sum().numpy()
        d[f"{nm}_len"] = l.numpy()
        names.append([nm, 100 + i, n])
    # silence (hits the 1e-10 clamp everywhere) and a loud square-ish wave
    z = torch.zeros(1, 8000)
    f, _ = fe(z, torch.tensor([8000]))
    d["silence_sub"] = f[0, ::25, :].numpy()
    np.savez_compressed(os.path.join(OUT, "frontend.npz"), meta=json.dumps(names), **d)
    print("frontend ok")


def rvq_case():
    cfg = synth.FULL
    AQ = sys.modules["taste_speech.modules_taste.audio_quantizer"]
    vq = AQ.RVQAudioQuantizer(codebook_dim=256, codebook_size=512, decay=0.99, dim=1280, kmeans_init=True,
                              kmeans_iters=100, num_quantizers=4, quantize_dropout=True).eval()
    W = synth.random_weights(cfg, 77)
    sd = {k[len("vq."):]: v for k, v in W.items() if k.startswith("vq.")}
    vq.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(9)
    z = torch.randn(3, 50, 1280, generator=g)
    z = z + 0.7 * torch.randn(1, 1, 1280, generator=g)
    lens = torch.tensor([50, 17, 1])
    mask = torch.arange(50)[None] < lens[:, None]
    r = vq(z, mask=mask)
    rvq = vq.rvq
    idx = r["quantized_indices"]
    np.savez_compressed(
        os.path.join(OUT, "rvq.npz"),
        meta=json.dumps(dict(weight_seed=77, z_seed=9, lens=[50, 17, 1])),
        quantized_feats=r["quantized_feats"].numpy(), quantized_indices=idx.numpy(),
        output_from_indices=rvq.get_output_from_indices(idx).numpy(),
        code_from_indices=rvq.get_code_from_indices(idx).numpy(),
        indices_from_code=rvq.get_indices_from_code(rvq.project_in(z), mask=mask).numpy(),
    )
    print("rvq ok")


def pooling_cases():
    JES = sys.modules["taste_speech.modules_taste.audio_joint_encoder_segmenter"]
    f = JES.WhisperAudioJointEncoderSegmenter._convert_word_ids_to_words_index
    cases = [
        ([[0, 0, 0, 0]], [2]), ([[0, 0]], [2]), ([[0, 0, 1, 1, 1, 2, 0, 0]], [7]), ([[0, 1, 2, 3]], [5]),
        ([[0, 0, 1, 1], [0, 1, 1, 0]], [5, 4]), ([[0, 1, 1, 2, 2, 2, 3, 3, 0, 0, 0]], [9]),
        ([[0, 0, 0, 1, 1]], [4]), ([[3, 3, 3]], [4]),
    ]
    res = []
    for wid, ln in cases:
        out = f(None, torch.tensor(wid, dtype=torch.int32), torch.tensor(ln))
        res.append(dict(word_ids=wid, lengths=ln, words_index=[list(map(int, t)) for t in out]))
    json.dump(res, open(os.path.join(OUT, "word_pooling.json"), "w"), indent=0)
    print("pooling ok")


def mapping_case():
    import importlib
    MT = sys.modules["taste_speech.modeling_taste"]
    cls = MT.TasteForCausalLM
    g = torch.Generator().manual_seed(3)
    B, T, L, Q = 3, 12, 16, 4
    asr_len = torch.tensor([12, 7, 1]); llm_len = torch.tensor([16, 9, 2])
    asr_wid = torch.zeros(B, T, dtype=torch.int32); llm_wid = torch.zeros(B, L, dtype=torch.int32)
    for b in range(B):
        nw = 0
        for t in range(int(asr_len[b])):
            if t > 0 and torch.rand(1, generator=g).item() < 0.55:
                nw += 1
            asr_wid[b, t] = nw
        # llm tokenisation of the same words with different sub-word splits
        per = [1 + int(torch.randint(0, 3, (1,), generator=g)) for _ in range(nw + 1)]
        seq = [w for w, c in enumerate(per) for _ in range(c)][: int(llm_len[b])]
        llm_len[b] = len(seq)
        llm_wid[b, : len(seq)] = torch.tensor(seq, dtype=torch.int32)
    llm_wid = llm_wid[:, : int(llm_len.max())].contiguous()          # collate pads to the longest row
    idx = torch.randint(0, 512, (B, T, Q), generator=g)
    idx = torch.where((torch.arange(T)[None] < asr_len[:, None])[..., None], idx, torch.full_like(idx, -1))
    sm = cls._get_word_start_mapping_matrix(None, asr_wid, llm_wid, asr_len, llm_len)
    # generate_mask_from_length truncates to max length: pad back to [B, L, T]
    llm = torch.bmm(sm, idx[:, : sm.shape[2]].float()) - (sm.sum(dim=-1, keepdim=True) == 0).float()
    llm = llm.to(idx.dtype)
    np.savez_compressed(os.path.join(OUT, "llm_mapping.npz"), asr_indices=idx.numpy(), asr_len=asr_len.numpy(),
                        llm_len=llm_len.numpy(), asr_wid=asr_wid.numpy(), llm_wid=llm_wid.numpy(),
                        llm_indices=llm.numpy())
    print("mapping ok", tuple(llm.shape))


def ingest_cases(fe):
    """(f)2: the reference's own `process_one_sample` (DS:37-113) and torchaudio's Resample on seeded PCM."""
    import importlib
    import torchaudio
    DS = importlib.import_module("taste_speech.data.dataset")
    d = {}
    specs = [("24k_mono", 24000, 1, 6007), ("24k_stereo", 24000, 2, 4801), ("44k1_mono", 44100, 1, 9001),
             ("8k_stereo", 8000, 2, 3001), ("48k_mono", 48000, 1, 7000), ("16k_stereo", 16000, 2, 2000),
             ("22k05_mono", 22050, 1, 5000), ("24k_tiny", 24000, 1, 2), ("24k_one", 24000, 1, 1)]
    meta = []
    for i, (nm, sr, ch, n) in enumerate(specs):
        x = synth.synth_pcm(500 + i, n, ch)
        pt = torch.tensor(x, dtype=torch.float32)
        if pt.dim() == 1:
            pt = pt.unsqueeze(0)
        y = torchaudio.transforms.Resample(orig_freq=sr, new_freq=16000)(pt).mean(0).squeeze(0).numpy()   # DS:55-60
        d[nm] = y
        meta.append([nm, 500 + i, sr, ch, n])

    class _Proc:
        tokenizer = synth.StubTokenizer(50257, 3, 1)
    llm_tok = synth.StubTokenizer(128256, 4, 2)
    samples = []
    for j, (sr, ch, n, nwords) in enumerate([(24000, 1, 48000, 9), (44100, 2, 30011, 5), (16000, 1, 20000, 1)]):
        x = synth.synth_pcm(600 + j, n, ch)
        text = "  " + synth.synth_text(700 + j, nwords) + " "
        sample = {"mp3": {"array": x, "sampling_rate": sr}, "json": {"text": text}, "s3_token": [1, 2, 3],
                  "spk_emb": [0.1] * 8}
        out = DS.process_one_sample(sample, resampler_dict={}, whisper_processor=_Proc(), llm_tokenizer=llm_tok,
                                    whisper_feature_extractor=fe)
        f = out["audio_features"]
        d[f"s{j}_feats_sub"] = f[0, ::20, :].numpy()
        d[f"s{j}_feats_sum"] = f.double().sum().numpy()
        d[f"s{j}_feat_len"] = out["audio_feature_lengths"].numpy()
        for k in ("asr_token_ids", "asr_word_ids", "llm_token_ids", "llm_word_ids"):
            d[f"s{j}_{k}"] = out[k][0].numpy()
        samples.append([600 + j, 700 + j, sr, ch, n, nwords])
    np.savez_compressed(os.path.join(OUT, "ingest.npz"), meta=json.dumps(dict(resample=meta, samples=samples)), **d)
    print("ingest ok")


if __name__ == "__main__":
    fe = ref_shim.build_reference_frontend()
    which = sys.argv[1:] or list(CASES) + ["frontend", "rvq", "pooling", "mapping", "ingest"]
    for name in which:
        if name in CASES:
            tower_case(name, *CASES[name], fe)
        if name in BIG_CASES:                    # only on request: minutes of CPU each
            tower_big_case(name, *BIG_CASES[name], fe)
    if "frontend" in which:
        frontend_cases(fe)
    if "ingest" in which:
        ingest_cases(fe)
    ref_shim.build_reference_tower(d_model=128, enc_layers=1, heads=2, ffn=128, vocab=51866)   # ensures modules imported
    if "rvq" in which:
        rvq_case()
    if "pooling" in which:
        pooling_cases()
    if "mapping" in which:
        mapping_case()
