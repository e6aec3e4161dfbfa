"""(f)4: the KV-cached `TasteSpokenLM.generate` against the reference's own loop (MT:1027-1199), on CPU.

The reference's `TasteSpokenLM` is instantiated where it lies (oracle/ref_shim.py) around a tiny random-init Llama
(no checkpoint exists offline; `__init__` would call `from_pretrained`, so the instance is assembled from the reference's
own sub-module classes: bridge fusion / extraction, TasteSampler, ResidualVQ).  Both loops run greedy with the same state:
the generated tokens, taste indices, word ids and lengths must be identical, and the cached loop must forward L + n
positions where the reference forwards sum_k (L + k)."""
import importlib
import importlib.util
import os
import sys

import pytest
import torch
from torch import nn

from oracle import ref_shim
from taste_spokenlm_b200 import synth
from taste_spokenlm_b200.generate import generate_kv_cached

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")
torch.set_grad_enabled(False)


def _real(name):
    """A reference sub-module the shim stubs out for the tower tests (bridge / sampler), loaded from its own file."""
    path = os.path.join(ref_shim.REF_ROOT, "taste_speech", "modules_taste", name + ".py")
    full = f"taste_speech.modules_taste._real_{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "taste_speech.modules_taste"
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


class _Tok:
    """Deterministic tokenizer stand-in for TasteSampler's vocabulary scans (sampler.py:31-58)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def decode(self, i):
        i = int(i[0]) if hasattr(i, "__len__") else int(i)
        if i % 7 == 0:
            return "."
        if i % 5 == 0:
            return "\n"                      # banned
        return (" w%d" if i % 2 == 0 else "x%d") % i


class _CountingBackbone(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model, self.positions = model, 0

    def forward(self, *a, **k):
        self.positions += k["inputs_embeds"].shape[1]
        return self.model(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("model"), name)


@pytest.fixture(scope="module")
def spoken_lm():
    from transformers import LlamaConfig, LlamaForCausalLM
    MT = ref_shim.import_modeling_taste()
    bridge, sampler = _real("bridge"), _real("sampler")
    torch.manual_seed(0)
    cfg = synth.TINY
    vocab, hid = 300, 64
    llama = LlamaForCausalLM(LlamaConfig(vocab_size=vocab, hidden_size=hid, intermediate_size=128, num_hidden_layers=3,
                                         num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512)).eval()
    for p in llama.parameters():
        p.mul_(3.0)                          # sharper logits: arg-max gaps far above fp32 re-association noise
    lm = MT.TasteSpokenLM.__new__(MT.TasteSpokenLM)
    nn.Module.__init__(lm)
    lm.language_model = llama
    lm._use_lora = False
    lm.fuse_for_bridge_in_llm = bridge.WeightedSumFusion(weight_init_type="balance", audio_dim=cfg.d_model, llm_dim=hid)
    lm.extract_for_bridge_out_llm = bridge.ContinueLatentLinearLastExtract(k=512, d=256, l=4, llm_dim=hid,
                                                                            llm_num_hidden_layers=3).eval()
    lm.sos_id, lm.k, lm.d = 1, 512, 256
    lm.delay, lm.delay_level, lm.audio_embed_conv_mode = 1, "word", "fill_forward"
    lm.pad_text_unit_embed = nn.Parameter(torch.randn(hid) * 0.1)
    lm.pad_audio_unit_embed = nn.Parameter(torch.randn(cfg.d_model) * 0.1)
    lm.taste_sampler = sampler.TasteSampler(1, "word", vocab, _Tok(vocab))
    lm.taste_sampler.ban_ids = [i for i in lm.taste_sampler.ban_ids if i < vocab]   # sampler.py:58 hard-codes Llama's 128001
    lm.eval()
    tower = ref_shim.build_reference_tower(d_model=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
                                           heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab)
    tower.load_state_dict(synth.random_weights(cfg, 7), strict=True)
    return MT, lm, tower.vq.rvq


def _run(fn, lm, vq, mode, **kw):
    counter = _CountingBackbone(lm.language_model.model)
    lm.language_model.model = counter
    try:
        out = fn(lm, vq, mode, **kw)
    finally:
        lm.language_model.model = counter.model
    return out, counter.positions


@pytest.mark.parametrize("mode", ["zero", "text", "audio"])
def test_kv_cached_generate_equals_reference_loop(spoken_lm, mode):
    MT, lm, vq = spoken_lm
    g = torch.Generator().manual_seed(3)
    L = 9
    kw = dict(extra_words=6)
    if mode != "zero":
        ids = torch.randint(2, 300, (1, L), generator=g)
        ids[0, 0] = lm.sos_id
        kw.update(llm_token_ids=ids, llm_token_lengths=torch.tensor([L]))
    if mode == "audio":
        wid = torch.tensor([[0, 0, 1, 1, 1, 2, 3, 3, 4]], dtype=torch.int32)
        idx = torch.randint(0, 512, (1, L, 4), generator=g)
        first = torch.diff(wid[0], prepend=torch.tensor([-1], dtype=torch.int32)) > 0
        idx[0, ~first] = -1                                  # llm_indices carry codes on word-start tokens only (MT:1877-1881)
        kw.update(llm_indices=idx, llm_word_ids=wid)
    ref, ref_pos = _run(MT.TasteSpokenLM.generate, lm, vq, mode, **kw)
    got, got_pos = _run(generate_kv_cached, lm, vq, mode, **kw)
    for a, b, name in zip(ref, got, ("llm_indices", "llm_token_ids", "llm_token_lengths", "llm_word_ids")):
        assert (a is None) == (b is None), name
        if a is not None:
            assert a.dtype == b.dtype and a.shape == b.shape, (name, a.dtype, b.dtype, a.shape, b.shape)
            assert torch.equal(a, b), name
    steps = int(ref[2]) + 2 if ref[2] is not None else 1
    assert got_pos < ref_pos and got_pos <= (L if mode != "zero" else 1) + steps + 8
    print(mode, "positions forwarded: reference", ref_pos, "kv-cached", got_pos)


def test_install_binds_generate(spoken_lm):
    MT, _, _ = spoken_lm
    from taste_spokenlm_b200 import tower as b200
    keep = (MT.TasteAudioTower, MT.TasteForCausalLM.extract_vq, MT.TasteSpokenLM.generate)
    try:
        b200.install(patch_frontend=False, patch_generate=True)
        assert MT.TasteSpokenLM.generate is generate_kv_cached
    finally:
        MT.TasteAudioTower, MT.TasteForCausalLM.extract_vq, MT.TasteSpokenLM.generate = keep
