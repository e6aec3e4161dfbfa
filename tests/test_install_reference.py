"""`tower.install()` against the REAL reference classes (build container only; skipped where /root/reference is absent).

The north star's "TasteForCausalLM, TasteProcessor and scripts/extract_vq_for_stage2_training.py are unchanged" rests on
three bindings, each exercised here on the reference's own modules, executed where they lie through oracle/ref_shim.py:
  * MT:1280   `TasteAudioTower(...)` is resolved at call time in `TasteForCausalLM.__init__` -> the class swap;
  * PT:20 / DS:18   `from ...whisper_frontend import WhisperFrontend` -> module-level rebinding;
  * MT:1859   `TasteForCausalLM.extract_vq` -> the patched method keeps the positional signature of MT:1870-1876.
The reference's constructor call is not re-typed here: its AST is cut out of modeling_taste.py and evaluated against the
patched class with the reference's own `TasteAudioTowerConfig` built from configs/model/taslm.json.
"""
import ast
import importlib
import inspect
import json
import os
import types

import pytest
import torch

from oracle import ref_shim
from taste_spokenlm_b200 import synth
from taste_spokenlm_b200 import tower as b200
from taste_spokenlm_b200.frontend import WhisperFrontendB200

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def installed():
    ref_shim.install()
    importlib.import_module("taste_speech.configuration_taste")        # the real config classes, not the shim's stubs
    MT = ref_shim.import_modeling_taste()
    PT = ref_shim.import_processing_taste()
    DS = importlib.import_module("taste_speech.data.dataset")
    ref_tower_cls, ref_extract, ref_fe = MT.TasteAudioTower, MT.TasteForCausalLM.extract_vq, PT.WhisperFrontend
    b200.install()
    yield MT, PT, DS, ref_tower_cls
    MT.TasteAudioTower = ref_tower_cls                                  # leave the process as the other tests expect it
    MT.TasteForCausalLM.extract_vq = ref_extract
    PT.WhisperFrontend = DS.WhisperFrontend = ref_fe
    importlib.import_module("taste_speech.modules_taste.cosyvoice.whisper_frontend").WhisperFrontend = ref_fe


def _reference_tower_call(MT):
    """The `TasteAudioTower(...)` call expression of `TasteForCausalLM.__init__` (MT:1280-1297), as an AST node."""
    src = inspect.getsource(MT.TasteForCausalLM.__init__)
    tree = ast.parse(inspect.cleandoc("\n" + src) if not src.startswith("def") else src)
    calls = [n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)
             and n.func.id == "TasteAudioTower"]
    assert len(calls) == 1
    return calls[0]


def test_class_swap_and_frontend_rebinding(installed):
    MT, PT, DS, ref_cls = installed
    assert MT.TasteAudioTower is b200.TasteAudioTowerB200 and ref_cls is not b200.TasteAudioTowerB200
    assert PT.WhisperFrontend is WhisperFrontendB200                     # PT:20 -> PT:164
    assert DS.WhisperFrontend is WhisperFrontendB200                     # DS:18 -> DS:132, DS:240
    assert MT.TasteForCausalLM.extract_vq is b200.extract_vq
    # the patched method keeps the reference's positional parameter order (MT:1859-1869; XV:24 calls it by keyword,
    # MT:1639 / 1695 / 1825 by keyword as well, MT:1870-1876 forwards positionally into the tower)
    ref_params = ["self", "asr_token_ids", "asr_token_lengths", "asr_word_ids", "llm_token_ids", "llm_token_lengths",
                  "llm_word_ids", "audio_features", "audio_feature_lengths"]
    ours = list(inspect.signature(b200.extract_vq).parameters)
    assert ours[1:] == ref_params[1:]
    # the frontend keeps the constructor keywords PT:164-168 / DS:132-136 pass
    fe_params = inspect.signature(WhisperFrontendB200.__init__).parameters
    for k in ("whisper_model", "do_pad_trim", "permute"):
        assert k in fe_params


def test_reference_call_site_constructs_the_b200_tower(installed):
    MT, _, _, ref_cls = installed
    CT = importlib.import_module("taste_speech.configuration_taste")
    with open(os.path.join(ref_shim.REF_ROOT, "configs", "model", "taslm.json")) as f:
        cfgd = json.load(f)
    call = _reference_tower_call(MT)
    # every keyword the reference passes is a keyword of the B200 constructor (and of the reference's own)
    kws = [k.arg for k in call.keywords]
    assert kws == ["encoder_input_size", "text_token_size", "audio_embed_dim", "quantization_on",
                   "is_joint_encoder_segmenter", "audio_dropout_ratio", "kwargs_audio_encoder", "kwargs_audio_segmenter",
                   "kwargs_for_joint_encoder_segmenter", "kwargs_for_quantizer"]
    ours = inspect.signature(b200.TasteAudioTowerB200.__init__).parameters
    theirs = inspect.signature(ref_cls.__init__).parameters
    for k in kws:
        assert k in ours and k in theirs, k
    # evaluate the reference's own call expression against the patched module namespace
    self_ns = types.SimpleNamespace(audio_tower_config=CT.TasteAudioTowerConfig(**cfgd["audio_tower_config"]))
    config = types.SimpleNamespace(asr_config=cfgd["asr_config"], _attn_implementation=cfgd["_attn_implementation"])
    expr = ast.Expression(call)
    ast.fix_missing_locations(expr)
    with torch.device("meta"):                                           # 759 M parameters: shapes only
        tower = eval(compile(expr, "<MT:1280>", "eval"), {"TasteAudioTower": MT.TasteAudioTower, "self": self_ns,
                                                           "config": config})
    assert isinstance(tower, b200.TasteAudioTowerB200)
    assert tower.quantization_on and tower.is_joint_encoder_segmenter and tower.add_eos
    assert tower.cfg == synth.FULL                                       # distil-large-v3 geometry + CFG:146-155
    spec = synth.state_dict_spec(synth.FULL)
    sd = tower.state_dict()
    assert list(sd.keys()) == list(spec.keys()) or set(sd.keys()) == set(spec.keys())
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k


def test_b200_tower_loads_the_reference_towers_state_dict(installed):
    MT, _, _, ref_cls = installed
    cfg = synth.TINY
    MT.TasteAudioTower = ref_cls                                         # build the REAL reference tower for comparison
    try:
        ref = ref_shim.build_reference_tower(d_model=cfg.d_model, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
                                             heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab)
    finally:
        MT.TasteAudioTower = b200.TasteAudioTowerB200
    assert type(ref) is ref_cls
    ref_sd = ref.state_dict()
    ours = b200.TasteAudioTowerB200.from_config(cfg).eval()
    epoch = ours._state_epoch
    res = ours.load_state_dict(ref_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert ours._state_epoch > epoch                                     # packed weights are rebuilt after a load
    for k, v in ref_sd.items():
        assert torch.equal(ours.state_dict()[k], v), k
    # and the other direction: a checkpoint saved from the B200 tower loads into the reference tower
    ref.load_state_dict(ours.state_dict(), strict=True)
    # the RVQ object the spoken-LM side reaches into (MT:681-689, bridge.py:413) resolves its tower lazily
    assert ours.vq.rvq._tower_ref() is ours
