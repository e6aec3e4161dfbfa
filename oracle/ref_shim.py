"""Import shim for the REAL reference (`/root/reference`) — test infrastructure only.

This module exists only in the build container: `/root/reference` is absent on the GPU box, so
nothing under `tests/ -m gpu`, `smoke()` or `bench.py` imports it.  It is used by
`tests/golden/make_golden.py` (to generate committed fixtures from the reference's own code) and by
the `not gpu` tests that pin `oracle/taste_oracle.py` against the reference when it is present.

What it does:
  1. registers empty `taste_speech` / `taste_speech.modules_taste` packages so submodules import
     without running `taste_speech/__init__.py` (which pulls onnxruntime/whisper/omegaconf/...);
  2. installs stub modules for absent third-party packages (einx, matplotlib, hyperpyyaml, librosa,
     whisper, onnxruntime, peft) and implements the four einx patterns the path uses;
  3. patches `load_whisper_whole_model` / `WhisperProcessor` / `WhisperTokenizer` in
     `audio_joint_encoder_segmenter` so the tower is built from a random-init
     `CustomWhisperModel(WhisperConfig(...))` instead of `./storage/pretrained_models/distil-large-v3`.
No reference source is copied; the modules are executed where they lie.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("TASTE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "taste_speech", "modules_taste"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_einx():
    import torch

    def get_at(pattern, src, idx):
        p = pattern.replace(" ", "")
        if p == "h[c]d,hbn->hbnd":            # VQ:534  embed [h,c,d], idx [h,b,n]
            h = src.shape[0]
            return torch.stack([src[i][idx[i]] for i in range(h)], 0)
        if p == "q[c]d,bnq->qbnd":            # RVQ:206 codebooks [q,c,d], idx [b,n,q]
            q = src.shape[0]
            return torch.stack([src[i][idx[..., i]] for i in range(q)], 0)
        raise NotImplementedError(pattern)

    def where(pattern, cond, a, b=None):
        p = pattern.replace(" ", "")
        if p == "bn,bnd,bnd->bnd":            # VQ:1198
            return torch.where(cond[..., None], a, b)
        if p == "bn,bn...,->bn...":           # VQ:1205
            c = cond
            while c.ndim < a.ndim:
                c = c[..., None]
            return torch.where(c, a, torch.as_tensor(b, dtype=a.dtype, device=a.device))
        raise NotImplementedError(pattern)

    _stub("einx", get_at=get_at, where=where)


def _install_whisper():
    import numpy as np
    import torch
    import torch.nn.functional as F
    from transformers.audio_utils import mel_filter_bank

    def mel_filters(device, n_mels):
        fb = mel_filter_bank(201, n_mels, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
        return torch.from_numpy(np.ascontiguousarray(fb.T)).to(torch.float32).to(device)

    def pad_or_trim(array, length=480000, *, axis=-1):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - array.shape[axis])
            array = F.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array

    audio = _stub("whisper.audio", N_FFT=400, HOP_LENGTH=160, N_SAMPLES=480000, mel_filters=mel_filters,
                  pad_or_trim=pad_or_trim)
    _stub("whisper", audio=audio, pad_or_trim=pad_or_trim)


_INSTALLED = False


def install() -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    for pkg, sub in (("taste_speech", "taste_speech"), ("taste_speech.modules_taste", "taste_speech/modules_taste")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF_ROOT, sub)]
        m.__spec__ = importlib.machinery.ModuleSpec(pkg, loader=None, is_package=True)
        m.__spec__.submodule_search_locations = m.__path__
        sys.modules[pkg] = m
    _install_einx()
    _install_whisper()
    plt = _stub("matplotlib.pyplot")
    _stub("matplotlib", pyplot=plt)
    _stub("hyperpyyaml", load_hyperpyyaml=lambda *a, **k: None)
    _stub("librosa")
    _stub("onnxruntime")
    _INSTALLED = True


def whisper_config(d_model=1280, enc_layers=32, dec_layers=2, heads=20, ffn=5120, vocab=51866):
    from transformers import WhisperConfig
    cfg = WhisperConfig(
        vocab_size=vocab, num_mel_bins=128, d_model=d_model,
        encoder_layers=enc_layers, encoder_attention_heads=heads, encoder_ffn_dim=ffn,
        decoder_layers=dec_layers, decoder_attention_heads=heads, decoder_ffn_dim=ffn,
        max_source_positions=1500, max_target_positions=448, use_cache=False,
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
    )
    cfg._attn_implementation = "eager"
    return cfg


def build_reference_tower(d_model=1280, enc_layers=32, dec_layers=2, heads=20, ffn=5120, vocab=51866,
                          codebook_dim=256, codebook_size=512, num_quantizers=4, target_hidden_layer=6):
    """Construct the reference `TasteAudioTower` (MT:33-95) with a random-init custom Whisper."""
    install()
    import torch
    JES = importlib.import_module("taste_speech.modules_taste.audio_joint_encoder_segmenter")
    CW = importlib.import_module("taste_speech.modules_taste.cosyvoice.customized_whisper")
    AQ = importlib.import_module("taste_speech.modules_taste.audio_quantizer")
    cfg = whisper_config(d_model, enc_layers, dec_layers, heads, ffn, vocab)

    def fake_loader(model_name_or_path="", attn_implementation="eager", dtype="float32", use_custom=False, **kw):
        model = CW.WhisperModel(cfg)
        return model, torch.float32

    class _FE:
        hop_length = 160
        nb_max_frames = 3000

    class _Proc:
        feature_extractor = _FE()

        @classmethod
        def from_pretrained(cls, *a, **k):
            return cls()

    JES.load_whisper_whole_model = fake_loader
    JES.WhisperProcessor = _Proc
    JES.WhisperTokenizer = _Proc
    tower = _RefTower(JES, AQ,
                      kwargs_for_joint_encoder_segmenter=dict(
                          dtype="float32", forward_type="asr_attn_pooling", is_word_level=True,
                          make_v_proj_identity=True, model_name_or_path="", skip_prefix_idx=4,
                          use_custom=True, target_hidden_layer=target_hidden_layer),
                      kwargs_for_quantizer=dict(
                          codebook_dim=codebook_dim, codebook_size=codebook_size, decay=0.99, dim=d_model,
                          kmeans_init=True, kmeans_iters=100, num_quantizers=num_quantizers,
                          quantize_dropout=True))
    return tower.eval()


def import_modeling_taste():
    """Import the reference's `taste_speech/modeling_taste.py` where it lies.

    The module imports the whole spoken-LM stack at its top; only `TasteAudioTower` / `TasteForCausalLM`'s class
    bodies are needed here, so the unrelated sub-modules (speech decoder, bridge, sampler ...) are stubbed.
    """
    install()
    name = "taste_speech.modeling_taste"
    if name not in sys.modules:
        try:                                       # the real config classes import cleanly (transformers only)
            importlib.import_module("taste_speech.configuration_taste")
        except Exception:
            _stub("taste_speech.configuration_taste", TasteConfig=object, TasteAudioTowerConfig=object,
                  TasteSpeechDecoderConfig=object, TasteSpokenLMConfig=object)
        for sub, attrs in (
            ("taste_speech.modules_taste.cosyvoice.encoder", dict(ConformerEncoder=object, TransformerEncoder=object)),
            ("taste_speech.modules_taste.cosyvoice.label_smoothing_loss", dict(LabelSmoothingLoss=object)),
            ("taste_speech.modules_taste.cosyvoice.utils", dict(IGNORE_ID=-1, th_accuracy=None)),
            ("taste_speech.modules_taste.audio_segmenter", dict(LocalAveragePoolingSegmenter=object)),
            ("taste_speech.modules_taste.bridge", dict(BRIDGE_FUSION_CLASSES={}, BRIDGE_EXTRACT_CLASSES={})),
            ("taste_speech.modules_taste.fusion", dict(TTS_INPUT_FUSION_CLASSES={})),
            ("taste_speech.modules_taste.sampler", dict(TasteSampler=object)),
        ):
            if sub not in sys.modules:
                _stub(sub, **attrs)
        importlib.import_module(name)
    return sys.modules[name]


def import_processing_taste():
    """Import the reference's `taste_speech/processing_taste.py` (binds `WhisperFrontend` at PT:20); the vocoder side
    it also imports (`inference_audio`, omegaconf) is stubbed."""
    install()
    name = "taste_speech.processing_taste"
    if name not in sys.modules:
        if "taste_speech.modules_taste.inference_audio" not in sys.modules:
            _stub("taste_speech.modules_taste.inference_audio", VoiceGenerator=object)
        try:
            import omegaconf  # noqa: F401
        except ImportError:
            _stub("omegaconf", DictConfig=dict)
        importlib.import_module(name)
    return sys.modules[name]


def _RefTower(JES, AQ, **kw):
    """Instantiate the reference's own `TasteAudioTower` class (MT:33-95)."""
    MT = import_modeling_taste()
    return MT.TasteAudioTower(is_joint_encoder_segmenter=True, quantization_on=True, **kw)


def build_reference_frontend():
    """The reference `WhisperFrontend(whisper_model='large-v3', do_pad_trim=True, permute=True)` (PT:164-168)."""
    install()
    WF = importlib.import_module("taste_speech.modules_taste.cosyvoice.whisper_frontend")
    return WF.WhisperFrontend(whisper_model="large-v3", do_pad_trim=True, permute=True)
