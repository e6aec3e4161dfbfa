"""Drop-in for the reference audio tower (`taste_speech.modeling_taste.TasteAudioTower`, MT:33-211).

Same constructor arguments, same parameter / buffer names (so `from_pretrained`, `load_state_dict` and
`load_from_cosyvoice_ckpt` (MT:97-106) keep working), same `forward` signature and result dict.  The module tree below
only HOLDS the state under the reference's key names; every FLOP of `forward` runs in libtaste_b200.so through
`TowerEngine`.  Training / autograd are outside the accelerated path (SURVEY §8(b) "Mode conventions"): in that case
the reference's own module must be used and this class raises.

Two ways to use it:
  * `TasteAudioTowerB200(...)`  — standalone module (what the tests, bench and smoke use; needs no reference code);
  * `install()`                 — inside a process that has the reference importable: replaces
                                  `taste_speech.modeling_taste.TasteAudioTower` before `TasteForCausalLM` is built
                                  (the class is resolved at call time, MT:1280) and the `WhisperFrontend` bindings in
                                  `processing_taste` / `data.dataset` (PT:20, DS:18).  See INTEGRATION.md.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib
from .engine import TowerEngine
from .synth import TowerConfig, FULL


class _Attention(nn.Module):
    """State of WhisperAttention (CW:277-321): k_proj has no bias."""

    def __init__(self, d):
        super().__init__()
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)


class _EncoderLayer(nn.Module):
    """State of WhisperEncoderLayer (CW:649-666)."""

    def __init__(self, d, ffn):
        super().__init__()
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class _DecoderLayer(nn.Module):
    """State of WhisperDecoderLayer (CW:719-749)."""

    def __init__(self, d, ffn):
        super().__init__()
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = _Attention(d)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class _Encoder(nn.Module):
    """State of WhisperEncoder (CW:1007-1040)."""

    def __init__(self, cfg: TowerConfig):
        super().__init__()
        d = cfg.d_model
        self.conv1 = nn.Conv1d(cfg.n_mels, d, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(d, d, kernel_size=3, stride=2, padding=1)
        self.embed_positions = nn.Embedding(cfg.max_source_positions, d)
        self.layers = nn.ModuleList([_EncoderLayer(d, cfg.ffn) for _ in range(cfg.enc_layers)])
        self.layer_norm = nn.LayerNorm(d)


class _Decoder(nn.Module):
    """State of WhisperDecoder (CW:1160-1192)."""

    def __init__(self, cfg: TowerConfig):
        super().__init__()
        d = cfg.d_model
        self.embed_tokens = nn.Embedding(cfg.vocab, d)
        self.embed_positions = nn.Embedding(cfg.max_target_positions, d)
        self.layers = nn.ModuleList([_DecoderLayer(d, cfg.ffn) for _ in range(cfg.dec_layers)])
        self.layer_norm = nn.LayerNorm(d)


class _Holder(nn.Module):
    def __init__(self, name, child):
        super().__init__()
        self.add_module(name, child)


class WhisperAudioJointEncoderSegmenterB200(nn.Module):
    """State of WhisperAudioJointEncoderSegmenter (JES:280-328) for forward_type='asr_attn_pooling'."""

    def __init__(self, cfg: TowerConfig, make_v_proj_identity: bool = False):
        super().__init__()
        self.audio_encoder = _Holder("encoder", _Encoder(cfg))          # WhisperAudioEncoderForJoint.encoder
        self.audio_segmenter = _Holder("decoder", _Decoder(cfg))        # WhisperCrossAttentionSegmenterForJoint.decoder
        if make_v_proj_identity:                                        # JES:320-322, JES:330-334
            with torch.no_grad():
                for l in range(min(2, cfg.dec_layers)):
                    vp = self.audio_segmenter.decoder.layers[l].encoder_attn.v_proj
                    vp.weight.copy_(torch.eye(cfg.d_model))
                    vp.bias.fill_(0.0)

    def to(self, device):                                               # JES:31-33 single-argument override
        self.device = device
        return super().to(device)

    _tower_ref = None            # weak reference to the owning TasteAudioTowerB200 (which owns the kernel engine)

    def __getstate__(self):      # weak references do not pickle / deep-copy: the owning tower re-binds them
        d = self.__dict__.copy()
        d["_tower_ref"] = None
        return d

    def forward(self, audio_features, audio_features_lengths, asr_token_ids=None, asr_token_lengths=None,
                asr_word_ids=None, whisper_text_token=None, whisper_text_token_len=None, words_index=None,
                word_ids=None, **kwargs):
        """Secondary interface of the reference (JES:336-416): `(encoded_results, segmented_results)`.

        `whisper_text_token` is the assembled sequence prefix(4) ++ ids ++ EOS (MT:144-151) and
        `whisper_text_token_len = T + 5`.  `segmented_feats` is `[B, Tmax + 1, D]` with lengths `T + 1` exactly as the
        reference returns it before the tower drops the EOS slot (MT:170-172); slot `T_b` holds the un-pooled decoder
        state of that position, which no caller consumes."""
        if self._tower_ref is None or self._tower_ref() is None:
            raise _lib.TasteError("WhisperAudioJointEncoderSegmenterB200 is not attached to a TasteAudioTowerB200")
        if words_index is not None:
            raise NotImplementedError("pass `word_ids`; explicit `words_index` lists are not supported")
        if whisper_text_token is None or whisper_text_token_len is None or word_ids is None:
            raise AssertionError("joint encoder segmenter is word-level, please pass `words_index` or `word_ids` properly!")
        tower = self._tower_ref()
        eng = tower.engine()
        dev = eng.device
        feats = audio_features.detach()
        if feats.shape[1] < _lib.N_FRAMES:
            feats = torch.nn.functional.pad(feats, (0, 0, 0, _lib.N_FRAMES - feats.shape[1]))
        if feats.dtype not in (torch.float32, eng.act_dtype):
            feats = feats.float()
        h_last, h_t = eng.encode(feats.to(dev).contiguous())
        ids = whisper_text_token.detach()[:, 4:-1].to(device=dev, dtype=torch.int64).contiguous()
        lens_host = (whisper_text_token_len.detach().cpu().numpy().astype("int64") - 5)
        wid = word_ids.detach().to(device=dev, dtype=torch.int32).contiguous()
        z, _ = eng.segment_and_quantize(h_last, h_t, ids, wid, lens_host, skip_vq=True)
        dec, cu = eng._last_decoder
        B, Tmax, D = z.shape
        seg = torch.zeros(B, Tmax + 1, D, dtype=z.dtype, device=dev)
        seg[:, :Tmax] = z
        rows = torch.as_tensor(cu[1:] - 1, device=dev, dtype=torch.int64)            # assembled position 4 + T_b
        seg[torch.arange(B, device=dev), torch.as_tensor(lens_host, device=dev)] = dec[rows]
        lens1 = torch.as_tensor(lens_host + 1, device=dev, dtype=whisper_text_token_len.dtype)
        encoded = {"encoded_feats": {"last_hidden": h_last, str(tower.cfg.target_hidden_layer): h_t},
                   "encoded_feats_lengths": audio_features_lengths // 2 if audio_features_lengths is not None else None}
        segmented = {"segmented_feats": seg, "segmented_feat_lengths": lens1}
        return encoded, segmented


class _Codebook(nn.Module):
    """Buffers of EuclideanCodebook (VQ:319-327)."""

    def __init__(self, k, dc, kmeans_init):
        super().__init__()
        self.register_buffer("initted", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.ones(1, k))
        self.register_buffer("embed_avg", torch.zeros(1, k, dc))
        self.register_buffer("embed", torch.zeros(1, k, dc))


class _VQLayer(nn.Module):
    def __init__(self, k, dc, kmeans_init):
        super().__init__()
        self._codebook = _Codebook(k, dc, kmeans_init)


class _UnusedMLP(nn.Module):
    """The QINCo MLPs ResidualVQ always constructs but never uses here (RVQ:155)."""

    def __init__(self, dc):
        super().__init__()
        self.proj_in = nn.Linear(2 * dc, dc)
        self.layers = nn.ModuleList([nn.Sequential(nn.Linear(dc, dc), nn.SiLU(), nn.Linear(dc, dc)) for _ in range(4)])


class ResidualVQB200(nn.Module):
    """`tower.vq.rvq`: state + the methods the spoken-LM side calls on it (MT:681-689, 884-904; bridge.py:413)."""

    def __init__(self, *, dim, num_quantizers, codebook_dim=None, codebook_size=512, kmeans_init=True,
                 quantize_dropout=False, **vq_kwargs):
        super().__init__()
        codebook_dim = codebook_dim or dim
        self.num_quantizers = num_quantizers
        self.quantize_dropout = quantize_dropout and num_quantizers > 1
        self.has_projections = codebook_dim != dim
        self.project_in = nn.Linear(dim, codebook_dim) if self.has_projections else nn.Identity()
        self.project_out = nn.Linear(codebook_dim, dim) if self.has_projections else nn.Identity()
        self.layers = nn.ModuleList([_VQLayer(codebook_size, codebook_dim, kmeans_init) for _ in range(num_quantizers)])
        self.mlps = nn.ModuleList([_UnusedMLP(codebook_dim) for _ in range(num_quantizers - 1)])
        self._tower_ref = None           # weak reference to the owning TasteAudioTowerB200 (set in its constructor)

    def __getstate__(self):      # weak references do not pickle / deep-copy: the owning tower re-binds them
        d = self.__dict__.copy()
        d["_tower_ref"] = None
        return d

    @property
    def codebook_size(self):
        return self.layers[0]._codebook.embed.shape[1]

    @property
    def codebook_dim(self):
        return self.layers[0]._codebook.embed.shape[2]

    @property
    def codebooks(self):                                                 # RVQ:176-181
        return torch.stack([l._codebook.embed[0] for l in self.layers], dim=0)

    def _engine(self) -> TowerEngine:
        """The owning tower's engine, resolved on EVERY call: packing is lazy (the spoken-LM side calls
        `get_output_from_indices` etc. with no tower forward in the process, MT:681-689, 884-904, bridge.py:413) and
        follows the tower's state key, so codebooks loaded or moved after the first forward are never stale."""
        tower = self._tower_ref() if self._tower_ref is not None else None
        if tower is None:
            raise _lib.TasteError("ResidualVQB200 is not attached to a TasteAudioTowerB200 (the kernels' packed "
                                  "codebooks live in the tower's engine)")
        return tower.engine()

    def _check_initted(self):
        for l in self.layers:                                            # VQ:349-351 would run k-means here
            if float(l._codebook.initted.reshape(-1)[0]) == 0.0:
                raise _lib.TasteError("codebook `initted` is 0: the reference would k-means-initialise on this batch "
                                      "(VQ:349-370); load a trained checkpoint first")

    def forward(self, x, mask=None, **kw):                               # RVQ:359-490 (eval)
        if self.training or torch.is_grad_enabled() and x.requires_grad:
            raise _lib.TasteError("training / autograd RVQ is outside the B200 path; use the reference module")
        self._check_initted()
        lengths = None if mask is None else mask.to(torch.int32).sum(-1).to(torch.int32)
        qz, idx = self._engine().rvq_encode(x.float().contiguous(), lengths)
        losses = torch.zeros(self.num_quantizers, device=x.device, dtype=torch.float32)   # eval: zeros (AQ:117)
        return qz, idx, losses

    def get_indices_from_code(self, code, mask=None, **kw):              # RVQ:258-357
        self._check_initted()
        lengths = None if mask is None else mask.to(torch.int32).sum(-1).to(torch.int32)
        _, idx = self._engine().rvq_encode(code.float().contiguous(), lengths, want_quantized=False)
        return idx

    def get_code_from_indices(self, indices):                            # RVQ:249-253
        return self._engine().rvq_decode(indices.contiguous(), project_out=False)

    def get_output_from_indices(self, indices):                          # RVQ:239-242
        return self._engine().rvq_decode(indices.contiguous(), project_out=True)


class RVQAudioQuantizerB200(nn.Module):
    """AQ:83-124."""

    def __init__(self, dim=1280, num_quantizers=4, codebook_dim=None, quantize_dropout=False, kmeans_init=True,
                 codebook_size=256, decay=0.99, **vq_kwargs):
        super().__init__()
        self.rvq = ResidualVQB200(dim=dim, num_quantizers=num_quantizers, codebook_dim=codebook_dim,
                                  quantize_dropout=quantize_dropout, kmeans_init=kmeans_init,
                                  codebook_size=codebook_size, decay=decay, **vq_kwargs)

    def forward(self, z, mask, **kwargs):
        quantized, indices, commit_loss = self.rvq(z, mask=mask)
        return {"quantized_feats": quantized, "quantized_indices": indices, "commit_loss": commit_loss.sum()}


def _whisper_geometry(model_name_or_path: Optional[str], override: Optional[dict]) -> dict:
    """Geometry of the Whisper checkpoint the reference would load (JES:294-299): `config.json` if present, else the
    public distil-large-v3 values."""
    g = dict(d_model=1280, encoder_layers=32, decoder_layers=2, encoder_attention_heads=20, encoder_ffn_dim=5120,
             vocab_size=51866, num_mel_bins=128, max_source_positions=1500, max_target_positions=448)
    if model_name_or_path:
        p = os.path.join(model_name_or_path, "config.json")
        if os.path.isfile(p):
            with open(p) as f:
                g.update({k: v for k, v in json.load(f).items() if k in g})
    if override:
        g.update(override)
    return g


def _invalidate_after_load(module, incompatible_keys):
    """load_state_dict post-hook (a module-level function so that the tower stays picklable / deep-copyable)."""
    module.invalidate_engine()


class TasteAudioTowerB200(nn.Module):
    def __init__(
        self,
        encoder_input_size: int = 512,
        text_token_size: int = 51866,
        audio_embed_dim: int = 1280,
        quantization_on=False,
        is_joint_encoder_segmenter=False,
        audio_dropout_ratio=0.0,
        kwargs_audio_encoder: Dict = None,
        kwargs_audio_segmenter: Dict = None,
        kwargs_for_joint_encoder_segmenter: Dict = None,
        kwargs_for_quantizer: Dict = None,
        whisper_geometry: Dict = None,
        precision: Optional[str] = None,
    ):
        """`precision` (not a reference argument; default TASTE_PRECISION or 'bf16'): the library flavour, i.e. the
        16-bit type of the tensor-core operands — 'bf16' (BASELINE config 2) or 'fp16' (the reference's own autocast
        dtype, JES:133; 8x smaller operand rounding, the flavour that meets the >= 99.5 % index-agreement bar)."""
        super().__init__()
        self.precision = _lib.resolve_precision(precision)
        if not is_joint_encoder_segmenter:
            raise NotImplementedError("only the joint encoder/segmenter tower (CFG:136) is on the B200 path")
        kj = dict(kwargs_for_joint_encoder_segmenter or {})
        if kj.get("forward_type", "add_and_norm") != "asr_attn_pooling" or not kj.get("is_word_level", False) \
                or kj.get("skip_prefix_idx", None) != 4:
            raise NotImplementedError("B200 path implements forward_type='asr_attn_pooling', is_word_level=True, "
                                      "skip_prefix_idx=4 (CFG:137-145)")
        g = _whisper_geometry(kj.get("model_name_or_path"), whisper_geometry)
        kq = dict(kwargs_for_quantizer) if kwargs_for_quantizer is not None else None
        if kq is not None and kq.pop("quantizer_class", "rvq") != "rvq":
            raise NotImplementedError("only quantizer_class='rvq' is on the B200 path")
        self.cfg = TowerConfig(
            d_model=g["d_model"], enc_layers=g["encoder_layers"], dec_layers=g["decoder_layers"],
            heads=g["encoder_attention_heads"], ffn=g["encoder_ffn_dim"], vocab=g["vocab_size"],
            n_mels=g["num_mel_bins"], max_source_positions=g["max_source_positions"],
            max_target_positions=g["max_target_positions"],
            codebook_dim=(kq or {}).get("codebook_dim") or (kq or {}).get("dim", 1280),
            codebook_size=(kq or {}).get("codebook_size", 512), num_quantizers=(kq or {}).get("num_quantizers", 4),
            target_hidden_layer=kj.get("target_hidden_layer", 6))
        self.is_joint_encoder_segmenter = True
        self.audio_joint_encoder_segmenter = WhisperAudioJointEncoderSegmenterB200(
            self.cfg, make_v_proj_identity=kj.get("make_v_proj_identity", False))
        self.affine_audio = False
        if kq is not None:
            self.vq = RVQAudioQuantizerB200(**kq)
            self.quantization_on = True
        else:
            self.quantization_on = False
        self.audio_dropout_ratio = audio_dropout_ratio
        self.add_eos = True
        self._bind_children()
        self._engine: Optional[TowerEngine] = None
        self._engine_key = None
        self._state_epoch = 0
        self._sentinels = None
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    def _bind_children(self) -> None:
        """The segmenter and the RVQ reach the kernels through their owning tower (weak references: no cycle in state)."""
        import weakref
        self.audio_joint_encoder_segmenter._tower_ref = weakref.ref(self)
        if self.quantization_on:
            self.vq.rvq._tower_ref = weakref.ref(self)

    def __getstate__(self):      # pickling / copy.deepcopy: the packed kernel-side weights are rebuilt on first use
        d = self.__dict__.copy()
        d["_engine"], d["_engine_key"], d["_sentinels"] = None, None, None
        return d

    def __setstate__(self, state):
        super().__setstate__(state)
        self._bind_children()

    # ---- construction helpers ----
    @classmethod
    def from_config(cls, cfg: TowerConfig = FULL, precision: Optional[str] = None) -> "TasteAudioTowerB200":
        return cls(
            precision=precision,
            is_joint_encoder_segmenter=True, quantization_on=True, audio_embed_dim=cfg.d_model,
            kwargs_for_joint_encoder_segmenter=dict(dtype="bfloat16", forward_type="asr_attn_pooling", is_word_level=True,
                                                    make_v_proj_identity=True, model_name_or_path="", skip_prefix_idx=4,
                                                    use_custom=True, target_hidden_layer=cfg.target_hidden_layer),
            kwargs_for_quantizer=dict(codebook_dim=cfg.codebook_dim, codebook_size=cfg.codebook_size, decay=0.99,
                                      dim=cfg.d_model, kmeans_init=True, kmeans_iters=100,
                                      num_quantizers=cfg.num_quantizers, quantize_dropout=True),
            whisper_geometry=dict(d_model=cfg.d_model, encoder_layers=cfg.enc_layers, decoder_layers=cfg.dec_layers,
                                  encoder_attention_heads=cfg.heads, encoder_ffn_dim=cfg.ffn, vocab_size=cfg.vocab,
                                  max_target_positions=cfg.max_target_positions))

    def load_from_cosyvoice_ckpt(self, pt_path):                          # MT:97-106
        loaded = torch.load(pt_path, map_location="cpu")
        converted = {}
        for name, param in loaded.items():
            if "audio_tokenizer" in name:
                new_name = name.split("audio_tokenizer.")[-1].replace("audio_quantizer", "vq")
                converted[new_name] = param
        self.load_state_dict(converted, strict=True)

    # ---- engine management ----
    def invalidate_engine(self) -> None:
        """Forget the packed kernel-side weights (they are rebuilt on the next call).  Called automatically after
        `load_state_dict` and `.to()` / `.cuda()` / `.float()`; call it by hand after mutating a parameter in place
        through a view the sentinel check below cannot see."""
        self._state_epoch += 1

    def _apply(self, fn, *a, **k):                                       # .to / .cuda / .cpu / dtype casts
        r = super()._apply(fn, *a, **k)
        self._state_epoch = getattr(self, "_state_epoch", 0) + 1
        self._sentinels = None
        return r

    def _state_key(self):
        """O(1) in the number of tensors: an epoch bumped by the hooks above, plus storage + version of a few sentinel
        tensors (first / last encoder weights, decoder embedding, RVQ projection and codebooks) so that in-place edits
        of the usual suspects (`copy_`, k-means re-init, optimizer steps on the whole module) are seen as well.  Walking
        all ~700 tensors on every forward cost 5 % of the B = 1 latency."""
        if self._sentinels is None:
            named = dict(self.named_parameters())
            named.update(dict(self.named_buffers()))
            keys = list(named)
            pick = [keys[0], keys[len(keys) // 2], keys[-1]] + [k for k in keys if k.endswith("_codebook.embed")
                                                                  or k.endswith("project_in.weight")
                                                                  or k.endswith("embed_tokens.weight")]
            self._sentinels = [named[k] for k in dict.fromkeys(pick)]
        ver = 0
        ptr = 0
        for t in self._sentinels:
            ver += t._version
            ptr ^= t.data_ptr()
        return (self._state_epoch, str(self._sentinels[0].device), ver, ptr, self.precision)

    def engine(self) -> TowerEngine:
        """Packed kernel-side weights; rebuilt when the state changed (see `_state_key`)."""
        key = self._state_key()
        if self._engine is None or key != self._engine_key:
            dev = next(self.parameters()).device
            eng = TowerEngine(self.cfg, dev, self.precision)
            eng.pack(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    # ---- MT:108-211 ----
    def forward(self, asr_token_ids, asr_token_lengths, audio_features, audio_feature_lengths,
                asr_token_alignments=None, kwargs_for_encoder=None, kwargs_for_segmenter=None,
                kwargs_for_joint_encoder_segmenter=None, **kwargs):
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            raise _lib.TasteError(
                "TasteAudioTowerB200 is the eval / no-grad tokenization path; stage-1 training (MT:1532-1557: autograd, "
                "EMA codebook updates, quantize dropout) runs the reference module. Call .eval() under torch.no_grad().")
        if self.audio_dropout_ratio > 0.0:
            raise _lib.TasteError("audio_dropout_ratio > 0 (MT:188-199) is a training-time feature")
        if kwargs.get("words_index", None) is not None:
            raise NotImplementedError("pass `asr_word_ids`; explicit `words_index` lists (MT:141) are not supported")
        asr_word_ids = kwargs.get("asr_word_ids", None)
        if asr_word_ids is None:
            raise AssertionError("joint encoder segmenter is word-level, please pass `words_index` or `word_ids` properly!")
        eng = self.engine()
        skip = bool(kwargs.get("skip_vq_in_audio_encoder", False)) or not self.quantization_on
        if not skip:
            self.vq.rvq._check_initted()
        dev = eng.device
        res = eng.tower_forward(asr_token_ids.detach().to(dev), asr_token_lengths.detach().to(dev),
                                audio_features.detach(), asr_word_ids.detach(), skip_vq=skip)
        result = {"audio_unit_embeds": res["audio_unit_embeds"], "audio_unit_lengths": res["audio_unit_lengths"]}
        if not skip:
            result["quantized_indices"] = res["quantized_indices"]
        return result


def extract_vq(model, asr_token_ids, asr_token_lengths, asr_word_ids, llm_token_ids, llm_token_lengths, llm_word_ids,
               audio_features, audio_feature_lengths):
    """Drop-in for `TasteForCausalLM.extract_vq` (MT:1859-1881) when `model.audio_tower` is a `TasteAudioTowerB200`:
    same arguments, same `(asr_indices, llm_indices)` result; the word-start mapping (MT:1438-1450) runs as one CUDA
    kernel instead of a [B,L,T] matrix, two cumsums and a float bmm.  `install(patch_extract_vq=True)` binds it."""
    tower = model.audio_tower if hasattr(model, "audio_tower") else model
    enc = tower(asr_token_ids, asr_token_lengths, audio_features, audio_feature_lengths, asr_word_ids=asr_word_ids)
    asr_indices = enc["quantized_indices"]
    llm_indices = tower.engine().map_to_llm_tokens(asr_indices, asr_word_ids, asr_token_lengths, llm_word_ids,
                                                   llm_token_lengths)
    return asr_indices, llm_indices


def install(patch_frontend: bool = True, patch_extract_vq: bool = True, patch_generate: bool = False) -> None:
    """Swap the reference's classes for the B200 ones (needs `taste_speech` importable).  See INTEGRATION.md.

    `patch_generate=True` additionally binds the KV-cached `generate.generate_kv_cached` as `TasteSpokenLM.generate`
    (off by default: it is outside the tokenizer)."""
    import importlib
    mt = importlib.import_module("taste_speech.modeling_taste")
    mt.TasteAudioTower = TasteAudioTowerB200
    if patch_extract_vq and hasattr(mt, "TasteForCausalLM"):
        mt.TasteForCausalLM.extract_vq = extract_vq
    if patch_generate and hasattr(mt, "TasteSpokenLM"):
        from .generate import generate_kv_cached
        mt.TasteSpokenLM.generate = generate_kv_cached
    if patch_frontend:
        from .frontend import WhisperFrontendB200
        for modname in ("taste_speech.processing_taste", "taste_speech.data.dataset",
                        "taste_speech.modules_taste.cosyvoice.whisper_frontend"):
            try:
                m = importlib.import_module(modname)
                m.WhisperFrontend = WhisperFrontendB200
            except Exception:
                pass
