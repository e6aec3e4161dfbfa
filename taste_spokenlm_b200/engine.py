"""Host side of the B200 path: packs reference-keyed weights for the kernels and sequences the C-ABI calls.

PyTorch is used for device memory, streams and a few index-building ops only; every FLOP of the path runs in
libtaste_b200.so.  There is no CPU or eager fallback: without the library or without a CUDA device this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib, mel
from .synth import TowerConfig

ENC = "audio_joint_encoder_segmenter.audio_encoder.encoder."
DEC = "audio_joint_encoder_segmenter.audio_segmenter.decoder."
RVQ = "vq.rvq."
PREFIX = (50258, 50259, 50360, 50364)     # MT:147
EOS = 50257                               # MT:149


def _on_device(fn):
    """Make the engine's device current around a C-ABI call: the library caches per-device state (shared-memory
    attributes, SM count) for the CURRENT device and launches on the stream handed in, so a tower on cuda:1 must not be
    driven while cuda:0 is current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise _lib.TasteError("the TASTE B200 path needs a CUDA device (sm_100a); there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class _Workspace:
    def __init__(self, device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.buf


class FrontendEngine:
    """Log-mel tables + handle (no model weights).  Stands in for WhisperFrontend's state (WF:7-49).

    `precision` ('bf16' default | 'fp16') selects the library flavour, i.e. the 16-bit type of the features handed to
    the encoder stem; the DFT itself runs on split-bf16 planes in both."""

    def __init__(self, device, precision=None):
        self.device = _require_cuda(device)
        self.precision = _lib.resolve_precision(precision)
        self.act_dtype = _lib.torch_dtype(self.precision)
        self.lib = _lib.load(self.precision)
        self._keep: Dict[str, torch.Tensor] = {}
        self.w = _lib.Weights()
        self._fill_tables(self.w)
        self.w.dims = _lib.Dims(d_model=128, heads=2, ffn=128, enc_layers=0, dec_layers=0, vocab=1, max_target_pos=1,
                                codebook_dim=256, codebook_size=512, num_quantizers=1, target_layer=0, reserved=0)
        self.handle = C.c_void_p()
        self._ck(self.lib.taste_handle_create(C.byref(self.w), C.byref(self.handle)), "taste_handle_create")
        self.ws = _Workspace(self.device)

    def _ck(self, rc: int, what: str = ""):
        _lib.check(rc, what, self.lib)

    def _dev(self, name: str, arr) -> C.c_void_p:
        t = torch.as_tensor(arr).contiguous().to(self.device)
        self._keep[name] = t
        return C.c_void_p(t.data_ptr())

    def _fill_tables(self, w: "_lib.Weights"):
        c, s = mel.dft_tables()
        st, cnt, wt = mel.sparse_filterbank()
        w.dft_cos = self._dev("dft_cos", c)
        w.dft_sin = self._dev("dft_sin", s)
        w.hann = self._dev("hann", mel.hann_periodic())
        w.mel_start = self._dev("mel_start", st)
        w.mel_count = self._dev("mel_count", cnt)
        w.mel_weight = self._dev("mel_weight", wt)
        wb = torch.from_numpy(mel.dft_gemm_weights().view("int16")).view(torch.bfloat16)
        w.dft_w_bf16 = self._dev("dft_w_bf16", wb)

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.taste_handle_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    @_on_device
    def logmel(self, wav: torch.Tensor, n_samples: torch.Tensor, want_f32: bool = True, want_bf16: bool = False):
        """wav fp32 [B, stride] on device; n_samples int32 [B] on device.  Returns (feats_f32|None, feats_bf16|None)."""
        assert wav.is_cuda and wav.dtype == torch.float32 and wav.dim() == 2 and wav.stride(1) == 1
        B = wav.shape[0]
        n_samples = n_samples.to(device=wav.device, dtype=torch.int32).contiguous()
        f32 = torch.empty(B, _lib.N_FRAMES, _lib.N_MELS, dtype=torch.float32, device=wav.device) if want_f32 else None
        b16 = torch.empty(B, _lib.N_FRAMES, _lib.N_MELS, dtype=self.act_dtype, device=wav.device) if want_bf16 else None
        nbytes = self.lib.taste_ws_bytes(self.handle, B, 0)
        ws = self.ws.get(nbytes)
        stream = C.c_void_p(torch.cuda.current_stream(wav.device).cuda_stream)
        self._ck(self.lib.taste_logmel_f32(self.handle, _lib.ptr(wav), _lib.ptr(n_samples), B, wav.stride(0),
                                             _lib.ptr(f32), _lib.ptr(b16), _lib.ptr(ws), ws.numel(), stream),
                   "taste_logmel_f32")
        return f32, b16


class TowerEngine(FrontendEngine):
    """Packed weights + handle for encoder, aggregator and RVQ (MT:33-211)."""

    def __init__(self, cfg: TowerConfig, device, precision=None):
        self.cfg = cfg
        self.device = _require_cuda(device)
        self.precision = _lib.resolve_precision(precision)
        self.act_dtype = _lib.torch_dtype(self.precision)
        self.lib = _lib.load(self.precision)
        self._keep = {}
        self.handle = C.c_void_p()
        self.ws = _Workspace(self.device)
        self._packed_version = None
        if cfg.d_model == cfg.codebook_dim:
            raise _lib.TasteError("d_model == codebook_dim (identity RVQ projections) is not supported")

    # ---- packing ---------------------------------------------------------------------------------------------
    @_on_device
    def pack(self, sd: Mapping[str, torch.Tensor]) -> None:
        """(Re)build the kernel-side weight set from a state_dict with the reference's key names."""
        cfg, dev = self.cfg, self.device
        if self.handle.value:
            self.lib.taste_handle_destroy(self.handle)
            self.handle = C.c_void_p()
        self._keep = {}
        D = cfg.d_model
        scale = (D // cfg.heads) ** -0.5

        def f32(name, t):
            return self._dev(name, t.detach().to(device=dev, dtype=torch.float32))

        def b16(name, t):          # the flavour's 16-bit operand type (bf16 or fp16)
            return self._dev(name, t.detach().to(device=dev, dtype=torch.float32).to(self.act_dtype))

        def qkv(prefix, p):
            wq = sd[p + "q_proj.weight"].detach().to(dev, torch.float32) * scale          # CW:342
            bq = sd[p + "q_proj.bias"].detach().to(dev, torch.float32) * scale
            wk = sd[p + "k_proj.weight"].detach().to(dev, torch.float32)                  # no bias, CW:315
            wv = sd[p + "v_proj.weight"].detach().to(dev, torch.float32)
            bv = sd[p + "v_proj.bias"].detach().to(dev, torch.float32)
            wcat = torch.cat([wq, wk, wv], 0)
            self._keep[prefix + "wqkv_f32"] = wcat          # transient: consumed by the LayerNorm fold (encoder layers)
            return (b16(prefix + "wqkv", wcat),
                    f32(prefix + "bqkv", torch.cat([bq, torch.zeros_like(bq), bv], 0)))

        w = _lib.Weights()
        w.dims = _lib.Dims(d_model=D, heads=cfg.heads, ffn=cfg.ffn, enc_layers=cfg.enc_layers,
                           dec_layers=cfg.dec_layers, vocab=cfg.vocab, max_target_pos=cfg.max_target_positions,
                           codebook_dim=cfg.codebook_dim, codebook_size=cfg.codebook_size,
                           num_quantizers=cfg.num_quantizers, target_layer=cfg.target_hidden_layer, reserved=0)
        self._fill_tables(w)
        c1 = sd[ENC + "conv1.weight"].detach().to(dev, torch.float32)                     # [D, 128, 3]
        c2 = sd[ENC + "conv2.weight"].detach().to(dev, torch.float32)                     # [D, D, 3]
        w.conv1_w = b16("conv1_w", c1.permute(0, 2, 1).reshape(D, -1))                    # k = tap*128 + c
        w.conv2_w = b16("conv2_w", c2.permute(0, 2, 1).reshape(D, -1))
        w.conv1_b = f32("conv1_b", sd[ENC + "conv1.bias"])
        w.conv2_b = f32("conv2_b", sd[ENC + "conv2.bias"])
        w.enc_pos = f32("enc_pos", sd[ENC + "embed_positions.weight"])
        enc = (_lib.EncLayer * max(cfg.enc_layers, 1))()
        for l in range(cfg.enc_layers):
            p, k = f"{ENC}layers.{l}.", f"enc{l}."
            L = enc[l]
            L.ln1_w = f32(k + "ln1_w", sd[p + "self_attn_layer_norm.weight"])
            L.ln1_b = f32(k + "ln1_b", sd[p + "self_attn_layer_norm.bias"])
            L.wqkv, L.bqkv = qkv(k, p + "self_attn.")
            L.wo = b16(k + "wo", sd[p + "self_attn.out_proj.weight"])
            L.bo = f32(k + "bo", sd[p + "self_attn.out_proj.bias"])
            L.ln2_w = f32(k + "ln2_w", sd[p + "final_layer_norm.weight"])
            L.ln2_b = f32(k + "ln2_b", sd[p + "final_layer_norm.bias"])
            L.w1 = b16(k + "w1", sd[p + "fc1.weight"])
            L.b1 = f32(k + "b1", sd[p + "fc1.bias"])
            L.w2 = b16(k + "w2", sd[p + "fc2.weight"])
            L.b2 = f32(k + "b2", sd[p + "fc2.bias"])
            # LayerNorm-folded copies (taste_enc_layer_t): W' = W * gamma (bf16), c = rowsum(W'), b' = b + W beta
            self._fold_ln(L, k, "qkv", self._keep[k + "wqkv_f32"], self._keep[k + "bqkv"],
                          sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
            del self._keep[k + "wqkv_f32"]
            self._fold_ln(L, k, "1", sd[p + "fc1.weight"].detach().to(dev, torch.float32), self._keep[k + "b1"],
                          sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
        w.enc = C.cast(enc, C.c_void_p)
        w.enc_ln_w = f32("enc_ln_w", sd[ENC + "layer_norm.weight"])
        w.enc_ln_b = f32("enc_ln_b", sd[ENC + "layer_norm.bias"])
        w.tok_emb = f32("tok_emb", sd[DEC + "embed_tokens.weight"])
        w.dec_pos = f32("dec_pos", sd[DEC + "embed_positions.weight"])
        dec = (_lib.DecLayer * max(cfg.dec_layers, 1))()
        for l in range(cfg.dec_layers):
            p, k = f"{DEC}layers.{l}.", f"dec{l}."
            L = dec[l]
            L.ln1_w = f32(k + "ln1_w", sd[p + "self_attn_layer_norm.weight"])
            L.ln1_b = f32(k + "ln1_b", sd[p + "self_attn_layer_norm.bias"])
            L.wqkv, L.bqkv = qkv(k, p + "self_attn.")
            self._keep.pop(k + "wqkv_f32", None)
            L.wo = b16(k + "wo", sd[p + "self_attn.out_proj.weight"])
            L.bo = f32(k + "bo", sd[p + "self_attn.out_proj.bias"])
            L.lnx_w = f32(k + "lnx_w", sd[p + "encoder_attn_layer_norm.weight"])
            L.lnx_b = f32(k + "lnx_b", sd[p + "encoder_attn_layer_norm.bias"])
            x = p + "encoder_attn."
            L.wq_x = b16(k + "wq_x", sd[x + "q_proj.weight"].detach().to(dev, torch.float32) * scale)
            L.bq_x = f32(k + "bq_x", sd[x + "q_proj.bias"].detach().to(dev, torch.float32) * scale)
            L.wk_x = b16(k + "wk_x", sd[x + "k_proj.weight"])
            L.wv_x = b16(k + "wv_x", sd[x + "v_proj.weight"])
            L.bv_x = f32(k + "bv_x", sd[x + "v_proj.bias"])
            L.wo_x = b16(k + "wo_x", sd[x + "out_proj.weight"])
            L.bo_x = f32(k + "bo_x", sd[x + "out_proj.bias"])
            L.ln2_w = f32(k + "ln2_w", sd[p + "final_layer_norm.weight"])
            L.ln2_b = f32(k + "ln2_b", sd[p + "final_layer_norm.bias"])
            L.w1 = b16(k + "w1", sd[p + "fc1.weight"])
            L.b1 = f32(k + "b1", sd[p + "fc1.bias"])
            L.w2 = b16(k + "w2", sd[p + "fc2.weight"])
            L.b2 = f32(k + "b2", sd[p + "fc2.bias"])
        w.dec = C.cast(dec, C.c_void_p)
        w.dec_ln_w = f32("dec_ln_w", sd[DEC + "layer_norm.weight"])
        w.dec_ln_b = f32("dec_ln_b", sd[DEC + "layer_norm.bias"])
        self._pack_rvq(w, sd)
        self.w = w
        self._enc_arr, self._dec_arr = enc, dec
        self._ck(self.lib.taste_handle_create(C.byref(w), C.byref(self.handle)), "taste_handle_create")

    def _fold_ln(self, L, key, which, w_f32, b_f32, gamma, beta):
        """Fill the `*_ln` fields of an encoder layer: Linear(LayerNorm(x)) = rstd * (x W'^T - mean * c) + b'."""
        dev = self.device
        g = gamma.detach().to(dev, torch.float32)
        bt = beta.detach().to(dev, torch.float32)
        wf = (w_f32 * g[None, :]).to(self.act_dtype)
        c = wf.float().sum(dim=1)                      # from the bf16-rounded W': what the tensor core multiplies
        bp = b_f32 + w_f32 @ bt
        names = ("wqkv_ln", "bqkv_ln", "cqkv_ln") if which == "qkv" else ("w1_ln", "b1_ln", "c1_ln")
        setattr(L, names[0], self._dev(key + names[0], wf))
        setattr(L, names[1], self._dev(key + names[1], bp))
        setattr(L, names[2], self._dev(key + names[2], c))

    def _pack_rvq(self, w, sd):
        cfg, dev = self.cfg, self.device
        Q = cfg.num_quantizers
        embed = torch.stack([sd[f"{RVQ}layers.{q}._codebook.embed"].detach().float().cpu()[0] for q in range(Q)])  # [Q,K,dc]
        w.rvq_code = self._dev("rvq_code", embed.to(dev))
        # split-bf16 planes for the tensor-core distance pass: e = hi + lo, [Q, 2, K, dc]
        hi = embed.to(torch.bfloat16)
        lo = (embed - hi.float()).to(torch.bfloat16)
        w.rvq_code_split = self._dev("rvq_code_split", torch.stack([hi, lo], dim=1).contiguous().to(dev))
        # |e|^2 with the reference's own reduction (VQ:46) on the host, so the bits match the fp32 reference
        w.rvq_code_sq = self._dev("rvq_code_sq", (embed ** 2).sum(-1).to(dev))
        w.rvq_win_t = self._dev("rvq_win_t", sd[RVQ + "project_in.weight"].detach().float().t().contiguous().to(dev))
        w.rvq_bin = self._dev("rvq_bin", sd[RVQ + "project_in.bias"].detach().float().to(dev))
        w.rvq_wout_t = self._dev("rvq_wout_t", sd[RVQ + "project_out.weight"].detach().float().t().contiguous().to(dev))
        w.rvq_bout = self._dev("rvq_bout", sd[RVQ + "project_out.bias"].detach().float().to(dev))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ws(self, batch: int, sum_tokens: int) -> torch.Tensor:
        return self.ws.get(self.lib.taste_ws_bytes(self.handle, batch, sum_tokens))

    # ---- stages ----------------------------------------------------------------------------------------------
    @_on_device
    def encode(self, feats: torch.Tensor):
        """feats [B,3000,128] fp32 or bf16 on device -> (h_last, h_target) bf16 [B,1500,D].   JES:133-223"""
        B, D = feats.shape[0], self.cfg.d_model
        assert feats.is_cuda and feats.is_contiguous() and feats.shape[1:] == (_lib.N_FRAMES, _lib.N_MELS)
        h_last = torch.empty(B, _lib.ENC_FRAMES, D, dtype=self.act_dtype, device=feats.device)
        h_t = torch.empty_like(h_last)
        ws = self._ws(B, 0)
        f32 = feats if feats.dtype == torch.float32 else None
        b16 = feats if feats.dtype == self.act_dtype else None
        if f32 is None and b16 is None:
            raise _lib.TasteError(f"encoder features must be fp32 or {self.act_dtype} (this engine's precision is "
                                  f"{self.precision}), got {feats.dtype}")
        self._ck(self.lib.taste_encoder_fwd(self.handle, _lib.ptr(f32), _lib.ptr(b16), B, _lib.ptr(h_last),
                                              _lib.ptr(h_t), _lib.ptr(ws), ws.numel(), self._stream()),
                   "taste_encoder_fwd")
        return h_last, h_t

    @staticmethod
    def assemble_tokens_host(ids: np.ndarray, lens: np.ndarray):
        """Packed assembled ids: per utterance prefix ++ ids[:T_b] ++ (ids[T_b] if T_b < Tmax else EOS).

        Row b of the reference's `whisper_text_token` (MT:144-151) truncated to T_b + 5 entries: by causality the
        decoder states the path consumes depend on nothing beyond them."""
        B, Tmax = ids.shape
        full = np.concatenate([np.tile(np.asarray(PREFIX, dtype=np.int64), (B, 1)), ids.astype(np.int64),
                               np.full((B, 1), EOS, dtype=np.int64)], axis=1)
        rows = [full[b, : int(lens[b]) + 5] for b in range(B)]
        cu = np.zeros(B + 1, dtype=np.int32)
        cu[1:] = np.cumsum([len(r) for r in rows])
        return np.concatenate(rows).astype(np.int32), cu

    @_on_device
    def aggregate(self, h_last, h_t, tokens_packed: torch.Tensor, cu_tokens: torch.Tensor, sum_tokens: int,
                  max_tokens: int) -> torch.Tensor:
        """-> decoder final-LN states fp32 [sum_tokens, D].   CW:1200-1437 with dict K/V (JES:377-388)"""
        B, D = h_last.shape[0], self.cfg.d_model
        out = torch.empty(sum_tokens, D, dtype=torch.float32, device=h_last.device)
        ws = self._ws(B, sum_tokens)
        self._ck(self.lib.taste_aggregator_fwd(self.handle, _lib.ptr(h_last), _lib.ptr(h_t), _lib.ptr(tokens_packed),
                                                 _lib.ptr(cu_tokens), B, sum_tokens, max_tokens, _lib.ptr(out),
                                                 _lib.ptr(ws), ws.numel(), self._stream()), "taste_aggregator_fwd")
        return out

    @_on_device
    def word_pool(self, dec_out, cu_tokens, word_ids, lengths, B, Tmax) -> torch.Tensor:
        z = torch.empty(B, Tmax, self.cfg.d_model, dtype=torch.float32, device=dec_out.device)
        self._ck(self.lib.taste_word_pool_f32(_lib.ptr(dec_out), _lib.ptr(cu_tokens), _lib.ptr(word_ids),
                                                _lib.ptr(lengths), B, Tmax, self.cfg.d_model, _lib.ptr(z),
                                                self._stream()), "taste_word_pool_f32")
        return z

    @_on_device
    def rvq_encode(self, z: torch.Tensor, lengths: Optional[torch.Tensor], want_quantized: bool = True):
        """z fp32 [B,T,in_dim] -> (quantized [B,T,D] | None, indices int64 [B,T,Q]).   RVQ:359-490 / RVQ:258-357"""
        assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 3
        B, T, in_dim = z.shape
        idx = torch.empty(B, T, self.cfg.num_quantizers, dtype=torch.int64, device=z.device)
        qz = torch.empty(B, T, self.cfg.d_model, dtype=torch.float32, device=z.device) if want_quantized else None
        if not hasattr(self, "_rvq_ws"):
            self._rvq_ws = _Workspace(self.device)
        ws = self._rvq_ws.get(self.lib.taste_rvq_ws_bytes(B * T))
        self._ck(self.lib.taste_rvq_encode_f32(self.handle, _lib.ptr(z), _lib.ptr(lengths), B, T, in_dim,
                                                 _lib.ptr(idx), _lib.ptr(qz), _lib.ptr(ws), ws.numel(), self._stream()),
                 "taste_rvq_encode_f32")
        return qz, idx

    @_on_device
    def rvq_decode(self, indices: torch.Tensor, project_out: bool = True) -> torch.Tensor:
        assert indices.is_cuda and indices.dtype == torch.int64
        shape = indices.shape[:-1]
        flat = indices.reshape(-1, indices.shape[-1]).contiguous()
        out = torch.empty(flat.shape[0], self.cfg.d_model if project_out else self.cfg.codebook_dim,
                          dtype=torch.float32, device=indices.device)
        self._ck(self.lib.taste_rvq_decode_f32(self.handle, _lib.ptr(flat), flat.shape[0], 1 if project_out else 0,
                                                 _lib.ptr(out), self._stream()), "taste_rvq_decode_f32")
        return out.reshape(*shape, out.shape[-1])

    @_on_device
    def map_to_llm_tokens(self, asr_indices, asr_word_ids, asr_token_lengths, llm_word_ids, llm_token_lengths):
        """extract_vq epilogue (MT:1438-1450, MT:1877-1881): asr-token indices [B,T,Q] -> llm-token indices [B,L,Q]
        (-1 where an llm token is not the first token of a word that also starts an asr word)."""
        dev = self.device
        idx = asr_indices.to(device=dev, dtype=torch.int64).contiguous()
        B, T, Q = idx.shape
        awid = asr_word_ids.to(device=dev, dtype=torch.int32)[:, :T].contiguous()
        lwid = llm_word_ids.to(device=dev, dtype=torch.int32).contiguous()
        L = lwid.shape[1]
        alen = asr_token_lengths.to(device=dev, dtype=torch.int32).contiguous()
        llen = llm_token_lengths.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty(B, L, Q, dtype=torch.int64, device=dev)
        self._ck(self.lib.taste_map_to_llm_tokens(_lib.ptr(idx), _lib.ptr(awid), _lib.ptr(alen), _lib.ptr(lwid),
                                                    _lib.ptr(llen), B, T, L, Q, _lib.ptr(out), self._stream()),
                   "taste_map_to_llm_tokens")
        return out

    @_on_device
    def assemble_tokens(self, ids_dev: torch.Tensor, lens32: torch.Tensor, cu: torch.Tensor, sum_tokens: int):
        """Packed assembled ids on the device (MT:144-152); see taste_assemble_tokens."""
        B, Tmax = ids_dev.shape
        tokens = torch.empty(sum_tokens, dtype=torch.int32, device=self.device)
        self._ck(self.lib.taste_assemble_tokens(_lib.ptr(ids_dev), _lib.ptr(lens32), _lib.ptr(cu), B, Tmax,
                                                  _lib.ptr(tokens), self._stream()), "taste_assemble_tokens")
        return tokens

    # ---- aggregator + pooling + RVQ on device-resident encoder states -------------------------------------------
    def segment_and_quantize(self, h_last, h_t, ids_dev, wid_dev, lengths_host: np.ndarray, skip_vq: bool = False,
                             want_quantized: bool = True):
        """MT:144-185 after the encoder: token assembly, aggregator, prefix skip + word pooling + EOS drop, RVQ.
        `lengths_host` is the host copy of asr_token_lengths (the only host-side quantity the launch geometry needs)."""
        dev = self.device
        B, Tmax = ids_dev.shape
        lengths_host = np.asarray(lengths_host).astype(np.int64)
        if (lengths_host < 0).any() or (lengths_host > Tmax).any():
            raise ValueError("asr_token_lengths out of range")
        cu_np = np.zeros(B + 1, dtype=np.int32)
        cu_np[1:] = np.cumsum(lengths_host + 5)
        # pinned staging from torch's caching host allocator (it keeps the block alive until the copy has run), so the
        # H2D of the two small index vectors is truly asynchronous
        meta_h = torch.empty(2 * B + 1, dtype=torch.int32, pin_memory=True)
        meta_np = meta_h.numpy()
        meta_np[: B + 1] = cu_np
        meta_np[B + 1:] = lengths_host
        meta = meta_h.to(dev, non_blocking=True)
        cu, lens32 = meta[: B + 1], meta[B + 1:]
        sum_tokens, max_tokens = int(cu_np[-1]), int(lengths_host.max()) + 5
        tokens = self.assemble_tokens(ids_dev, lens32, cu, sum_tokens)
        dec = self.aggregate(h_last, h_t, tokens, cu, sum_tokens, max_tokens)
        z = self.word_pool(dec, cu, wid_dev, lens32, B, Tmax)
        self._last_decoder = (dec, cu_np)        # for the secondary WhisperAudioJointEncoderSegmenter interface
        if skip_vq:                                                                      # MT:180,205
            return z, None
        Tm = int(lengths_host.max())             # generate_mask_from_length width (modules_taste/utils.py:5-8)
        if Tm != Tmax:
            raise ValueError(f"padded width {Tmax} != longest transcript {Tm}: the reference's mask/feature shapes "
                             f"disagree in this case (MT:181-184)")
        return self.rvq_encode(z, lens32, want_quantized=want_quantized)

    # ---- waveform -> indices with everything resident on the device (corpus driver path) ------------------------
    def tokenize_device(self, wav, n_samples, ids_dev, wid_dev, lengths_host, want_quantized: bool = True):
        """wav fp32 [B, N] + n_samples int32 [B] + padded ids int64 / word ids int32 [B,Tmax], all on the device.
        Returns (quantized [B,Tmax,D] fp32, indices [B,Tmax,Q] int64).  WF:87-113 -> MT:108-211."""
        _, feats = self.logmel(wav, n_samples, want_f32=False, want_bf16=True)
        h_last, h_t = self.encode(feats)
        return self.segment_and_quantize(h_last, h_t, ids_dev, wid_dev, lengths_host, want_quantized=want_quantized)

    # ---- the whole tower: MT:108-211 --------------------------------------------------------------------------
    def tower_forward(self, asr_token_ids, asr_token_lengths, audio_features, asr_word_ids, skip_vq: bool = False,
                      lengths_host: Optional[np.ndarray] = None):
        dev = self.device
        if lengths_host is None:
            lengths_host = asr_token_lengths.detach().cpu().numpy()      # the reference syncs here too (MT:210)
        feats = audio_features
        if feats.shape[1] < _lib.N_FRAMES:                                              # JES:164-168
            feats = torch.nn.functional.pad(feats, (0, 0, 0, _lib.N_FRAMES - feats.shape[1]))
        elif feats.shape[1] > _lib.N_FRAMES:                                            # JES:169-172
            raise ValueError(f"Whisper expects the mel input features to be of length {_lib.N_FRAMES}, "
                             f"but found {feats.shape[1]}")
        if feats.dtype not in (torch.float32, self.act_dtype):
            feats = feats.float()
        feats = feats.to(dev).contiguous()
        h_last, h_t = self.encode(feats)
        ids = asr_token_ids.to(device=dev, dtype=torch.int64).contiguous()
        wid = asr_word_ids.to(device=dev, dtype=torch.int32).contiguous()
        qz, idx = self.segment_and_quantize(h_last, h_t, ids, wid, lengths_host, skip_vq=skip_vq)
        out = {"audio_unit_lengths": asr_token_lengths.clone(), "audio_unit_embeds": qz}
        if not skip_vq:
            out["quantized_indices"] = idx
        return out
