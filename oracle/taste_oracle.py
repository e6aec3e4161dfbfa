"""CPU oracle for the TASTE speech-tokenization hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain restatement (torch CPU tensor arithmetic, explicit loops where the reference loops) of what the
reference computes on the path  waveform -> log-mel -> Whisper encoder -> text-aligned aggregator ->
word pooling -> RVQ  (SURVEY.md §8(a) rows R1-R9, Appendix A1-A7).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it, and only
as the checker.  The product (`taste_spokenlm_b200`) never imports this module.

Parity status: the reference ships no golden vectors or tests for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by importing
the reference modules through `oracle/ref_shim.py` (`tests/golden/make_golden.py` is the committed
generator; `tests/golden/*.npz` are the fixtures; `tests/test_oracle_vs_reference.py` re-checks live when
`/root/reference` is present).  Third-party arithmetic the reference calls but does not vendor:
openai-whisper==20231117 (`mel_filters`, `pad_or_trim`; restated in `mel_filterbank()` below from the
published Slaney formula = `librosa.filters.mel(sr=16000, n_fft=400, n_mels=128)`), torch.stft,
transformers' GELU (exact erf).

All weights are addressed by the reference's own state_dict keys (SURVEY.md §8(b) "State / ownership").

Reference citations use the SURVEY.md abbreviations:
  WF  taste_speech/modules_taste/cosyvoice/whisper_frontend.py
  JES taste_speech/modules_taste/audio_joint_encoder_segmenter.py
  CW  taste_speech/modules_taste/cosyvoice/customized_whisper.py
  MT  taste_speech/modeling_taste.py
  AQ  taste_speech/modules_taste/audio_quantizer.py
  RVQ taste_speech/modules_taste/vq/residual_vq.py
  VQ  taste_speech/modules_taste/vq/vector_quantize_pytorch.py
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_MELS = 128
PREFIX = (50258, 50259, 50360, 50364)   # MT:147  <sot>,<en>,<transcribe>,<notimestamps>
EOS = 50257                             # MT:149
ENC = "audio_joint_encoder_segmenter.audio_encoder.encoder."
DEC = "audio_joint_encoder_segmenter.audio_segmenter.decoder."
RVQK = "vq.rvq."


# --------------------------------------------------------------------------------------------------
# R1  log-mel front-end                                                                   WF:56-113
# --------------------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mels = f / (200.0 / 3)
    log_t = f >= 1000.0
    mels = np.where(log_t, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / (np.log(6.4) / 27.0), mels)
    return mels


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = m * (200.0 / 3)
    log_t = m >= 15.0
    return np.where(log_t, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)


def mel_filterbank(n_mels: int = N_MELS, n_fft: int = N_FFT, sr: int = 16000) -> np.ndarray:
    """Slaney-scale, Slaney-normalised triangular filterbank [n_mels, n_fft//2+1] (fp32).

    openai-whisper's `mel_filters.npz['mel_128']` == librosa.filters.mel(sr=16000, n_fft=400, n_mels=128)
    (call site WF:44,66-70).  Restated from the published formula.
    """
    n_bins = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fft_freqs[None, :]
    w = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def log_mel(wav: torch.Tensor, n_samples: Optional[List[int]] = None, dtype=torch.float32
            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """wav [B, N] -> (feats [B, 3000, 128], lens [B]).   WF:87-113 -> WF:56-85, Appendix A1.

    pad_or_trim to 480000 (WF:98-99); reflect-pad 200 each side (torch.stft center=True); periodic Hann(400);
    400-point DFT, bins 0..200; drop frame 3000 (WF:65); |X|^2; mel (WF:70); log10(clamp 1e-10) (WF:72);
    floor at per-utterance max - 8 (WF:79-82); (x+4)/4 (WF:83); time-major (permute=True, WF:111-112).
    `lens` follows the reference's quirk of using input_lengths[0] for every row (WF:102).
    """
    B, N = wav.shape
    x = wav.to(dtype)
    if N > N_SAMPLES:
        x = x[:, :N_SAMPLES]
    elif N < N_SAMPLES:
        x = F.pad(x, (0, N_SAMPLES - N))
    xp = F.pad(x[:, None, :], (N_FFT // 2, N_FFT // 2), mode="reflect")[:, 0]        # [B, 480400]
    frames = xp.unfold(-1, N_FFT, HOP)[:, :N_FRAMES]                                   # [B, 3000, 400]
    n = torch.arange(N_FFT, dtype=torch.float64)
    window = (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / N_FFT)).to(dtype)
    spec = torch.fft.rfft(frames * window, n=N_FFT, dim=-1)                            # [B, 3000, 201]
    power = spec.real ** 2 + spec.imag ** 2
    fb = torch.from_numpy(mel_filterbank()).to(dtype)                                  # [128, 201]
    mel = power @ fb.T                                                                 # [B, 3000, 128]
    logs = torch.clamp(mel, min=1e-10).log10()
    gmax = logs.reshape(B, -1).max(dim=-1)[0]
    logs = torch.maximum(logs, gmax[:, None, None] - 8.0)
    feats = (logs + 4.0) / 4.0
    if n_samples is None:
        n_samples = [N] * B
    lens = torch.tensor([int(n_samples[0]) // HOP for _ in range(B)])
    return feats, lens


# --------------------------------------------------------------------------------------------------
# shared blocks
# --------------------------------------------------------------------------------------------------
def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))     # exact erf GELU (ACT2FN['gelu'])


def _attn(W: Dict[str, torch.Tensor], pre: str, xq, xk, xv, heads: int, causal: bool):
    """CW:324-409 eager attention.  q = (Wq x + bq) * hd^-0.5 (CW:342); k has no bias (CW:315)."""
    B, Tq, D = xq.shape
    hd = D // heads
    q = (xq @ W[pre + "q_proj.weight"].T + W[pre + "q_proj.bias"]) * (hd ** -0.5)
    k = xk @ W[pre + "k_proj.weight"].T
    v = xv @ W[pre + "v_proj.weight"].T + W[pre + "v_proj.bias"]
    q = q.view(B, Tq, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    s = q @ k.transpose(2, 3)                                           # CW:377
    if causal:                                                          # CW:1344-1350, CW:1440-1509
        Tk = s.shape[-1]
        m = torch.full((Tq, Tk), torch.finfo(s.dtype).min, dtype=s.dtype).triu(1)
        s = s + m
    p = torch.softmax(s, dim=-1)                                        # CW:383
    o = (p @ v).transpose(1, 2).reshape(B, Tq, D)                       # CW:394-405
    return o @ W[pre + "out_proj.weight"].T + W[pre + "out_proj.bias"]  # CW:407


# --------------------------------------------------------------------------------------------------
# R2/R3  encoder                                                   JES:133-223, CW:668-716, A2-A3
# --------------------------------------------------------------------------------------------------
def encoder(W: Dict[str, torch.Tensor], feats: torch.Tensor, heads: int, n_layers: int,
            target_hidden_layer: int = 6, return_all: bool = False):
    """feats [B, <=3000, 128] -> (h_last [B,1500,D], h_target [B,1500,D]).

    conv1(k3,p1)+GELU, conv2(k3,s2,p1)+GELU (JES:174-175), + embed_positions (JES:178-180), n_layers pre-LN
    layers (CW:668-716) with NO attention mask (JES:198-203), capture hidden ENTERING layer
    `target_hidden_layer` (JES:192-193), final layer_norm (JES:210-216).
    """
    x = feats.transpose(1, 2)
    if x.shape[-1] < N_FRAMES:                                          # JES:164-168
        x = F.pad(x, (0, N_FRAMES - x.shape[-1]))
    elif x.shape[-1] > N_FRAMES:
        raise ValueError("Whisper expects 3000 mel frames")             # JES:169-172
    x = _gelu(F.conv1d(x, W[ENC + "conv1.weight"], W[ENC + "conv1.bias"], padding=1))
    x = _gelu(F.conv1d(x, W[ENC + "conv2.weight"], W[ENC + "conv2.bias"], stride=2, padding=1))
    h = x.permute(0, 2, 1) + W[ENC + "embed_positions.weight"]
    h_target = None
    hs = []
    for l in range(n_layers):
        if l == target_hidden_layer:
            h_target = h
        if return_all:
            hs.append(h)
        p = f"{ENC}layers.{l}."
        a = _ln(h, W[p + "self_attn_layer_norm.weight"], W[p + "self_attn_layer_norm.bias"])
        h = h + _attn(W, p + "self_attn.", a, a, a, heads, causal=False)
        m = _ln(h, W[p + "final_layer_norm.weight"], W[p + "final_layer_norm.bias"])
        m = _gelu(m @ W[p + "fc1.weight"].T + W[p + "fc1.bias"])
        h = h + (m @ W[p + "fc2.weight"].T + W[p + "fc2.bias"])
    h_last = _ln(h, W[ENC + "layer_norm.weight"], W[ENC + "layer_norm.bias"])
    if return_all:
        return h_last, h_target, hs
    return h_last, h_target


# --------------------------------------------------------------------------------------------------
# R4/R5  token assembly + aggregator (2-layer Whisper decoder, K from h_last, V from h_target)
#                                                              MT:144-152, CW:1281-1437, CW:751-833, A4
# --------------------------------------------------------------------------------------------------
def assemble_tokens(asr_token_ids: torch.Tensor) -> torch.Tensor:
    """[B,Tmax] -> [B,Tmax+5]: prefix ++ ids ++ EOS after the PADDED width (MT:144-151)."""
    B = asr_token_ids.shape[0]
    pre = torch.tensor([list(PREFIX)] * B, dtype=asr_token_ids.dtype)
    eos = torch.tensor([[EOS]] * B, dtype=asr_token_ids.dtype)
    return torch.cat([pre, asr_token_ids, eos], dim=1)


def aggregator(W: Dict[str, torch.Tensor], tokens: torch.Tensor, h_last: torch.Tensor, h_target: torch.Tensor,
               heads: int, n_layers: int = 2) -> torch.Tensor:
    """tokens [B,T'] -> decoder final-LN state [B,T',D].  CW:1300-1415."""
    Tp = tokens.shape[1]
    d = W[DEC + "embed_tokens.weight"][tokens] + W[DEC + "embed_positions.weight"][:Tp]     # CW:1300,1328-1341
    for l in range(n_layers):
        p = f"{DEC}layers.{l}."
        a = _ln(d, W[p + "self_attn_layer_norm.weight"], W[p + "self_attn_layer_norm.bias"])
        d = d + _attn(W, p + "self_attn.", a, a, a, heads, causal=True)                      # CW:786-797
        c = _ln(d, W[p + "encoder_attn_layer_norm.weight"], W[p + "encoder_attn_layer_norm.bias"])
        d = d + _attn(W, p + "encoder_attn.", c, h_last, h_target, heads, causal=False)      # CW:801-813, CW:361-366
        m = _ln(d, W[p + "final_layer_norm.weight"], W[p + "final_layer_norm.bias"])
        m = _gelu(m @ W[p + "fc1.weight"].T + W[p + "fc1.bias"])
        d = d + (m @ W[p + "fc2.weight"].T + W[p + "fc2.bias"])                              # CW:818-824
    return _ln(d, W[DEC + "layer_norm.weight"], W[DEC + "layer_norm.bias"])                  # CW:1415


# --------------------------------------------------------------------------------------------------
# R6  prefix skip + word pooling                                                JES:393-458, A5
# --------------------------------------------------------------------------------------------------
def word_runs(word_ids_row: torch.Tensor, length: int) -> List[Tuple[int, int]]:
    """Maximal runs of equal consecutive word ids on the PADDED row with run-length > 1 and end <= length.

    JES:437-458 (`unique_consecutive` counts, `counts > 1`, `cumsum <= token_len`), incl. the padded-row quirk
    (SURVEY.md §8(a) R6): a last word whose id equals the pad value 0 merges with the padding and is skipped.
    """
    ids = [int(v) for v in word_ids_row.tolist()]
    runs, s = [], 0
    for i in range(1, len(ids) + 1):
        if i == len(ids) or ids[i] != ids[s]:
            if i - s > 1 and i <= length:
                runs.append((s, i))
            s = i
    return runs


def word_pool(x: torch.Tensor, word_ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """x [B,T,D]; replace every qualifying run by its mean, computed from the un-averaged input (JES:418-435)."""
    out = x.clone()
    for b in range(x.shape[0]):
        for (s, e) in word_runs(word_ids[b], int(lengths[b])):
            if e > x.shape[1]:
                raise ValueError("Invalid segment indices")                                  # JES:427-428
            out[b, s:e] = x[b, s:e].mean(dim=0, keepdim=True)
    return out


# --------------------------------------------------------------------------------------------------
# R7/R9  RVQ                                     AQ:109-124, RVQ:359-490, VQ:44-48, VQ:462-566, VQ:955-1217, A6
# --------------------------------------------------------------------------------------------------
def rvq_encode(W: Dict[str, torch.Tensor], z: torch.Tensor, mask: Optional[torch.Tensor], n_q: int = 4,
               project_in: bool = True, return_residuals: bool = False):
    """z [B,T,D] (fp32), mask [B,T] bool -> (quantized [B,T,D], indices [B,T,n_q] int64 with -1 at pads).

    x = project_in(z) (RVQ:371); per level: dist = -sqrt(clamp((|r|^2 + |e|^2) + (-2 r.e), 0)) (VQ:44-48,511),
    idx = first argmax (VQ:102), quant = embed[idx] (VQ:534); masked rows -> quant 0, idx -1 (VQ:1192-1210);
    residual -= quant; out += quant (RVQ:455-456); project_out(out) (RVQ:470).
    """
    x = z.float() if z.dtype != torch.float64 else z
    if project_in:
        x = x @ W[RVQK + "project_in.weight"].T.to(x.dtype) + W[RVQK + "project_in.bias"].to(x.dtype)
    residual = x
    out = torch.zeros_like(x)
    all_idx, residuals = [], []
    for q in range(n_q):
        e = W[f"{RVQK}layers.{q}._codebook.embed"][0].to(x.dtype)        # [K, dc]
        r = residual
        if return_residuals:
            residuals.append(r.clone())
        x2 = (r ** 2).sum(-1)
        y2 = (e ** 2).sum(-1)
        xy = torch.einsum("bid,jd->bij", r, e) * -2
        dist = -((x2[..., None] + y2[None, None, :]) + xy).clamp(min=0).sqrt()
        idx = dist.argmax(dim=-1)
        quant = e[idx]
        if mask is not None:
            quant = torch.where(mask[..., None], quant, torch.zeros_like(quant))
            idx = torch.where(mask, idx, torch.full_like(idx, -1))
        residual = residual - quant
        out = out + quant
        all_idx.append(idx)
    indices = torch.stack(all_idx, dim=-1)
    quantized = out @ W[RVQK + "project_out.weight"].T.to(x.dtype) + W[RVQK + "project_out.bias"].to(x.dtype)
    if return_residuals:
        return quantized, indices, residuals
    return quantized, indices


def rvq_codes_from_indices(W: Dict[str, torch.Tensor], indices: torch.Tensor) -> torch.Tensor:
    """sum_q C_q[idx_q] with -1 -> 0.   RVQ:183-237 + get_code_from_indices."""
    n_q = indices.shape[-1]
    dc = W[RVQK + "layers.0._codebook.embed"].shape[-1]
    acc = torch.zeros(*indices.shape[:-1], dc, dtype=torch.float32)
    for q in range(n_q):
        e = W[f"{RVQK}layers.{q}._codebook.embed"][0]
        i = indices[..., q]
        m = i == -1
        c = e[i.masked_fill(m, 0)]
        acc = acc + c.masked_fill(m[..., None], 0.0)
    return acc


def rvq_output_from_indices(W: Dict[str, torch.Tensor], indices: torch.Tensor) -> torch.Tensor:
    """RVQ:239-242."""
    return rvq_codes_from_indices(W, indices) @ W[RVQK + "project_out.weight"].T + W[RVQK + "project_out.bias"]


# --------------------------------------------------------------------------------------------------
# R4+R8  the boundary: TasteAudioTower.forward                                           MT:108-211
# --------------------------------------------------------------------------------------------------
def tower_forward(W: Dict[str, torch.Tensor], asr_token_ids: torch.Tensor, asr_token_lengths: torch.Tensor,
                  audio_features: torch.Tensor, asr_word_ids: torch.Tensor, heads: int, enc_layers: int,
                  dec_layers: int = 2, n_q: int = 4, target_hidden_layer: int = 6, skip_vq: bool = False,
                  stages: bool = False):
    W = {k: v for k, v in W.items()}
    h_last, h_t = encoder(W, audio_features, heads, enc_layers, target_hidden_layer)
    tokens = assemble_tokens(asr_token_ids)                                                  # MT:144-152
    dec = aggregator(W, tokens, h_last, h_t, heads, dec_layers)
    x = dec[:, 4:, :]                                                                        # JES:393-396
    lens = asr_token_lengths.to(torch.int64) + 5 - 4
    x = word_pool(x, asr_word_ids, lens)                                                     # JES:398-402
    seg = x[:, :-1, :]                                                                       # MT:170-172
    seg_len = lens - 1
    out = {"audio_unit_lengths": seg_len.to(asr_token_lengths.dtype)}
    if skip_vq:
        out["audio_unit_embeds"] = seg
    else:
        Tm = int(seg_len.max())
        mask = torch.arange(Tm)[None, :] < seg_len[:, None]                                  # modules_taste/utils.py:5-8
        quantized, indices = rvq_encode(W, seg, mask, n_q)
        out["audio_unit_embeds"] = quantized
        out["quantized_indices"] = indices
    assert int((out["audio_unit_lengths"].to(torch.int64) - asr_token_lengths.to(torch.int64)).sum()) == 0   # MT:210
    if stages:
        out["_h_last"], out["_h_target"], out["_aggregated"] = h_last, h_t, seg
    return out


# --------------------------------------------------------------------------------------------------
# (f)1  extract_vq epilogue: asr-token indices -> llm-token indices            MT:1438-1450, MT:1877-1881, A7
# --------------------------------------------------------------------------------------------------
def map_indices_to_llm_tokens(asr_indices: torch.Tensor, asr_token_lengths: torch.Tensor, asr_word_ids: torch.Tensor,
                              llm_token_lengths: torch.Tensor, llm_word_ids: torch.Tensor) -> torch.Tensor:
    """asr_indices [B,T,Q] -> llm_indices [B,L,Q] (-1 where an llm token is not a word start / has no match).

    M[b,l,t] = (asr_word_ids[b,t] == llm_word_ids[b,l]) & t<T_b & l<L_b; keep for each l the first matching t
    (cumsum over t == 1), then for each t the first l (cumsum over l == 1, re-masked by M) (MT:1438-1450);
    llm_indices = M' @ asr_indices - [row has no match] (MT:1878-1880).
    """
    B, T, Q = asr_indices.shape
    L = llm_word_ids.shape[1]
    out = torch.empty((B, L, Q), dtype=asr_indices.dtype)
    for b in range(B):
        tm = torch.arange(T) < int(asr_token_lengths[b])
        lm = torch.arange(L) < int(llm_token_lengths[b])
        M = ((asr_word_ids[b][None, :] == llm_word_ids[b][:, None]) & tm[None, :] & lm[:, None]).to(torch.int64)
        W1 = (torch.cumsum(M, dim=-1) == 1).to(torch.int64) * M
        W2 = (torch.cumsum(W1, dim=-2) == 1).to(torch.int64) * M
        out[b] = W2 @ asr_indices[b].to(torch.int64) - (W2.sum(-1, keepdim=True) == 0).to(torch.int64)
    return out


# --------------------------------------------------------------------------------------------------
# (f)2  corpus ingest: resample to 16 kHz + channel mean + transcript word split            DS:46-95
# --------------------------------------------------------------------------------------------------
# Third-party arithmetic: torchaudio==2.3.1 (requirements.txt:114) `transforms.Resample(orig, new)` with its defaults
# (sinc_interp_hann, lowpass_filter_width=6, rolloff=0.99), i.e. functional._get_sinc_resample_kernel +
# _apply_sinc_resample_kernel.  Restated here from the published algorithm; pinned by tests/golden/resample.npz,
# which tests/golden/make_golden.py produces with torchaudio itself running the reference's own call
# `resampler(speech_pt).mean(0)` (DS:52-60).
def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99
                         ) -> Tuple[np.ndarray, int, int, int]:
    """(kernel fp32 [new, 2*width+orig], width, orig, new) with orig / new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torch.arange(0, -new, -1) / new is evaluated in float32 (the default dtype) before it meets the float64 `idx`
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = (phase + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * (window * (base / orig))
    return k.astype(np.float32), width, orig, new


def resample_mean(x: np.ndarray, orig_freq: int, new_freq: int = 16000) -> np.ndarray:
    """x fp32 [C, n] -> fp32 [ceil(new*n/orig)]: torchaudio Resample on every channel, then `.mean(0)` (DS:52-60)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None]                                   # DS:53-54
    C, n = x.shape
    if int(orig_freq) == int(new_freq):               # transforms.Resample.forward returns the input unchanged
        return x.mean(0, dtype=np.float32) if C > 1 else x[0].copy()
    k, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    K = k.shape[1]
    xp = np.zeros((C, n + 2 * width + orig), dtype=np.float32)
    xp[:, width: width + n] = x
    frames = (xp.shape[1] - K) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, K, axis=1)[:, ::orig][:, :frames]      # [C, frames, K]
    y = np.einsum("cfk,pk->cfp", win, k, dtype=np.float32).reshape(C, frames * new)
    target = int(math.ceil(new * n / orig))
    y = y[:, :target]
    return (y.sum(0, dtype=np.float32) / np.float32(C)).astype(np.float32) if C > 1 else y[0]


def split_transcript(text: str, asr_encode, llm_encode):
    """DS:71-95: (asr_token_ids, asr_word_ids, llm_token_ids, llm_word_ids) of one transcript.

    `asr_encode(word)` / `llm_encode(word)` stand for `tokenizer.encode(word, add_special_tokens=False)`.  Words are
    the whitespace-split pieces of the stripped text, each but the first carrying a leading space.
    """
    import re
    text = text.strip()
    words = [" " + w for w in re.split(r"\s", text)]
    words[0] = words[0].lstrip()
    a_ids, a_wid, l_ids, l_wid = [], [], [], []
    for i, word in enumerate(words):
        for t in asr_encode(word):
            a_ids.append(int(t))
            a_wid.append(i)
        for t in llm_encode(word):
            l_ids.append(int(t))
            l_wid.append(i)
    return a_ids, a_wid, l_ids, l_wid
