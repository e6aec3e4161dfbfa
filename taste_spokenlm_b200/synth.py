"""Synthetic weights and inputs for the TASTE tokenization path (no checkpoints or datasets exist offline).

`random_weights(cfg, seed)` returns a state_dict keyed exactly like the reference `TasteAudioTower`
(`audio_joint_encoder_segmenter.audio_encoder.encoder.*`,
`audio_joint_encoder_segmenter.audio_segmenter.decoder.*`, `vq.rvq.*`; constructed at MT:34-95, JES:281-328,
RVQ:102-170).  Every tensor is drawn from its own generator seeded by (seed, crc32(key)) so the same values
are produced here, in `tests/golden/make_golden.py` (where they are loaded into the real reference), in the
oracle and on the GPU box, independent of construction order.

The distribution is a "well-conditioned random init" (SURVEY.md §7 hard part 3): fan-in-scaled linears so that
attention logits have O(1) spread, identity cross-attention `v_proj` as the reference constructs it
(JES:320-322), and Gaussian codebooks whose scale is matched to the projected residuals so that all 512 codes of
all 4 levels are in play.  It is NOT the HF default init (std 0.02), under which every token collapses onto
~4 codes.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict

import torch

ENC = "audio_joint_encoder_segmenter.audio_encoder.encoder."
DEC = "audio_joint_encoder_segmenter.audio_segmenter.decoder."
RVQ = "vq.rvq."


@dataclass(frozen=True)
class TowerConfig:
    """Geometry of the audio tower.  Defaults = distil-large-v3 + CFG:146-155."""
    d_model: int = 1280
    enc_layers: int = 32
    dec_layers: int = 2
    heads: int = 20
    ffn: int = 5120
    vocab: int = 51866
    n_mels: int = 128
    max_source_positions: int = 1500
    max_target_positions: int = 448
    codebook_dim: int = 256
    codebook_size: int = 512
    num_quantizers: int = 4
    target_hidden_layer: int = 6

    def as_dict(self):
        return asdict(self)


FULL = TowerConfig()
TINY = TowerConfig(d_model=128, enc_layers=8, heads=2, ffn=256)
SMALL = TowerConfig(d_model=384, enc_layers=8, heads=6, ffn=1024)


def state_dict_spec(cfg: TowerConfig) -> "OrderedDict[str, tuple]":
    D, FF, dc, K = cfg.d_model, cfg.ffn, cfg.codebook_dim, cfg.codebook_size
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def attn(p):
        s[p + "k_proj.weight"] = (D, D)
        s[p + "v_proj.weight"] = (D, D)
        s[p + "v_proj.bias"] = (D,)
        s[p + "q_proj.weight"] = (D, D)
        s[p + "q_proj.bias"] = (D,)
        s[p + "out_proj.weight"] = (D, D)
        s[p + "out_proj.bias"] = (D,)

    def ln(p):
        s[p + "weight"] = (D,)
        s[p + "bias"] = (D,)

    def mlp(p):
        s[p + "fc1.weight"] = (FF, D)
        s[p + "fc1.bias"] = (FF,)
        s[p + "fc2.weight"] = (D, FF)
        s[p + "fc2.bias"] = (D,)

    s[ENC + "conv1.weight"] = (D, cfg.n_mels, 3)
    s[ENC + "conv1.bias"] = (D,)
    s[ENC + "conv2.weight"] = (D, D, 3)
    s[ENC + "conv2.bias"] = (D,)
    s[ENC + "embed_positions.weight"] = (cfg.max_source_positions, D)
    for l in range(cfg.enc_layers):
        p = f"{ENC}layers.{l}."
        attn(p + "self_attn.")
        ln(p + "self_attn_layer_norm.")
        mlp(p)
        ln(p + "final_layer_norm.")
    ln(ENC + "layer_norm.")
    s[DEC + "embed_tokens.weight"] = (cfg.vocab, D)
    s[DEC + "embed_positions.weight"] = (cfg.max_target_positions, D)
    for l in range(cfg.dec_layers):
        p = f"{DEC}layers.{l}."
        attn(p + "self_attn.")
        ln(p + "self_attn_layer_norm.")
        attn(p + "encoder_attn.")
        ln(p + "encoder_attn_layer_norm.")
        mlp(p)
        ln(p + "final_layer_norm.")
    ln(DEC + "layer_norm.")
    s[RVQ + "project_in.weight"] = (dc, D)
    s[RVQ + "project_in.bias"] = (dc,)
    s[RVQ + "project_out.weight"] = (D, dc)
    s[RVQ + "project_out.bias"] = (D,)
    for q in range(cfg.num_quantizers):
        p = f"{RVQ}layers.{q}._codebook."
        s[p + "initted"] = (1,)
        s[p + "cluster_size"] = (1, K)
        s[p + "embed_avg"] = (1, K, dc)
        s[p + "embed"] = (1, K, dc)
    for m in range(cfg.num_quantizers - 1):          # unused-but-present QINCo MLPs (RVQ:155)
        p = f"{RVQ}mlps.{m}."
        s[p + "proj_in.weight"] = (dc, 2 * dc)
        s[p + "proj_in.bias"] = (dc,)
        for j in range(4):
            for k in (0, 2):
                s[p + f"layers.{j}.{k}.weight"] = (dc, dc)
                s[p + f"layers.{j}.{k}.bias"] = (dc,)
    return s


def sinusoids(length: int, channels: int) -> torch.Tensor:
    """Whisper encoder position table (CW:117-126): [sin(t*w_i) | cos(t*w_i)]."""
    inc = math.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def _randn(key: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    return torch.randn(shape, generator=g, dtype=torch.float32)


CODEBOOK_STD = 0.19     # minimises E min_j |r - c_j|^2 for r~N(0,I_256), 512 Gaussian codes (see module docstring)


def random_weights(cfg: TowerConfig, seed: int = 1234, only=None) -> Dict[str, torch.Tensor]:
    """`only`: optional predicate on the key, to draw a subset (values do not depend on which keys are drawn)."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        if only is not None and not only(key):
            continue
        r = None
        if key.endswith("embed_positions.weight") and key.startswith(ENC):
            r = sinusoids(shape[0], shape[1])
        elif key.endswith("_codebook.initted"):
            r = torch.ones(shape)
        elif key.endswith("_codebook.cluster_size"):
            r = torch.ones(shape)
        elif key.endswith("_codebook.embed") or key.endswith("_codebook.embed_avg"):
            r = _randn(key.replace("embed_avg", "embed"), shape, seed) * CODEBOOK_STD
        elif "layer_norm" in key:
            r = 1.0 + 0.1 * _randn(key, shape, seed) if key.endswith("weight") else 0.1 * _randn(key, shape, seed)
        elif key.endswith("encoder_attn.v_proj.weight"):
            r = torch.eye(shape[0])                      # JES:320-322 make_v_proj_identity
        elif key.endswith("encoder_attn.v_proj.bias"):
            r = torch.zeros(shape)
        elif key.endswith("embed_tokens.weight"):
            r = _randn(key, shape, seed)
        elif key.endswith("embed_positions.weight"):
            r = 0.5 * _randn(key, shape, seed)
        elif key.endswith(".bias"):
            r = 0.05 * _randn(key, shape, seed)
        elif key.endswith(".weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 1.0
            if "encoder_attn.q_proj" in key or "encoder_attn.k_proj" in key:
                gain = 3.0                               # sharp cross-attention: token/audio-specific pooling
            elif "encoder_attn.out_proj" in key:
                gain = 0.3                               # keeps the frame-mean component from swamping tokens
            elif "conv" in key:
                gain = 1.5
            elif "out_proj" in key or "fc2" in key:
                gain = 0.7
            r = (gain / math.sqrt(fan_in)) * _randn(key, shape, seed)
        else:
            raise KeyError(key)
        out[key] = r.contiguous()
    return out


def default_init_weights(cfg: TowerConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """The second weight set of SURVEY.md §7 hard part 3: what the reference's constructors leave behind without a
    checkpoint.  HF Whisper `_init_weights` (Linear / Conv1d / Embedding ~ N(0, 0.02^2), biases 0, LayerNorm 1 / 0,
    sinusoidal encoder positions), identity cross-attention v_proj (JES:320-322), torch's default Linear init for
    `project_in` / `project_out` (U(+-1/sqrt(fan_in))), N(0,1) codebooks marked initted.  Ill-conditioned on purpose:
    only a handful of codes are in play, so index agreement on it is REPORTED, never asserted."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        if key.endswith("embed_positions.weight") and key.startswith(ENC):
            r = sinusoids(shape[0], shape[1])
        elif key.endswith("_codebook.initted") or key.endswith("_codebook.cluster_size"):
            r = torch.ones(shape)
        elif key.endswith("_codebook.embed") or key.endswith("_codebook.embed_avg"):
            r = _randn(key.replace("embed_avg", "embed"), shape, seed)
        elif "layer_norm" in key:
            r = torch.ones(shape) if key.endswith("weight") else torch.zeros(shape)
        elif key.endswith("encoder_attn.v_proj.weight"):
            r = torch.eye(shape[0])
        elif key.startswith(RVQ):
            fan_in = shape[1] if len(shape) > 1 else {"project_in.bias": cfg.d_model, "project_out.bias": cfg.codebook_dim}.get(
                key[len(RVQ):], cfg.codebook_dim)
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
            r = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        elif key.endswith(".bias"):
            r = torch.zeros(shape)
        else:
            r = 0.02 * _randn(key, shape, seed)
        out[key] = r.contiguous()
    return out


def synth_waveform(seed: int, n_samples: int, total: int = None) -> torch.Tensor:
    """Deterministic speech-like test signal: 8 amplitude-modulated sinusoids 80-7600 Hz + white noise."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / 16000.0
    f = 80.0 + (7600.0 - 80.0) * torch.rand(8, generator=g, dtype=torch.float64)
    ph = 2 * math.pi * torch.rand(8, generator=g, dtype=torch.float64)
    am = 0.5 + 4.5 * torch.rand(8, generator=g, dtype=torch.float64)
    amp = 0.02 + 0.1 * torch.rand(8, generator=g, dtype=torch.float64)
    x = torch.zeros(n_samples, dtype=torch.float64)
    for i in range(8):
        env = 0.5 * (1.0 + torch.sin(2 * math.pi * am[i] * t + ph[i]))
        x += amp[i] * env * torch.sin(2 * math.pi * f[i] * t + ph[(i + 3) % 8])
    x += 0.01 * torch.randn(n_samples, generator=g, dtype=torch.float64)
    x = x.to(torch.float32)
    if total is not None and total > n_samples:
        x = torch.cat([x, torch.zeros(total - n_samples)])
    return x


def synth_transcript(seed: int, T: int, Tmax: int = None):
    """T token ids uniform in [0, 50257) and word ids = cumsum of Bernoulli(0.55) word starts."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    ids = torch.randint(0, 50257, (T,), generator=g, dtype=torch.int64)
    starts = (torch.rand(T, generator=g) < 0.55).to(torch.int32)
    starts[0] = 0
    wid = torch.cumsum(starts, 0).to(torch.int32)
    if Tmax is not None and Tmax > T:
        ids = torch.cat([ids, torch.zeros(Tmax - T, dtype=torch.int64)])
        wid = torch.cat([wid, torch.zeros(Tmax - T, dtype=torch.int32)])
    return ids, wid


def synth_batch(seed: int, durations_s, token_counts, pad_wave_to: int = None):
    """Batch in the boundary's input schema (MT:108-119 / DS sample schema): padded ids, lengths, word ids, waves."""
    B = len(durations_s)
    Tmax = max(token_counts)
    n_samples = [int(round(d * 16000)) for d in durations_s]
    total = pad_wave_to or max(n_samples)
    wav = torch.stack([synth_waveform(seed * 7919 + b, n_samples[b], total) for b in range(B)])
    ids, wids = zip(*[synth_transcript(seed * 104729 + b, token_counts[b], Tmax) for b in range(B)])
    return {
        "wav": wav, "n_samples": torch.tensor(n_samples, dtype=torch.int32),
        "asr_token_ids": torch.stack(ids), "asr_token_lengths": torch.tensor(token_counts, dtype=torch.int32),
        "asr_word_ids": torch.stack(wids),
    }


class StubTokenizer:
    """Deterministic stand-in for the Whisper / Llama tokenizers (no tokenizer files exist offline): a word is cut into
    pieces of `piece` characters and each piece hashed into [0, vocab).  Has the two call forms process_one_sample uses
    (DS:71-95): `.encode(word, add_special_tokens=False)` and `tok(list_of_words, add_special_tokens=False).input_ids`."""

    def __init__(self, vocab: int = 50257, piece: int = 3, salt: int = 0):
        self.vocab, self.piece, self.salt = int(vocab), int(piece), int(salt)

    def encode(self, word: str, add_special_tokens: bool = False):
        import zlib
        return [(zlib.crc32((word[i: i + self.piece] + str(self.salt)).encode()) % self.vocab)
                for i in range(0, len(word), self.piece)]

    def __call__(self, words, add_special_tokens: bool = False):
        class _Enc:
            pass
        e = _Enc()
        e.input_ids = [self.encode(w) for w in words]
        return e


_WORDS = ("the quick brown fox jumps over a lazy dog while seventeen purple elephants quietly contemplate "
          "extraordinary circumstances beyond their immediate understanding").split()


def synth_text(seed: int, n_words: int) -> str:
    g = torch.Generator().manual_seed(int(seed))
    idx = torch.randint(0, len(_WORDS), (n_words,), generator=g).tolist()
    return " ".join(_WORDS[i] for i in idx)


def synth_pcm(seed: int, n: int, channels: int = 1):
    """Decoded-PCM stand-in for sample['mp3']['array'] (DS:46): band-limited noise + tones, fp32 [n] or [C, n]."""
    import numpy as np
    g = torch.Generator().manual_seed(int(seed))
    t = torch.arange(n, dtype=torch.float64)
    x = 0.05 * torch.randn(channels, n, generator=g, dtype=torch.float64)
    for _ in range(4):
        f = float(torch.rand(1, generator=g)) * 0.2 + 0.002            # cycles per sample, below every Nyquist used
        ph = float(torch.rand(1, generator=g)) * 6.283185307179586
        x += 0.1 * torch.sin(2 * 3.141592653589793 * f * t + ph)[None] * torch.rand(channels, 1, generator=g, dtype=torch.float64)
    x = x.to(torch.float32).numpy()
    return x[0] if channels == 1 else np.ascontiguousarray(x)


class SynthCorpus:
    """BASELINE configs 3 / 4: a host-resident corpus of ragged utterances (no dataset exists offline).

    Durations ~ U[1, 30] s; transcript length T = clip(round(2.7 * duration + N(0, 2)), 1, 443); word ids
    as in `synth_transcript`; an llm tokenisation of the same words with 1-3 sub-word tokens per word (the columns the
    reference's extract_vq job writes, XV:51-58).  Decoded audio lives in ordinary host memory as the data loader would
    leave it: `pool` distinct 30 s waveforms, utterance u plays `pool[u % pool][:n_samples[u]]` (holding 16 384 distinct
    waveforms would take 31 GB and change nothing the path can see).  Implements both corpus protocols of
    `shard.tokenize_corpus`: `fetch_host(indices, slot)` (pinned staging, the corpus job) and `load(indices)` (device).
    """

    wav_stride = 480000

    def __init__(self, n_utts: int, seed: int = 4, pool: int = 64, max_tokens: int = 443):
        import numpy as np
        rng = np.random.default_rng(seed)
        self.n = int(n_utts)
        self.durations = rng.uniform(1.0, 30.0, self.n)
        tc = np.clip(np.round(2.7 * self.durations + rng.normal(0.0, 2.0, self.n)), 1, max_tokens).astype(np.int64)
        self.token_counts = tc.tolist()
        self.n_samples = np.minimum((self.durations * 16000).astype(np.int64), self.wav_stride)
        self.pool = np.stack([synth_waveform(seed * 7919 + p, self.wav_stride).numpy() for p in range(int(pool))])
        self.off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(tc, out=self.off[1:])
        total = int(self.off[-1])
        self.ids = rng.integers(0, 50257, total, dtype=np.int64)
        starts = (rng.random(total) < 0.55)
        starts[self.off[:-1]] = False
        c = np.cumsum(starts)
        self.wid = (c - np.repeat(c[self.off[:-1]], tc)).astype(np.int32)            # word id of every asr token
        n_words = self.wid[self.off[1:] - 1].astype(np.int64) + 1
        self.word_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(n_words, out=self.word_off[1:])
        pieces = rng.integers(1, 4, int(self.word_off[-1]))
        local_word = np.arange(int(self.word_off[-1])) - np.repeat(self.word_off[:-1], n_words)
        self.llm_wid = np.repeat(local_word, pieces).astype(np.int32)
        lcount = np.add.reduceat(pieces, self.word_off[:-1])
        self.llm_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(lcount, out=self.llm_off[1:])
        self.llm_ids = rng.integers(0, 128256, int(self.llm_off[-1]), dtype=np.int64)
        self.llm_counts = lcount
        self.max_llm_tokens = int(lcount.max())
        self.audio_seconds = float(self.n_samples.sum() / 16000.0)

    def fetch_host(self, indices, slot):
        import numpy as np
        idx = np.asarray(indices, dtype=np.int64)
        B = len(idx)
        tc = np.asarray([self.token_counts[i] for i in idx], dtype=np.int64)
        T, L = int(tc.max()), int(self.llm_counts[idx].max())
        wav, ids, wid, lwid = slot.wav_h.numpy(), slot.ids_h.numpy(), slot.wid_h.numpy(), slot.lwid_h.numpy()
        with_llm = lwid.shape[1] >= L                 # the driver sizes the llm staging only when it maps to llm tokens
        ids[:B, :T] = 0
        wid[:B, :T] = 0
        if with_llm:
            lwid[:B, :L] = 0
        llm_ids = np.zeros((B, L), dtype=np.int64)
        for r, u in enumerate(idx):
            n = int(self.n_samples[u])
            wav[r, :n] = self.pool[u % len(self.pool), :n]
            a0, a1 = self.off[u], self.off[u + 1]
            ids[r, : a1 - a0] = self.ids[a0:a1]
            wid[r, : a1 - a0] = self.wid[a0:a1]
            if with_llm:
                l0, l1 = self.llm_off[u], self.llm_off[u + 1]
                lwid[r, : l1 - l0] = self.llm_wid[l0:l1]
                llm_ids[r, : l1 - l0] = self.llm_ids[l0:l1]
        slot.ns_h.numpy()[:B] = self.n_samples[idx]
        slot.len_h.numpy()[0, :B] = tc
        slot.len_h.numpy()[1, :B] = self.llm_counts[idx]
        return {"B": B, "T": T, "L": L, "lengths_host": tc, "llm_lengths_host": self.llm_counts[idx].copy(),
                "max_n": int(self.n_samples[idx].max()), "llm_ids": llm_ids}

    def load(self, indices, device="cuda"):
        """Device-resident protocol (small jobs / tests): the same rows as tensors on `device`."""
        import numpy as np
        idx = np.asarray(indices, dtype=np.int64)
        tc = np.asarray([self.token_counts[i] for i in idx], dtype=np.int64)
        T = int(tc.max())
        wav = torch.zeros(len(idx), self.wav_stride)
        ids = torch.zeros(len(idx), T, dtype=torch.int64)
        wid = torch.zeros(len(idx), T, dtype=torch.int32)
        for r, u in enumerate(idx):
            n = int(self.n_samples[u])
            wav[r, :n] = torch.from_numpy(self.pool[u % len(self.pool), :n])
            ids[r, : tc[r]] = torch.from_numpy(self.ids[self.off[u]: self.off[u + 1]])
            wid[r, : tc[r]] = torch.from_numpy(self.wid[self.off[u]: self.off[u + 1]])
        return {"wav": wav.to(device), "n_samples": torch.from_numpy(self.n_samples[idx].astype(np.int32)).to(device),
                "ids": ids.to(device), "wid": wid.to(device), "lengths_host": tc}
