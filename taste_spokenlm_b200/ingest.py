"""GPU-side ingest for the corpus job (reference: taste_speech/data/dataset.py `process_one_sample`,
DS:37-113, as driven by scripts/extract_vq_for_stage2_training.py, XV:39-74, XV:137-165).

The reference turns every arrow row into model inputs on the CPU, one sample at a time, inside two DataLoader
workers per GPU (XV:143): `torchaudio.transforms.Resample(orig_sr, 16000)(speech_pt).mean(0)` (DS:52-60), the Whisper
log-mel front-end (DS:63-64) and a per-word tokenizer loop (DS:71-95).  At GPU speed that is the bottleneck by more
than 20x.  Here the decoded PCM arrays of a whole batch are packed into one pinned buffer, copied once, and

    resample + channel mean   `taste_resample_mean_f32`  (csrc/resample.cu)
    log-mel                   `taste_logmel_f32`
    tower                     `TowerEngine.segment_and_quantize`
    llm-token mapping         `taste_map_to_llm_tokens`

run back to back on the device; the rows go to `shard.ShardWriter` (the reference's column names, streaming and
resumable).  The transcript side (`split_transcript`, DS:71-95) stays on the host: it is tokenizer calls.

Nothing here falls back to the CPU: without the CUDA library or a device, construction raises `TasteError`.
"""
from __future__ import annotations

import ctypes as C
import math
import re
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import TasteError

TARGET_SR = 16_000                    # DS:39
N_SAMPLES = _lib.N_SAMPLES


# --------------------------------------------------------------------------------------------------------------
# polyphase tables: torchaudio.functional._get_sinc_resample_kernel (third party, torchaudio==2.3.1) in sparse form
# --------------------------------------------------------------------------------------------------------------
def polyphase_taps(orig_freq: int, new_freq: int = TARGET_SR, lowpass_filter_width: int = 6, rolloff: float = 0.99,
                   eps: float = 1e-20) -> Dict[str, object]:
    """The windowed-sinc kernel `transforms.Resample(orig_freq, new_freq)` builds (sinc_interp_hann), as the kernel wants it.

    Returns dict(orig, new, width, knz, knz_ld, taps fp32 [new, knz_ld], kstart int32 [new]).  Per phase the dense run
    starts at the first tap with |tap| > eps: what is dropped is the clamped Hann window's tail (cos(pi/2)^2 ~ 4e-33 in
    float64), below 1e-20 of the signal.
    """
    orig_freq, new_freq = int(orig_freq), int(new_freq)
    if orig_freq <= 0 or new_freq <= 0:
        raise ValueError("sampling rates must be positive integers")
    g = math.gcd(orig_freq, new_freq)
    orig, new = orig_freq // g, new_freq // g
    if orig == new:
        return dict(orig=1, new=1, width=0, knz=1, knz_ld=1, taps=np.ones((1, 1), np.float32), kstart=np.zeros(1, np.int32))
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torchaudio evaluates arange(0, -new, -1) / new in float32 before adding the float64 grid
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = np.clip((phase + idx) * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    full = (k * (window * (base / orig))).astype(np.float32)                  # [new, 2*width+orig]
    sig = np.abs(full) > eps
    first = sig.argmax(1)
    last = full.shape[1] - 1 - sig[:, ::-1].argmax(1)
    knz = int((last - first + 1).max())
    kstart = np.minimum(first, full.shape[1] - knz).astype(np.int32)          # keep every run inside the kernel
    knz_ld = knz | 1                                                          # odd row stride: conflict-free smem rows
    taps = np.zeros((new, knz_ld), np.float32)
    for p in range(new):
        taps[p, :knz] = full[p, kstart[p]: kstart[p] + knz]
    return dict(orig=orig, new=new, width=int(width), knz=knz, knz_ld=int(knz_ld), taps=taps, kstart=kstart)


# --------------------------------------------------------------------------------------------------------------
# DS:71-95  transcript -> word-aligned asr / llm token ids
# --------------------------------------------------------------------------------------------------------------
def split_transcript(text: str, whisper_tokenizer, llm_tokenizer) -> Tuple[List[int], List[int], List[int], List[int]]:
    """(asr_token_ids, asr_word_ids, llm_token_ids, llm_word_ids) exactly as process_one_sample builds them."""
    text = text.strip()
    words = [" " + w for w in re.split(r"\s", text)]
    words[0] = words[0].lstrip()
    a_ids: List[int] = []
    a_wid: List[int] = []
    l_ids: List[int] = []
    l_wid: List[int] = []
    for i, word in enumerate(words):
        for t in whisper_tokenizer.encode(word, add_special_tokens=False):
            a_ids.append(int(t))
            a_wid.append(i)
        for t in llm_tokenizer.encode(word, add_special_tokens=False):
            l_ids.append(int(t))
            l_wid.append(i)
    return a_ids, a_wid, l_ids, l_wid


def _pad_rows(rows: Sequence[Sequence[int]], dtype, fill: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    lens = np.asarray([len(r) for r in rows], dtype=np.int32)
    width = max(int(lens.max()) if len(rows) else 0, 1)
    out = np.full((len(rows), width), fill, dtype=dtype)
    for i, r in enumerate(rows):
        out[i, : len(r)] = r
    return out, lens


# --------------------------------------------------------------------------------------------------------------
# device side
# --------------------------------------------------------------------------------------------------------------
class ResampleMeanB200:
    """`resampler(speech_pt).mean(0)` (DS:52-60) for a batch of PCM arrays, on the device.

    Mirrors the reference's `resampler_dict`: one table set per source rate, built on first use.
    """

    def __init__(self, device, new_freq: int = TARGET_SR, wav_stride: int = N_SAMPLES):
        device = torch.device(device)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise TasteError("ResampleMeanB200 needs a CUDA device (there is no CPU fallback)")
        self.device = device
        self.lib = _lib.load("bf16")          # the resampler is fp32 end to end: either flavour would do
        self.new_freq = int(new_freq)
        self.wav_stride = int(wav_stride)
        self._tables: Dict[int, Dict[str, object]] = {}
        self._pinned: Optional[torch.Tensor] = None
        self._dev_in: Optional[torch.Tensor] = None

    def tables(self, orig_freq: int) -> Dict[str, object]:
        tb = self._tables.get(int(orig_freq))
        if tb is None:
            tb = polyphase_taps(orig_freq, self.new_freq)
            tb["taps_dev"] = torch.from_numpy(tb["taps"]).to(self.device)
            tb["kstart_dev"] = torch.from_numpy(tb["kstart"]).to(self.device)
            self._tables[int(orig_freq)] = tb
        return tb

    def _staging(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Two pinned / device staging pairs used in rotation.  The H2D copy is asynchronous, so before the host writes
        a pinned buffer again it waits for the event recorded after that buffer's previous copy; the device buffer is
        protected by stream order (copy and kernel are issued on the same stream).  With two pairs the host packs
        batch i + 1 while batch i is still in flight."""
        self._slot = (getattr(self, "_slot", -1) + 1) % 2
        if not hasattr(self, "_slots"):
            self._slots = [None, None]
        st = self._slots[self._slot]
        if st is None or st[0].numel() < n:
            if st is not None and st[2] is not None:
                st[2].synchronize()
            cap = max(n, 1 << 20)
            st = [torch.empty(cap, dtype=torch.float32).pin_memory(),
                  torch.empty(cap, dtype=torch.float32, device=self.device), None]
            self._slots[self._slot] = st
        if st[2] is not None:
            st[2].synchronize()               # the previous DMA out of this pinned buffer has finished
        self._pinned, self._dev_in = st[0], st[1]
        return st[0], st[1]

    def output_lengths(self, n_in: Sequence[int], orig_freq: int) -> np.ndarray:
        tb = self.tables(orig_freq)
        n = np.asarray(n_in, dtype=np.int64)
        return (tb["new"] * n + tb["orig"] - 1) // tb["orig"]

    def run_device(self, packed: torch.Tensor, offsets: np.ndarray, channels: np.ndarray, n_in: np.ndarray,
                   orig_freq: int, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """packed fp32 on the device (utterance b = [channels[b], n_in[b]] at offsets[b]).  Returns (wav [B, stride],
        n_samples int32 [B] on the device, clipped to the 30 s window as `pad_or_trim` does, WF:98-99)."""
        tb = self.tables(orig_freq)
        B = len(n_in)
        if out is None:
            out = torch.empty(B, self.wav_stride, dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.shape[0] >= B and out.stride(1) == 1
        n_out = torch.empty(B, dtype=torch.int32, device=self.device)
        meta = torch.from_numpy(np.concatenate([np.asarray(offsets, np.int64),
                                                np.asarray(channels, np.int64), np.asarray(n_in, np.int64)]))
        meta = meta.to(self.device, non_blocking=True)
        off_d = meta[: B + 1]
        ch_d = meta[B + 1: 2 * B + 1].to(torch.int32)
        nin_d = meta[2 * B + 1:].to(torch.int32)
        tgt = self.output_lengths(n_in, orig_freq)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):            # the library configures and launches on the CURRENT device
            _lib.check(self.lib.taste_resample_mean_f32(
                _lib.ptr(packed), _lib.ptr(off_d), _lib.ptr(ch_d), _lib.ptr(nin_d), B, tb["orig"], tb["new"], tb["width"],
                _lib.ptr(tb["taps_dev"]), _lib.ptr(tb["kstart_dev"]), tb["knz"], tb["knz_ld"], int(tgt.max()) if B else 0,
                int(offsets[-1]), int(np.minimum(tgt, self.wav_stride).sum()), _lib.ptr(out), out.stride(0),
                _lib.ptr(n_out), stream), "taste_resample_mean_f32")
        return out, torch.clamp(n_out, max=self.wav_stride)

    def __call__(self, arrays: Sequence[np.ndarray], orig_freq: int, out: Optional[torch.Tensor] = None):
        """arrays: decoded PCM, each [n] or [C, n] (sample['mp3']['array'], DS:46).  One pinned pack + one H2D copy."""
        arrs = [np.asarray(a, dtype=np.float32) for a in arrays]
        arrs = [a[None] if a.ndim == 1 else a for a in arrs]                   # DS:53-54
        channels = np.asarray([a.shape[0] for a in arrs], dtype=np.int64)
        n_in = np.asarray([a.shape[1] for a in arrs], dtype=np.int64)
        offsets = np.zeros(len(arrs) + 1, dtype=np.int64)
        np.cumsum(channels * n_in, out=offsets[1:])
        total = int(offsets[-1])
        pinned, dev_in = self._staging(max(total, 1))
        host = pinned.numpy()
        for a, o in zip(arrs, offsets[:-1]):
            host[o: o + a.size] = a.reshape(-1)
        dev_in[:total].copy_(pinned[:total], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._slots[self._slot][2] = ev
        return self.run_device(dev_in, offsets, channels, n_in, orig_freq, out=out)


class CorpusIngestB200:
    """arrow rows -> llm-aligned RVQ indices, written through `ShardWriter` (the XV:39-74 loop, batched on the GPU).

    `tower`: a `TasteAudioTowerB200` on the device (eval).  Tokenizers: objects with
    `.encode(word, add_special_tokens=False)` (the reference passes `WhisperProcessor.tokenizer` and the Llama tokenizer).
    Rows are dicts with the reference's arrow schema: row['mp3']['array'], row['mp3']['sampling_rate'],
    row['json']['text'] (DS:46-50); s3 tokens / speaker embeddings are not needed for the VQ extraction columns.
    """

    def __init__(self, tower, whisper_tokenizer, llm_tokenizer, batch_size: int = 64, max_asr_tokens: int = 443):
        self.tower = tower
        self.engine = tower.engine()
        self.device = self.engine.device
        self.resample = ResampleMeanB200(self.device)
        self.whisper_tokenizer = whisper_tokenizer
        self.llm_tokenizer = llm_tokenizer
        self.batch_size = int(batch_size)
        self.max_asr_tokens = int(max_asr_tokens)         # 448 decoder positions - 4 prefix - 1 eos (MT:144-152)

    def text_side(self, rows: Sequence[dict]):
        sp = [split_transcript(r["json"]["text"], self.whisper_tokenizer, self.llm_tokenizer) for r in rows]
        return sp

    def tokenize_batch(self, rows: Sequence[dict], texts=None):
        """One batch of rows (any mix of sampling rates).  Returns per-row (llm_indices [L,Q] int64 np, llm_token_ids,
        llm_word_ids) in row order."""
        if texts is None:
            texts = self.text_side(rows)
        B = len(rows)
        wav = torch.empty(B, self.resample.wav_stride, dtype=torch.float32, device=self.device)
        n_samples = torch.empty(B, dtype=torch.int32, device=self.device)
        by_rate: Dict[int, List[int]] = {}
        for i, r in enumerate(rows):
            by_rate.setdefault(int(r["mp3"]["sampling_rate"]), []).append(i)
        for rate, idxs in by_rate.items():
            if len(by_rate) == 1:
                _, ns = self.resample([rows[i]["mp3"]["array"] for i in idxs], rate, out=wav)
                n_samples = ns
            else:
                w, ns = self.resample([rows[i]["mp3"]["array"] for i in idxs], rate)
                sel = torch.as_tensor(idxs, device=self.device)
                wav.index_copy_(0, sel, w)
                n_samples.index_copy_(0, sel, ns)
        a_ids, a_len = _pad_rows([t[0] for t in texts], np.int64)
        a_wid, _ = _pad_rows([t[1] for t in texts], np.int32)
        l_ids, l_len = _pad_rows([t[2] for t in texts], np.int64)
        l_wid, _ = _pad_rows([t[3] for t in texts], np.int32)
        if int(a_len.max()) > self.max_asr_tokens:
            raise TasteError(f"transcript of {int(a_len.max())} asr tokens exceeds the aggregator's {self.max_asr_tokens}")
        if int(a_len.min()) < 1:
            raise TasteError("empty transcript")
        ids_d = torch.from_numpy(a_ids).to(self.device, non_blocking=True)
        wid_d = torch.from_numpy(a_wid).to(self.device, non_blocking=True)
        _, idx = self.engine.tokenize_device(wav, n_samples, ids_d, wid_d, a_len, want_quantized=False)
        llm_idx = self.engine.map_to_llm_tokens(idx, wid_d, torch.from_numpy(a_len), torch.from_numpy(l_wid),
                                                torch.from_numpy(l_len)).cpu().numpy()
        return [(llm_idx[i, : l_len[i]], l_ids[i, : l_len[i]], l_wid[i, : l_len[i]]) for i in range(B)]

    def run(self, rows: Iterable[dict], writer, utt_ids: Optional[Iterable[int]] = None, world_size: int = 1,
            rank: int = 0) -> int:
        """Stream `rows` (this rank takes every world_size-th row, DistributedSampler-style as XV:137-146), skip what
        the writer already holds, tokenize in batches, add to the writer.  Returns the number of rows written."""
        done = 0
        batch_rows: List[dict] = []
        batch_ids: List[int] = []

        def flush():
            nonlocal done
            if not batch_rows:
                return
            texts = self.text_side(batch_rows)
            order = np.argsort([len(t[0]) for t in texts], kind="stable")      # tight padding inside the batch
            out = self.tokenize_batch([batch_rows[i] for i in order], [texts[i] for i in order])
            for pos, i in enumerate(order):
                li, lt, lw = out[pos]
                writer.add(batch_ids[i], li, lt, lw)
            done += len(batch_rows)
            batch_rows.clear()
            batch_ids.clear()

        ids = iter(utt_ids) if utt_ids is not None else None
        for n, row in enumerate(rows):
            uid = int(next(ids)) if ids is not None else n
            if n % world_size != rank or writer.is_done(uid):
                continue
            batch_rows.append(row)
            batch_ids.append(uid)
            if len(batch_rows) == self.batch_size:
                flush()
        flush()
        writer.flush()
        return done
