"""Drop-in for `WhisperFrontend` (WF:7-113): 16 kHz waveform -> Whisper log-mel, computed by the fused CUDA kernel.

Same constructor and `forward(input, input_lengths) -> (feats, feats_lens)` contract as the reference, including its
quirks: every row is padded / trimmed to 30 s (WF:98-99), `feats_lens` is `input_lengths[0] // 160` for EVERY row
(WF:102), output is time-major `[B, 3000, 128]` when `permute=True` (WF:111-112) and is returned on the input's
device (the reference runs on CPU inside DataLoader workers, PT:246 / DS:64; here the tensor makes a round trip to the
GPU).  `forward_device` is the zero-copy entry the corpus driver uses.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from .engine import FrontendEngine


class WhisperFrontendB200(nn.Module):
    def __init__(self, fs: int = 16000, whisper_model: str = None, do_pad_trim: bool = True, n_mels: int = 80,
                 permute: bool = False, **kwargs):
        super().__init__()
        assert fs == 16000
        self.fs = fs
        self.n_fft = 400
        self.win_length = 400
        self.hop_length = 160
        self.pad_samples = _lib.N_SAMPLES
        self.frame_shift = int(self.hop_length / self.fs * 1000)
        self.lfr_n = 1
        self.n_mels = n_mels
        if whisper_model == "large-v3" or whisper_model == "large":
            self.n_mels = 128
        if self.n_mels != 128:
            raise NotImplementedError("the B200 log-mel kernel is built for the 128-bin large-v3 filterbank (PT:164-168)")
        if kwargs.get("filters_path", None) is not None:
            raise NotImplementedError("custom filters_path (WF:37-42) is not supported")
        if not do_pad_trim:
            raise NotImplementedError("do_pad_trim=False is not used by TASTE (PT:166, DS:242)")
        self.do_pad_trim = do_pad_trim
        self.permute = permute
        self.precision = _lib.resolve_precision(kwargs.get("precision"))   # 16-bit dtype of forward_device's features
        self._device_hint: Optional[torch.device] = None
        self._engine: Optional[FrontendEngine] = None

    def output_size(self) -> int:
        return self.n_mels

    def to(self, *args, **kwargs):
        dev = args[0] if args else kwargs.get("device")
        if dev is not None and not isinstance(dev, torch.dtype):
            self._device_hint = torch.device(dev)
        return super().to(*args, **kwargs)

    def engine(self, device=None) -> FrontendEngine:
        if not torch.cuda.is_available():
            raise _lib.TasteError("WhisperFrontendB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.device(device or self._device_hint or "cuda")
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._engine is None or self._engine.device != device:
            self._engine = FrontendEngine(device, self.precision)
        return self._engine

    def forward_device(self, wav: torch.Tensor, n_samples: torch.Tensor, want_f32=True, want_bf16=False):
        """wav fp32 [B, N] already on the GPU; n_samples int32 [B].  Returns (feats_f32|None, feats_bf16|None)."""
        return self.engine(wav.device).logmel(wav, n_samples, want_f32, want_bf16)

    def forward(self, input: torch.Tensor, input_lengths, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        batch_size = input.size(0)
        src_device = input.device
        eng = self.engine(src_device if src_device.type == "cuda" else None)
        wav = input.to(device=eng.device, dtype=torch.float32)
        if wav.dim() != 2:
            raise ValueError("expected [B, N] waveforms")
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        n = min(wav.shape[1], self.pad_samples)                       # rows are full width (WF:98: trim to 30 s)
        n_samples = torch.full((batch_size,), n, dtype=torch.int32, device=eng.device)
        f32, _ = eng.logmel(wav, n_samples, True, False)
        feats = f32 if self.permute else f32.permute(0, 2, 1)
        first_len = int(input_lengths[0]) if not torch.is_tensor(input_lengths) else int(input_lengths.reshape(-1)[0])
        feats_lens = torch.as_tensor([first_len // self.hop_length for _ in range(batch_size)])   # WF:74-75,102
        return feats.to(src_device), feats_lens
