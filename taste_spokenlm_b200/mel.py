"""Host-side tables of the log-mel kernel: Slaney mel filterbank (sparse form), periodic Hann window and the folded
DFT twiddles.  Replaces `whisper.audio.mel_filters` / `torch.hann_window` at WF:44,61,66-70.

openai-whisper ships `mel_filters.npz['mel_128']` = librosa.filters.mel(sr=16000, n_fft=400, n_mels=128) (Slaney mel
scale, Slaney area normalisation, fmin 0, fmax 8000).  The filterbank is rebuilt here from that definition in fp64
and rounded to fp32 once.
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
N_BINS = 201
N_MELS = 128
SR = 16000
DFT_LD = 224
MEL_MAXW = 16

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    lin = m * _F_SP
    return np.where(m >= _MIN_LOG_MEL, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), lin)


def _hz_to_mel(f: float) -> float:
    return f / _F_SP if f < _MIN_LOG_HZ else _MIN_LOG_MEL + np.log(f / _MIN_LOG_HZ) / _LOGSTEP


def slaney_filterbank(n_mels: int = N_MELS) -> np.ndarray:
    """Dense [n_mels, 201] fp32 filterbank."""
    freqs = np.arange(N_BINS, dtype=np.float64) * (SR / N_FFT)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SR / 2.0), n_mels + 2))
    fb = np.zeros((n_mels, N_BINS), dtype=np.float64)
    for m in range(n_mels):
        lo, ce, hi = edges[m], edges[m + 1], edges[m + 2]
        up = (freqs - lo) / (ce - lo)
        down = (hi - freqs) / (hi - ce)
        fb[m] = np.clip(np.minimum(up, down), 0.0, None) * (2.0 / (hi - lo))
    return fb.astype(np.float32)


def sparse_filterbank(n_mels: int = N_MELS):
    """(start [n_mels] i32, count [n_mels] i32, weight [n_mels, MEL_MAXW] f32): the contiguous non-zero span of each filter."""
    fb = slaney_filterbank(n_mels)
    start = np.zeros(n_mels, dtype=np.int32)
    count = np.zeros(n_mels, dtype=np.int32)
    weight = np.zeros((n_mels, MEL_MAXW), dtype=np.float32)
    for m in range(n_mels):
        nz = np.nonzero(fb[m])[0]
        if len(nz) == 0:
            continue
        s, e = int(nz[0]), int(nz[-1]) + 1
        if e - s > MEL_MAXW:
            raise ValueError(f"mel filter {m} spans {e - s} bins > {MEL_MAXW}")
        start[m], count[m] = s, e - s
        weight[m, : e - s] = fb[m, s:e]
    return start, count, weight


def hann_periodic() -> np.ndarray:
    n = np.arange(N_FFT, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)).astype(np.float32)


DFT_K = 448      # TASTE_DFT_K
DFT_N = 512      # TASTE_DFT_N


def _bf16_round(x32: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bf16, returned as float32 values."""
    u = x32.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def dft_gemm_weights() -> np.ndarray:
    """B operand of the tensor-core DFT: uint16 (bf16 bits) [DFT_N, 3 * DFT_K].

    Row j is output column j of the GEMM: Re X[j] for j = 0..200, Im X[j - 256] (up to sign) for j = 256..456, zero
    otherwise.  With t[j][n] = hann[n] * cos|sin(2 pi bin n / 400) in fp32 (n < 400; zero up to DFT_K), the three K slabs
    are hi(t), lo(t), hi(t), to be multiplied with the frame slabs hi(x), hi(x), lo(x): x t ~= hi hi + hi lo + lo hi."""
    n = np.arange(N_FFT, dtype=np.int64)[None, :]
    k = np.arange(N_BINS, dtype=np.int64)[:, None]
    ang = 2.0 * np.pi * ((n * k) % N_FFT).astype(np.float64) / N_FFT
    win = hann_periodic().astype(np.float64)[None, :]
    t = np.zeros((DFT_N, DFT_K), dtype=np.float32)
    t[:N_BINS, :N_FFT] = (win * np.cos(ang)).astype(np.float32)
    t[256:256 + N_BINS, :N_FFT] = (win * np.sin(ang)).astype(np.float32)
    hi = _bf16_round(t)
    lo = _bf16_round(t - hi)
    w = np.concatenate([hi, lo, hi], axis=1)
    return (w.view(np.uint32) >> 16).astype(np.uint16)


def dft_tables():
    """cos/sin(2*pi*k*n/400) for n = 1..199 (row n-1; row 199 is zero) and k = 0..200, leading dim DFT_LD."""
    n = np.arange(1, 200, dtype=np.int64)[:, None]
    k = np.arange(N_BINS, dtype=np.int64)[None, :]
    ang = 2.0 * np.pi * ((n * k) % N_FFT).astype(np.float64) / N_FFT
    c = np.zeros((200, DFT_LD), dtype=np.float32)
    s = np.zeros((200, DFT_LD), dtype=np.float32)
    c[:199, :N_BINS] = np.cos(ang)
    s[:199, :N_BINS] = np.sin(ang)
    return c, s
