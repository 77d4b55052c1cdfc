"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's conv-STFT frontends.

numpy builds the tables (independently of the product's torch tables), torch-CPU fp32 does
the heavy contractions so this file can also serve as the multi-threaded CPU baseline.

Follows:
  windows ........ FSMN/STFT_Process.py:36-58; FireRedVAD/STFT_Process.py:89-114;
                   NVIDIA_*/STFT_Process.py:94; DFSMN/near_and_far_end_audio/STFT_Process.py:37-60
  conv kernels ... FSMN/STFT_Process.py:87-98 (v1), FireRedVAD/STFT_Process.py:203-216 (v2)
  zero centre pad  FSMN/STFT_Process.py:144-157; FireRedVAD/STFT_Process.py:264-278
  mel banks ...... FireRedVAD/Export_FireRedVAD.py:122-169; torchaudio melscale_fbanks
                   (FSMN/Export_FSMN_VAD.py:63, NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:186-189)
"""
from __future__ import annotations

import math

import numpy as np
import torch

f32 = np.float32


def window(kind: str, n: int) -> np.ndarray:
    k = np.arange(n, dtype=np.float64)
    if kind in ("hamming", "dfsmn_bartlett"):
        w = 0.54 - 0.46 * np.cos(2 * np.pi * k / n)
    elif kind == "hann":
        w = 0.5 - 0.5 * np.cos(2 * np.pi * k / n)
    elif kind == "hann_sym":
        w = 0.5 - 0.5 * np.cos(2 * np.pi * k / (n - 1))
    elif kind == "povey":
        w = (0.5 - 0.5 * np.cos(2 * np.pi * k / (n - 1))).astype(f32).astype(np.float64) ** 0.85
    else:
        raise ValueError(kind)
    return w.astype(f32)


def stft_kernel(n_fft: int, win_length: int, kind: str, flavour: str) -> np.ndarray:
    """[2F, n_fft] conv kernel (cos rows then -sin rows), window centred in n_fft."""
    F = n_fft // 2 + 1
    w = np.zeros(n_fft, f32)
    pl = (n_fft - win_length) // 2
    w[pl:pl + win_length] = window(kind, win_length)
    t = np.arange(n_fft, dtype=f32)[None, :]
    f = np.arange(F, dtype=f32)[:, None]
    if flavour == "v1":
        om = (f32(2 * np.pi) * f * t) / f32(n_fft)
    else:
        om = (f32(2.0 * np.pi / n_fft) * f) * t
    om = om.astype(f32)
    return np.concatenate([np.cos(om).astype(f32) * w, -np.sin(om).astype(f32) * w], 0).astype(f32)


def stft_power(x: torch.Tensor, kernel: np.ndarray, hop: int, center_pad: bool) -> torch.Tensor:
    """x [N,1,L] fp32 -> power [N,F,T] (re^2+im^2), the reference's strided conv1d."""
    k = torch.from_numpy(kernel).unsqueeze(1)
    F = kernel.shape[0] // 2
    if center_pad:
        half = kernel.shape[1] // 2
        x = torch.nn.functional.pad(x, (half, half))
    y = torch.nn.functional.conv1d(x, k, stride=hop)
    re, im = y[:, :F], y[:, F:]
    return re * re + im * im


def kaldi_like_bank(n_fft: int, n_mels: int, sr: int, low: float = 20.0, high: float = 0.0) -> np.ndarray:
    if high <= 0:
        high = sr / 2.0 + high
    mel = lambda h: h if h < 1000.0 else 1000.0 + 1000.0 * math.log(h / 1000.0) / math.log(2.0)
    inv = lambda m: m if m < 1000.0 else 1000.0 * math.exp((m - 1000.0) * math.log(2.0) / 1000.0)
    nb = n_fft // 2 + 1
    cm = torch.linspace(mel(low), mel(high), n_mels + 2)
    hz = np.array([inv(float(v)) for v in cm], dtype=f32)
    fr = torch.linspace(0, sr / 2.0, nb).numpy()
    fb = np.zeros((n_mels, nb), f32)
    for i in range(n_mels):
        lo, ce, up = hz[i], hz[i + 1], hz[i + 2]
        for j in range(nb):
            q = fr[j]
            if lo <= q <= ce and ce > lo:
                fb[i, j] = (q - lo) / (ce - lo)
            elif ce < q <= up and up > ce:
                fb[i, j] = (up - q) / (up - ce)
    return fb


def torchaudio_bank(n_freqs, f_min, f_max, n_mels, sr, norm, scale) -> np.ndarray:
    import torchaudio
    return torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sr, norm, scale) \
        .transpose(0, 1).contiguous().numpy()
