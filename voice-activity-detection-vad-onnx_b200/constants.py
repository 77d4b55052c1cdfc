"""Host-side constant tables for the frontend kernels (windows, DFT bases, mel banks).

These are computed ONCE on the host in float32 with torch-CPU, in the same
operation order as the reference builds its conv kernels, because the
reference's bases are themselves fp32-rounded (the phase 2*pi*f*t/n_fft reaches
~1.2e3 rad, where one fp32 ulp is ~1e-4 rad): to match its frame
probabilities we have to contract against the same numbers, not against an
ideal DFT.  The tables are uploaded to the device by `vadx_set_tensor`.

Reference anchors:
  windows / padded window .... FSMN/STFT_Process.py:36-58, FireRedVAD/STFT_Process.py:89-114,
                               NVIDIA_*/STFT_Process.py:94, DFSMN/*/STFT_Process.py:37-60
  DFT conv kernels ........... FSMN/STFT_Process.py:87-98 ("v1": 2*pi*f*t/n_fft),
                               FireRedVAD/STFT_Process.py:203-216 ("v2": (2*pi/n_fft)*f*t)
  Kaldi-style mel bank ....... FireRedVAD/Export_FireRedVAD.py:122-169
  HTK / slaney mel banks ..... torchaudio.functional.melscale_fbanks as called at
                               FSMN/Export_FSMN_VAD.py:63 and NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:186-189
"""
from __future__ import annotations

import math

import torch

_F32 = torch.float32


def make_window(kind: str, win_length: int) -> torch.Tensor:
    """Analysis window of length `win_length` (fp32), by reference window name."""
    if kind in ("hamming", "dfsmn_bartlett"):
        # DFSMN's table maps 'bartlett' to a periodic hamming window (DFSMN/*/STFT_Process.py:38)
        return torch.hamming_window(win_length, periodic=True, dtype=_F32)
    if kind == "hann":
        return torch.hann_window(win_length, periodic=True, dtype=_F32)
    if kind == "hann_sym":
        return torch.hann_window(win_length, periodic=False, dtype=_F32)
    if kind == "hann_sqrt":
        return torch.hann_window(win_length, periodic=False, dtype=_F32).pow(0.5)
    if kind == "povey":
        return torch.hann_window(win_length, periodic=False, dtype=_F32).pow(0.85)
    if kind == "blackman":
        return torch.blackman_window(win_length, periodic=True, dtype=_F32)
    if kind == "kaiser":
        return torch.kaiser_window(win_length, periodic=True, beta=12.0, dtype=_F32)
    raise ValueError(f"unknown window kind {kind!r}")


def window_support(win_length: int, n_fft: int) -> tuple[int, int]:
    """(first_tap, n_taps): where the centred window sits inside the n_fft frame."""
    if win_length >= n_fft:
        return 0, n_fft
    return (n_fft - win_length) // 2, win_length


def dft_basis(n_fft: int, win_length: int, window: str, flavour: str) -> tuple[torch.Tensor, int]:
    """Windowed real-DFT basis restricted to the window support.

    Returns (basis [2*F, n_taps] fp32, first_tap) with rows 0..F-1 = w*cos and rows
    F..2F-1 = -w*sin, F = n_fft//2+1.  Taps outside the window support are
    identically zero in the reference's conv kernel, so they are dropped here
    (the device kernel never multiplies by them): K shrinks from 512 to 400
    for the FSMN/MarbleNet frontends.
    """
    f_bins = n_fft // 2 + 1
    win = make_window(window, win_length)
    first, n_taps = window_support(win_length, n_fft)
    if win_length > n_fft:
        start = (win_length - n_fft) // 2
        win = win[start:start + n_fft]
    t = torch.arange(n_fft, dtype=_F32).unsqueeze(0)
    f = torch.arange(f_bins, dtype=_F32).unsqueeze(1)
    if flavour == "v1":
        omega = 2 * torch.pi * f * t / n_fft
    elif flavour == "v2":
        omega = (2.0 * torch.pi / n_fft) * f * t
    else:
        raise ValueError(flavour)
    omega = omega[:, first:first + n_taps]
    w = win.unsqueeze(0)
    cos_k = torch.cos(omega) * w
    sin_k = -torch.sin(omega) * w
    return torch.cat([cos_k, sin_k], dim=0).contiguous(), first


def kaldi_like_mel_bank(n_fft: int, n_mels: int, sample_rate: int,
                        low_freq: float = 20.0, high_freq: float = 0.0) -> torch.Tensor:
    """[n_mels, F] triangular bank on the piecewise (linear < 1 kHz, log2 above) scale
    used by FireRedVAD/Export_FireRedVAD.py:122-169 (triangles evaluated in Hz)."""
    if high_freq <= 0:
        high_freq = sample_rate / 2.0 + high_freq

    def to_mel(hz: float) -> float:
        return hz if hz < 1000.0 else 1000.0 + 1000.0 * math.log(hz / 1000.0) / math.log(2.0)

    def to_hz(m: float) -> float:
        return m if m < 1000.0 else 1000.0 * math.exp((m - 1000.0) * math.log(2.0) / 1000.0)

    n_bins = n_fft // 2 + 1
    edges_mel = torch.linspace(to_mel(low_freq), to_mel(high_freq), n_mels + 2)
    edges_hz = torch.tensor([to_hz(v.item()) for v in edges_mel], dtype=_F32)
    bin_hz = torch.linspace(0, sample_rate / 2.0, n_bins)
    lo = edges_hz[:-2].unsqueeze(1)
    mid = edges_hz[1:-1].unsqueeze(1)
    hi = edges_hz[2:].unsqueeze(1)
    fq = bin_hz.unsqueeze(0)
    rising = (fq - lo) / (mid - lo)
    falling = (hi - fq) / (hi - mid)
    bank = torch.zeros(n_mels, n_bins, dtype=_F32)
    up_mask = (lo <= fq) & (fq <= mid) & (mid > lo)
    dn_mask = (mid < fq) & (fq <= hi) & (hi > mid)
    bank = torch.where(up_mask, rising, bank)
    bank = torch.where(dn_mask & ~up_mask, falling, bank)
    return bank.contiguous()


def _hz_to_mel(freq: float, scale: str) -> float:
    if scale == "htk":
        return 2595.0 * math.log10(1.0 + freq / 700.0)
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    if freq >= min_log_hz:
        mels = min_log_hz / f_sp + math.log(freq / min_log_hz) / (math.log(6.4) / 27.0)
    return mels


def _mel_to_hz(mels: torch.Tensor, scale: str) -> torch.Tensor:
    if scale == "htk":
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_mel = 1000.0 / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = 1000.0 * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def torchaudio_mel_bank(n_freqs: int, f_min: float, f_max: float, n_mels: int,
                        sample_rate: int, norm: str | None, scale: str) -> torch.Tensor:
    """[n_mels, n_freqs] bank with the arithmetic of torchaudio.functional.melscale_fbanks
    (what the FSMN / DFSMN / MarbleNet wrappers call); re-derived here so the product has
    no torchaudio dependency.  tests/test_constants.py checks bit-equality against torchaudio."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel(f_min, scale)
    m_max = _hz_to_mel(f_max, scale)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = _mel_to_hz(m_pts, scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    if norm == "slaney":
        enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
        fb = fb * enorm.unsqueeze(0)
    return fb.transpose(0, 1).contiguous()


def ceps_bases(n_fft: int):
    """CepsUnit tables (DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py:104-131): rectangular-window DFT
    kernels [n_fft//2+1, n_fft] (cos, -sin) and the pinv inverse basis [2*(n_fft//2+1), n_fft]."""
    half = n_fft // 2
    t = torch.arange(n_fft, dtype=_F32).unsqueeze(0)
    f = torch.arange(half + 1, dtype=_F32).unsqueeze(1)
    omega = 2 * torch.pi * f * t / n_fft
    fb = torch.fft.fft(torch.eye(n_fft, dtype=_F32))
    fb_ri = torch.vstack([torch.real(fb[:half + 1]), torch.imag(fb[:half + 1])]).float()
    return torch.cos(omega), -torch.sin(omega), torch.linalg.pinv(fb_ri).T.contiguous()


def istft_tables(n_fft: int, hop: int, max_frames: int):
    """NET tables (:186-209): windowed pinv synthesis basis [2*(n_fft//2+1), n_fft] and the inverse
    overlap-add window sum."""
    half = n_fft // 2
    window = torch.hamming_window(n_fft)
    fb = torch.fft.fft(torch.eye(n_fft, dtype=_F32))
    fb_ri = torch.vstack([torch.real(fb[:half + 1]), torch.imag(fb[:half + 1])]).float()
    inv = (torch.linalg.pinv((fb_ri * n_fft) / hop).T * window.view(1, -1)).contiguous()
    out_len = (max_frames - 1) * hop + n_fft
    wsum = torch.zeros(out_len, dtype=_F32)
    wsq = window ** 2
    for i in range(max_frames):
        s = i * hop
        n = min(n_fft, out_len - s)
        if n <= 0:
            break
        wsum[s:s + n] += wsq[:n]
    return inv, n_fft / (wsum * hop + 1e-6)
