"""Synthetic 16 kHz int16 audio for benchmarks and parity tests (no datasets offline).

Base signal: int16 uniform in [-8000, 8000) -- the range the reference's own export
validation draws from (FireRedVAD/Export_FireRedVAD.py:1539-1543) -- multiplied by an
on/off burst envelope (bursts and gaps of 0.3-2 s) so energy gates and the post-processing
state machines switch.  numpy's legacy RandomState keeps it bit-reproducible everywhere.
"""
from __future__ import annotations

import numpy as np


def burst_envelope(rs: np.random.RandomState, n_samples: int, sr: int = 16000,
                   quiet_gain: float = 0.02) -> np.ndarray:
    env = np.empty(n_samples, np.float32)
    pos, on = 0, bool(rs.randint(0, 2))
    while pos < n_samples:
        dur = int(rs.uniform(0.3, 2.0) * sr)
        env[pos:pos + dur] = 1.0 if on else quiet_gain
        pos += dur
        on = not on
    return env


def synth_streams(n_streams: int, n_samples: int, seed: int = 1234, enveloped: bool = True) -> np.ndarray:
    """[n_streams, n_samples] int16."""
    rs = np.random.RandomState(seed)
    x = rs.randint(-8000, 8000, size=(n_streams, n_samples)).astype(np.float32)
    if enveloped:
        for s in range(n_streams):
            x[s] *= burst_envelope(rs, n_samples)
    return x.astype(np.int16)


def synth_chunks_fast(n_chunks: int, chunk_len: int, seed: int = 1234) -> np.ndarray:
    """Large benchmark batches: a 64-stream enveloped pool tiled with per-row circular shifts
    (generating 1 M distinct seconds with RandomState would take minutes; the arithmetic the
    kernels do is identical)."""
    pool = synth_streams(64, chunk_len, seed)
    out = np.empty((n_chunks, chunk_len), np.int16)
    for i in range(n_chunks):
        out[i] = np.roll(pool[i % 64], (i // 64) * 37 % chunk_len)
    return out
