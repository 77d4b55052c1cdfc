"""Device-layout versions of the frontend constant tables (host-built, uploaded once).

`constants.py` builds the tables with the reference's own arithmetic; this module only
re-packs them for the kernels:
  * DFT basis  -> [n_taps][ld] fp32, (re, im) interleaved along the row, ld = 2F rounded up to 4
  * mel bank   -> sparse triangles: start[m], len[m], w[m][max_len]
"""
from __future__ import annotations

import numpy as np

from . import constants


def interleaved_basis(n_fft: int, win_length: int, window: str, flavour: str):
    """-> (basis [n_taps, ld] fp32, first_tap, n_bins)."""
    b, first = constants.dft_basis(n_fft, win_length, window, flavour)
    b = b.numpy()
    f_bins = n_fft // 2 + 1
    n_taps = b.shape[1]
    ld = (2 * f_bins + 3) // 4 * 4
    out = np.zeros((n_taps, ld), np.float32)
    out[:, 0:2 * f_bins:2] = b[:f_bins].T
    out[:, 1:2 * f_bins:2] = b[f_bins:].T
    return out, first, f_bins


def sparse_bank(bank: np.ndarray):
    """dense [n_mels, F] -> (start int32[n_mels], len int32[n_mels], w fp32[n_mels, max_len])."""
    bank = np.asarray(bank, np.float32)
    n_mels = bank.shape[0]
    start = np.zeros(n_mels, np.int32)
    length = np.ones(n_mels, np.int32)
    for m in range(n_mels):
        nz = np.flatnonzero(bank[m])
        if nz.size:
            start[m] = nz[0]
            length[m] = nz[-1] - nz[0] + 1
    max_len = int(length.max())
    w = np.zeros((n_mels, max_len), np.float32)
    for m in range(n_mels):
        w[m, :length[m]] = bank[m, start[m]:start[m] + length[m]]
    return start, length, w
