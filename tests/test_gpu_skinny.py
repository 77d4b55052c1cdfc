"""The skinny (<= 16 rows) exact-fp32 contraction path against torch fp32, every epilogue it serves."""
import ctypes as C

import numpy as np
import pytest
import torch

import vadx
from vadx import lib

pytestmark = pytest.mark.gpu


def _linear(x, w, b, res, act):
    """through the C ABI: y = act(x @ w.T + b) (+ res)"""
    L = lib.load()
    M, K = x.shape
    N = w.shape[0]
    ldw = (N + 3) // 4 * 4
    wt = torch.zeros((K, ldw), dtype=torch.float32, device=x.device)
    wt[:, :N] = w.t()
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    lib.check(L.vadx_linear_f32(x.data_ptr(), K, wt.data_ptr(), ldw, lib.ptr(b), lib.ptr(res), N, y.data_ptr(), N, M, K, N,
                                act, None))
    return y


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8, 9, 14, 16])
@pytest.mark.parametrize("K,N", [(400, 140), (140, 250), (250, 128), (80, 256), (128, 248), (33, 9)])
def test_skinny_linear_matches_fp32(cuda, M, K, N):
    g = torch.Generator(device="cpu").manual_seed(M * 1000 + K + N)
    x = torch.randn((M, K), generator=g).to(cuda)
    w = (torch.randn((N, K), generator=g) / K ** 0.5).to(cuda)
    b = torch.randn((N,), generator=g).to(cuda)
    res = torch.randn((M, N), generator=g).to(cuda)
    ref = x.double() @ w.double().t() + b.double()
    for act, fn in ((0, lambda v: v), (1, torch.relu), (2, torch.sigmoid)):
        y = _linear(x, w, b, None, act)
        assert (y.double() - fn(ref)).abs().max().item() <= 2e-5
    y = _linear(x, w, b, res, 1)
    assert (y.double() - (torch.relu(ref) + res.double())).abs().max().item() <= 2e-5
    y = _linear(x, w, None, res, 1 | 16)      # residual before the activation
    assert (y.double() - torch.relu(ref - b.double() + res.double())).abs().max().item() <= 2e-5


@pytest.mark.parametrize("S,T", [(1, 1), (1, 3), (2, 7), (1, 14), (3, 5)])
def test_skinny_framed_dft_power(cuda, S, T):
    L = lib.load()
    n_taps, hop, n_bins = 400, 160, 201
    ld_basis, ld_power = 404, 204
    g = torch.Generator(device="cpu").manual_seed(S * 100 + T)
    stride = (T - 1) * hop + n_taps + 8
    sig = torch.randn((S, stride), generator=g).to(cuda)
    basis = torch.zeros((n_taps, ld_basis), dtype=torch.float32)
    basis[:, :2 * n_bins] = torch.randn((n_taps, 2 * n_bins), generator=g) / 20
    basis = basis.to(cuda)
    power = torch.zeros((S * T, ld_power), dtype=torch.float32, device=cuda)
    lib.check(L.vadx_stft_power_f32(sig.data_ptr(), stride, S, T, hop, n_taps, basis.data_ptr(), ld_basis, n_bins,
                                    power.data_ptr(), ld_power, None))
    frames = torch.stack([sig[s, t * hop:t * hop + n_taps] for s in range(S) for t in range(T)]).double()
    z = frames @ basis[:, :2 * n_bins].double()
    ref = z[:, 0::2] ** 2 + z[:, 1::2] ** 2
    err = ((power[:, :n_bins].double() - ref).abs() / (1 + ref)).max().item()
    assert err <= 2e-5
    assert power[:, n_bins:].abs().max().item() == 0.0
