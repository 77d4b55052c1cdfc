"""Per-kernel parity (GPU): each C-ABI entry point against the oracle / a torch-CPU fp32 reference
on the same seeded inputs."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vadx
from vadx import constants, lib, postprocess as PP, synth, tables, weights as W
from oracle import frontend as OF, postproc as OP
from oracle.firered import FireRedOracle

pytestmark = pytest.mark.gpu


def _st():
    return lib.stream_ptr()


def test_prep_audio_modes(cuda):
    l = lib.load()
    x = torch.from_numpy(synth.synth_streams(5, 4001, seed=3))
    xd = x.to(cuda)
    for scale, dc, mode, pad in [(1.0, 0, 1, 0), (1.0 / 32768, 0, 1, 256), (1.0, 1, 2, 256), (1.0 / 32768, 1, 2, 159),
                                 (1.0, 0, 0, 3)]:
        stride = (pad + 4001 + pad + 3) // 4 * 4
        out = torch.full((5, stride), 7.0, device=cuda)
        lib.check(l.vadx_prep_audio(xd.data_ptr(), lib.DT_I16, 5, 4001, 4001, scale, dc, mode, 0.97, pad,
                                    out.data_ptr(), stride, _st()))
        a = x.float() * np.float32(scale)
        if dc:
            a = a - a.mean(dim=1, keepdim=True)
        if mode == 1:
            a = torch.cat([a[:, :1], a[:, 1:] - np.float32(0.97) * a[:, :-1]], 1)
        elif mode == 2:
            a = torch.cat([a[:, :1], a[:, 1:] - np.float32(0.97) * a[:, :-1]], 1)
        ref = F.pad(a, (pad, stride - pad - 4001))
        err = (out.cpu() - ref).abs().max().item()
        assert err <= 2e-3 * max(1.0, scale * 32768) * (1 if scale == 1.0 else 1e-4), (scale, dc, mode, pad, err)
    # fp32 input (Silero-style)
    xf = (x.float() / 32768).to(cuda)
    out = torch.empty((5, 4004), device=cuda)
    lib.check(l.vadx_prep_audio(xf.data_ptr(), lib.DT_F32, 5, 4001, 4001, 1.0, 0, 0, 0.0, 0, out.data_ptr(), 4004, _st()))
    assert torch.equal(out[:, :4001].cpu(), xf.cpu())


@pytest.mark.parametrize("n_fft,win,window,flavour,center", [(400, 400, "povey", "v2", False),
                                                               (512, 400, "hamming", "v1", True),
                                                               (512, 400, "hann_sym", "v2", True)])
def test_stft_power(cuda, n_fft, win, window, flavour, center):
    l = lib.load()
    S, L, hop = 3, 16000, 160
    x = torch.from_numpy(synth.synth_streams(S, L, seed=11)).float()
    basis, first, nb = tables.interleaved_basis(n_fft, win, window, flavour)
    n_taps = basis.shape[0]
    pad = n_fft // 2 if center else 0
    T = (L + 2 * pad - n_fft) // hop + 1
    # frame t starts at sample t*hop - pad + first (window support only)
    pad_left = pad - first if center else 0
    stride = (pad_left + L + pad + 3) // 4 * 4
    assert (T - 1) * hop + n_taps <= stride
    sig = torch.zeros((S, stride))
    sig[:, pad_left:pad_left + L] = x
    sd, bd = sig.to(cuda), torch.from_numpy(basis).to(cuda)
    ldp = (nb + 1) // 2 * 2
    out = torch.zeros((S * T, ldp), device=cuda)
    lib.check(l.vadx_stft_power_f32(sd.data_ptr(), stride, S, T, hop, n_taps, bd.data_ptr(), basis.shape[1], nb,
                                    out.data_ptr(), ldp, _st()))
    ker = OF.stft_kernel(n_fft, win, "hamming" if window == "hamming" else window, flavour)
    ref = OF.stft_power(x.unsqueeze(1), ker, hop, center)          # [S, F, T]
    ref = ref.permute(0, 2, 1).reshape(S * T, nb)
    got = out[:, :nb].cpu()
    scale = ref.max(dim=1, keepdim=True).values
    assert ((got - ref).abs() / scale).max().item() <= 2e-6


def test_mel_log(cuda):
    l = lib.load()
    rows, nb = 1000, 201
    g = torch.Generator().manual_seed(5)
    power = (torch.rand((rows, 202), generator=g) * 1e9).float()
    bank = constants.kaldi_like_mel_bank(400, 80, 16000)
    st, ln, w = tables.sparse_bank(bank.numpy())
    out = torch.zeros((rows, 80), device=cuda)
    args = [t.to(cuda) for t in (power, torch.from_numpy(st), torch.from_numpy(ln), torch.from_numpy(w))]
    for mode, floor in [(0, 1e-7), (1, 1e-7)]:
        lib.check(l.vadx_mel_log_f32(args[0].data_ptr(), 202, rows, nb, 80, args[1].data_ptr(), args[2].data_ptr(),
                                     args[3].data_ptr(), w.shape[1], mode, floor, out.data_ptr(), 80, _st()))
        m = power[:, :nb] @ bank.T
        ref = torch.clamp(m, min=floor).log() if mode == 0 else (m + floor).log()
        assert (out.cpu() - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("rows,n_in,n_out,act,res", [(1000, 80, 256, 1, False), (777, 256, 128, 0, True),
                                                      (98 * 3, 128, 256, 1, False), (500, 400, 140, 0, False),
                                                      (131, 140, 250, 1, False), (64, 250, 248, 2, False),
                                                      (300, 256, 1, 2, False), (300, 96, 3, 2, False)])
def test_linear(cuda, rows, n_in, n_out, act, res):
    l = lib.load()
    g = torch.Generator().manual_seed(rows + n_in)
    x = torch.randn((rows, n_in), generator=g)
    w = torch.randn((n_out, n_in), generator=g) / np.sqrt(n_in)
    b = torch.randn((n_out,), generator=g)
    r = torch.randn((rows, n_out), generator=g) if res else None
    ldw = (n_out + 3) // 4 * 4
    wt = torch.zeros((n_in, ldw))
    wt[:, :n_out] = w.T
    xd, wd, bd = x.to(cuda), wt.to(cuda), b.to(cuda)
    rd = r.to(cuda) if res else None
    y = torch.zeros((rows, n_out), device=cuda)
    lib.check(l.vadx_linear_f32(xd.data_ptr(), n_in, wd.data_ptr(), ldw, bd.data_ptr(), lib.ptr(rd), n_out,
                                y.data_ptr(), n_out, rows, n_in, n_out, act, _st()))
    ref = x.double() @ w.double().T + b.double()
    ref = torch.relu(ref) if act == 1 else torch.sigmoid(ref) if act == 2 else ref
    if res:
        ref = ref + r.double()
    assert (y.cpu().double() - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("T,C,n1,s1,n2,s2,cache", [(98, 128, 20, 1, 20, 1, False), (98, 64, 5, 2, 3, 2, False),
                                                    (14, 128, 20, 1, 0, 1, True), (1, 128, 20, 1, 20, 1, False),
                                                    (200, 128, 20, 1, 0, 1, True), (101, 128, 20, 1, 0, 1, True)])
def test_fsmn_memory(cuda, T, C, n1, s1, n2, s2, cache):
    l = lib.load()
    S = 4
    g = torch.Generator().manual_seed(T * C)
    p = torch.randn((S, T, C), generator=g)
    res = torch.randn((S, T, C), generator=g)
    wl = torch.randn((C, 1, n1), generator=g) * 0.2
    wr = torch.randn((C, 1, max(n2, 1)), generator=g) * 0.2
    halo = (n1 - 1) * s1
    cin = torch.randn((S, C, halo), generator=g) if cache else None
    # reference (channel-first), FireRedVAD/Export_FireRedVAD.py:213-236 / :496-515
    x = p.permute(0, 2, 1)
    xin = torch.cat([cin, x], 2) if cache else F.pad(x, (halo, 0))
    ref = x + F.conv1d(xin, wl, dilation=s1, groups=C)
    if n2 > 0 and T > 1:
        ref = ref + F.conv1d(F.pad(x, (0, n2 * s2)), wr, dilation=s2, groups=C)[:, :, s2:]
    ref = ref.permute(0, 2, 1) + res
    cref = xin[:, :, -halo:] if halo else None
    out = torch.zeros((S, T, C), device=cuda)
    cout = torch.zeros((S, C, halo), device=cuda)
    d = [t.to(cuda) if t is not None else None for t in (p, wl.reshape(C, n1), wr.reshape(C, -1), res, cin)]
    lib.check(l.vadx_fsmn_memory_f32(d[0].data_ptr(), C, d[1].data_ptr(), n1, s1, d[2].data_ptr() if n2 else None, n2, s2,
                                     d[3].data_ptr(), C, out.data_ptr(), C, S, T, C, lib.ptr(d[4]),
                                     cout.data_ptr() if halo else None, _st()))
    assert (out.cpu() - ref).abs().max().item() <= 2e-5
    if halo:
        assert torch.equal(cout.cpu(), cref.contiguous())


def test_postprocess_frames_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "postproc.npz"))
    for i in range(7):
        p = torch.from_numpy(g[f"fr{i}_probs"]).to(cuda).unsqueeze(0)
        ws, thr, msp, mxs, msi, mrg, ext = g[f"fr{i}_params"]
        cfg = PP.FramePostConfig(int(ws), float(thr), int(msp), int(mxs), int(msi), int(mrg), int(ext))
        dec, cnt, seg = PP.postprocess_frames(p, cfg)
        assert np.array_equal(dec[0].cpu().numpy(), g[f"fr{i}_dec"]), f"case {i}"
        k = int(cnt[0])
        ts = PP.segments_to_seconds(seg[0, :k].cpu().numpy(), p.shape[1], cfg, float(g[f"fr{i}_dur"]))
        assert np.array_equal(np.array(ts, np.float64).reshape(-1, 2), g[f"fr{i}_seg"]), f"case {i}"


def test_postprocess_frames_batch_ragged(cuda):
    rs = np.random.RandomState(4)
    S, T = 300, 1500
    lvl = np.clip(0.5 + np.cumsum(rs.normal(0, 0.05, size=(S, T)), 1), 0, 1)
    gate = (np.sin(np.arange(T)[None] / rs.uniform(10, 80, size=(S, 1))) > 0).astype(np.float32)
    p = np.clip(0.1 + 0.8 * gate * lvl + rs.normal(0, 0.05, (S, T)), 0, 1).astype(np.float32)
    n = rs.randint(0, T + 1, size=S).astype(np.int32)
    n[:3] = [0, 1, T]
    cfg = PP.FramePostConfig(5, 0.4, 20, 300, 20, 5, 0)
    dec, cnt, seg = PP.postprocess_frames(torch.from_numpy(p).to(cuda), cfg, torch.from_numpy(n).to(cuda))
    dec, cnt, seg = dec.cpu().numpy(), cnt.cpu().numpy(), seg.cpu().numpy()
    for s in range(S):
        ref = OP.frame_decisions(p[s, :n[s]], 5, 0.4, 20, 300, 20, 5, 0)
        assert np.array_equal(dec[s, :n[s]], ref), s
        padded = np.concatenate(([0], ref, [0])).astype(np.int8)
        d = np.diff(padded)
        pairs = np.stack([np.flatnonzero(d == 1), np.flatnonzero(d == -1)], 1)
        assert cnt[s] == len(pairs) and np.array_equal(seg[s, :cnt[s]], pairs), s


@pytest.mark.parametrize("cfg_t", [
    (5, 0.4, 20, 2000, 20, 5, 0), (3, 0.5, 10, 1000, 10, 3, 0), (5, 0.4, 20, 120, 20, 5, 0), (1, 0.5, 0, 2000, 0, 0, 0),
    (4, 0.45, 5, 60, 7, 4, 3), (1, 0.5, 3, 50, 0, 0, 2), (7, 0.35, 0, 33, 4, 9, 1), (2, 0.6, 1, 2000, 1, 1, 0),
    (5, 0.4, 1, 7, 1, 0, 0), (64, 0.4, 30, 500, 30, 5, 0)])
def test_postprocess_frames_run_based_matches_sequential_oracle(cuda, cfg_t):
    """The run-based state machines (find-first-set over the flag mask) against the frame-by-frame
    oracle on lively tracks, for window / min-speech / min-silence / merge / extend / split settings
    including the degenerate ones; every length from 0 up."""
    rs = np.random.RandomState(hash(cfg_t) % (2 ** 31))
    S, T = 160, 700
    lvl = np.clip(0.5 + np.cumsum(rs.normal(0, 0.07, size=(S, T)), 1), 0, 1)
    gate = (np.sin(np.arange(T)[None] / rs.uniform(3, 60, size=(S, 1)) + rs.uniform(0, 6, size=(S, 1))) > rs.uniform(-0.7, 0.7, size=(S, 1)))
    p = np.clip(0.1 + 0.8 * gate * lvl + rs.normal(0, 0.08, (S, T)), 0, 1).astype(np.float32)
    p[5] = 0.9                                   # one unbroken segment: exercises the split chain
    p[6, ::2] = 0.9; p[6, 1::2] = 0.1            # alternating flags
    n = rs.randint(0, T + 1, size=S).astype(np.int32)
    n[:8] = [0, 1, 2, T, 31, T, T, 33]
    cfg = PP.FramePostConfig(*cfg_t)
    dec, cnt, seg = PP.postprocess_frames(torch.from_numpy(p).to(cuda), cfg, torch.from_numpy(n).to(cuda))
    dec, cnt, seg = dec.cpu().numpy(), cnt.cpu().numpy(), seg.cpu().numpy()
    n_seg = 0
    for s in range(S):
        ref = OP.frame_decisions(p[s, :n[s]], *cfg_t)
        assert np.array_equal(dec[s, :n[s]], ref), (s, int(n[s]))
        d = np.diff(np.concatenate(([0], ref, [0])).astype(np.int8))
        pairs = np.stack([np.flatnonzero(d == 1), np.flatnonzero(d == -1)], 1)
        assert cnt[s] == len(pairs) and np.array_equal(seg[s, :cnt[s]], pairs), s
        n_seg += len(pairs)
    assert n_seg > S // 4


@pytest.mark.parametrize("T,n2,with_res,S", [(98, 20, True, 333), (98, 20, False, 40), (98, 0, True, 149), (57, 20, True, 64),
                                             (2, 20, True, 32), (131, 0, False, 33)])
def test_fsmn_memory_whole_chunk_kernel(cuda, T, n2, with_res, S):
    """The persistent shared-memory kernel (>= 32 streams, 128 channels, no cache): every stream slot of the
    two-deep ring is reused many times (S >> SM count for the first case)."""
    l = lib.load()
    C, n1 = 128, 20
    g = torch.Generator().manual_seed(T * 7 + n2 + S)
    p = torch.randn((S, T, C), generator=g)
    res = torch.randn((S, T, C), generator=g) if with_res else None
    wl = torch.randn((C, 1, n1), generator=g) * 0.2
    wr = torch.randn((C, 1, max(n2, 1)), generator=g) * 0.2
    x = p.permute(0, 2, 1)
    ref = x + F.conv1d(F.pad(x, (n1 - 1, 0)), wl, groups=C)
    if n2 > 0:
        ref = ref + F.conv1d(F.pad(x, (0, n2)), wr, groups=C)[:, :, 1:]
    ref = ref.permute(0, 2, 1)
    if with_res:
        ref = ref + res
    out = torch.full((S, T, C), float("nan"), device=cuda)
    d = [t.to(cuda) if t is not None else None for t in (p, wl.reshape(C, n1), wr.reshape(C, -1), res)]
    lib.check(l.vadx_fsmn_memory_f32(d[0].data_ptr(), C, d[1].data_ptr(), n1, 1, d[2].data_ptr() if n2 else None, n2, 1,
                                     lib.ptr(d[3]), C, out.data_ptr(), C, S, T, C, None, None, _st()))
    assert (out.cpu() - ref).abs().max().item() <= 2e-5
