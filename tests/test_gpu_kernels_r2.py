"""Unit parity of the round-2 kernels through the C ABI, each against a plain PyTorch fp32/fp64 restatement of the same
operation (the model-level tests cover them end to end; these pin the entry points' own contracts: layouts, strides, edges)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from vadx import lib

pytestmark = pytest.mark.gpu


_KEEP = []


def _ptr(t):
    """data pointer of a tensor that stays alive until the end of the test session (a temporary `x.to(cuda)` passed inline
    would be released -- and its block reused by the next temporary -- before the call it was made for has even been enqueued)"""
    _KEEP.append(t)
    return t.data_ptr()


def test_layernorm_perm_all_orders_and_padding(cuda):
    """vadx_layernorm_perm_f32: (x - mean) / (unbiased std + eps) * w + b over a [n1][n2][n3] row, written in any axis order,
    tables in input or output order, optional zero padding before / after each output row
    (DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py:163-167)."""
    l = lib.load()
    g = torch.Generator().manual_seed(3)
    rows, n1, n2, n3 = 7, 5, 2, 9
    x = torch.randn((rows, n1, n2, n3), generator=g) * 1.7 + 0.3
    w, b = torch.rand((n1, n2, n3), generator=g) + 0.5, torch.randn((n1, n2, n3), generator=g) * 0.1
    xd, wd, bd = x.to(cuda), w.to(cuda), b.to(cuda)
    flat = x.reshape(rows, -1)
    base = (x - flat.mean(1).view(-1, 1, 1, 1)) / (flat.std(1).view(-1, 1, 1, 1) + 1e-6)
    for q in ((0, 1, 2), (1, 0, 2), (2, 1, 0), (2, 0, 1)):
        for w_out in (0, 1):
            if w_out:
                wq, bq = w.permute(*q).contiguous().to(cuda), b.permute(*q).contiguous().to(cuda)
                ref = base.permute(0, *[a + 1 for a in q]) * w.permute(*q) + b.permute(*q)
            else:
                wq, bq = wd, bd
                ref = (base * w + b).permute(0, *[a + 1 for a in q])
            out = torch.full((rows, n1 * n2 * n3), float("nan"), device=cuda)
            lib.check(l.vadx_layernorm_perm_f32(_ptr(xd), rows, n1, n2, n3, *q, _ptr(wq), _ptr(bq), w_out, 1e-6, _ptr(out), 0, 0,
                                                lib.stream_ptr()))
            assert (out.cpu() - ref.reshape(rows, -1)).abs().max().item() <= 2e-5, (q, w_out)
    # padded rows: `pad` zeros before and after, rows out_stride apart
    pad, D = 6, n1 * n2 * n3
    out = torch.full((rows, D + 2 * pad + 3), 7.0, device=cuda)
    lib.check(l.vadx_layernorm_perm_f32(_ptr(xd), rows, n1, n2, n3, 0, 1, 2, _ptr(wd), _ptr(bd), 0, 1e-6, _ptr(out), D + 2 * pad + 3, pad,
                                        lib.stream_ptr()))
    o = out.cpu()
    assert (o[:, pad:pad + D] - (base * w + b).reshape(rows, -1)).abs().max().item() <= 2e-5
    assert o[:, :pad].abs().max().item() == 0 and o[:, pad + D:pad + D + pad].abs().max().item() == 0
    assert (o[:, D + 2 * pad:] == 7.0).all()


def test_cepstral_gain_and_transposed_add(cuda):
    """vadx_ceps_cmul_t_f32 and vadx_add_transposed_f32 against the literal permute + complex product + permute sequence
    (CepsUnit, Export_DFSMN_VAD.py:145-146)."""
    l = lib.load()
    g = torch.Generator().manual_seed(5)
    B, C, cb, Fb = 6, 20, 81, 160
    q = torch.randn((B, cb, 2, C), generator=g)            # cepstral-major: [re(C) | im(C)] per cepstral bin
    spec = torch.randn((B, C, 2, cb), generator=g)         # the cepstral DFT layer's native output
    out = torch.empty_like(spec).to(cuda)
    lib.check(l.vadx_ceps_cmul_t_f32(_ptr(q.to(cuda)), _ptr(spec.to(cuda)), _ptr(out), B, C, cb, lib.stream_ptr()))
    p = spec.permute(0, 3, 2, 1)                           # [B][cb][2][C]
    re = q[:, :, 0] * p[:, :, 0] - q[:, :, 1] * p[:, :, 1]
    im = q[:, :, 0] * p[:, :, 1] + q[:, :, 1] * p[:, :, 0]
    ref = torch.stack([re, im], 2).permute(0, 3, 2, 1)     # back to [B][C][2][cb]
    assert (out.cpu() - ref).abs().max().item() <= 1e-6
    a = torch.randn((B, Fb + 2, C), generator=g)           # padded block rows (block stride (F + 2) * C)
    t = torch.randn((B, C, Fb), generator=g)
    y = torch.empty((B, Fb, C), device=cuda)
    lib.check(l.vadx_add_transposed_f32(_ptr(a.to(cuda)), (Fb + 2) * C, _ptr(t.to(cuda)), _ptr(y), B, Fb, C, lib.stream_ptr()))
    assert torch.equal(y.cpu(), a[:, :Fb] + t.permute(0, 2, 1))
    lib.check(l.vadx_add_transposed_f32(_ptr(a[:, :Fb].contiguous().to(cuda)), 0, _ptr(t.to(cuda)), _ptr(y), B, Fb, C, lib.stream_ptr()))
    assert torch.equal(y.cpu(), a[:, :Fb] + t.permute(0, 2, 1))


@pytest.mark.parametrize("cin", [20, 40])
def test_gated_block_front_against_the_literal_sequence(cuda, cin):
    """vadx_cfb_front_f32 = LN0 -> sigmoid(W_g . + b_g), W_i x + b_i -> gating -> LN1 (padded rows) and LN2 (transposed), against
    the same steps in float64 (CFB, Export_DFSMN_VAD.py:87-93,163-167)."""
    l = lib.load()
    assert l.vadx_cfb_front_supported(cin, 20, 160) == 1 and l.vadx_cfb_front_supported(24, 20, 160) == 0
    g = torch.Generator().manual_seed(cin)
    B, Fb, C = 9, 160, 20
    x = torch.randn((B, Fb, cin), generator=g) * 1.3

    def tab(n):
        return torch.rand((Fb, n), generator=g) + 0.5, torch.randn((Fb, n), generator=g) * 0.1

    (w0, b0), (w1, b1), (w2, b2) = tab(cin), tab(C), tab(C)
    wg, bg = torch.randn((C, cin), generator=g) / cin ** 0.5, torch.randn((C,), generator=g) * 0.1
    wi, bi = torch.randn((C, cin), generator=g) / cin ** 0.5, torch.randn((C,), generator=g) * 0.1
    d = [t.to(cuda).contiguous() for t in (x, w0, b0, wg, bg, wi, bi, w1, b1, w2, b2)]
    col = torch.full((B, Fb + 2, C), float("nan"), device=cuda)
    z = torch.full((B, C, Fb), float("nan"), device=cuda)
    lib.check(l.vadx_cfb_front_f32(_ptr(d[0]), B, Fb, cin, C, _ptr(d[1]), _ptr(d[2]), _ptr(d[3]), _ptr(d[4]), _ptr(d[5]), _ptr(d[6]),
                                   _ptr(d[7]), _ptr(d[8]), _ptr(d[9]), _ptr(d[10]), 1e-6, _ptr(col), _ptr(z), lib.stream_ptr()))

    def ln(v, w, b):
        f = v.reshape(B, -1)
        return (v - f.mean(1).view(-1, 1, 1)) / (f.std(1).view(-1, 1, 1) + 1e-6) * w + b

    xd = x.double()
    gate = torch.sigmoid(ln(xd, w0.double(), b0.double()) @ wg.double().t() + bg.double())
    xi = xd @ wi.double().t() + bi.double()
    gx = gate * xi
    dd = xi - gx
    ref_col = ln(gx, w1.double(), b1.double())
    ref_z = ln(dd, w2.double(), b2.double()).permute(0, 2, 1)
    c = col.cpu().double()
    assert c[:, 0].abs().max().item() == 0 and c[:, Fb + 1].abs().max().item() == 0
    assert (c[:, 1:Fb + 1] - ref_col).abs().max().item() <= 2e-5
    assert (z.cpu().double() - ref_z).abs().max().item() <= 2e-5


@pytest.mark.parametrize("n_in,H,bi", [(40, 20, True), (20, 40, False), (4, 20, True), (40, 40, False)])
def test_split_lstm_against_torch(cuda, n_in, H, bi):
    """The split LSTM (dense input projection with permuted gate rows + vadx_lstm_recurrence_f32) against torch.nn.LSTM, in
    both sequence layouts the echo estimator uses: rows-consecutive sequences and time sequences over [S][T][F][.]."""
    l = lib.load()
    assert l.vadx_lstm_recurrence_supported(H) == 1 and l.vadx_lstm_recurrence_supported(24) == 0
    g = torch.Generator().manual_seed(n_in + H)
    m = torch.nn.LSTM(n_in, H, batch_first=True, bidirectional=bi)
    nd = 2 if bi else 1
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    S, T, Fb = 3, 13, 37
    x = torch.randn((S, T, Fb, n_in), generator=g)
    R = S * T * Fb
    # permuted, direction-stacked projection and recurrent weights (csrc/model_dfsmn.cu: finalize)
    wp = torch.empty((nd * 4 * H, n_in))
    bp = torch.empty((nd * 4 * H,))
    hp = []
    for dname, d in (("", 0), ("_reverse", 1))[:nd]:
        wih, whh = sd["weight_ih_l0" + dname], sd["weight_hh_l0" + dname]
        b = sd["bias_ih_l0" + dname] + sd["bias_hh_l0" + dname]
        idx = torch.tensor([gg * H + j for j in range(H) for gg in range(4)])
        wp[d * 4 * H:(d + 1) * 4 * H] = wih[idx]
        bp[d * 4 * H:(d + 1) * 4 * H] = b[idx]
        hp.append(whh[idx].contiguous().to(cuda))
    G = nd * 4 * H
    gates = ((x.reshape(R, n_in).double() @ wp.double().t()) + bp.double()).float().to(cuda).contiguous()
    for layout in ("rows", "time"):
        y = torch.zeros((R, nd * H), device=cuda)
        if layout == "rows":      # sequences = (s, t) blocks of Fb consecutive rows (the frequency / cepstral LSTMs)
            n_seq, n_inner, L, ro, ri, rs = S * T, 1, Fb, Fb, 0, 1
            with torch.inference_mode():
                ref = m(x.reshape(S * T, Fb, n_in))[0].reshape(R, nd * H)
        else:                     # sequences = (s, f) over the frames (the time LSTMs)
            n_seq, n_inner, L, ro, ri, rs = S * Fb, Fb, T, T * Fb, 1, Fb
            with torch.inference_mode():
                ref = m(x.permute(0, 2, 1, 3).reshape(S * Fb, T, n_in))[0].reshape(S, Fb, T, nd * H).permute(0, 2, 1, 3).reshape(R, nd * H)
        for d in range(nd):
            lib.check(l.vadx_lstm_recurrence_f32(gates.data_ptr() + 4 * d * 4 * H, ro * G, ri * G, rs * G, y.data_ptr() + 4 * d * H,
                                                 ro * nd * H, ri * nd * H, rs * nd * H, _ptr(hp[d]), n_seq, n_inner, L, H, d,
                                                 lib.stream_ptr()))
        err = (y.cpu() - ref).abs().max().item()
        assert err <= 2e-5, (layout, err)


def test_silero_recurrence_in_one_launch(cuda):
    """vadx_silero_lstm_windows_f32 against a torch LSTMCell loop + relu + 1-output sigmoid head over W windows with state carry
    (Silero/modeling_modified/utils_vad.py:114-123 is the per-window contract)."""
    l = lib.load()
    g = torch.Generator().manual_seed(8)
    S, W, H = 61, 9, 128                      # 61 streams: three CTAs, the last one ragged
    cell = torch.nn.LSTMCell(H, H)
    w_hh = cell.weight_hh.detach()
    gin = torch.randn((W, S, 4 * H), generator=g) * 0.5
    st = torch.randn((2, S, H), generator=g) * 0.3
    hw, hb = torch.randn((H,), generator=g) / H ** 0.5, 0.1
    wt = w_hh.t().contiguous().to(cuda)       # [H][4H]
    out_state = torch.empty((2, S, H), device=cuda)
    probs = torch.empty((W, S), device=cuda)
    lib.check(l.vadx_silero_lstm_windows_f32(_ptr(gin.to(cuda)), _ptr(wt), 4 * H, _ptr(st.to(cuda)), _ptr(out_state), _ptr(hw.to(cuda)),
                                             hb, _ptr(probs), S, W, H, lib.stream_ptr()))
    h, c = st[0].double(), st[1].double()
    ref = []
    for w in range(W):
        gt = gin[w].double() + h @ w_hh.double().t()
        i, f, gg, o = gt.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        ref.append(torch.sigmoid(torch.relu(h) @ hw.double() + hb))
    assert (probs.cpu().double() - torch.stack(ref)).abs().max().item() <= 2e-6
    assert (out_state.cpu().double() - torch.stack([h, c])).abs().max().item() <= 5e-6


def test_window_helpers(cuda):
    """vadx_gather_windows_i16, vadx_reflect_windows_f32, vadx_stft_mag_compact_f32, vadx_affine_f32 against indexing in torch."""
    l = lib.load()
    g = torch.Generator().manual_seed(2)
    # gather: vector path (all offsets multiples of 8) and scalar path
    for S, W, stride, L in ((5, 4, 11040, 16000), (3, 7, 353, 513)):
        n = L + (W - 1) * stride + (8 if stride % 8 == 0 else 3)
        a = torch.randint(-30000, 30000, (S, n), generator=g, dtype=torch.int16)
        out = torch.zeros((S * W, L), dtype=torch.int16, device=cuda)
        lib.check(l.vadx_gather_windows_i16(_ptr(a.to(cuda)), n, S, W, stride, L, _ptr(out), lib.stream_ptr()))
        ref = torch.stack([a[s, w * stride:w * stride + L] for s in range(S) for w in range(W)])
        assert torch.equal(out.cpu(), ref)
    # reflect pads of all windows: out[w][s] = x[s][w*step : w*step + n_in] + reflected tail
    S, W, n_in, pad, step = 6, 5, 576, 64, 512
    x = torch.randn((S, (W - 1) * step + n_in), generator=g)
    out = torch.empty((W, S, n_in + pad), device=cuda)
    lib.check(l.vadx_reflect_windows_f32(_ptr(x.to(cuda)), x.shape[1], S, W, step, n_in, pad, _ptr(out), lib.stream_ptr()))
    for w in range(W):
        seg = x[:, w * step:w * step + n_in]
        ref = F.pad(seg[:, None, :], (0, pad), mode="reflect")[:, 0]
        assert torch.equal(out[w].cpu(), ref)
    # magnitudes of (re, im)-interleaved DFT rows with junk rows between windows
    n_win, per, T, Fb, ld = 11, 5, 4, 129, 260
    y = torch.randn(((n_win - 1) * per + T, ld), generator=g)
    mag = torch.empty((n_win, T * Fb), device=cuda)
    lib.check(l.vadx_stft_mag_compact_f32(_ptr(y.to(cuda)), ld, n_win, per, T, Fb, _ptr(mag), lib.stream_ptr()))
    rows = torch.stack([y[r * per + t] for r in range(n_win) for t in range(T)]).reshape(n_win, T, ld)
    ref = torch.sqrt(rows[:, :, 0:2 * Fb:2] ** 2 + rows[:, :, 1:2 * Fb:2] ** 2).reshape(n_win, T * Fb)
    assert (mag.cpu() - ref).abs().max().item() <= 1e-6
    v = torch.randn((1000,), generator=g)
    o = torch.empty((1000,), device=cuda)
    lib.check(l.vadx_affine_f32(_ptr(v.to(cuda)), -1.0, 1.0, _ptr(o), 1000, lib.stream_ptr()))
    assert torch.equal(o.cpu(), 1.0 - v)


def test_dense_layer_over_overlapping_rows(cuda):
    """vadx_linear_tc_f32 with ldx < n_in: the input rows overlap (hop-strided frames of a signal = the framed DFT as a dense
    layer; the 3-tap frequency conv over padded rows) -- against unfold + matmul in float64."""
    l = lib.load()
    g = torch.Generator().manual_seed(4)
    hop, K, N, M = 128, 256, 128, 700
    sig = torch.randn(((M - 1) * hop + K,), generator=g)
    w = torch.randn((N, K), generator=g) / K ** 0.5
    img = torch.from_numpy(lib.pack_weight_tc(w.numpy())).to(cuda)
    y = torch.empty((M, N), device=cuda)
    lib.check(l.vadx_linear_tc_f32(_ptr(sig.to(cuda)), hop, _ptr(img), None, None, 0, _ptr(y), N, M, K, N, 0, lib.stream_ptr()))
    ref = sig.double().unfold(0, K, hop) @ w.double().t()
    assert (y.cpu().double() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_kernel_profile_records(cuda):
    """vadx_profile_collect_kernels: every entry point reports its kernel's name and the algorithmic bytes / flops of the call."""
    l = lib.load()
    x = torch.randn((4096, 128), device=cuda)
    y = torch.empty_like(x)
    lib.profile_enable(True)
    lib.profile_collect_kernels()
    for _ in range(3):
        lib.check(l.vadx_affine_f32(_ptr(x), 2.0, 0.5, _ptr(y), x.numel(), lib.stream_ptr()))
    torch.cuda.synchronize()
    k = lib.profile_collect_kernels()
    lib.profile_enable(False)
    assert k["affine_kernel"]["calls"] == 3 and k["affine_kernel"]["bytes"] == 3 * 8.0 * x.numel() and k["affine_kernel"]["ms"] > 0
    assert lib.profile_collect_kernels() == {}
