"""DFSMN AEC-VAD on the GPU through the C ABI: building blocks against torch, the echo estimator
against the golden output of the reference's own NET, the whole graph against the golden output of
the reference's DFSMN_VAD wrapper and against the oracle."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import dfsmn_aec, lib, postprocess as PP, synth, weights as W
from oracle import postproc as OP
from oracle.dfsmn_aec import DfsmnAecOracle

pytestmark = pytest.mark.gpu
TOL = 1e-3
BOUND = 5e-5  # what this path is ASSERTED to: measured 1.5e-5; TOL stays the contract and the decision-margin test


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "dfsmn_aec.npz"))


@pytest.fixture(scope="module")
def wts():
    return W.dfsmn_aec_random_init(W.DfsmnAecConfig(), 0)


def test_layernorm_permute_lstm(cuda):
    l = lib.load()
    g = torch.Generator().manual_seed(1)
    x = torch.randn((37, 3200), generator=g) * 2 + 0.5
    w, b = torch.rand((3200,), generator=g) + 0.5, torch.randn((3200,), generator=g) * 0.1
    out = torch.empty((37, 3200), device=cuda)
    d = [t.to(cuda) for t in (x, w, b)]
    lib.check(l.vadx_layernorm_f32(d[0].data_ptr(), 37, 3200, d[1].data_ptr(), d[2].data_ptr(), 1e-6, out.data_ptr(),
                                   lib.stream_ptr()))
    ref = (x - x.mean(1, keepdim=True)) / (x.std(1, keepdim=True) + 1e-6) * w + b
    assert (out.cpu() - ref).abs().max().item() <= 2e-5
    # permute
    a = torch.randn((5, 7, 2, 9), generator=g)
    for perm in ((0, 3, 2, 1), (0, 2, 1, 3), (3, 1, 0, 2)):
        o = torch.empty(a.numel(), device=cuda)
        ad = a.to(cuda)
        lib.check(l.vadx_permute4_f32(ad.data_ptr(), o.data_ptr(), 5, 7, 2, 9, *perm, lib.stream_ptr()))
        assert torch.equal(o.cpu(), a.permute(*perm).contiguous().reshape(-1))
    # LSTM sequences (forward and reverse, two-level sequence indexing)
    for n_in, H in ((4, 20), (40, 20), (20, 40), (40, 40)):
        m = torch.nn.LSTM(n_in, H, batch_first=True, bidirectional=True)
        S, T, Fb = 2, 11, 6
        xs = torch.randn((S, T, Fb, n_in), generator=g)
        with torch.inference_mode():
            ref = m(xs.permute(0, 2, 1, 3).reshape(S * Fb, T, n_in))[0].reshape(S, Fb, T, 2 * H).permute(0, 2, 1, 3)
        y = torch.zeros((S, T, Fb, 2 * H), device=cuda)
        xd = xs.to(cuda)
        sd = {k: v.detach().to(cuda).contiguous() for k, v in m.state_dict().items()}
        for rev, suf, off in ((0, "", 0), (1, "_reverse", H)):
            lib.check(l.vadx_lstm_seq_f32(xd.data_ptr(), T * Fb * n_in, n_in, Fb * n_in, y.data_ptr() + 4 * off,
                                          T * Fb * 2 * H, 2 * H, Fb * 2 * H, sd["weight_ih_l0" + suf].data_ptr(),
                                          sd["weight_hh_l0" + suf].data_ptr(), sd["bias_ih_l0" + suf].data_ptr(),
                                          sd["bias_hh_l0" + suf].data_ptr(), S * Fb, Fb, T, n_in, H, rev, lib.stream_ptr()))
        assert (y.cpu() - ref).abs().max().item() <= 2e-5, (n_in, H)


def test_echo_estimator_matches_reference_net(cuda, gold, wts):
    cfg = W.DfsmnAecConfig()
    x = torch.from_numpy(gold["iccrn_in"])                     # [1,4,160,24]
    T = x.shape[-1]
    sess = vadx.DfsmnAecSession(wts, cfg, chunk_len=(T - 1) * 160 + 1)
    x4 = x[0].permute(2, 1, 0).contiguous().reshape(-1, 4).to(cuda)      # [T][F][4]
    dummy = torch.zeros((1, sess.chunk_len), dtype=torch.int16, device=cuda)
    aec = sess.echo_estimate(dummy, dummy, x4_override=x4)
    ref = gold["iccrn_out"]
    err = np.abs(aec[0].cpu().numpy() - ref).max()
    print(f"ICCRN echo estimate: max abs err {err:.2e} (signal max {np.abs(ref).max():.3f})")
    assert err <= 1e-4 * max(1.0, np.abs(ref).max())


def test_whole_graph_against_reference_wrapper(cuda, gold, wts, measured):
    cfg = W.DfsmnAecConfig()
    sess = vadx.DfsmnAecSession(wts, cfg, chunk_len=31841)
    near, far = torch.from_numpy(gold["near"]).to(cuda), torch.from_numpy(gold["far"]).to(cuda)
    p = sess.run_batch(near, far).cpu().numpy()
    assert p.shape == (2, 100)
    for s in range(2):
        err = np.abs(p[s] - gold[f"probs{s}"]).max()
        measured("dfsmn_aec: stream", err, BOUND)
    # ORT-shaped single-stream call
    out = sess.run(["vad_results"], {"near_end_audio": gold["near"][:1, None, :], "far_end_audio": gold["far"][:1, None, :]})[0]
    assert out.shape == (100,) and np.abs(out - gold["probs0"]).max() <= BOUND
    with pytest.raises(ValueError):
        sess.run(None, {"near_end_audio": gold["near"][:1, None, :1000], "far_end_audio": gold["far"][:1, None, :1000]})


def test_stream_loop_and_hysteresis(cuda, wts):
    """Overlapping windows + probability-mode look-ahead hysteresis vs the oracle on the same windows."""
    cfg = W.DfsmnAecConfig()
    L = 31841
    sess = vadx.DfsmnAecSession(wts, cfg, chunk_len=L)
    orc = DfsmnAecOracle(wts, cfg)
    n = L + 26721 - 500
    far = synth.synth_streams(1, n, seed=51)[0]
    near = np.clip(synth.synth_streams(1, n, seed=52)[0].astype(np.float32) + 0.4 * np.roll(far, 300).astype(np.float32),
                   -32768, 32767).astype(np.int16)
    r = dfsmn_aec.run_vad(near, far, sess, rng=np.random.RandomState(3), keep_trace=True)
    from vadx import audio_io
    n16, f16 = audio_io.normalize_to_int16(near.astype(np.float32)), audio_io.normalize_to_int16(far.astype(np.float32))
    na, stride, _ = audio_io.align_overlapping(n16, L, 15, 320, np.random.RandomState(3))
    rs = np.random.RandomState(3)
    rs.normal(size=(len(na) - len(n16),))      # run_vad draws the near-end pad first, then the far-end pad
    fa, _, _ = audio_io.align_overlapping(f16, L, 15, 320, rs)
    assert len(r.probs) == (len(na) - L) // stride + 1 == 2
    chunks = []
    for wdx in range(2):
        ref = orc.forward(na[wdx * stride:wdx * stride + L], fa[wdx * stride:wdx * stride + L]).numpy()
        assert np.abs(r.probs[wdx] - ref).max() <= BOUND
        chunks.append(r.probs[wdx])
    # hysteresis in probability mode: oracle machine on the SAME (device) probabilities
    flags = []
    silence, lb = True, 15
    for ci, pr in enumerate(chunks):
        for i in range(len(pr) - lb):
            if silence:
                if pr[i] >= 0.5:
                    votes = 1 + sum(1 for j in range(1, lb) if pr[i + j] >= 0.5)
                    silence = not (votes * (1.0 / lb) >= 0.5)
            else:
                if pr[i] <= 0.5:
                    votes = 1 + sum(1 for j in range(1, lb) if pr[i + j] <= 0.5)
                    silence = not (votes * (1.0 / lb) <= 0.5)
                else:
                    silence = False
            flags.append(silence)
    for i in range(len(chunks[-1]) - lb, len(chunks[-1])):
        pr = chunks[-1]
        silence = (not pr[i] >= 0.5) if silence else (pr[i] <= 0.5)
        flags.append(silence)
    assert np.array_equal(r.saved, np.array(flags))


def test_near_end_only_variant(cuda, golden_dir, wts, measured):
    """DFSMN/only_near_end_audio: one input, the far end replaced by the graph's constant noise buffers."""
    g = np.load(os.path.join(golden_dir, "dfsmn_near.npz"))
    cfg = W.DfsmnAecConfig()
    sess = vadx.DfsmnAecSession(wts, cfg, chunk_len=31841, far_noise=W.dfsmn_near_noise(cfg, seed=77))
    assert [i.name for i in sess.get_inputs()] == ["audio"]
    near = torch.from_numpy(g["near"]).to(cuda)
    p = sess.run_batch(near).cpu().numpy()
    assert p.shape == (2, 100)
    for s in range(2):
        err = np.abs(p[s] - g[f"probs{s}"]).max()
        measured("dfsmn_aec: near-only stream", err, BOUND)
    out = sess.run(["vad_results"], {"audio": g["near"][:1, None, :]})[0]
    assert out.shape == (100,) and np.abs(out - g["probs0"]).max() <= BOUND
    with pytest.raises(ValueError):
        sess.run(None, {"near_end_audio": g["near"][:1, None, :], "far_end_audio": g["near"][:1, None, :]})
    with pytest.raises(ValueError):
        sess.run_batch(near, near)
    with pytest.raises(ValueError):
        vadx.DfsmnAecSession(wts, cfg, chunk_len=31841, far_noise=(np.zeros((3, 3, 3)), np.zeros((2, 3, 3))))
    # the whole single-recording script loop (overlapping windows, hysteresis, timestamps) runs
    r = dfsmn_aec.run_vad(g["near"][0], None, sess, rng=np.random.RandomState(1))
    assert len(r.saved) > 0


def test_cuda_graph_replay_is_bit_identical(cuda, gold, wts):
    cfg = W.DfsmnAecConfig()
    sess = vadx.DfsmnAecSession(wts, cfg, chunk_len=31841)
    near, far = torch.from_numpy(gold["near"]).to(cuda), torch.from_numpy(gold["far"]).to(cuda)
    a = sess.run_batch(near, far).clone()
    b = sess.run_batch_graph(near, far).clone()
    c = sess.run_batch_graph(far, near).clone()           # second replay, different data
    d = sess.run_batch(far, near)
    assert torch.equal(a, b) and torch.equal(c, d)
