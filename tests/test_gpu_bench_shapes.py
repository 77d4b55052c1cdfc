"""Parity at the BENCH shapes of BASELINE configs[1], [2] and [4] (the sizes bench.py's `families` runs time; FireRed's
configs[3] twin lives in test_gpu_firered.py::test_full_step_is_batch_independent).

The oracles cannot run these sizes in seconds, so each test uses the size-independent property of the path -- a stream's
probabilities do not depend on what else is in the batch: slices of the big call must equal the same streams run in a
small call BIT FOR BIT -- and checks a sample of streams of the big call against the oracle."""
import numpy as np
import pytest
import torch

import vadx
from vadx import dfsmn_aec, marblenet_vad, synth, weights as W

pytestmark = pytest.mark.gpu


def test_silero_4096_streams_x_32_windows(cuda, measured):
    """configs[1]: 4096 parallel streams, 32 ms windows with LSTM state carry (Silero/modeling_modified/utils_vad.py:87-146)."""
    from oracle.silero import OnnxWrapperOracle, SileroNetOracle
    cfg = W.SileroConfig()
    wts = W.silero_random_init(cfg, 0)
    sess = vadx.SileroSession(wts, cfg)
    S, n_win = 4096, 32
    host = synth.synth_chunks_fast(S, n_win * 512, seed=13)
    audio = torch.from_numpy(host).to(cuda).float() * 0.000030517578
    big = sess.speech_probs(audio)
    assert big.shape == (S, n_win) and torch.isfinite(big).all()
    for lo in (0, 2016, S - 64):
        small = sess.speech_probs(audio[lo:lo + 64].contiguous())
        assert torch.equal(big[lo:lo + 64], small), f"streams {lo}..{lo + 63} depend on the batch"
    idx = [0, 1, 2047, 2048, S - 1]
    ref = OnnxWrapperOracle(SileroNetOracle(wts, cfg)).audio_forward(torch.from_numpy(host[idx]).float() * 0.000030517578)
    measured(f"silero bench shape {S} x {n_win}: sample of {len(idx)} streams vs oracle",
             (big[idx].cpu() - ref).abs().max().item(), 5e-5)


@pytest.mark.parametrize("B", [64, 256])
def test_marblenet_60s_clips(cuda, measured, B):
    """configs[2]: a batch of 60 s clips (NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:222-275); 64 is bench.py's batch, 256 the
    survey's per-GPU shape."""
    from oracle.marblenet import MarbleNetOracle
    cfg = W.MarbleNetConfig()
    wts = W.marblenet_random_init(cfg, 0)
    sess = vadx.MarbleNetSession(wts, cfg)
    host = synth.synth_chunks_fast(B, 960000, seed=12)
    clips = torch.from_numpy(host).to(cuda)
    probs, dec, cnt, seg = marblenet_vad.run_vad_clips(sess, clips)
    assert probs.shape == (B, 3000) and torch.isfinite(probs).all()
    for lo in (0, B // 2 - 1, B - 2):
        p2, d2, c2, s2 = marblenet_vad.run_vad_clips(sess, clips[lo:lo + 2].contiguous())
        assert torch.equal(probs[lo:lo + 2], p2), f"clips {lo}, {lo + 1} depend on the batch"
        assert torch.equal(dec[lo:lo + 2], d2) and torch.equal(cnt[lo:lo + 2], c2)
    idx = [0, B - 1]
    _, act, n = MarbleNetOracle(wts, cfg).forward(host[idx])
    assert n == 3000
    measured(f"marblenet bench shape {B} x 60 s: sample of {len(idx)} clips vs oracle",
             np.abs(probs[idx].cpu().numpy() - act.numpy()[:, :3000, 0]).max(), 1e-4)


@pytest.mark.parametrize("S", [32, 256])
def test_dfsmn_aec_stream_pairs(cuda, measured, S):
    """configs[4]: near-end + far-end pairs, one 31841-sample chunk each (DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py:
    317-354)."""
    from oracle.dfsmn_aec import DfsmnAecOracle
    cfg = W.DfsmnAecConfig()
    wts = W.dfsmn_aec_random_init(cfg, 0)
    sess = vadx.DfsmnAecSession(wts, cfg, chunk_len=31841)
    far_h = synth.synth_chunks_fast(S, 31841, seed=14)
    near_h = synth.synth_chunks_fast(S, 31841, seed=15)
    far, near = torch.from_numpy(far_h).to(cuda), torch.from_numpy(near_h).to(cuda)
    big = sess.run_batch(near, far)
    assert big.shape[0] == S and torch.isfinite(big).all()
    for lo in (0, S // 2, S - 4):
        small = sess.run_batch(near[lo:lo + 4].contiguous(), far[lo:lo + 4].contiguous())
        assert torch.equal(big[lo:lo + 4], small), f"pairs {lo}..{lo + 3} depend on the batch"
    orc = DfsmnAecOracle(wts, cfg)
    worst = 0.0
    for i in (0, S - 1):
        ref = orc.forward(near_h[i], far_h[i]).numpy()
        worst = max(worst, float(np.abs(big[i].cpu().numpy().reshape(-1) - ref.reshape(-1)).max()))
    measured(f"dfsmn_aec bench shape {S} pairs: 2 pairs vs oracle", worst, 5e-5)
