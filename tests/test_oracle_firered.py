"""The oracle restatement of the FireRedVAD graph against outputs of the REAL reference modules
(frozen in tests/golden/firered.npz by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import synth, weights as W
from oracle.firered import FireRedOracle

TOL = 1e-5  # the reference's own PyTorch-vs-ORT tolerance (FireRedVAD/Export_FireRedVAD.py:1553)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "firered.npz"))


@pytest.fixture(scope="module")
def oracle():
    cfg = W.FireRedConfig()
    return FireRedOracle(W.firered_random_init(cfg, 0), cfg)


def test_reference_validation_recipe(gold, oracle):
    # seed 1234, randint(-8000, 8000) -- FireRedVAD/Export_FireRedVAD.py:1539-1543
    np.random.seed(1234)
    a = np.random.randint(-8000, 8000, size=(1, 1, 16000)).astype(np.int16)
    p = oracle.forward(a).numpy()
    assert p.shape == (1, 1, 98)
    assert np.abs(p - gold["recipe_probs"]).max() <= TOL
    assert p.max() - p.min() > 0.5, "random-init head should spread probabilities"


def test_synthetic_chunks_and_intermediates(gold, oracle):
    chunks = synth.synth_streams(6, 16000, seed=1234)
    p = oracle.forward(chunks).numpy()
    assert np.abs(p - gold["synth_probs"]).max() <= TOL
    lm = oracle.logmel(torch.from_numpy(chunks[:1]).unsqueeze(1)).numpy()[0]
    assert np.abs(lm - gold["synth_logmel0"]).max() <= 1e-4


@pytest.mark.parametrize("L", [5000, 400])
def test_ragged_lengths(gold, oracle, L):
    a = synth.synth_streams(1, L, seed=77)
    p = oracle.forward(a).numpy()
    assert p.shape == gold[f"len{L}_probs"].shape
    assert np.abs(p - gold[f"len{L}_probs"]).max() <= TOL


def test_aed_head_with_strides(gold):
    cfg = W.FireRedConfig(R=3, H=96, P=64, N1=5, S1=2, N2=3, S2=2, odim=3)
    o = FireRedOracle(W.firered_random_init(cfg, 3), cfg)
    chunks = synth.synth_streams(6, 16000, seed=1234)
    p = o.forward(chunks[1:2]).numpy()
    assert np.abs(p - gold["aed_probs"]).max() <= TOL


def test_streaming_twin_cache_carry(gold):
    cfg = W.FireRedConfig(N2=0, S2=0, streaming=True)
    o = FireRedOracle(W.firered_random_init(cfg, 5), cfg)
    chunks = synth.synth_streams(6, 16000, seed=1234)
    caches = torch.zeros(cfg.R, 1, cfg.P, (cfg.N1 - 1) * cfg.S1)
    for i in range(4):
        p, caches = o.forward(chunks[2][i * 2560:(i + 1) * 2560][None], caches)
        assert np.abs(p.numpy()[0, 0] - gold["stream_probs"][i]).max() <= TOL
    assert np.abs(caches.numpy()[:, 0, ::16, :] - gold["stream_caches_last"]).max() <= 1e-4


# ---- the UNMODIFIED Inference_FireRed_ONNX.py on vad_sample.wav (tests/golden/firered_script.npz) ----
@pytest.fixture(scope="module")
def gold_script(golden_dir):
    return np.load(os.path.join(golden_dir, "firered_script.npz"))


def test_script_vad_and_aed_sections(gold_script, golden_dir, oracle):
    """VAD then AED, as the script runs them: both tails are padded from the same global numpy stream."""
    from vadx import audio_io
    from oracle import postproc as OP
    audio = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"]
    rng = np.random.RandomState(1234)
    chunks, n = audio_io.align_non_overlapping(audio, 16000, rng)
    nv = OP.valid_frames(n)
    p = oracle.forward(chunks).numpy().reshape(-1)[:nv]
    assert np.abs(p - gold_script["vad_probs"]).max() <= TOL
    dec = OP.frame_decisions(gold_script["vad_probs"], 5, 0.4, 20, 2000, 20, 5, 0)
    ts = OP.segments_from_decisions(dec, 0.01, 0.025, n / 16000, True)
    assert np.array_equal(np.array(ts, np.float64).reshape(-1, 2), gold_script["vad_timestamps"])
    # AED
    cfg = W.FireRedConfig(odim=3)
    aed = FireRedOracle(W.firered_random_init(cfg, 2), cfg)
    chunks2, _ = audio_io.align_non_overlapping(audio, 16000, rng)
    pa = aed.forward(chunks2).numpy()                       # [n_chunks, 3, 98]
    pa = np.concatenate(list(pa), axis=1)[:, :nv]
    assert np.abs(pa - gold_script["aed_probs"]).max() <= TOL
    for e, (ev, thr) in enumerate((("speech", 0.4), ("singing", 0.5), ("music", 0.5))):
        d = OP.frame_decisions(gold_script["aed_probs"][e], 5, thr, 20, 2000, 20, 5, 0)
        t = OP.segments_from_decisions(d, 0.01, 0.025, n / 16000, True)
        assert np.array_equal(np.array(t, np.float64).reshape(-1, 2), gold_script[f"aed_{ev}_timestamps"])
        assert round(float(np.mean(gold_script["aed_probs"][e] >= thr)), 3) == float(gold_script[f"aed_{ev}_ratio"])


def test_script_stream_section(gold_script, golden_dir):
    audio = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"]
    cfg = W.FireRedConfig(N2=0, S2=0, streaming=True)
    o = FireRedOracle(W.firered_random_init(cfg, 5), cfg)
    caches = torch.zeros(cfg.R, 1, cfg.P, (cfg.N1 - 1) * cfg.S1)
    out = []
    for pos in range(0, len(audio), 2560):
        c = audio[pos:pos + 2560]
        if len(c) < 400:
            c = np.pad(c, (0, 400 - len(c)))
        p, caches = o.forward(torch.from_numpy(np.ascontiguousarray(c)).view(1, 1, -1), caches)
        out.append(p.numpy()[0, 0])
    p = np.concatenate(out)[:557]
    assert p.shape == gold_script["stream_probs"].shape
    assert np.abs(p - gold_script["stream_probs"]).max() <= TOL
    assert np.abs(caches.numpy()[:, 0, ::16, :] - gold_script["stream_caches_last"]).max() <= 1e-4


@pytest.mark.parametrize("rate", [8000, 48000, 22050])
def test_in_graph_resampler(golden_dir, rate):
    """IN_SAMPLE_RATE != 16000 (tests/golden/firered_rates.npz: the reference wrapper built with that rate)"""
    g = np.load(os.path.join(golden_dir, "firered_rates.npz"))
    cfg = W.FireRedConfig()
    o = FireRedOracle(W.firered_random_init(cfg, 0), cfg, in_sample_rate=rate)
    a = synth.synth_streams(2, rate, seed=rate)
    p = o.forward(a).numpy()
    assert p.shape == g[f"r{rate}_probs"].shape
    # 1e-4: an upsampled signal has almost no energy in the upper mel bands, so the log amplifies fp32
    # summation-order differences between this batch-of-2 call and the reference's batch-1 calls
    assert np.abs(p - g[f"r{rate}_probs"]).max() <= 1e-4
