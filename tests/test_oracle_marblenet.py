"""MarbleNet oracle (restated wrapper graph on the NeMo-shaped stand-in) against outputs of the
reference's own wrapper + BN-folding code run on the same stand-in (tests/golden/marblenet.npz)."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import synth, weights as W
from oracle import postproc as OP
from oracle.marblenet import MarbleNetOracle


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "marblenet.npz"))


@pytest.fixture(scope="module")
def oracle():
    cfg = W.MarbleNetConfig()
    return MarbleNetOracle(W.marblenet_random_init(cfg, 0), cfg)


@pytest.mark.parametrize("L", [16000, 48000, 160000])
def test_reference_validation_recipe(gold, oracle, L):
    # seed 1234, randint(-32768, 32767): NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:392-400; tolerance :414
    sil, act, n = oracle.forward(gold[f"recipe{L}_audio"][None])
    assert n == int(gold[f"recipe{L}_signal_len"]) == L // 320
    assert np.abs(act.numpy()[0, :, 0] - gold[f"recipe{L}_active"]).max() <= 1e-5
    assert np.abs(sil.numpy()[0, :, 0] - gold[f"recipe{L}_silence"]).max() <= 1e-5


def test_synthetic_clips_and_folding(gold, oracle):
    clips = synth.synth_streams(3, 160000, seed=1234)
    _, act, n = oracle.forward(clips)
    assert np.abs(act.numpy()[:, :, 0] - gold["synth_active"]).max() <= 1e-5
    assert gold["synth_active"].max() - gold["synth_active"].min() > 0.9
    # host-side folding used by the product == reference folding (checked through the outputs above
    # in the GPU tests); here: folded shapes
    f = W.marblenet_fold(W.MarbleNetConfig(), W.marblenet_random_init(W.MarbleNetConfig(), 0))
    assert f["b0.r0.dw"].shape == (80, 11) and f["b0.r0.pw"].shape == (128, 80) and f["b4.r0.dw"].shape == (64, 29)
    assert f["b1.res"].shape == (64, 128) and f["decoder.weight"].shape == (2, 128)


def test_vad_sample_script_record(gold, oracle, golden_dir):
    audio = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"]
    _, act, n = oracle.forward(audio[None])
    p = act.numpy()[0, :n, 0]
    assert p.shape == gold["sample_probs"].shape
    assert np.abs(p - gold["sample_probs"]).max() <= 1e-5
    dec = OP.frame_decisions(gold["sample_probs"], 3, 0.5, 10, 1000, 10, 3, 0)
    assert np.array_equal(dec, gold["sample_decisions"])
    seg = OP.segments_from_decisions(dec, 0.02, 0.025, len(audio) / 16000, False)
    assert np.array_equal(np.array(seg, np.float64).reshape(-1, 2), gold["sample_timestamps"])


@pytest.mark.parametrize("rate", [8000, 48000])
def test_in_graph_resampler(golden_dir, rate):
    """IN_SAMPLE_RATE != 16000 (tests/golden/marblenet_rates.npz: the reference's BN-folded wrapper built with that rate)"""
    from vadx import synth
    from oracle.marblenet import MarbleNetOracle
    g = np.load(os.path.join(golden_dir, "marblenet_rates.npz"))
    cfg = W.MarbleNetConfig()
    o = MarbleNetOracle(W.marblenet_random_init(cfg, 0), cfg, in_sample_rate=rate)
    a = synth.synth_streams(2, 2 * rate, seed=rate + 1)
    sil, act, n = o.forward(a)
    assert n == int(g[f"r{rate}_signal_len"])
    assert np.abs(act.numpy()[:, :, 0] - g[f"r{rate}_active"]).max() <= 1e-4
