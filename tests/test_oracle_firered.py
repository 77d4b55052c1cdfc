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
