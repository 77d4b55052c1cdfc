"""DFSMN AEC-VAD oracle restatement against the reference's own NET / AlphaPredictor / DFSMN_VAD /
UniDeepFsmn modules (tests/golden/dfsmn_aec.npz, seeded weights)."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import weights as W
from oracle.dfsmn_aec import DfsmnAecOracle


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "dfsmn_aec.npz"))


@pytest.fixture(scope="module")
def oracle():
    cfg = W.DfsmnAecConfig()
    return DfsmnAecOracle(W.dfsmn_aec_random_init(cfg, 0), cfg)


def test_iccrn_matches_reference_net(gold, oracle):
    with torch.inference_mode():
        y, n = oracle.iccrn(torch.from_numpy(gold["iccrn_in"]))
    assert n == gold["iccrn_out"].shape[0] == (24 - 1) * 160 + 319 - 2 * 159
    assert np.abs(y.numpy()[0, 0] - gold["iccrn_out"]).max() <= 1e-5 * max(1.0, np.abs(gold["iccrn_out"]).max())


def test_whole_graph_matches_reference_wrapper(gold, oracle):
    p = oracle.forward(gold["near"][0], gold["far"][0]).numpy()
    assert p.shape == (100,)
    assert np.abs(p - gold["probs0"]).max() <= 1e-4
    assert gold["probs0"].max() - gold["probs0"].min() > 0.3
