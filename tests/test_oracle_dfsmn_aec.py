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


def test_near_end_only_variant_matches_reference_wrapper(golden_dir, oracle):
    """DFSMN/only_near_end_audio: the far end replaced by the graph's constant noise buffers (seeded, rebuilt here)."""
    g = np.load(os.path.join(golden_dir, "dfsmn_near.npz"))
    cfg = W.DfsmnAecConfig()
    pow_far, far_comp = W.dfsmn_near_noise(cfg, seed=77)
    assert np.array_equal(pow_far[:2, :3].astype(np.float32), g["pow_far_head"])
    assert np.array_equal(far_comp[:, :2, :3].astype(np.float32), g["far_comp_head"])
    p = oracle.forward(g["near"][0], None, far_noise=(pow_far, far_comp)).numpy()
    assert p.shape == (100,)
    assert np.abs(p - g["probs0"]).max() <= 1e-4
