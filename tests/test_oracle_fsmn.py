"""Oracle restatement of the FSMN-VAD graph + chunk loop against the golden record of the
reference's UNMODIFIED Inference_FSMN_VAD_ONNX.py (run with its own PyTorch graph, seeded weights)."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import weights as W
from oracle import fsmn as OFS, postproc as OP


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "fsmn.npz"))


@pytest.mark.parametrize("tag,L,lookback", [("c16000", 16000, 0.3), ("c512", 512, 0.0)])
def test_chunk_loop_matches_reference_script(gold, golden_dir, tag, L, lookback):
    cfg = W.FsmnConfig()
    orc = OFS.FsmnOracle(W.fsmn_random_init(cfg, 0), cfg, L)
    audio = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"].astype(np.float32)
    # loader normalisation of the reference (FSMN/Inference_FSMN_VAD_ONNX.py:60-69)
    a16 = (audio * float(32767.0 / np.max(np.abs(audio)))).astype(np.int16)
    noise = np.random.RandomState(1234).normal(loc=0.0, scale=1.0, size=(20000,))
    r = OFS.run_stream(orc, a16, look_backward_s=lookback, noise=noise)
    assert np.array_equal(r["aligned"], gold[f"{tag}_aligned_audio"])
    scores = np.stack(r["chunks"])
    assert scores.shape == gold[f"{tag}_scores"].shape
    mism = int((scores != gold[f"{tag}_scores"]).sum())
    assert mism == 0, f"{mism} frame flags differ from the reference graph"
    nd = np.array([t[3] for t in r["trace"]], np.float32)
    ok = ~np.isnan(gold[f"{tag}_noisy_dB"])
    assert np.array_equal(np.isnan(nd), ~ok)
    assert np.abs(nd[ok] - gold[f"{tag}_noisy_dB"][ok]).max() <= 1e-5
    assert np.abs(np.array([t[2] for t in r["trace"]]) - gold[f"{tag}_noise_avg_in"]).max() <= 1e-5
    assert np.array_equal(np.array(r["saved"]), gold[f"{tag}_saved"])
    assert np.array_equal(np.array(r["timestamps"], np.float64).reshape(-1, 2), gold[f"{tag}_timestamps"])
    sec = "".join(f"{OP.clock_string(a)} --> {OP.clock_string(b)}\n" for a, b in r["timestamps"])
    idx = "".join(f"{int(a * 16000)} --> {int(b * 16000)}\n" for a, b in r["timestamps"])
    assert sec == str(gold[f"{tag}_file_second"]) and idx == str(gold[f"{tag}_file_indices"])
