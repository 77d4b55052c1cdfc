"""Drop-in boundary, stage A (see oracle/dropin.py): the session calls the reference's UNMODIFIED inference scripts make
-- recorded in tests/golden/dropin_transcript.npz from all-reference runs -- go through the PRODUCT sessions' ORT surface
(get_inputs()[i].name, _inputs_meta, run(names, feed)) in the scripts' order, with the state each script feeds back taken
from the product's own outputs.  Outputs are checked against the reference's and written to
gpurun_out/dropin_vadx_outputs.npz for stage B (tests/test_dropin_replay.py), which replays them into the scripts."""
import os

import numpy as np
import pytest

import vadx
from vadx import weights as W

pytestmark = pytest.mark.gpu
OUT = {}


@pytest.fixture(scope="module")
def tr(golden_dir):
    return np.load(os.path.join(golden_dir, "dropin_transcript.npz"))


def test_firered_script_calls(cuda, tr, measured):
    """FireRedVAD/Inference_FireRed_ONNX.py:523-572"""
    cfg = W.FireRedConfig()
    sess = vadx.FireRedSession(W.firered_random_init(cfg, 0), cfg, chunk_len=16000)
    in_name, out_name = sess.get_inputs()[0].name, sess.get_outputs()[0].name
    assert sess._inputs_meta[0].shape[-1] == 16000            # :540: a static axis -> INPUT_AUDIO_LENGTH_RUN
    probs = []
    for k in range(tr["firered_audio"].shape[0]):
        p = sess.run([out_name], {in_name: tr["firered_audio"][k][None, None, :]})[0]
        assert p.shape == (1, 1, 98) and p.dtype == np.float32
        probs.append(p[0, 0])                                  # :571 all_vad_probs.append(probs[0, 0])
    OUT["firered_probs"] = np.stack(probs)
    measured("dropin firered: script calls vs the reference wrapper", np.abs(OUT["firered_probs"] - tr["firered_probs"]).max(), 3e-4)


def test_fsmn_script_calls(cuda, tr, measured):
    """FSMN/Inference_FSMN_VAD_ONNX.py:40-57 (names), :156-187 (state and call), :224-225 (running noise level)"""
    cfg = W.FsmnConfig()
    sess = vadx.FsmnSession(W.fsmn_random_init(cfg, 0), cfg, chunk_len=16000)
    ins, outs = [i.name for i in sess.get_inputs()], [o.name for o in sess.get_outputs()]
    assert len(ins) == 7 and len(outs) == 6
    assert sess._inputs_meta[0].shape[-1] == 16000
    cache = [np.zeros((1, 128, 19, 1), np.float32)] * 4
    noise = tr["fsmn_noise_avg_in"][0].copy()                  # the script's initial value (:163,:172)
    snr = float(tr["fsmn_snr_threshold"])
    fed_noise, scores, caches_out, noisy_out = [], [], [], []
    for k in range(tr["fsmn_audio"].shape[0]):
        fed_noise.append(np.asarray(noise, np.float32).copy())
        feed = {ins[0]: tr["fsmn_audio"][k][None, None, :], ins[1]: cache[0], ins[2]: cache[1], ins[3]: cache[2], ins[4]: cache[3],
                ins[5]: tr["fsmn_one_minus"][k], ins[6]: noise}
        score, c0, c1, c2, c3, noisy_dB = sess.run(outs, feed)
        assert score.dtype == np.uint8 and score.shape == tr["fsmn_score"][k].shape
        assert c0.shape == (1, 128, 19, 1) and c0.dtype == np.float32
        cache = [c0, c1, c2, c3]
        if noisy_dB > 0.0:                                     # :224-225, numpy float32 arithmetic like the script's
            noise = 0.5 * (noise + noisy_dB + snr)
        scores.append(score); caches_out.append(np.stack(cache)); noisy_out.append(np.float32(noisy_dB))
    OUT["fsmn_score"] = np.stack(scores)
    OUT["fsmn_caches"] = np.stack(caches_out)
    OUT["fsmn_noisy_dB"] = np.array(noisy_out, np.float32)
    OUT["fsmn_noise_avg_fed"] = np.stack(fed_noise).astype(np.float32)
    # flags are decisions: they may only differ where the reference's decision variable is within the contract of its
    # threshold; on this recording none is (tests/test_gpu_fsmn.py checks the margins), so they are equal
    assert np.array_equal(OUT["fsmn_score"], tr["fsmn_score"])
    measured("dropin fsmn: noisy_dB per window vs the reference wrapper", np.abs(OUT["fsmn_noisy_dB"] - tr["fsmn_noisy_dB"]).max(), 1e-4)
    measured("dropin fsmn: last cache of layer 3 vs the reference wrapper",
             np.abs(caches_out[-1][3][0, :, :, 0] - tr["fsmn_cache3_last"]).max() / max(1.0, float(np.abs(tr["fsmn_cache3_last"]).max())), 1e-4)


def test_silero_wrapper_calls(cuda, tr, measured):
    """Silero/modeling_modified/utils_vad.py:119-128: session.run(None, {'input', 'state', 'sr'}) -> (out, state)"""
    cfg = W.SileroConfig()
    sess = vadx.SileroSession(W.silero_random_init(cfg, 0), cfg)
    assert [i.name for i in sess.get_inputs()] == ["input", "state", "sr"]
    state = np.zeros((2, 1, 128), np.float32)
    outs, states = [], []
    for k in range(tr["silero_input"].shape[0]):
        out, state = sess.run(None, {"input": tr["silero_input"][k][None, :], "state": state,
                                     "sr": np.array(int(tr["silero_sr"][k]), dtype="int64")})
        assert out.shape == (1, 1) and state.shape == (2, 1, 128) and state.dtype == np.float32
        outs.append(out[0, 0]); states.append(state.copy())
    OUT["silero_output"] = np.array(outs, np.float32)
    OUT["silero_state"] = np.stack(states)
    measured("dropin silero: 175 wrapper calls vs the restated network", np.abs(OUT["silero_output"] - tr["silero_output"]).max(), 5e-5)
    measured("dropin silero: last LSTM state", np.abs(states[-1][:, 0, :] - tr["silero_state_last"]).max(), 1e-4)


def test_write_stage_a_outputs(cuda):
    need = {"firered_probs", "fsmn_score", "fsmn_caches", "fsmn_noisy_dB", "fsmn_noise_avg_fed", "silero_output", "silero_state"}
    assert need <= set(OUT), "the three script-call tests must run first"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(root, "gpurun_out", "dropin_vadx_outputs.npz"), **OUT)
