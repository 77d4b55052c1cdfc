"""Literal drop-in check of the ORT boundary, in two stages (TEST INFRASTRUCTURE).

The reference's UNMODIFIED Inference_*.py scripts cannot travel to the GPU box (/root/reference does not exist there) and
this container has no GPU, so "unmodified script + vadx session" is closed across the two machines:

  record  (here, reference present)  the scripts run under oracle/ref_runner.py with the reference's own PyTorch wrappers
          behind the fake onnxruntime; every session call is recorded -> tests/golden/dropin_transcript.npz
          (python -m oracle.make_golden dropin): the feeds the scripts build and the reference's outputs.
  stage A (GPU box, tests/test_gpu_dropin.py)  the transcript's calls go through the PRODUCT sessions' ORT surface
          (get_inputs()[i].name, _inputs_meta, run) in the scripts' call order, state fed back from the product's OWN
          outputs the way each script does it; outputs are checked against the reference's and written to
          gpurun_out/dropin_vadx_outputs.npz (committed as tests/golden/dropin_vadx_outputs.npz).
  stage B (here, tests/test_dropin_replay.py)  the unmodified scripts run again with a ReplaySession that answers every
          run() with what vadx produced on the B200 -- after checking that the feed the script built is the one stage A
          fed -- and the two text files must equal the goldens of the all-reference run.

FireRed:  FireRedVAD/Inference_FireRed_ONNX.py:523-613 (RUN_VAD section)
FSMN:     FSMN/Inference_FSMN_VAD_ONNX.py:40-57,156-234 (16000-sample windows, LOOK_BACKWARD 0.3)
Silero:   Silero/Inference_Silero_VAD_ONNX.py:80-96 through the reference's OnnxWrapper (utils_vad.py:87-146)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

from . import ref_loader as RL
from . import ref_runner as RR

FAMILIES = ("firered", "fsmn", "silero")


def _load_wav(p, sr):
    import vadx  # noqa: F401
    from vadx import audio_io
    return audio_io.load_wav_int16(os.path.realpath(p), sr)


def run_firered_script(factory):
    wav = os.path.join(RL.REF_ROOT, "FireRedVAD", "vad_sample.wav")
    return RR.run_script("FireRedVAD/Inference_FireRed_ONNX.py", factory, _load_wav, seed=1234,
                         overrides={"RUN_AED": False, "RUN_STREAM_VAD": False}, files_to_link={"vad_sample.wav": wav})


def run_fsmn_script(factory):
    wav = os.path.join(RL.REF_ROOT, "FSMN", "vad_sample.wav")
    return RR.run_script("FSMN/Inference_FSMN_VAD_ONNX.py", factory, _load_wav, overrides={"LOOK_BACKWARD": 0.3}, seed=1234,
                         files_to_link={"vad_sample.wav": wav})


def run_silero_script(session_factory):
    """The script calls silero_vad.load_silero_vad(onnx=True) -> the reference's OnnxWrapper around an
    onnxruntime.InferenceSession; only that session is ours."""
    uv = RL.import_file("Silero/modeling_modified/utils_vad.py", "silero_utils_vad_ref")
    fake_ort = RR._fake_onnxruntime(session_factory)
    sv = types.ModuleType("silero_vad")
    sv.get_speech_timestamps = uv.get_speech_timestamps

    def load_silero_vad(onnx=False, opset_version=16, use_cpu=True, path=""):
        saved = sys.modules.get("onnxruntime")
        sys.modules["onnxruntime"] = fake_ort
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return uv.OnnxWrapper(str(path), force_onnx_cpu=use_cpu)
        finally:
            if saved is not None:
                sys.modules["onnxruntime"] = saved

    sv.load_silero_vad = load_silero_vad
    wav = os.path.join(RL.REF_ROOT, "Silero", "vad_sample.wav")
    return RR.run_script("Silero/Inference_Silero_VAD_ONNX.py", session_factory, _load_wav, seed=1234,
                         files_to_link={"vad_sample.wav": wav}, extra_modules={"silero_vad": sv})


# ------------------------------------------------------------------------------------------------ record
def record_transcripts() -> dict:
    """All-reference runs; returns the flat dict saved as tests/golden/dropin_transcript.npz."""
    import torch
    import vadx  # noqa: F401
    from vadx import weights as W
    from oracle import make_golden as MG
    from oracle.silero import SileroNetOracle

    out = {}
    # FireRed, RUN_VAD only
    cfg = W.FireRedConfig()
    ref_v, _ = MG.firered_reference(cfg, W.firered_random_init(cfg, 0))
    box = {}

    def fr_factory(_path):
        def fn(feed):
            with torch.inference_mode():
                return [ref_v(torch.from_numpy(feed["audio"])).numpy()]
        box["s"] = RR.FakeSession([RR.NodeArg("audio", [1, 1, 16000], "tensor(int16)")],
                                  [RR.NodeArg("probs", [1, 1, 98], "tensor(float)")], fn)
        return box["s"]

    ns, files = run_firered_script(fr_factory)
    calls = box["s"].calls
    out["firered_audio"] = np.stack([c[0]["audio"][0, 0] for c in calls]).astype(np.int16)
    out["firered_probs"] = np.stack([c[1][0][0, 0] for c in calls]).astype(np.float32)
    out["firered_file_second"] = np.array(files["timestamps_second.txt"])
    out["firered_file_indices"] = np.array(files["timestamps_indices.txt"])
    # FSMN, 16000-sample windows
    cfg = W.FsmnConfig()
    w = W.fsmn_random_init(cfg, 0)

    def fs_factory(_path):
        box["s"] = MG.fsmn_fake_session(cfg, w, 16000)
        return box["s"]

    ns, files = run_fsmn_script(fs_factory)
    calls = box["s"].calls
    out["fsmn_audio"] = np.stack([c[0]["audio"][0, 0] for c in calls]).astype(np.int16)
    out["fsmn_one_minus"] = np.stack([c[0]["one_minus_speech_threshold"] for c in calls]).astype(np.float32)
    out["fsmn_noise_avg_in"] = np.stack([c[0]["noise_average_dB"] for c in calls]).astype(np.float32)
    out["fsmn_score"] = np.stack([c[1][0] for c in calls]).astype(np.uint8)
    out["fsmn_noisy_dB"] = np.array([c[1][5] for c in calls], np.float32)
    out["fsmn_cache3_last"] = calls[-1][1][4][0, :, :, 0]
    out["fsmn_snr_threshold"] = np.array(ns["SNR_THRESHOLD"], np.float64)       # already scaled by 0.1 (:173)
    for k in range(1, len(calls)):   # the script's protocol: caches pass straight through
        for i in range(4):
            assert np.array_equal(calls[k][0][f"cache_{i}"], calls[k - 1][1][1 + i])
    out["fsmn_file_second"] = np.array(files["timestamps_second.txt"])
    out["fsmn_file_indices"] = np.array(files["timestamps_indices.txt"])
    # Silero
    cfg = W.SileroConfig()
    net = SileroNetOracle(W.silero_random_init(cfg, 0), cfg)

    def si_factory(_path):
        ins = [RR.NodeArg("input", [None, 576], "tensor(float)"), RR.NodeArg("state", [2, None, 128], "tensor(float)"),
               RR.NodeArg("sr", [], "tensor(int64)")]
        outs = [RR.NodeArg("output", [None, 1], "tensor(float)"), RR.NodeArg("stateN", [2, None, 128], "tensor(float)")]

        def fn(feed):
            o, s = net.step(torch.from_numpy(feed["input"]), torch.from_numpy(feed["state"]))
            return [o.numpy(), s.numpy()]

        box["s"] = RR.FakeSession(ins, outs, fn)
        return box["s"]

    ns, files = run_silero_script(si_factory)
    calls = box["s"].calls
    out["silero_input"] = np.stack([c[0]["input"][0] for c in calls]).astype(np.float32)
    out["silero_sr"] = np.array([int(c[0]["sr"]) for c in calls], np.int64)
    out["silero_output"] = np.array([c[1][0][0, 0] for c in calls], np.float32)
    out["silero_state_last"] = calls[-1][1][1][:, 0, :]
    for k in range(1, len(calls)):
        assert np.array_equal(calls[k][0]["state"], calls[k - 1][1][1])
    assert not calls[0][0]["state"].any()
    out["silero_file_second"] = np.array(files["timestamps_second.txt"])
    out["silero_file_indices"] = np.array(files["timestamps_indices.txt"])
    return out


# ------------------------------------------------------------------------------------------------ stage B
class ReplaySession:
    """ORT-shaped session that answers the k-th run() with recorded outputs after checking the feed the script built.
    `expect(k, feed)` raises AssertionError on a mismatch; `outputs[k]` is the list in the session's output order."""

    def __init__(self, inputs, outputs_meta, outputs, expect):
        self._inputs_meta, self._outputs_meta = inputs, outputs_meta
        self._outputs, self._expect, self.k = outputs, expect, 0

    def get_inputs(self):
        return self._inputs_meta

    def get_outputs(self):
        return self._outputs_meta

    def get_providers(self):
        return ["B200ExecutionProvider (replayed)"]

    def run(self, output_names, feed, run_options=None):
        assert self.k < len(self._outputs), f"the script makes more than the {len(self._outputs)} recorded calls"
        assert set(feed) == {i.name for i in self._inputs_meta}, sorted(feed)
        self._expect(self.k, feed)
        outs = self._outputs[self.k]
        self.k += 1
        names = [o.name for o in self._outputs_meta]
        if output_names is None:
            return list(outs)
        return [outs[names.index(n)] for n in output_names]


def replay_scripts(tr: dict, vx: dict) -> dict:
    """Stage B.  tr = dropin_transcript.npz, vx = dropin_vadx_outputs.npz -> {family: (ns, files, calls_made)}."""
    res = {}
    # FireRed
    n = tr["firered_audio"].shape[0]

    def fr_expect(k, feed):
        a = feed["audio"]
        assert a.dtype == np.int16 and a.shape == (1, 1, 16000) and np.array_equal(a[0, 0], tr["firered_audio"][k])

    fr = ReplaySession([RR.NodeArg("audio", [1, 1, 16000], "tensor(int16)")], [RR.NodeArg("probs", [1, 1, 98], "tensor(float)")],
                       [[vx["firered_probs"][k][None, None, :]] for k in range(n)], fr_expect)
    ns, files = run_firered_script(lambda _p: fr)
    res["firered"] = (ns, files, fr.k)
    # FSMN: caches and the running noise level are fed back from the product's own outputs by the SCRIPT; stage A fed
    # exactly those values, so the script's feeds must reproduce stage A's bit for bit
    n = tr["fsmn_audio"].shape[0]
    T = vx["fsmn_score"].shape[1]

    def fs_expect(k, feed):
        assert np.array_equal(feed["audio"][0, 0], tr["fsmn_audio"][k]) and feed["audio"].dtype == np.int16
        assert np.array_equal(feed["one_minus_speech_threshold"], tr["fsmn_one_minus"][k])
        got = np.asarray(feed["noise_average_dB"], np.float32)
        assert np.array_equal(got, vx["fsmn_noise_avg_fed"][k]), (k, got, vx["fsmn_noise_avg_fed"][k])
        for i in range(4):
            want = vx["fsmn_caches"][k - 1, i] if k else np.zeros((1, 128, 19, 1), np.float32)
            assert np.array_equal(feed[f"cache_{i}"], want), (k, i)

    ins = [RR.NodeArg("audio", [1, 1, 16000], "tensor(int16)")]
    ins += [RR.NodeArg(f"cache_{i}", [1, 128, 19, 1], "tensor(float)") for i in range(4)]
    ins += [RR.NodeArg("one_minus_speech_threshold", [1], "tensor(float)"), RR.NodeArg("noise_average_dB", [1], "tensor(float)")]
    outs_meta = [RR.NodeArg("score", [T], "tensor(uint8)")]
    outs_meta += [RR.NodeArg(f"cache_{i}_out", [1, 128, 19, 1], "tensor(float)") for i in range(4)]
    outs_meta += [RR.NodeArg("noisy_dB", [], "tensor(float)")]
    fs = ReplaySession(ins, outs_meta,
                       [[vx["fsmn_score"][k]] + [vx["fsmn_caches"][k, i] for i in range(4)] + [np.float32(vx["fsmn_noisy_dB"][k])]
                        for k in range(n)], fs_expect)
    ns, files = run_fsmn_script(lambda _p: fs)
    res["fsmn"] = (ns, files, fs.k)
    # Silero: the reference's OnnxWrapper builds input = cat(context, window) and passes stateN back as state
    n = tr["silero_input"].shape[0]

    def si_expect(k, feed):
        assert np.array_equal(feed["input"][0], tr["silero_input"][k]) and feed["input"].shape == (1, 576)
        assert int(feed["sr"]) == int(tr["silero_sr"][k])
        want = vx["silero_state"][k - 1] if k else np.zeros((2, 1, 128), np.float32)
        assert np.array_equal(feed["state"], want), k

    si_ins = [RR.NodeArg("input", [None, 576], "tensor(float)"), RR.NodeArg("state", [2, None, 128], "tensor(float)"),
              RR.NodeArg("sr", [], "tensor(int64)")]
    si_outs = [RR.NodeArg("output", [None, 1], "tensor(float)"), RR.NodeArg("stateN", [2, None, 128], "tensor(float)")]
    si = ReplaySession(si_ins, si_outs, [[vx["silero_output"][k].reshape(1, 1), vx["silero_state"][k]] for k in range(n)], si_expect)
    ns, files = run_silero_script(lambda _p: si)
    res["silero"] = (ns, files, si.k)
    return res
