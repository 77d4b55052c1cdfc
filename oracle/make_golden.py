"""Generate tests/golden/*.npz by running the REAL reference modules from /root/reference.

TEST INFRASTRUCTURE.  Only runs in the authoring container (needs /root/reference).  The
reference classes are AST-loaded (oracle/ref_loader.py), never copied; the same seeded weights
the product uses (vadx.weights.*_random_init) are loaded into them with load_state_dict.

    python -m oracle.make_golden [firered] [postproc] [fsmn] [marblenet] [dfsmn] ...
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader as RL  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _sd_from_numpy(w):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}


# ------------------------------------------------------------------------------ FireRed
def firered_reference(cfg, weights, streaming=False, in_sample_rate=16000):
    stft = RL.import_file("FireRedVAD/STFT_Process.py", "STFT_Process")
    ns = RL.extract("FireRedVAD/Export_FireRedVAD.py", {"STFT_Process": stft.STFT_Process})
    args = RL.Args(idim=cfg.idim, R=cfg.R, M=cfg.M, H=cfg.H, P=cfg.P, N1=cfg.N1, S1=cfg.S1,
                   N2=cfg.N2, S2=cfg.S2, odim=cfg.odim)
    if streaming:
        dm = ns["DetectModel_Streaming"](args).eval()
        dm.load_state_dict(_sd_from_numpy(weights), strict=True)
        return ns["FireRedStreamVAD_ONNX"](dm, cfg.n_fft, cfg.hop, cfg.win_length, cfg.n_mels, 16000,
                                           cfg.pre_emphasis, cfg.window).eval(), ns
    dm = ns["DetectModel"](args).eval()
    dm.load_state_dict(_sd_from_numpy(weights), strict=True)
    return ns["FireRedVAD_ONNX"](dm, cfg.n_fft, cfg.hop, cfg.win_length, cfg.n_mels, 16000,
                                 cfg.pre_emphasis, cfg.window, in_sample_rate).eval(), ns


def gen_firered():
    import vadx  # noqa: F401
    from vadx import synth, weights as W

    out = {}
    # (1) default architecture, the reference validation recipe: seed 1234, randint(-8000, 8000)
    cfg = W.FireRedConfig()
    w = W.firered_random_init(cfg, seed=0)
    ref, ns = firered_reference(cfg, w)
    np.random.seed(1234)
    a0 = np.random.randint(-8000, 8000, size=(1, 1, 16000)).astype(np.int16)
    with torch.inference_mode():
        out["recipe_probs"] = ref(torch.from_numpy(a0)).numpy()
    # (2) enveloped synthetic chunks (the bench workload), batch-1 calls like the reference
    chunks = synth.synth_streams(6, 16000, seed=1234)
    with torch.inference_mode():
        out["synth_probs"] = np.stack([ref(torch.from_numpy(c).view(1, 1, -1)).numpy()[0] for c in chunks])
        # intermediate: log-mel of chunk 0, by replaying the wrapper's frontend
        x = torch.from_numpy(chunks[0]).view(1, 1, -1).float()
        x = torch.nn.functional.conv1d(torch.nn.functional.pad(x, (1, 0)), ref.preemph_kernel)
        re, im = ref.stft(x)
        pw = re * re + im * im
        out["synth_power0"] = pw.numpy()[0, :, ::7]
        mel = torch.clamp(torch.nn.functional.conv1d(pw, ref.fbank_conv), min=1e-7).log()
        out["synth_logmel0"] = mel.numpy()[0]
    # (3) ragged length (dynamic axis), 5000 samples -> 29 frames; and minimal 400 -> 1 frame
    for L in (5000, 400):
        a = synth.synth_streams(1, L, seed=77)[0]
        with torch.inference_mode():
            out[f"len{L}_probs"] = ref(torch.from_numpy(a).view(1, 1, -1)).numpy()
    # (4) AED head (odim 3) on a small architecture with strides > 1
    cfg3 = W.FireRedConfig(R=3, H=96, P=64, N1=5, S1=2, N2=3, S2=2, odim=3)
    w3 = W.firered_random_init(cfg3, seed=3)
    ref3, _ = firered_reference(cfg3, w3)
    with torch.inference_mode():
        out["aed_probs"] = ref3(torch.from_numpy(chunks[1]).view(1, 1, -1)).numpy()
    # (5) streaming twin: 2560-sample chunks with cache carry over 4 chunks
    cfgs = W.FireRedConfig(N2=0, S2=0, streaming=True)
    ws = W.firered_random_init(cfgs, seed=5)
    refs, _ = firered_reference(cfgs, ws, streaming=True)
    caches = torch.zeros(cfgs.R, 1, cfgs.P, (cfgs.N1 - 1) * cfgs.S1)
    pr = []
    with torch.inference_mode():
        for i in range(4):
            p, caches = refs(torch.from_numpy(chunks[2][i * 2560:(i + 1) * 2560].copy()).view(1, 1, -1), caches)
            pr.append(p.numpy()[0, 0])
    out["stream_probs"] = np.stack(pr)
    out["stream_caches_last"] = caches.numpy()[:, 0, ::16, :]
    np.savez_compressed(os.path.join(GOLD, "firered.npz"), **out)
    print("firered.npz:", {k: v.shape for k, v in out.items()})


def gen_firered_rates():
    """IN_SAMPLE_RATE != 16000: the wrapper's in-graph linear resampler (Export_FireRedVAD.py:389-393,431-449)
    for a lower, a higher and a non-integer-ratio input rate."""
    import vadx  # noqa: F401
    from vadx import synth, weights as W
    cfg = W.FireRedConfig()
    w = W.firered_random_init(cfg, seed=0)
    out = {}
    for rate in (8000, 48000, 22050):
        ref, _ = firered_reference(cfg, w, in_sample_rate=rate)
        a = synth.synth_streams(2, rate, seed=rate)          # one second at the input rate
        with torch.inference_mode():
            out[f"r{rate}_probs"] = np.stack([ref(torch.from_numpy(c).view(1, 1, -1)).numpy()[0] for c in a])
        print(rate, out[f"r{rate}_probs"].shape)
    np.savez_compressed(os.path.join(GOLD, "firered_rates.npz"), **out)


# ------------------------------------------------------------------------------ FireRed: whole script
def gen_firered_script():
    """The reference's UNMODIFIED FireRedVAD/Inference_FireRed_ONNX.py on vad_sample.wav with all three
    sections on (VAD, AED, Stream-VAD), its own PyTorch wrappers behind a fake onnxruntime; plus the
    StreamVadPostprocessor class on synthetic probability tracks, one call and chunked calls."""
    import vadx  # noqa: F401
    from vadx import audio_io, weights as W
    from oracle import ref_runner as RR

    cfg_v = W.FireRedConfig()
    cfg_a = W.FireRedConfig(odim=3)
    cfg_s = W.FireRedConfig(N2=0, S2=0, streaming=True)
    w_v, w_a, w_s = W.firered_random_init(cfg_v, 0), W.firered_random_init(cfg_a, 2), W.firered_random_init(cfg_s, 5)
    ref_v, _ = firered_reference(cfg_v, w_v)
    ref_a, _ = firered_reference(cfg_a, w_a)
    ref_s, _ = firered_reference(cfg_s, w_s, streaming=True)
    Lb = (cfg_s.N1 - 1) * cfg_s.S1

    def static(ref, odim):
        def fn(feed):
            with torch.inference_mode():
                return [ref(torch.from_numpy(feed["audio"])).numpy()]
        return RR.FakeSession([RR.NodeArg("audio", [1, 1, 16000], "tensor(int16)")],
                              [RR.NodeArg("probs", [1, odim, 98], "tensor(float)")], fn)

    def streaming():
        def fn(feed):
            with torch.inference_mode():
                p, c = ref_s(torch.from_numpy(feed["audio"]), torch.from_numpy(feed["caches_in"]))
            return [p.numpy(), c.numpy()]
        return RR.FakeSession([RR.NodeArg("audio", [1, 1, "audio_len"], "tensor(int16)"),
                               RR.NodeArg("caches_in", [cfg_s.R, 1, cfg_s.P, Lb], "tensor(float)")],
                              [RR.NodeArg("probs", [1, 1, "T"], "tensor(float)"),
                               RR.NodeArg("caches_out", [cfg_s.R, 1, cfg_s.P, Lb], "tensor(float)")], fn)

    def factory(path):
        if path.endswith("FireRedVAD.onnx"):
            return static(ref_v, 1)
        if path.endswith("FireRedAED.onnx"):
            return static(ref_a, 3)
        return streaming()

    wav = os.path.join(RL.REF_ROOT, "FireRedVAD", "vad_sample.wav")
    ns, files = RR.run_script("FireRedVAD/Inference_FireRed_ONNX.py", factory,
                              lambda p, sr: audio_io.load_wav_int16(os.path.realpath(p), sr), seed=1234,
                              files_to_link={"vad_sample.wav": wav})
    out = {}
    out["vad_probs"] = np.asarray(ns["all_vad_probs"], np.float32)
    out["vad_timestamps"] = np.array(ns["timestamps"], np.float64).reshape(-1, 2)
    out["vad_file_second"] = np.array(files["timestamps_second.txt"])
    out["vad_file_indices"] = np.array(files["timestamps_indices.txt"])
    out["aed_probs"] = np.asarray(ns["all_probs"], np.float32)
    for ev in ("speech", "singing", "music"):
        out[f"aed_{ev}_timestamps"] = np.array(ns["event2timestamps"][ev], np.float64).reshape(-1, 2)
        out[f"aed_{ev}_ratio"] = np.array(ns["event2ratio"][ev], np.float64)
    out["stream_probs"] = np.asarray(ns["all_stream_probs"], np.float32)
    out["stream_timestamps"] = np.array(ns["stream_timestamps"], np.float64).reshape(-1, 2)
    out["stream_caches_last"] = np.asarray(ns["caches"], np.float32)[:, 0, ::16, :]
    print("script: vad", out["vad_probs"].shape, out["vad_timestamps"].tolist())
    print("        aed", out["aed_probs"].shape, {ev: out[f"aed_{ev}_timestamps"].shape[0] for ev in ("speech", "singing", "music")})
    print("        stream", out["stream_probs"].shape, out["stream_timestamps"].tolist())

    # StreamVadPostprocessor on synthetic tracks: (ws, thr, pad_start, min_speech, max_speech, min_silence, split)
    cls = RL.extract("FireRedVAD/Inference_FireRed_ONNX.py")["StreamVadPostprocessor"]
    rs = np.random.RandomState(7)
    for i, (n, prm, split) in enumerate([
            (1400, (5, 0.4, 5, 8, 2000, 20), 14),
            (3000, (5, 0.4, 5, 8, 150, 20), 14),      # max-speech re-arm
            (900, (1, 0.5, 0, 1, 2000, 1), 7),        # no smoothing, immediate transitions
            (2000, (3, 0.45, 9, 4, 90, 6), 0),        # pad_start > window, one call
            (40, (5, 0.4, 5, 8, 2000, 20), 14),       # ends inside a segment
            (2, (5, 0.4, 5, 8, 2000, 20), 0)]):
        steps = rs.normal(0, 0.08, size=n)
        lvl = np.clip(0.5 + np.cumsum(steps) * 0.5, 0, 1)
        gate = (np.sin(np.arange(n) / rs.uniform(15, 60)) > rs.uniform(-0.5, 0.5)).astype(np.float64)
        p = np.clip(0.12 + 0.75 * gate * (0.35 + 0.65 * lvl) + rs.normal(0, 0.06, n), 0, 1).astype(np.float32)
        if i == 1:
            p[100:2500] = np.clip(p[100:2500] + 0.5, 0, 1)
        if i == 4:
            p[10:] = 0.9
        whole = cls(*prm).process_batch(p.copy())
        out[f"sp{i}_probs"] = p
        out[f"sp{i}_params"] = np.array(prm + (split,), np.float64)
        out[f"sp{i}_whole"] = np.array(whole, np.float64).reshape(-1, 2)
        if split:
            pp = cls(*prm)
            per_call = [pp.process_batch(p[j:j + split].copy()) for j in range(0, n, split)]
            out[f"sp{i}_chunked"] = np.array([t for c in per_call for t in c], np.float64).reshape(-1, 2)
            out[f"sp{i}_chunked_counts"] = np.array([len(c) for c in per_call], np.int32)
        print(f"sp{i}: n={n} segments {len(whole)}" + (f", chunked total {len(out[f'sp{i}_chunked'])}" if split else ""))
    np.savez_compressed(os.path.join(GOLD, "firered_script.npz"), **out)


# ------------------------------------------------------------------------------ post-processing
def gen_postproc():
    ns = RL.extract("FireRedVAD/Inference_FireRed_ONNX.py")
    nsn = RL.extract("NVIDIA_Frame_VAD_Multilingual_MarbleNet/Inference_NVIDIA_MarbleNet_VAD_ONNX.py")
    nsf = RL.extract("FSMN/Inference_FSMN_VAD_ONNX.py")
    rs = np.random.RandomState(99)
    out = {}
    cases = []
    # random-walk probability tracks with plateaus so every state transition fires
    for i, (n, ws, thr, msp, mxs, msi, mrg, ext) in enumerate([
            (588, 5, 0.4, 20, 2000, 20, 5, 0),
            (3000, 3, 0.5, 10, 1000, 10, 3, 0),
            (4000, 5, 0.4, 20, 300, 20, 5, 0),     # forces max-speech splits
            (700, 1, 0.5, 0, 2000, 0, 0, 0),       # plain threshold path
            (900, 4, 0.45, 5, 120, 7, 4, 3),       # extend > 0
            (3, 5, 0.4, 20, 2000, 20, 5, 0),       # shorter than the smoothing window
            (1, 5, 0.4, 20, 2000, 20, 5, 0)]):
        steps = rs.normal(0, 0.08, size=n)
        lvl = np.clip(0.5 + np.cumsum(steps) * 0.5, 0, 1)
        gate = (np.sin(np.arange(n) / rs.uniform(15, 60)) > rs.uniform(-0.5, 0.5)).astype(np.float64)
        p = np.clip(0.15 + 0.7 * gate * lvl + rs.normal(0, 0.05, n), 0, 1).astype(np.float32)
        if i == 2:
            p[200:3500] = np.clip(p[200:3500] + 0.5, 0, 1)
        cases.append((p, (ws, thr, msp, mxs, msi, mrg, ext)))
    for i, (p, prm) in enumerate(cases):
        pp = ns["VadPostprocessor"](*prm)
        dec = pp.process(p.copy())
        dur = len(p) * 0.01 + 0.012
        seg = pp.decision_to_segment(dec, dur)
        out[f"fr{i}_probs"] = p
        out[f"fr{i}_params"] = np.array(prm, np.float64)
        out[f"fr{i}_dec"] = dec
        out[f"fr{i}_seg"] = np.array(seg, np.float64).reshape(-1, 2)
        out[f"fr{i}_dur"] = np.array(dur)
        # MarbleNet copy of the class (frame shift argument, open tail without +frame_length)
        try:
            ppn = nsn["VadPostprocessor"](*prm, 0.02)
        except TypeError:
            ppn = nsn["VadPostprocessor"](*prm)
        decn = ppn.process(p.copy())
        segn = ppn.decision_to_segment(decn, len(p) * 0.02 + 0.012)
        out[f"nv{i}_dec"] = decn
        out[f"nv{i}_seg"] = np.array(segn, np.float64).reshape(-1, 2)
    # format_time + timestamp fusing
    ts = [0.0, 2.28, 2.279999, 36480 / 16000, 59.9996, 3599.9999, 3600.5, 0.001, 1.0005, 12.3456789]
    out["clock_in"] = np.array(ts)
    out["clock_out"] = np.array([nsf["format_time"](t) for t in ts])
    flags = (rs.uniform(size=2000) < 0.5)
    flags = np.repeat(flags[:200], rs.randint(1, 40, size=200))[:2000]
    raw = nsf["vad_to_timestamps"](list(flags), 0.01)
    fused = nsf["process_timestamps"](list(raw), 0.3, 0.2)
    out["runs_flags"] = flags
    out["runs_raw"] = np.array(raw, np.float64).reshape(-1, 2)
    out["runs_fused"] = np.array(fused, np.float64).reshape(-1, 2)
    out["valid_frames_in"] = np.array([0, 399, 400, 559, 560, 16000, 89431, 960000])
    out["valid_frames_out"] = np.array([ns["valid_frame_count"](int(v)) for v in out["valid_frames_in"]])
    np.savez_compressed(os.path.join(GOLD, "postproc.npz"), **out)
    print("postproc.npz written:", len(out), "arrays")


# ------------------------------------------------------------------------------ audio fixture
def gen_audio():
    """vad_sample.wav (48 kHz stereo) decoded the way pydub does it (wave -> audioop.tomono ->
    audioop.ratecv), frozen as 16 kHz mono int16 so GPU-box tests need no /root/reference."""
    import vadx  # noqa: F401
    from vadx import audio_io
    a = audio_io.load_wav_int16(os.path.join(RL.REF_ROOT, "FireRedVAD", "vad_sample.wav"), 16000)
    np.savez_compressed(os.path.join(GOLD, "vad_sample_16k.npz"), audio=a)
    print("vad_sample_16k.npz:", a.shape, a.dtype, int(np.abs(a).max()))
    # the raw PCM of the first 1.5 s (48 kHz stereo, interleaved) with the loader's output for it: the
    # fixture of the on-device ingest (audioop.tomono + audioop.ratecv)
    import audioop
    import wave
    with wave.open(os.path.join(RL.REF_ROOT, "FireRedVAD", "vad_sample.wav"), "rb") as w:
        ch, width, sr = w.getnchannels(), w.getsampwidth(), w.getframerate()
        w.setpos(int(0.5 * sr))
        raw = w.readframes(int(1.5 * sr))
    assert width == 2
    mono = audioop.tomono(raw, 2, 0.5, 0.5) if ch == 2 else raw
    out = {"pcm": np.frombuffer(raw, np.int16), "channels": np.array(ch), "rate": np.array(sr)}
    for r in (16000, 8000, 22050):
        out[f"mono_{r}"] = np.frombuffer(audioop.ratecv(mono, 2, 1, sr, r, None)[0], np.int16)
    np.savez_compressed(os.path.join(GOLD, "vad_sample_pcm_head.npz"), **out)
    print("vad_sample_pcm_head.npz:", {k: v.shape for k, v in out.items()})


# ------------------------------------------------------------------------------ FSMN
def fsmn_reference(cfg, weights, input_audio_len):
    """The reference's own FSMN_VAD wrapper (FSMN/Export_FSMN_VAD.py:56-101) around its own FunASR
    FSMN encoder (FSMN/modeling_modified/encoder.py:159-217), with our seeded weights."""
    stft = RL.import_file("FSMN/STFT_Process.py", "STFT_Process")
    enc = RL.import_file("FSMN/modeling_modified/encoder.py", "fsmn_encoder_ref")
    ns = RL.extract("FSMN/Export_FSMN_VAD.py", {"STFT_Process": stft.STFT_Process})
    m = enc.FSMN(cfg.input_dim, cfg.input_affine_dim, cfg.fsmn_layers, cfg.linear_dim, cfg.proj_dim, cfg.lorder,
                 cfg.rorder, cfg.lstride, cfg.rstride, cfg.output_affine_dim, cfg.output_dim).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items() if not k.startswith("cmvn")},
                      strict=True)
    hop = min(cfg.hop, input_audio_len)
    T = input_audio_len // cfg.hop + 1
    custom_stft = stft.STFT_Process(model_type='stft_B', n_fft=cfg.n_fft, hop_len=hop, win_length=cfg.win_length,
                                    max_frames=0, window_type=cfg.window).eval()
    means = torch.from_numpy(weights["cmvn_means"]).view(1, 1, -1)
    vars_ = torch.from_numpy(weights["cmvn_vars"]).view(1, 1, -1)
    wrap = ns["FSMN_VAD"](m, custom_stft, cfg.n_fft, T, cfg.n_mels, 16000, cfg.pre_emphasis, cfg.lfr_m, cfg.lfr_n,
                          (T + cfg.lfr_n - 1) // cfg.lfr_n, cfg.speech_2_noise_ratio, input_audio_len, hop, means,
                          vars_).eval()
    return wrap, T


def fsmn_fake_session(cfg, weights, input_audio_len):
    from oracle import ref_runner as RR
    wrap, T = fsmn_reference(cfg, weights, input_audio_len)
    ins = [RR.NodeArg("audio", [1, 1, input_audio_len], "tensor(int16)")]
    ins += [RR.NodeArg(f"cache_{i}", [1, 128, 19, 1], "tensor(float)") for i in range(4)]
    ins += [RR.NodeArg("one_minus_speech_threshold", [1], "tensor(float)"),
            RR.NodeArg("noise_average_dB", [1], "tensor(float)")]
    outs = [RR.NodeArg("score", [T], "tensor(uint8)")]
    outs += [RR.NodeArg(f"cache_{i}_out", [1, 128, 19, 1], "tensor(float)") for i in range(4)]
    outs += [RR.NodeArg("noisy_dB", [], "tensor(float)")]

    def fn(feed):
        with torch.inference_mode():
            r = wrap(torch.from_numpy(feed["audio"]), *[torch.from_numpy(feed[f"cache_{i}"]) for i in range(4)],
                     torch.from_numpy(feed["one_minus_speech_threshold"]), torch.from_numpy(feed["noise_average_dB"]))
        return [np.asarray(t.numpy()) for t in r]

    return RR.FakeSession(ins, outs, fn)


def gen_fsmn():
    """Run the reference's UNMODIFIED FSMN/Inference_FSMN_VAD_ONNX.py on vad_sample.wav with its own
    PyTorch graph behind a fake onnxruntime (oracle/ref_runner.py), for both chunkings."""
    import vadx  # noqa: F401
    from vadx import audio_io, weights as W
    from oracle import ref_runner as RR

    cfg = W.FsmnConfig()
    w = W.fsmn_random_init(cfg, 0)
    wav = os.path.join(RL.REF_ROOT, "FSMN", "vad_sample.wav")
    out = {}
    for tag, L, lookback in (("c16000", 16000, 0.3), ("c512", 512, 0.0)):
        sess_box = {}

        def factory(_path, L=L):
            sess_box["s"] = fsmn_fake_session(cfg, w, L)
            return sess_box["s"]

        ns, files = RR.run_script("FSMN/Inference_FSMN_VAD_ONNX.py", factory,
                                  lambda p, sr: audio_io.load_wav_int16(os.path.realpath(p), sr),
                                  overrides={"LOOK_BACKWARD": lookback}, seed=1234,
                                  files_to_link={"vad_sample.wav": wav})
        calls = sess_box["s"].calls
        out[f"{tag}_saved"] = np.array(ns["saved"], np.bool_)
        out[f"{tag}_timestamps"] = np.array(ns["timestamps"], np.float64).reshape(-1, 2)
        out[f"{tag}_file_second"] = np.array(files["timestamps_second.txt"])
        out[f"{tag}_file_indices"] = np.array(files["timestamps_indices.txt"])
        out[f"{tag}_aligned_audio"] = ns["audio"].reshape(-1).astype(np.int16)
        out[f"{tag}_scores"] = np.stack([c[1][0] for c in calls]).astype(np.uint8)
        out[f"{tag}_noisy_dB"] = np.array([c[1][5] for c in calls], np.float32)
        out[f"{tag}_noise_avg_in"] = np.array([c[0]["noise_average_dB"][0] for c in calls], np.float32)
        out[f"{tag}_cache0_last"] = calls[-1][1][1][0, :, :, 0]
        out[f"{tag}_cache3_last"] = calls[-1][1][4][0, :, :, 0]
        print(tag, "windows", len(calls), "flags", len(ns["saved"]), "speech frames", int((~out[f'{tag}_saved']).sum()),
              "segments", out[f"{tag}_timestamps"].tolist())
    np.savez_compressed(os.path.join(GOLD, "fsmn.npz"), **out)


# ------------------------------------------------------------------------------ MarbleNet
def marblenet_reference(cfg, weights, optimized=True, in_sample_rate=16000):
    """The reference's OWN wrapper (+ its own BatchNorm folding) around the NeMo-shaped stand-in
    (oracle/marblenet.py) carrying our seeded weights."""
    from oracle import marblenet as OM
    stft = RL.import_file("NVIDIA_Frame_VAD_Multilingual_MarbleNet/STFT_Process.py", "STFT_Process")
    ns = RL.extract("NVIDIA_Frame_VAD_Multilingual_MarbleNet/Export_NVIDIA_MarbleNet_VAD.py",
                    {"STFT_Process": stft.STFT_Process})
    custom_stft = stft.STFT_Process(model_type='stft_B', n_fft=cfg.n_fft, hop_len=cfg.hop, win_length=cfg.win_length,
                                    max_frames=0, window_type=cfg.window, center_pad=True, pad_mode='constant').eval()
    net = OM.StandIn(cfg, weights)
    cls = ns["NVIDIA_VAD_Optimized"] if optimized else ns["NVIDIA_VAD_Reference"]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(net, custom_stft, cfg.n_fft, cfg.n_mels, 16000, cfg.pre_emphasis, in_sample_rate).eval()
    return m


def marblenet_fake_session(cfg, weights):
    from oracle import ref_runner as RR
    m = marblenet_reference(cfg, weights, optimized=True)
    ins = [RR.NodeArg("audio", [1, 1, "audio_len"], "tensor(int16)")]
    outs = [RR.NodeArg("score_silence", [1, "signal_len", 1], "tensor(float)"),
            RR.NodeArg("score_active", [1, "signal_len", 1], "tensor(float)"),
            RR.NodeArg("signal_len", [1], "tensor(int32)")]

    def fn(feed):
        with torch.inference_mode():
            r = m(torch.from_numpy(feed["audio"]))
        return [np.asarray(t.numpy()) for t in r]

    return RR.FakeSession(ins, outs, fn)


def gen_marblenet_rates():
    """IN_SAMPLE_RATE != 16000: the MarbleNet wrapper's in-graph linear resampler (BN-folded wrapper)."""
    import vadx  # noqa: F401
    from vadx import synth, weights as W
    cfg = W.MarbleNetConfig()
    w = W.marblenet_random_init(cfg, 0)
    out = {}
    for rate in (8000, 48000):
        ref = marblenet_reference(cfg, w, optimized=True, in_sample_rate=rate)
        a = synth.synth_streams(2, 2 * rate, seed=rate + 1)          # two seconds at the input rate
        with torch.inference_mode():
            r = [ref(torch.from_numpy(c).view(1, 1, -1)) for c in a]
        out[f"r{rate}_active"] = np.stack([x[1].numpy()[0, :, 0] for x in r])
        out[f"r{rate}_signal_len"] = np.array(int(r[0][2]))
        print(rate, out[f"r{rate}_active"].shape, int(r[0][2]))
    np.savez_compressed(os.path.join(GOLD, "marblenet_rates.npz"), **out)


def gen_marblenet():
    import random
    import vadx  # noqa: F401
    from vadx import audio_io, synth, weights as W
    from oracle import ref_runner as RR

    cfg = W.MarbleNetConfig()
    w = W.marblenet_random_init(cfg, 0)
    out = {}
    # (1) the reference's own validation recipe (seed 1234, randint(-32768, 32767), three lengths)
    ref_m = marblenet_reference(cfg, w, optimized=False)
    opt_m = marblenet_reference(cfg, w, optimized=True)
    random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
    for L in (16000, 48000, 160000):
        a = torch.randint(-32768, 32767, (1, 1, L), dtype=torch.int16)
        with torch.inference_mode():
            rs, ra, rl = ref_m(a)
            os_, oa, ol = opt_m(a)
        assert int(rl) == int(ol)
        print(f"recipe L={L}: ref-vs-optimized max err {float((ra - oa).abs().max()):.2e}, signal_len {int(ol)}")
        out[f"recipe{L}_audio"] = a.numpy()[0, 0]
        out[f"recipe{L}_active"] = oa.numpy()[0, :, 0]
        out[f"recipe{L}_silence"] = os_.numpy()[0, :, 0]
        out[f"recipe{L}_active_unfolded"] = ra.numpy()[0, :, 0]
        out[f"recipe{L}_signal_len"] = np.array(int(ol))
    # (2) enveloped synthetic clips (the bench workload shape, shortened)
    clips = synth.synth_streams(3, 10 * 16000, seed=1234)
    with torch.inference_mode():
        out["synth_active"] = np.stack([opt_m(torch.from_numpy(c).view(1, 1, -1))[1].numpy()[0, :, 0] for c in clips])
    # (3) the unmodified inference script on vad_sample.wav
    wav = os.path.join(RL.REF_ROOT, "NVIDIA_Frame_VAD_Multilingual_MarbleNet", "vad_sample.wav")
    box = {}

    def factory(_p):
        box["s"] = marblenet_fake_session(cfg, w)
        return box["s"]

    ns, files = RR.run_script("NVIDIA_Frame_VAD_Multilingual_MarbleNet/Inference_NVIDIA_MarbleNet_VAD_ONNX.py", factory,
                              lambda p, sr: audio_io.load_wav_int16(os.path.realpath(p), sr), seed=1234,
                              files_to_link={"vad_sample.wav": wav})
    out["sample_probs"] = np.asarray(ns["all_vad_probs"], np.float32)
    out["sample_decisions"] = np.asarray(ns["vad_decisions"], np.int8)
    out["sample_timestamps"] = np.array(ns["timestamps"], np.float64).reshape(-1, 2)
    out["sample_file_second"] = np.array(files["timestamps_second.txt"])
    out["sample_file_indices"] = np.array(files["timestamps_indices.txt"])
    print("vad_sample:", out["sample_probs"].shape, out["sample_timestamps"].tolist())
    np.savez_compressed(os.path.join(GOLD, "marblenet.npz"), **out)


# ------------------------------------------------------------------------------ Silero
def gen_silero():
    """(1) the reference's UNMODIFIED Silero/Inference_Silero_VAD_ONNX.py with its own OnnxWrapper and
    get_speech_timestamps (Silero/modeling_modified/utils_vad.py) -- only the ORT session underneath
    is replaced by the restated v5 network; (2) get_speech_timestamps on synthetic probability tracks
    that exercise the max-speech split paths."""
    import types
    import vadx  # noqa: F401
    from vadx import audio_io, synth, weights as W
    from oracle import ref_runner as RR
    from oracle.silero import SileroNetOracle

    cfg = W.SileroConfig()
    w = W.silero_random_init(cfg, 0)
    net = SileroNetOracle(w, cfg)
    uv = RL.import_file("Silero/modeling_modified/utils_vad.py", "silero_utils_vad_ref")

    def session_factory(_path):
        ins = [RR.NodeArg("input", [None, 576], "tensor(float)"), RR.NodeArg("state", [2, None, 128], "tensor(float)"),
               RR.NodeArg("sr", [], "tensor(int64)")]
        outs = [RR.NodeArg("output", [None, 1], "tensor(float)"), RR.NodeArg("stateN", [2, None, 128], "tensor(float)")]

        def fn(feed):
            o, s = net.step(torch.from_numpy(feed["input"]), torch.from_numpy(feed["state"]))
            return [o.numpy(), s.numpy()]

        return RR.FakeSession(ins, outs, fn)

    fake_ort = RR._fake_onnxruntime(session_factory)
    sv = types.ModuleType("silero_vad")
    sv.get_speech_timestamps = uv.get_speech_timestamps
    box = {}

    def load_silero_vad(onnx=False, opset_version=16, use_cpu=True, path=""):
        saved = sys.modules.get("onnxruntime")
        sys.modules["onnxruntime"] = fake_ort
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                box["m"] = uv.OnnxWrapper(str(path), force_onnx_cpu=use_cpu)
        finally:
            if saved is not None:
                sys.modules["onnxruntime"] = saved
        return box["m"]

    sv.load_silero_vad = load_silero_vad
    wav = os.path.join(RL.REF_ROOT, "Silero", "vad_sample.wav")
    ns, files = RR.run_script("Silero/Inference_Silero_VAD_ONNX.py", session_factory,
                              lambda p, sr: audio_io.load_wav_int16(os.path.realpath(p), sr), seed=1234,
                              files_to_link={"vad_sample.wav": wav}, extra_modules={"silero_vad": sv})
    calls = box["m"].session.calls
    out = {"sample_probs": np.array([c[1][0][0, 0] for c in calls], np.float32),
           "sample_state_last": calls[-1][1][1][:, 0, :],
           "sample_timestamps": np.array(ns["timestamps"], np.float64).reshape(-1, 2),
           "sample_file_second": np.array(files["timestamps_second.txt"]),
           "sample_file_indices": np.array(files["timestamps_indices.txt"])}
    print("silero vad_sample:", len(calls), "windows; probs", float(out["sample_probs"].min()),
          float(out["sample_probs"].max()), "timestamps", out["sample_timestamps"].tolist())

    # (2) get_speech_timestamps on probability tracks (replay model)
    class Replay:
        def __init__(self, probs):
            self.p, self.i = probs, 0

        def reset_states(self):
            self.i = 0

        def __call__(self, chunk, sr):
            v = self.p[self.i]
            self.i += 1
            return torch.tensor([[v]])

    rs = np.random.RandomState(7)
    cases = []
    for n, max_s, min_sil, use_max in [(400, 20, 250, True), (3000, 6, 250, True), (3000, 6, 100, False),
                                       (1200, float("inf"), 250, True), (5, 20, 250, True)]:
        lvl = np.clip(0.5 + np.cumsum(rs.normal(0, 0.06, n)), 0, 1)
        gate = (np.sin(np.arange(n) / rs.uniform(8, 60)) > rs.uniform(-0.6, 0.3)).astype(np.float64)
        p = np.clip(0.1 + 0.85 * gate * (0.5 + 0.5 * lvl) + rs.normal(0, 0.04, n), 0, 1).astype(np.float32)
        cases.append((p, max_s, min_sil, use_max))
    for i, (p, max_s, min_sil, use_max) in enumerate(cases):
        n_samples = len(p) * 512 - int(rs.randint(0, 511))
        audio = torch.zeros(n_samples)
        for sec in (True, False):
            r = uv.get_speech_timestamps(audio, Replay([float(v) for v in p]), threshold=0.5,
                                         max_speech_duration_s=max_s, min_speech_duration_ms=250,
                                         min_silence_duration_ms=min_sil, return_seconds=sec,
                                         use_max_poss_sil_at_max_speech=use_max)
            out[f"ts{i}_{'sec' if sec else 'smp'}"] = np.array([(d["start"], d["end"]) for d in r], np.float64).reshape(-1, 2)
        out[f"ts{i}_probs"] = p
        out[f"ts{i}_params"] = np.array([n_samples, max_s if np.isfinite(max_s) else -1, min_sil, 1 if use_max else 0], np.float64)
        print(f"ts{i}: {len(p)} windows -> {len(out[f'ts{i}_smp'])} segments")
    np.savez_compressed(os.path.join(GOLD, "silero.npz"), **out)


def gen_silero_iterator():
    """The reference's VADIterator (Silero/modeling_modified/utils_vad.py:494-585) driven window by window
    with replayed probabilities: the online start/end events, in samples and in seconds."""
    uv = RL.import_file("Silero/modeling_modified/utils_vad.py", "silero_utils_vad_ref")

    class Replay:
        def __init__(self, probs):
            self.p, self.i = probs, 0

        def reset_states(self):
            self.i = 0

        def __call__(self, chunk, sr):
            v = self.p[self.i]
            self.i += 1
            return torch.tensor([[v]])

    rs = np.random.RandomState(11)
    out = {}
    for i, (n, thr, min_sil, pad) in enumerate([(600, 0.5, 100, 30), (900, 0.4, 250, 0), (300, 0.6, 32, 100)]):
        lvl = np.clip(0.5 + np.cumsum(rs.normal(0, 0.06, n)), 0, 1)
        gate = (np.sin(np.arange(n) / rs.uniform(8, 40)) > rs.uniform(-0.6, 0.3)).astype(np.float64)
        p = np.clip(0.1 + 0.85 * gate * (0.5 + 0.5 * lvl) + rs.normal(0, 0.05, n), 0, 1).astype(np.float32)
        for sec in (False, True):
            it = uv.VADIterator(Replay([float(v) for v in p]), threshold=thr, sampling_rate=16000,
                                min_silence_duration_ms=min_sil, speech_pad_ms=pad)
            ev = []
            for w in range(n):
                r = it(torch.zeros(512), return_seconds=sec, time_resolution=2)
                if r is not None:
                    (k, v), = r.items()
                    ev.append((w, 0 if k == "start" else 1, v))
            out[f"it{i}_{'sec' if sec else 'smp'}"] = np.array(ev, np.float64).reshape(-1, 3)
        out[f"it{i}_probs"] = p
        out[f"it{i}_params"] = np.array([thr, min_sil, pad], np.float64)
        print(f"it{i}: {n} windows -> {len(ev)} events")
    np.savez_compressed(os.path.join(GOLD, "silero_iter.npz"), **out)


# ------------------------------------------------------------------------------ DFSMN AEC-VAD
def dfsmn_aec_reference(cfg, weights, variant="near_and_far_end_audio"):
    """The reference's own DFSMN_VAD wrapper around its own NET / AlphaPredictor / UniDeepFsmn modules
    (variant = the DFSMN sub-directory: near_and_far_end_audio or only_near_end_audio)."""
    import importlib.util
    import types
    stft = RL.import_file(f"DFSMN/{variant}/STFT_Process.py", "STFT_Process")
    pkg = types.ModuleType("dfsmn_ref_pkg")
    pkg.__path__ = []
    sys.modules["dfsmn_ref_pkg"] = pkg
    lb = types.ModuleType("dfsmn_ref_pkg.layer_base")

    class LayerBase(torch.nn.Module):
        def __init__(self):
            super().__init__()

    lb.LayerBase = LayerBase
    lb.to_kaldi_matrix = lb.expect_kaldi_matrix = lb.expect_token_number = lambda *a, **k: None
    sys.modules["dfsmn_ref_pkg.layer_base"] = lb
    spec = importlib.util.spec_from_file_location(
        "dfsmn_ref_pkg.uni_deep_fsmn",
        os.path.join(RL.REF_ROOT, f"DFSMN/{variant}/modeling_modified/uni_deep_fsmn.py"))
    udf = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = udf
    spec.loader.exec_module(udf)
    ns = RL.extract(f"DFSMN/{variant}/Export_DFSMN_VAD.py", {"STFT_Process": stft.STFT_Process})
    tw = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    net = ns["NET"](max_frames=cfg.max_frames)
    missing, unexpected = net.load_state_dict({k[6:]: v for k, v in tw.items() if k.startswith("iccrn.")}, strict=False)
    assert not unexpected and all(("kernel" in m or "basis" in m or "window_sum" in m) for m in missing), (missing, unexpected)
    net = net.float().eval()
    ap = ns["AlphaPredictor"](cfg.alpha_k)
    ap.load_state_dict({k[6:]: v for k, v in tw.items() if k.startswith("alpha.")}, strict=True)

    class MaskNet(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.linear1 = torch.nn.Linear(3 * cfg.n_mels, cfg.mask_hidden)
            self.relu = torch.nn.ReLU()
            self.deepfsmn = torch.nn.Sequential(*[udf.UniDeepFsmn(cfg.mask_hidden, cfg.mask_hidden, cfg.mask_lorder,
                                                                  cfg.mask_inner) for _ in range(cfg.mask_layers)])
            self.linear3 = torch.nn.Linear(cfg.mask_hidden, 1)

    mask = MaskNet()
    mask.load_state_dict({k[5:]: v for k, v in tw.items() if k.startswith("mask.")}, strict=True)
    feature = types.SimpleNamespace(shift=tw["shift"], scale=tw["scale"])
    pipe = types.SimpleNamespace(model=mask.eval(), preprocessor=types.SimpleNamespace(feature=feature))
    mk = lambda n, w_, h: stft.STFT_Process(model_type='stft_B', n_fft=n, hop_len=h, win_length=w_, max_frames=0,
                                            window_type='hamming').eval()
    wrap = ns["DFSMN_VAD"](pipe, net, ap.float().eval(), mk(cfg.n_fft_a, cfg.win_a, cfg.hop_a), mk(640, cfg.win_a, cfg.hop_a),
                           mk(cfg.n_fft_b, cfg.n_fft_b, cfg.hop_b), cfg.n_fft_a, cfg.n_fft_b, cfg.alpha_k, cfg.max_frames,
                           cfg.pre_emphasis, 16000, cfg.n_mels).eval()
    return wrap, net


def gen_dfsmn_aec():
    import vadx  # noqa: F401
    from vadx import synth, weights as W
    cfg = W.DfsmnAecConfig()
    w = W.dfsmn_aec_random_init(cfg, 0)
    wrap, net = dfsmn_aec_reference(cfg, w)
    L = 31841
    rs = np.random.RandomState(5)
    far = synth.synth_streams(2, L, seed=41)
    near_clean = synth.synth_streams(2, L, seed=42)
    # near = own speech bursts + a delayed, attenuated copy of the far end (echo)
    near = np.clip(near_clean.astype(np.float32) + 0.5 * np.roll(far, 240, axis=1).astype(np.float32), -32768, 32767).astype(np.int16)
    out = {"near": near, "far": far}
    for s in range(2):
        with torch.inference_mode():
            p = wrap(torch.from_numpy(near[s]).view(1, 1, -1), torch.from_numpy(far[s]).view(1, 1, -1))
        out[f"probs{s}"] = p.numpy()
        print(f"stream {s}: probs", p.shape, float(p.min()), float(p.max()))
    # ICCRN alone on a seeded input (stage-level pin)
    torch.manual_seed(3)
    x = torch.randn(1, 4, 160, 24) * 0.05
    with torch.inference_mode():
        y, n = net(x)
    out["iccrn_in"] = x.numpy()
    out["iccrn_out"] = y.numpy()[0, 0]
    np.savez_compressed(os.path.join(GOLD, "dfsmn_aec.npz"), **out)


def gen_dfsmn_near():
    """Near-end-only variant (DFSMN/only_near_end_audio): the wrapper's two white-noise buffers are drawn under
    a fixed seed and frozen with the probabilities -- they are constants of the exported graph."""
    import vadx  # noqa: F401
    from vadx import synth, weights as W
    cfg = W.DfsmnAecConfig()
    w = W.dfsmn_aec_random_init(cfg, 0)
    torch.manual_seed(77)
    wrap, _ = dfsmn_aec_reference(cfg, w, variant="only_near_end_audio")
    L = 31841
    near = synth.synth_streams(2, L, seed=43)
    # the wrapper draws its two constant noise buffers with torch.randn at construction (:309-310); replace them
    # by buffers of the same shape / dtype / distribution from a numpy seed, so tests can rebuild them without a
    # megabyte fixture (vadx.weights.dfsmn_near_noise)
    pow_far, far_comp = W.dfsmn_near_noise(cfg, seed=77)
    assert wrap.pow_far_white_noise.shape[2:] == pow_far.shape and wrap.far_comp_white_noise.shape[2:] == far_comp.shape
    wrap.pow_far_white_noise = torch.from_numpy(pow_far).view(wrap.pow_far_white_noise.shape)
    wrap.far_comp_white_noise = torch.from_numpy(far_comp).view(wrap.far_comp_white_noise.shape)
    out = {"near": near, "pow_far_head": pow_far[:2, :3].astype(np.float32), "far_comp_head": far_comp[:, :2, :3].astype(np.float32)}
    # the wrapper's forward only runs under tracing (it calls .unsqueeze on a shape element, :322), which is
    # how the reference itself executes it (torch.onnx.export): trace it, then run the traced module
    with torch.no_grad():
        traced = torch.jit.trace(wrap, (torch.from_numpy(near[0]).view(1, 1, -1),), check_trace=False)
        for s in range(2):
            p = traced(torch.from_numpy(near[s]).view(1, 1, -1))
            out[f"probs{s}"] = p.numpy()
            print(f"near-only stream {s}: probs", p.shape, float(p.min()), float(p.max()))
    np.savez_compressed(os.path.join(GOLD, "dfsmn_near.npz"), **out)


def gen_marblenet_windows():
    """The unmodified MarbleNet inference script in its STATIC-axis mode (a 32000-sample export: vad_sample.wav becomes three
    non-overlapping windows, the last padded with RMS-matched noise, :130-147) -- the same code path that splits recordings
    longer than one hour in the dynamic mode."""
    import vadx  # noqa: F401
    from vadx import audio_io, weights as W
    from oracle import ref_runner as RR
    cfg = W.MarbleNetConfig()
    w = W.marblenet_random_init(cfg, 0)
    wav = os.path.join(RL.REF_ROOT, "NVIDIA_Frame_VAD_Multilingual_MarbleNet", "vad_sample.wav")
    box = {}

    def factory(_p):
        s = marblenet_fake_session(cfg, w)
        s._inputs_meta[0].shape = [1, 1, 32000]
        box["s"] = s
        return s

    ns, files = RR.run_script("NVIDIA_Frame_VAD_Multilingual_MarbleNet/Inference_NVIDIA_MarbleNet_VAD_ONNX.py", factory,
                              lambda p, sr: audio_io.load_wav_int16(os.path.realpath(p), sr), seed=1234,
                              files_to_link={"vad_sample.wav": wav})
    out = {"probs": np.asarray(ns["all_vad_probs"], np.float32), "decisions": np.asarray(ns["vad_decisions"], np.int8),
           "timestamps": np.array(ns["timestamps"], np.float64).reshape(-1, 2),
           "file_second": np.array(files["timestamps_second.txt"]), "file_indices": np.array(files["timestamps_indices.txt"]),
           "n_calls": np.array(len(box["s"].calls)), "window": np.array(32000)}
    print("marblenet static 32000:", len(box["s"].calls), "windows,", out["probs"].shape, out["timestamps"].tolist())
    np.savez_compressed(os.path.join(GOLD, "marblenet_windows.npz"), **out)


def gen_dropin():
    """The session-call transcript of the three unmodified inference scripts (oracle/dropin.py, stage "record")."""
    from oracle import dropin
    out = dropin.record_transcripts()
    for fam in dropin.FAMILIES:
        print(fam, {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith(fam + "_") and "file" not in k})
    np.savez_compressed(os.path.join(GOLD, "dropin_transcript.npz"), **out)


GENERATORS = {"dropin": gen_dropin, "marblenet_windows": gen_marblenet_windows, "dfsmn_near": gen_dfsmn_near, "firered_rates": gen_firered_rates, "marblenet_rates": gen_marblenet_rates, "firered": gen_firered, "firered_script": gen_firered_script, "postproc": gen_postproc, "audio": gen_audio, "fsmn": gen_fsmn,
              "marblenet": gen_marblenet, "silero": gen_silero, "silero_iterator": gen_silero_iterator, "dfsmn_aec": gen_dfsmn_aec}


def main(argv):
    if not RL.reference_available():
        raise SystemExit("needs /root/reference (authoring container only)")
    os.makedirs(GOLD, exist_ok=True)
    todo = argv or list(GENERATORS)
    for name in todo:
        GENERATORS[name]()


if __name__ == "__main__":
    main(sys.argv[1:])
