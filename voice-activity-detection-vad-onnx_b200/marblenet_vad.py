"""MarbleNet entry point -- the B200 twin of NVIDIA_Frame_VAD_Multilingual_MarbleNet/
Inference_NVIDIA_MarbleNet_VAD_ONNX.py: raw audio in, speech timestamps out, same two text files.
Config names and defaults follow the reference's module-level constants (:15-30)."""
from __future__ import annotations

import numpy as np

from . import audio_io, postprocess as PP, weights as W
from .firered_vad import VadResult
from .session import MarbleNetSession

IN_SAMPLE_RATE = 16000
OUTPUT_FRAME_LENGTH = 320
OUTPUT_FRAME_SHIFT_S = OUTPUT_FRAME_LENGTH / 16000
SPEAKING_SCORE = 0.5
SMOOTH_WINDOW_SIZE = 3
MIN_SPEECH_FRAME = 10
MAX_SPEECH_FRAME = 1000
MIN_SILENCE_FRAME = 10
MERGE_SILENCE_FRAME = 3
EXTEND_SPEECH_FRAME = 0

POST_DEFAULT = PP.FramePostConfig(SMOOTH_WINDOW_SIZE, SPEAKING_SCORE, MIN_SPEECH_FRAME, MAX_SPEECH_FRAME,
                                  MIN_SILENCE_FRAME, MERGE_SILENCE_FRAME, EXTEND_SPEECH_FRAME, OUTPUT_FRAME_SHIFT_S,
                                  0.025, False)


def normalise_audio(audio: np.ndarray, target_rms: float = 8192.0) -> np.ndarray:
    """Optional loader normalisation (:112-120)."""
    a = audio.astype(np.float32)
    rms = np.sqrt(np.mean(a * a, dtype=np.float32), dtype=np.float32)
    if rms > 0:
        a *= (target_rms / (rms + 1e-7))
        np.clip(a, -32768.0, 32767.0, out=a)
        return a.astype(np.int16)
    return audio


def run_vad_clips(session: MarbleNetSession, clips, post: PP.FramePostConfig = POST_DEFAULT, stream=None):
    """clips: CUDA int16 [S, L] (one window per stream, the dynamic-axis mode of the reference).
    -> (active probs [S, T'-1] view, decisions, seg_count, segments) on the device."""
    scores = session.run_batch(clips, stream=stream)
    n = scores.shape[2] - 1                       # signal_len = T' - 1 (:366-373)
    probs = scores[1, :, :n]
    dec, cnt, seg = PP.postprocess_frames(probs, post, None, stream=stream)
    return probs, dec, cnt, seg


def run_vad_windows(session: MarbleNetSession, windows, post: PP.FramePostConfig = POST_DEFAULT, stream=None):
    """windows: CUDA int16 [S, n_windows, L] -- every stream's recording cut into non-overlapping windows of L samples
    (audio_io.align_non_overlapping; the reference's static-axis mode and its split of recordings longer than one hour,
    :130-147).  All S * n_windows windows go through the network as ONE batch; each window contributes its signal_len
    valid frames, the per-stream concatenation (:358-378) is post-processed on the device.
    -> (probs [S, n_windows * (T'-1)], decisions, seg_count, segments)."""
    S, n_win, L = windows.shape
    scores = session.run_batch(windows.reshape(S * n_win, L), stream=stream)        # [2, S*n_win, T']
    n = scores.shape[2] - 1                                                         # signal_len = T' - 1 (:366-373)
    probs = scores[1, :, :n].reshape(S, n_win * n)
    dec, cnt, seg = PP.postprocess_frames(probs, post, None, stream=stream)
    return probs, dec, cnt, seg


def run_vad(audio, session: MarbleNetSession, post: PP.FramePostConfig = POST_DEFAULT, normalize: bool = False,
            save_timestamps_second: str | None = None, save_timestamps_indices: str | None = None,
            input_audio_length: int | None = None, rng=None) -> VadResult:
    """One recording, the reference's behaviour (:121-147, :355-403).  input_audio_length = the export's static audio axis
    (the reference reads it from _inputs_meta[0].shape[-1]); None = the dynamic axis: one window of min(one hour, the
    recording).  Longer recordings are cut into non-overlapping windows, the last one padded with RMS-matched noise (`rng`
    makes the padding reproducible), all windows run as one batch and their valid frames are concatenated."""
    import torch
    if isinstance(audio, str):
        audio = audio_io.load_wav_int16(audio, IN_SAMPLE_RATE)
    audio = np.asarray(audio, np.int16)
    if normalize:
        audio = normalise_audio(audio)
    audio_len = len(audio)
    L = int(input_audio_length) if input_audio_length else min(IN_SAMPLE_RATE * 3600, audio_len)
    if audio_len == L:
        d = torch.from_numpy(audio).cuda().unsqueeze(0)
        probs, dec, cnt, seg = run_vad_clips(session, d, post)
    else:
        windows, _ = audio_io.align_non_overlapping(audio, L, rng)
        probs, dec, cnt, seg = run_vad_windows(session, torch.from_numpy(windows).cuda().unsqueeze(0), post)
    n = probs.shape[1]
    pairs = PP.take_segments(cnt, seg, 0)
    ts = PP.segments_to_seconds(pairs, n, post, audio_len / IN_SAMPLE_RATE)
    sec, idx = PP.timestamp_lines(ts, IN_SAMPLE_RATE)
    if save_timestamps_second and save_timestamps_indices:
        PP.write_timestamp_files(ts, save_timestamps_second, save_timestamps_indices, IN_SAMPLE_RATE)
    return VadResult(ts, probs[0].cpu().numpy(), dec[0, :n].cpu().numpy(), sec, idx)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="MarbleNet Frame-VAD on B200 (random-init weights unless --weights is given)")
    ap.add_argument("audio")
    ap.add_argument("--weights", help=".npz with the NeMo state_dict")
    ap.add_argument("--out-second", default="./timestamps_second.txt")
    ap.add_argument("--out-indices", default="./timestamps_indices.txt")
    a = ap.parse_args(argv)
    cfg = W.MarbleNetConfig()
    w = dict(np.load(a.weights)) if a.weights else W.marblenet_random_init(cfg, 0)
    r = run_vad(a.audio, MarbleNetSession(w, cfg), save_timestamps_second=a.out_second,
                save_timestamps_indices=a.out_indices)
    print("\nTimestamps in Second:")
    print("".join(r.lines_second), end="")
    print("\nTimestamps in Indices:")
    print("".join(r.lines_indices), end="")


if __name__ == "__main__":
    main()
