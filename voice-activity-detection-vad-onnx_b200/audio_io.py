"""Audio loading equivalent to the reference's pydub call
``AudioSegment.from_file(p).set_channels(1).set_frame_rate(16000).get_array_of_samples()``
(FSMN/Inference_FSMN_VAD_ONNX.py:68, FireRedVAD/Inference_FireRed_ONNX.py:535) for WAV input:
pydub reads PCM with `wave`, down-mixes with audioop.tomono(0.5, 0.5) and resamples with
audioop.ratecv -- the same three stdlib calls are made here (audioop exists up to Python 3.12).
"""
from __future__ import annotations

import wave

import numpy as np


def load_wav_int16(path: str, sample_rate: int = 16000) -> np.ndarray:
    import audioop  # noqa: deprecated in 3.11, removed in 3.13
    with wave.open(path, "rb") as w:
        ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width != 2:
        raw = audioop.lin2lin(raw, width, 2)
        width = 2
    if ch == 2:
        raw = audioop.tomono(raw, width, 0.5, 0.5)
    elif ch != 1:
        raise ValueError(f"{path}: {ch} channels are not supported")
    if sr != sample_rate:
        raw, _ = audioop.ratecv(raw, width, 1, sr, sample_rate, None)
    return np.frombuffer(raw, dtype=np.int16).copy()


def align_non_overlapping(audio: np.ndarray, chunk_len: int, rng: np.random.RandomState | None = None):
    """FireRed / MarbleNet static-axis chunker (FireRedVAD/Inference_FireRed_ONNX.py:547-559):
    non-overlapping windows, the tail padded with RMS-matched Gaussian noise cast to int16.
    The reference draws the noise from the unseeded global numpy generator; pass `rng` to make
    it reproducible.  Returns (chunks [n, chunk_len] int16, original_length)."""
    audio = np.asarray(audio, np.int16).reshape(-1)
    rng = rng if rng is not None else np.random
    n = audio.shape[0]
    if n > chunk_len:
        num = int(np.ceil((n - chunk_len) / chunk_len)) + 1
        pad = (num - 1) * chunk_len + chunk_len - n
        if pad > 0:
            tail = audio[-pad:].astype(np.float32)
            noise = (np.sqrt(np.mean(tail * tail)) * rng.normal(loc=0.0, scale=1.0, size=(pad,))).astype(np.int16)
            audio = np.concatenate((audio, noise))
    elif n < chunk_len:
        af = audio.astype(np.float32)
        noise = (np.sqrt(np.mean(af * af)) * rng.normal(loc=0.0, scale=1.0, size=(chunk_len - n,))).astype(np.int16)
        audio = np.concatenate((audio, noise))
    return audio.reshape(-1, chunk_len), n
