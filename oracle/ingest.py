"""Oracle (TEST INFRASTRUCTURE): the reference loader's down-mix and rate conversion.

The reference loads audio with pydub (`AudioSegment.from_file(p).set_channels(1).set_frame_rate(sr)`,
FSMN/Inference_FSMN_VAD_ONNX.py:68, FireRedVAD/Inference_FireRed_ONNX.py:535), which for PCM calls the
CPython stdlib: audioop.tomono(data, 2, 0.5, 0.5) and audioop.ratecv(data, 2, 1, in_rate, out_rate, None)
(pydub 0.25.x audio_segment.py set_channels / set_frame_rate; CPython Modules/audioop.c).  `pydub_chain`
runs exactly those calls (audioop ships with Python <= 3.12); `closed_form` restates ratecv's
sequential recurrence (default weights: 16-bit samples widened to 32 bits, linear interpolation with an
integer phase, the high half taken back with an arithmetic shift = floor) as independent outputs, and is
pinned against audioop by tests/test_oracle_ingest.py.
"""
from __future__ import annotations

import math

import numpy as np


def pydub_chain(pcm: np.ndarray, n_channels: int, in_rate: int, out_rate: int) -> np.ndarray:
    import audioop  # noqa: deprecated in 3.11, removed in 3.13
    raw = np.ascontiguousarray(pcm, np.int16).tobytes()
    if n_channels == 2:
        raw = audioop.tomono(raw, 2, 0.5, 0.5)
    if in_rate != out_rate:
        raw, _ = audioop.ratecv(raw, 2, 1, in_rate, out_rate, None)
    return np.frombuffer(raw, np.int16)


def closed_form(pcm: np.ndarray, n_channels: int, in_rate: int, out_rate: int) -> np.ndarray:
    x = np.asarray(pcm, np.int64)
    m = (x[0::2] + x[1::2]) >> 1 if n_channels == 2 else x          # floor(l*0.5 + r*0.5)
    if in_rate == out_rate or len(m) == 0:
        return m.astype(np.int16)
    g = math.gcd(in_rate, out_rate)
    a, b = in_rate // g, out_rate // g
    k = np.arange(((len(m) - 1) * b) // a + 1, dtype=np.int64)
    n = -((-k * a) // b) + 1                  # inputs consumed when output k is produced
    d = (n - 1) * b - k * a                   # phase in [0, b)
    cur = m[n - 1]
    prev = np.where(n >= 2, m[np.maximum(n - 2, 0)], 0)
    return ((prev * d + cur * (b - d)) // b).astype(np.int16)
