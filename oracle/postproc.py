"""Oracle (TEST INFRASTRUCTURE): restatement of the reference's post-processing state machines.

Pure numpy / Python, sequential on purpose: np.cumsum in float32 is a sequential sum and the
reference's decisions depend on that rounding order.

Follows:
  frame_decisions / segments_from_decisions ... FireRedVAD/Inference_FireRed_ONNX.py:102-305
        (VadPostprocessor.process / decision_to_segment; MarbleNet copy
         NVIDIA_*/Inference_NVIDIA_MarbleNet_VAD_ONNX.py:160-353 differs in frame shift and open tail)
  lookahead_hysteresis_* ...................... FSMN/Inference_FSMN_VAD_ONNX.py:188-234;
                                                DFSMN/near_and_far_end_audio/Inference_DFSMN_VAD_ONNX.py:231-273
  runs_to_timestamps / fuse_timestamps ........ FSMN/Inference_FSMN_VAD_ONNX.py:102-141
  clock_string ................................ FSMN/Inference_FSMN_VAD_ONNX.py:144-153
  valid_frames ................................ FireRedVAD/Inference_FireRed_ONNX.py:84-89
  StreamPost .................................. FireRedVAD/Inference_FireRed_ONNX.py:307-490
        (StreamVadPostprocessor: ring-buffer mean, 1-based frame counter, pad_start, max-speech re-arm,
         the open segment re-reported at the end of every call)
"""
from __future__ import annotations

from datetime import timedelta

import numpy as np

f32 = np.float32


def valid_frames(num_samples: int, win: int = 400, hop: int = 160, in_sr: int = 16000) -> int:
    n = int(num_samples * 16000 / in_sr)
    return 0 if n < win else 1 + (n - win) // hop


def smooth_probs(p: np.ndarray, ws: int) -> np.ndarray:
    """float32 moving average through a sequential float32 running sum."""
    n = p.shape[0]
    if ws <= 1:
        return p
    run = np.zeros(n + 1, f32)
    acc = f32(0.0)
    for i in range(n):
        acc = f32(acc + p[i])
        run[i + 1] = acc
    out = np.empty(n, f32)
    inv = f32(1.0 / ws)
    for i in range(n):
        if i < ws - 1:
            out[i] = f32(run[i + 1] / f32(i + 1))
        else:
            out[i] = f32(f32(run[i + 1] - run[i + 1 - ws]) * inv)
    return out


def frame_decisions(probs, ws, thr, min_speech, max_speech, min_silence, merge_silence, extend=0):
    p = np.asarray(probs, f32)
    n = p.shape[0]
    dec = np.zeros(n, np.int8)
    if n == 0:
        return dec
    ws = max(1, ws)
    sm = smooth_probs(p, ws)
    thr = f32(thr)
    SIL, MAYBE_SP, SP, MAYBE_SIL = 0, 1, 2, 3
    if min_speech <= 0 and min_silence <= 0:
        dec[:] = sm >= thr
    else:
        st, t_sp, t_si = SIL, 0, 0
        for t in range(n):
            hot = bool(sm[t] >= thr)
            if st == SIL:
                if hot:
                    st, t_sp = MAYBE_SP, t
            elif st == MAYBE_SP:
                if not hot:
                    st = SIL
                elif t - t_sp >= min_speech:
                    st = SP
                    dec[t_sp:t] = 1
            elif st == SP:
                if not hot:
                    st, t_si = MAYBE_SIL, t
            else:
                if hot:
                    st = SP
                elif t - t_si >= min_silence:
                    st = SIL
            dec[t] = 1 if st >= SP else 0
    # rising edges move left by ws
    if ws > 1:
        for t in range(1, n):
            if dec[t] == 1 and dec[t - 1] == 0:
                dec[max(0, t - ws):t] = 1
    # short gaps are filled
    if merge_silence > 0:
        gap = -1
        for t in range(1, n):
            a, b = dec[t - 1], dec[t]
            if a == 1 and b == 0 and gap < 0:
                gap = t
            elif a == 0 and b == 1 and gap >= 0:
                if t - gap < merge_silence:
                    dec[gap:t] = 1
                gap = -1
    if extend > 0:
        for order in (range(n), range(n - 1, -1, -1)):
            d = extend + 1
            for t in order:
                if dec[t]:
                    d = 0
                else:
                    d += 1
                    if d <= extend:
                        dec[t] = 1
    # over-long runs are cut at the least likely frame of the back half of each max window
    half = max_speech >> 1
    t = 0
    while t < n:
        if not dec[t]:
            t += 1
            continue
        s = t
        while t < n and dec[t]:
            t += 1
        if t - s > max_speech:
            pos, end = s, t
            while pos + max_speech < end:
                a, b = pos + half, min(pos + max_speech, end)
                if a >= b:
                    break
                cut = a + int(np.argmin(p[a:b]))
                dec[cut] = 0
                pos = cut + 1
    return dec


def segments_from_decisions(dec, frame_shift=0.01, frame_length=0.025, wav_dur=None, tail_adds_length=True):
    """-> list[(start_s, end_s)] with the reference's float32 products and round(,3)."""
    dec = np.asarray(dec, np.int8)
    n = dec.shape[0]
    if n == 0:
        return []
    padded = np.zeros(n + 2, np.int8)
    padded[1:-1] = dec
    d = np.diff(padded)
    st = np.flatnonzero(d == 1).astype(f32)
    en = np.flatnonzero(d == -1).astype(f32)
    if st.shape[0] == 0:
        return []
    fs, fl = f32(frame_shift), f32(frame_length)
    seg = np.empty((st.shape[0], 2), f32)
    seg[:, 0] = st * fs
    seg[:, 1] = en * fs
    if dec[-1] != 0:
        e = n * fs + fl if tail_adds_length else n * fs
        if wav_dur is not None and wav_dur < e:
            e = wav_dur
        seg[-1, 1] = e
    return [(round(a, 3), round(b, 3)) for a, b in seg.tolist()]


# ------------------------------------------------------------------ FSMN / DFSMN look-ahead
def lookahead_hysteresis_flags(chunks, look_backward, speaking_score, silence_score):
    """FSMN flavour: `chunks` is a list of uint8 score arrays (one per window); every window but
    the tail of the last contributes len-look_backward decisions.  Returns list[bool] `silence`."""
    lb = look_backward if look_backward != 0 else 1
    inv = float(1.0 / lb)
    saved, silence = [], True
    score = None
    for score in chunks:
        rng = len(score) - look_backward
        for i in range(rng):
            if silence:
                if score[i] != 0:
                    votes = 1 + sum(1 for j in range(1, lb) if score[i + j] != 0)
                    silence = not (votes * inv >= speaking_score)
            else:
                if score[i] != 1:
                    votes = 1 + sum(1 for j in range(1, lb) if score[i + j] != 1)
                    silence = not (votes * inv <= silence_score)
            saved.append(silence)
    if score is not None:
        for i in range(len(score) - look_backward, len(score)):
            silence = (score[i] == 0) if silence else (score[i] != 1)
            saved.append(silence)
    return saved


def lookahead_hysteresis_probs(chunks, look_backward, speaking_score, silence_score):
    """DFSMN flavour (DFSMN/near_and_far_end_audio/Inference_DFSMN_VAD_ONNX.py:231-273): `chunks` is a list of float32
    probability arrays.  Each frame is tested against SPEAKING_SCORE / SILENCE_SCORE (numpy float32 against a Python
    float: float32 comparison under NEP 50), and so is the vote ratio (Python floats)."""
    lb = look_backward if look_backward != 0 else 1
    inv = float(1.0 / lb)
    sp, si = np.float32(speaking_score), np.float32(silence_score)
    saved, silence = [], True
    pr = None
    for pr in chunks:
        pr = np.asarray(pr, np.float32)
        for i in range(len(pr) - look_backward):
            if silence:
                if pr[i] >= sp:
                    votes = 1 + sum(1 for j in range(1, lb) if pr[i + j] >= sp)
                    silence = not (votes * inv >= speaking_score)
                else:
                    silence = True
            else:
                if pr[i] <= si:
                    votes = 1 + sum(1 for j in range(1, lb) if pr[i + j] <= si)
                    silence = not (votes * inv <= silence_score)
                else:
                    silence = False
            saved.append(silence)
    if pr is not None:
        for i in range(len(pr) - look_backward, len(pr)):
            silence = (not pr[i] >= sp) if silence else bool(pr[i] <= si)
            saved.append(silence)
    return saved


def runs_to_timestamps(silence_flags, frame_duration):
    out, start = [], None
    for i, s in enumerate(silence_flags):
        if s:
            if start is not None:
                out.append((start, i * frame_duration + frame_duration))
                start = None
        elif start is None:
            start = i * frame_duration
    if start is not None:
        out.append((start, len(silence_flags) * frame_duration))
    return out


def fuse_timestamps(ts, fusion_threshold=1.0, min_duration=0.5):
    cur = [(a, b) for a, b in ts if (b - a) >= min_duration]
    for _ in range(2):
        nxt = []
        for a, b in cur:
            if nxt and (a - nxt[-1][1] <= fusion_threshold):
                nxt[-1] = (nxt[-1][0], b)
            else:
                nxt.append((a, b))
        cur = nxt
    return cur


def clock_string(seconds: float) -> str:
    tot = timedelta(seconds=seconds).total_seconds()
    whole = int(tot)
    ms = int((tot - whole) * 1000)
    return f"{whole // 3600:02}:{(whole % 3600) // 60:02}:{whole % 60:02}.{ms:03}"


class StreamPost:
    """Call-by-call streaming segmenter (FireRedVAD/Inference_FireRed_ONNX.py:307-490), pinned by
    tests/golden/firered_script.npz.  `feed(probs)` returns the (start_s, end_s) pairs of that call:
    every segment closed inside the call, then -- if a segment is open when the call ends -- that open
    segment up to the last frame seen (so a chunked caller sees it again, longer, on the next call)."""

    SIL, MAYBE_SP, SP, MAYBE_SIL = 0, 1, 2, 3

    def __init__(self, ws, thr, pad_start, min_speech, max_speech, min_silence, fps=100):
        self.ws = max(1, int(ws))
        self.thr = f32(thr)
        self.pad = max(self.ws, int(pad_start))
        self.min_sp, self.max_sp, self.min_si = int(min_speech), int(max_speech), int(min_silence)
        self.inv_fps = 1.0 / fps
        self.ring = np.zeros(self.ws, f32)
        self.acc = f32(0.0)
        self.head = 0
        self.fill = 0
        self.frame = 0           # 1-based index of the last frame consumed
        self.mode = self.SIL
        self.n_sp = 0
        self.n_si = 0
        self.rearm = False       # the previous frame hit max_speech: a new segment starts on this one
        self.open_from = -1      # 1-based first frame of the open segment, -1 when none
        self.closed_at = -1      # 1-based last frame of the latest closed segment

    def _mean(self, p):
        if self.ws <= 1:
            return p
        gone = self.ring[self.head]
        self.ring[self.head] = p
        self.acc = f32(self.acc + f32(p - gone))
        self.head = (self.head + 1) % self.ws
        self.fill = min(self.fill + 1, self.ws)
        return f32(self.acc / f32(self.fill))

    def feed(self, probs):
        probs = np.asarray(probs, f32)
        if probs.shape[0] == 0:
            return []
        out = []
        for p in probs:
            self.frame += 1
            k = self.frame
            speech = bool(self._mean(p) >= self.thr)
            begin = end = -1
            if self.rearm:
                begin = self.open_from = k
                self.rearm = False
            force_close = False
            if self.mode == self.SIL:
                if speech:
                    self.mode, self.n_sp = self.MAYBE_SP, 1
                else:
                    self.n_si += 1
                    self.n_sp = 0
            elif self.mode == self.MAYBE_SP:
                if speech:
                    self.n_sp += 1
                    if self.n_sp >= self.min_sp:
                        self.mode = self.SP
                        begin = self.open_from = max(1, k - self.n_sp + 1 - self.pad, self.closed_at + 1)
                        self.n_si = 0
                else:
                    self.mode, self.n_si, self.n_sp = self.SIL, 1, 0
            elif self.mode == self.SP:
                self.n_sp += 1
                if speech:
                    self.n_si = 0
                    force_close = self.n_sp >= self.max_sp
                else:
                    self.mode, self.n_si = self.MAYBE_SIL, 1
            else:
                self.n_sp += 1
                if speech:
                    self.mode, self.n_si = self.SP, 0
                    force_close = self.n_sp >= self.max_sp
                else:
                    self.n_si += 1
                    if self.n_si >= self.min_si:
                        self.mode = self.SIL
                        begin, end = self.open_from, k
                        self.open_from, self.closed_at, self.n_sp = -1, k, 0
            if force_close:
                self.rearm = True
                self.n_sp = 0
                begin, end = self.open_from, k
                self.open_from, self.closed_at = -1, k
            if begin > 0 and end > 0:
                out.append((max(0, begin - 1) * self.inv_fps, max(0, end - 1) * self.inv_fps))
        if self.open_from > 0:
            out.append((max(0, self.open_from - 1) * self.inv_fps, (self.frame - 1) * self.inv_fps))
        return out
