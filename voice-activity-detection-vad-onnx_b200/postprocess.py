"""Post-processing: device state machines (through the C ABI) + the host-side formatting that
turns frame-index segments into the reference's seconds / sample indices / text files.

Reference anchors:
  VadPostprocessor.process / decision_to_segment ... FireRedVAD/Inference_FireRed_ONNX.py:102-305
  MarbleNet copy (frame_shift arg, open tail) ....... NVIDIA_Frame_VAD_Multilingual_MarbleNet/
                                                      Inference_NVIDIA_MarbleNet_VAD_ONNX.py:160-353
  format_time / file writing ........................ FireRedVAD/Inference_FireRed_ONNX.py:593-609,
                                                      FSMN/Inference_FSMN_VAD_ONNX.py:144-153,244-258
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from datetime import timedelta

import numpy as np

from . import lib


@dataclass(frozen=True)
class FramePostConfig:
    """Defaults = FireRed VAD (FireRedVAD/Inference_FireRed_ONNX.py:40-53)."""
    smooth_window_size: int = 5
    prob_threshold: float = 0.4
    min_speech_frame: int = 20
    max_speech_frame: int = 2000
    min_silence_frame: int = 20
    merge_silence_frame: int = 5
    extend_speech_frame: int = 0
    frame_shift_s: float = 0.01
    frame_length_s: float = 0.025
    tail_adds_frame_length: bool = True   # FireRed: open tail ends at n*shift + length; MarbleNet: n*shift

    def c_struct(self) -> lib.PostCfg:
        return lib.PostCfg(max(1, self.smooth_window_size), float(np.float32(self.prob_threshold)),
                           self.min_speech_frame, self.max_speech_frame, self.min_silence_frame,
                           self.merge_silence_frame, self.extend_speech_frame)


def postprocess_frames(probs, cfg: FramePostConfig, n_frames=None, max_segments: int | None = None, stream=None):
    """probs: CUDA fp32 [S, T] (row stride may exceed T).  n_frames: optional CUDA int32 [S].
    Returns (decisions int8 [S,T], seg_count int32 [S], segments int32 [S,max_segments,2]) on device."""
    import torch
    if not (torch.is_tensor(probs) and probs.is_cuda and probs.dtype == torch.float32 and probs.dim() == 2):
        raise ValueError("postprocess_frames: probs must be a CUDA fp32 tensor [S, T]")
    if probs.stride(1) != 1:
        raise ValueError("postprocess_frames: probs rows must be contiguous")
    S, T = probs.shape
    if max_segments is None:
        max_segments = T // 2 + 1
    dec = torch.empty((S, max(T, 1)), dtype=torch.int8, device=probs.device)
    cnt = torch.empty((S,), dtype=torch.int32, device=probs.device)
    seg = torch.empty((S, max_segments, 2), dtype=torch.int32, device=probs.device)
    if n_frames is not None and not (n_frames.is_cuda and n_frames.dtype == torch.int32 and n_frames.numel() == S):
        raise ValueError("postprocess_frames: n_frames must be a CUDA int32 tensor [S]")
    c = cfg.c_struct()
    lib.check(lib.load().vadx_postprocess_frames(probs.data_ptr(), probs.stride(0) if S > 0 else T, lib.ptr(n_frames),
                                                 S, T, C.byref(c), dec.data_ptr(), cnt.data_ptr(), seg.data_ptr(),
                                                 max_segments, lib.stream_ptr(stream)))
    return dec[:, :T], cnt, seg


def segments_to_seconds(pairs: np.ndarray, n_frames: int, cfg: FramePostConfig, wav_dur: float | None):
    """(start, end_exclusive) frame pairs of ONE stream -> [(start_s, end_s)] with the reference's
    float32 products and round(, 3)  (decision_to_segment, :146-179)."""
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    if pairs.shape[0] == 0 or n_frames == 0:
        return []
    fs, fl = np.float32(cfg.frame_shift_s), np.float32(cfg.frame_length_s)
    seg = np.empty((pairs.shape[0], 2), np.float32)
    seg[:, 0] = pairs[:, 0].astype(np.float32) * fs
    seg[:, 1] = pairs[:, 1].astype(np.float32) * fs
    if pairs[-1, 1] == n_frames:  # last decision is speech: open tail
        end_time = n_frames * fs + fl if cfg.tail_adds_frame_length else n_frames * fs
        if wav_dur is not None and wav_dur < end_time:
            end_time = wav_dur
        seg[-1, 1] = end_time
    return [(round(a, 3), round(b, 3)) for a, b in seg.tolist()]


def format_time(seconds: float) -> str:
    """hh:mm:ss.mmm with truncated milliseconds (FSMN/Inference_FSMN_VAD_ONNX.py:144-153)."""
    td = timedelta(seconds=seconds)
    total = td.total_seconds()
    whole = int(total)
    ms = int((total - whole) * 1000)
    h, rem = divmod(whole, 3600)
    m, s = divmod(rem, 60)
    return f"{h:02}:{m:02}:{s:02}.{ms:03}"


def timestamp_lines(timestamps, sample_rate: int = 16000):
    """-> (lines_seconds, lines_indices) exactly as the reference writes them."""
    sec = [f"{format_time(a)} --> {format_time(b)}\n" for a, b in timestamps]
    idx = [f"{int(a * sample_rate)} --> {int(b * sample_rate)}\n" for a, b in timestamps]
    return sec, idx


def write_timestamp_files(timestamps, path_second: str, path_indices: str, sample_rate: int = 16000):
    sec, idx = timestamp_lines(timestamps, sample_rate)
    with open(path_second, "w", encoding="UTF-8") as f:
        f.writelines(sec)
    with open(path_indices, "w", encoding="UTF-8") as f:
        f.writelines(idx)
