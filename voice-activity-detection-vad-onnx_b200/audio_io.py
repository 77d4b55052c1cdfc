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


def ingest_pcm16_device(pcm, n_channels: int, in_rate: int, out_rate: int = 16000, n_frames=None, stream=None):
    """Device twin of the loader's down-mix + resample: `pcm` CUDA int16 [S, n_frames * n_channels]
    (interleaved; a 1-D tensor is one stream), `n_frames` optional CUDA int64 [S] valid frames per stream.
    -> (mono CUDA int16 [S, n_out], counts CUDA int64 [S]); bit-identical to audioop.tomono + ratecv."""
    import torch
    from . import lib
    if not (torch.is_tensor(pcm) and pcm.is_cuda and pcm.dtype == torch.int16 and pcm.dim() in (1, 2)):
        raise ValueError("ingest_pcm16_device: pcm must be a CUDA int16 tensor [S, n_frames * n_channels]")
    x = (pcm.reshape(1, -1) if pcm.dim() == 1 else pcm).contiguous()
    S, width = x.shape
    if width % n_channels:
        raise ValueError(f"ingest_pcm16_device: row length {width} is not a multiple of {n_channels} channels")
    n_in = width // n_channels
    L = lib.load()
    n_out = int(L.vadx_ingest_out_frames(n_in, in_rate, out_rate))
    out = torch.empty((S, n_out), dtype=torch.int16, device=x.device)
    cnt = torch.empty((S,), dtype=torch.int64, device=x.device)
    if n_frames is not None and not (n_frames.is_cuda and n_frames.dtype == torch.int64 and n_frames.numel() == S):
        raise ValueError("ingest_pcm16_device: n_frames must be a CUDA int64 tensor [S]")
    lib.check(L.vadx_ingest_pcm16(x.data_ptr(), width, lib.ptr(n_frames), S, n_in, n_channels, in_rate, out_rate,
                                  out.data_ptr(), max(n_out, 0), cnt.data_ptr(), lib.stream_ptr(stream)))
    return out, cnt


def load_wav_int16_device(path: str, sample_rate: int = 16000):
    """load_wav_int16 with the down-mix and the rate conversion on the GPU: only the raw PCM crosses PCIe.
    -> CUDA int16 [n]"""
    import torch
    with wave.open(path, "rb") as w:
        ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width != 2:
        raise ValueError(f"{path}: {8 * width}-bit PCM is not supported by the device loader (use load_wav_int16)")
    pcm = torch.frombuffer(bytearray(raw), dtype=torch.int16).cuda()
    mono, _ = ingest_pcm16_device(pcm, ch, sr, sample_rate)
    return mono[0]


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


def normalize_to_int16(audio: np.ndarray) -> np.ndarray:
    """Peak-normalise to 32767 like the FSMN / DFSMN loaders (FSMN/Inference_FSMN_VAD_ONNX.py:60-63)."""
    audio = np.asarray(audio, np.float32)
    max_val = np.max(np.abs(audio))
    scaling_factor = 32767.0 / max_val if max_val > 0 else 1.0
    return (audio * float(scaling_factor)).astype(np.int16)


def align_overlapping(audio: np.ndarray, chunk_len: int, look_backward_frames: int, frame_len: int,
                      rng: np.random.RandomState | None = None):
    """FSMN / DFSMN chunker (FSMN/Inference_FSMN_VAD_ONNX.py:79-99): windows of chunk_len samples that
    advance by chunk_len - (look_backward+1)*frame_len, tail padded with RMS-matched Gaussian noise
    cast to int16.  Returns (aligned int16 [n], stride, original_length)."""
    audio = np.asarray(audio, np.int16).reshape(-1)
    rng = rng if rng is not None else np.random
    n = audio.shape[0]
    stride = chunk_len - (look_backward_frames + 1) * frame_len
    if stride <= 0:
        raise ValueError(f"chunk of {chunk_len} samples is too short for a look-backward of {look_backward_frames} frames")
    pad = 0
    if n > chunk_len:
        num = int(np.ceil((n - chunk_len) / stride)) + 1
        pad = (num - 1) * stride + chunk_len - n
        tail = audio[-pad:].astype(np.float32)
    elif n < chunk_len:
        pad = chunk_len - n
        tail = audio.astype(np.float32)
    if pad > 0:
        noise = (np.sqrt(np.mean(tail * tail)) * rng.normal(loc=0.0, scale=1.0, size=(pad,))).astype(np.int16)
        audio = np.concatenate((audio, noise))
    return audio, stride, n
