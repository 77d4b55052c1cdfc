"""FireRedVAD entry point -- the B200 twin of FireRedVAD/Inference_FireRed_ONNX.py (RUN_VAD section,
:523-613): raw audio in, speech timestamps (seconds + sample indices) out, same two text files.

Config names and defaults follow the reference's module-level constants (:26-53).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import audio_io, postprocess as PP, weights as W
from .session import FireRedSession

IN_SAMPLE_RATE = 16000
INPUT_AUDIO_LENGTH = 16000           # static export axis of the reference graph
SPEAKING_SCORE = 0.4
SMOOTH_WINDOW_SIZE = 5
MIN_SPEECH_FRAME = 20
MAX_SPEECH_FRAME = 2000
MIN_SILENCE_FRAME = 20
MERGE_SILENCE_FRAME = 5
EXTEND_SPEECH_FRAME = 0

POST_DEFAULT = PP.FramePostConfig(SMOOTH_WINDOW_SIZE, SPEAKING_SCORE, MIN_SPEECH_FRAME, MAX_SPEECH_FRAME,
                                  MIN_SILENCE_FRAME, MERGE_SILENCE_FRAME, EXTEND_SPEECH_FRAME, 0.01, 0.025, True)


def valid_frame_count(num_samples: int, in_sr: int = IN_SAMPLE_RATE) -> int:
    """FireRedVAD/Inference_FireRed_ONNX.py:84-89"""
    n = int(num_samples * 16000 / in_sr)
    return 0 if n < 400 else 1 + (n - 400) // 160


@dataclass
class VadResult:
    timestamps: list            # [(start_s, end_s)]
    probs: np.ndarray           # kept frame probabilities (valid_frame_count of them)
    decisions: np.ndarray       # int8 per kept frame
    lines_second: list
    lines_indices: list


def run_vad_streams(session: FireRedSession, chunks, lengths, post: PP.FramePostConfig = POST_DEFAULT, stream=None,
                    n_valid=None):
    """Batched core: `chunks` CUDA int16 [S, n_chunks, chunk_len] (already aligned/padded),
    `lengths` the original sample counts per stream.  Everything up to the segment frame pairs
    runs on the device.  Returns (probs [S, n_chunks*T] cuda, decisions, seg_count, segments, n_valid)."""
    import torch
    S, n_chunks, L = chunks.shape
    T = session.frames(L)
    probs = session.run_batch(chunks.reshape(S * n_chunks, L), stream=stream)      # [S*n_chunks, odim, T]
    if session.cfg.odim != 1:
        raise ValueError("run_vad_streams expects the VAD head (odim == 1)")
    probs = probs.reshape(S, n_chunks * T)
    if n_valid is None:   # pass a cached CUDA int32 tensor to keep the call free of host->device copies
        n_valid = torch.tensor([min(valid_frame_count(int(n)), n_chunks * T) for n in lengths], dtype=torch.int32,
                               device=chunks.device)
    dec, cnt, seg = PP.postprocess_frames(probs, post, n_valid, stream=stream)
    return probs, dec, cnt, seg, n_valid


class HostBatchPipeline:
    """Host-buffer front door for serving: pinned int16 batches in, segment frame pairs back in pinned
    host memory.  The H2D copy of batch i+1 runs on a copy stream while batch i computes (two device
    input buffers, event hand-over), so a steady stream of batches costs max(copy, compute) per batch
    instead of their sum; results (seg_count, segments) are copied back asynchronously."""

    def __init__(self, session: FireRedSession, n_streams: int, n_chunks: int, post: PP.FramePostConfig = POST_DEFAULT,
                 device=None):
        import torch
        self.session, self.post = session, post
        self.S, self.n_chunks, self.L = n_streams, n_chunks, session.chunk_len
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.T = session.frames(self.L)
        self.max_seg = (n_chunks * self.T) // 2 + 1
        self.d_in = [torch.empty((n_streams, n_chunks, self.L), dtype=torch.int16, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.h_cnt = [torch.empty((n_streams,), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.h_seg = [torch.empty((n_streams, self.max_seg, 2), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.n_valid = torch.full((n_streams,), min(valid_frame_count(n_chunks * self.L), n_chunks * self.T),
                                  dtype=torch.int32, device=dev)
        self._i = 0
        self._prefetched = False

    @property
    def h2d_bytes(self) -> int:
        return self.S * self.n_chunks * self.L * 2

    @property
    def d2h_bytes(self) -> int:
        return self.h_cnt[0].numel() * 4 + self.h_seg[0].numel() * 4

    def prefetch(self, pinned):
        """enqueue the H2D copy of the NEXT batch (returns immediately)"""
        import torch
        b = self._i % 2
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[b])
            self.d_in[b].copy_(pinned.view(self.S, self.n_chunks, self.L), non_blocking=True)
            self.copied[b].record(self.copy_stream)
        self._prefetched = True

    def run(self, pinned, next_pinned=None):
        """process one batch; if `next_pinned` is given its copy overlaps this batch's compute.
        Returns the pinned (seg_count, segments) buffers this batch's results are being copied into
        (valid after the current stream is synchronised)."""
        import torch
        if not self._prefetched:
            self.prefetch(pinned)
        b = self._i % 2
        main = torch.cuda.current_stream()
        main.wait_event(self.copied[b])
        self._i += 1
        self._prefetched = False
        if next_pinned is not None:
            self.prefetch(next_pinned)
        _, _, cnt, seg, _ = run_vad_streams(self.session, self.d_in[b], None, self.post, n_valid=self.n_valid)
        self.consumed[b].record(main)
        self.h_cnt[b].copy_(cnt, non_blocking=True)
        self.h_seg[b].copy_(seg, non_blocking=True)
        return self.h_cnt[b], self.h_seg[b]


def run_vad(audio, session: FireRedSession, post: PP.FramePostConfig = POST_DEFAULT, rng=None,
            save_timestamps_second: str | None = None, save_timestamps_indices: str | None = None) -> VadResult:
    """One stream, the reference's behaviour: `audio` is a path to a wav file or an int16 array."""
    import torch
    if isinstance(audio, str):
        audio = audio_io.load_wav_int16(audio, IN_SAMPLE_RATE)
    chunk_len = session.chunk_len or min(IN_SAMPLE_RATE * 3600, len(audio))
    chunks, audio_len = audio_io.align_non_overlapping(audio, chunk_len, rng)
    d = torch.from_numpy(chunks).cuda().unsqueeze(0)
    probs, dec, cnt, seg, n_valid = run_vad_streams(session, d, [audio_len], post)
    n = int(n_valid[0].item())
    k = int(cnt[0].item())
    pairs = seg[0, :k].cpu().numpy()
    ts = PP.segments_to_seconds(pairs, n, post, audio_len / IN_SAMPLE_RATE)
    sec, idx = PP.timestamp_lines(ts, IN_SAMPLE_RATE)
    if save_timestamps_second and save_timestamps_indices:
        PP.write_timestamp_files(ts, save_timestamps_second, save_timestamps_indices, IN_SAMPLE_RATE)
    return VadResult(ts, probs[0, :n].cpu().numpy(), dec[0, :n].cpu().numpy(), sec, idx)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="FireRedVAD on B200 (random-init weights unless --weights is given)")
    ap.add_argument("audio")
    ap.add_argument("--weights", help=".npz with the DetectModel state_dict")
    ap.add_argument("--out-second", default="./timestamps_second.txt")
    ap.add_argument("--out-indices", default="./timestamps_indices.txt")
    a = ap.parse_args(argv)
    cfg = W.FireRedConfig()
    w = dict(np.load(a.weights)) if a.weights else W.firered_random_init(cfg, 0)
    sess = FireRedSession(w, cfg, chunk_len=INPUT_AUDIO_LENGTH)
    r = run_vad(a.audio, sess, save_timestamps_second=a.out_second, save_timestamps_indices=a.out_indices)
    print("\nTimestamps in Second:")
    print("".join(r.lines_second), end="")
    print("\nTimestamps in Indices:")
    print("".join(r.lines_indices), end="")


if __name__ == "__main__":
    main()
