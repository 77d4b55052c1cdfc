"""DFSMN AEC-VAD (near-end + far-end input) -- the B200 twin of
DFSMN/near_and_far_end_audio/{Export,Inference}_DFSMN_VAD*.py.

Graph (Export_DFSMN_VAD.py:317-354): both inputs scaled + DC-removed -> STFT-B (319/160) ->
AlphaPredictor scaling of the far end -> ICCRN echo estimator (freq bi-LSTM, 10 gated conv blocks
with cepstral bi-LSTMs, 2-layer time LSTM bottleneck, time LSTM, ISTFT) -> pre-emphasis, echo =
near - 1.15*aec -> 3 x STFT-A (1024/640/320) power -> HTK mel -> log -> (x+shift)*scale ->
mask-net (linear1, relu, N x UniDeepFsmn, linear3, sigmoid) -> one probability per 20 ms frame.

The whole graph is one vadx_forward of the native "dfsmn_aec" model (csrc/model_dfsmn.cu sequences the
kernels for S streams at once on [stream][frame][bin][channel] activations); this
module re-lays the reference state dict into that model's tensors and mirrors the script's host loop.  Post-processing = the look-ahead hysteresis in probability mode
(Inference_DFSMN_VAD_ONNX.py:231-273), on the device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import audio_io, constants, lib, postprocess as PP, tables, weights as W
from .session import NodeArg

SAMPLE_RATE = 16000
OUTPUT_FRAME_LENGTH = 320
FUSION_THRESHOLD = 0.3
MIN_SPEECH_DURATION = 0.2
SPEAKING_SCORE = 0.5
SILENCE_SCORE = 0.5
LOOK_BACKWARD = 0.3


def _t(a):
    return np.ascontiguousarray(np.asarray(a, np.float32))


class DfsmnAecSession:
    """I/O contract of the reference graph (Export_DFSMN_VAD.py:380-393): near_end_audio, far_end_audio
    int16 (1,1,L) -> vad_results fp32 (T,), T = L // 320 + 1 (100 for L = 31841).

    The whole graph is ONE vadx_forward of the native "dfsmn_aec" model (csrc/model_dfsmn.cu): this class only re-lays
    the reference state dict into the tensors that model declares and owns the I/O buffers."""

    def __init__(self, weights: dict, cfg: W.DfsmnAecConfig = W.DfsmnAecConfig(), chunk_len: int = 31841,
                 tensor_cores: bool = True, far_noise=None):
        """far_noise = (pow_far [F, max_frames, k], far_comp [2, F, max_frames]) selects the near-end-only graph
        (DFSMN/only_near_end_audio/Export_DFSMN_VAD.py:291-352): one input `audio`, the far end replaced by those
        two constant buffers (see weights.dfsmn_near_noise)."""
        import torch
        from .session import _Engine
        self._torch = torch
        lib.load()
        lib.require_device()
        self.cfg, self.chunk_len = cfg, int(chunk_len)
        if (self.chunk_len - 1) % cfg.hop_b != 0:
            raise ValueError("DfsmnAecSession: chunk_len must be 1 + a multiple of 160 (the ISTFT of the echo estimator "
                             "then returns exactly chunk_len samples, like the reference's 31841)")
        self.T_b = (self.chunk_len - 1) // cfg.hop_b + 1
        if self.T_b > cfg.max_frames:
            raise ValueError(f"DfsmnAecSession: {self.T_b} STFT frames exceed max_frames={cfg.max_frames}")
        self.T = self.chunk_len // cfg.hop_a + 1
        spec = W.dfsmn_aec_spec(cfg)
        for name in spec:
            if name not in weights:
                raise KeyError(f"DfsmnAecSession: weight '{name}' missing from the state dict")
            if tuple(np.shape(weights[name])) != tuple(spec[name]):
                raise ValueError(f"DfsmnAecSession: '{name}' has shape {np.shape(weights[name])}, expected {spec[name]}")
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.near_only = far_noise is not None
        _, first_a, _ = tables.interleaved_basis(cfg.n_fft_a, cfg.win_a, "hamming", "v1")
        hp = [cfg.channels, cfg.n_fft_b, cfg.hop_b, cfg.alpha_k, cfg.n_fft_a, cfg.win_a, cfg.hop_a, cfg.n_mels, cfg.mask_hidden,
              cfg.mask_layers, cfg.mask_inner, cfg.mask_lorder, cfg.max_frames, 1 if self.near_only else 0, int(first_a)]
        self._e = _Engine("dfsmn_aec", hp)
        self._e.set_scalar("engine.use_tc", 1.0 if tensor_cores else 0.0)
        self._build_constants({k: _t(v) for k, v in weights.items()})
        if self.near_only:
            pf, fc = (np.ascontiguousarray(np.asarray(a, np.float32)) for a in far_noise)
            want_pf, want_fc = (cfg.n_bins_b, cfg.max_frames, cfg.alpha_k), (2, cfg.n_bins_b, cfg.max_frames)
            if pf.shape != want_pf or fc.shape != want_fc:
                raise ValueError(f"DfsmnAecSession: far_noise shapes {pf.shape}, {fc.shape}; expected {want_pf}, {want_fc}")
            self._e.set_tensor("far.pow", pf)
            self._e.set_tensor("far.comp", fc)
            self._inputs_meta = [NodeArg("audio", [1, 1, self.chunk_len], "tensor(int16)")]
        else:
            self._inputs_meta = [NodeArg("near_end_audio", [1, 1, self.chunk_len], "tensor(int16)"),
                                 NodeArg("far_end_audio", [1, 1, self.chunk_len], "tensor(int16)")]
        self._outputs_meta = [NodeArg("vad_results", [self.T], "tensor(float)")]

    # ------------------------------------------------------------------ constants
    def _put(self, name, arr):
        self._e.set_tensor(name, np.ascontiguousarray(np.asarray(arr, np.float32)))

    def _put_linear(self, name, w, bias=None):
        """w [out, in] in the reference layout; the native model transposes / packs the tensor-core image itself"""
        self._e.set_tensor(name + ".weight", _t(w))
        if bias is not None:
            self._e.set_tensor(name + ".bias", _t(bias))

    def _put_lstm(self, name, w, prefix, layers=1, bi=False):
        for l in range(layers):
            for suf in ([""] + (["_reverse"] if bi else [])):
                for part in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    self._put(f"{name}.{part}_l{l}{suf}", w[f"{prefix}.{part}_l{l}{suf}"])

    def _build_constants(self, w):
        cfg = self.cfg
        F = cfg.n_bins_b
        basis_b, _, _ = tables.interleaved_basis(cfg.n_fft_b, cfg.n_fft_b, "hamming", "v1")
        basis_a, _, _ = tables.interleaved_basis(cfg.n_fft_a, cfg.win_a, "hamming", "v1")
        self._put("basis_b", basis_b)
        self._put("basis_a", basis_a)
        bank = constants.torchaudio_mel_bank(cfg.n_fft_a // 2 + 1, 20.0, 8000.0, cfg.n_mels, 16000, None, "htk").numpy()
        st, ln, mw = tables.sparse_bank(bank)
        self._e.set_tensor("mel_start", st)
        self._e.set_tensor("mel_len", ln)
        self._put("mel_w", mw)
        cos_k, sin_k, ceps_inv = constants.ceps_bases(F)
        self._put_linear("ceps.dft", np.concatenate([cos_k.numpy(), sin_k.numpy()], 0))           # [162, 160]
        self._put_linear("ceps.idft", ceps_inv.numpy().T)                                           # [160, 162]
        inv_basis, wsum_inv = constants.istft_tables(cfg.n_fft_b, cfg.hop_b, cfg.max_frames)
        self._put_linear("istft", inv_basis.numpy().T)                                              # [319, 320]
        self._put("wsum_inv", wsum_inv.numpy())
        self._put("alpha.w2", w["alpha.linear2.weight"].reshape(-1))
        for k, v in (("alpha.w1_far", w["alpha.linear1.weight"][0, 0]), ("alpha.w1_mix", w["alpha.linear1.weight"][0, 1]),
                     ("alpha.b1", w["alpha.linear1.bias"][0]), ("alpha.b2", w["alpha.linear2.bias"][0])):
            self._e.set_scalar(k, float(v))
        for k, v in (("pre_emphasis", cfg.pre_emphasis), ("echo_factor", cfg.echo_factor), ("log_floor", cfg.log_floor)):
            self._e.set_scalar(k, float(v))
        self._put_lstm("in_lstm", w, "iccrn.in_ch_lstm.lstm2", bi=True)
        self._put_linear("in_lstm.linear", w["iccrn.in_ch_lstm.linear.weight"], w["iccrn.in_ch_lstm.linear.bias"])
        self._put_linear("in_conv", w["iccrn.in_conv.weight"][:, :, 0, 0], w["iccrn.in_conv.bias"])
        names = [f"cfb_e{i}" for i in range(1, 6)] + [f"cfb_d{i}" for i in range(5, 0, -1)]
        for n in names:
            p = f"iccrn.{n}."
            self._put_linear(n + ".gate", w[p + "conv_gate.weight"][:, :, 0, 0], w[p + "conv_gate.bias"])
            self._put_linear(n + ".input", w[p + "conv_input.weight"][:, :, 0, 0], w[p + "conv_input.bias"])
            cw = w[p + "conv.weight"][:, :, :, 0]                                                   # [co, ci, 3]
            self._put_linear(n + ".conv", np.ascontiguousarray(cw.transpose(0, 2, 1)).reshape(cw.shape[0], -1),
                             w[p + "conv.bias"])                                                    # K index = j*C + ci
            for lnn in ("LN0", "LN1", "LN2"):
                self._put(f"{n}.{lnn}.w", w[p + lnn + ".w"][0, :, :, 0].T)                          # [F][C]
                self._put(f"{n}.{lnn}.b", w[p + lnn + ".b"][0, :, :, 0].T)
            self._put(n + ".cLN.w", w[p + "ceps_unit.LN.w"][0, :, :, 0].T)                          # [bin][part*C + c]
            self._put(n + ".cLN.b", w[p + "ceps_unit.LN.b"][0, :, :, 0].T)
            self._put_lstm(n + ".clstm", w, p + "ceps_unit.ch_lstm_f.lstm2", bi=True)
            self._put_linear(n + ".clstm.linear", w[p + "ceps_unit.ch_lstm_f.linear.weight"],
                             w[p + "ceps_unit.ch_lstm_f.linear.bias"])
        self._put("ln.w", w["iccrn.ln.w"][0, :, :, 0].T)
        self._put("ln.b", w["iccrn.ln.b"][0, :, :, 0].T)
        self._put_lstm("mid_lstm", w, "iccrn.ch_lstm.lstm2", layers=2)
        self._put_linear("mid_lstm.linear", w["iccrn.ch_lstm.linear.weight"], w["iccrn.ch_lstm.linear.bias"])
        self._put_lstm("out_lstm", w, "iccrn.out_ch_lstm.lstm2")
        self._put_linear("out_lstm.linear", w["iccrn.out_ch_lstm.linear.weight"], w["iccrn.out_ch_lstm.linear.bias"])
        self._put_linear("out_conv", w["iccrn.out_conv.weight"][:, :, 0, 0], w["iccrn.out_conv.bias"])
        # mask-net: (feat + shift') * scale folded into linear1 (shift' = shift + ln(32768^2), :291)
        shift = w["shift"].astype(np.float64) + float(np.float32(np.log(np.float32(32768.0 ** 2))))
        scale = w["scale"].astype(np.float64)
        w1 = w["mask.linear1.weight"].astype(np.float64)
        self._put_linear("mask.linear1", (w1 * scale[None, :]).astype(np.float32),
                         (w["mask.linear1.bias"].astype(np.float64) + w1 @ (shift * scale)).astype(np.float32))
        for i in range(cfg.mask_layers):
            p = f"mask.deepfsmn.{i}."
            self._put_linear(f"mask.{i}.linear", w[p + "linear.weight"], w[p + "linear.bias"])
            self._put_linear(f"mask.{i}.project", w[p + "project.weight"])
            self._put(f"mask.{i}.conv", w[p + "conv1.weight"][:, 0, :, 0])                          # [C][lorder]
        self._put_linear("mask.linear3", w["mask.linear3.weight"], w["mask.linear3.bias"])

    # ------------------------------------------------------------------ whole graph
    def _check_inputs(self, near, far):
        torch = self._torch
        if self.near_only != (far is None):
            raise ValueError("run_batch: the near-end-only graph takes one input, the near+far graph two")
        for a in ((near,) if far is None else (near, far)):
            if not (torch.is_tensor(a) and a.is_cuda and a.dtype == torch.int16 and a.dim() == 2 and a.is_contiguous()):
                raise ValueError("run_batch: near/far must be contiguous CUDA int16 tensors [S, L]")
        if (far is not None and near.shape != far.shape) or near.shape[1] != self.chunk_len:
            raise ValueError(f"InvalidArgument: inputs must both have shape [S, {self.chunk_len}]")

    def run_batch(self, near, far=None, out=None, aec_out=None, x4_override=None, stream=None):
        """near, far: CUDA int16 [S, L] -> probabilities CUDA fp32 [S, T] (far is None for the near-end-only graph).
        aec_out (optional, fp32 [S, L]) receives the echo estimate; x4_override injects the 4-channel ICCRN input
        [S*T_b*F, 4] instead of the computed one (stage-level parity tests).  One vadx_forward, asynchronous."""
        torch = self._torch
        self._check_inputs(near, far)
        S, L = near.shape
        probs = out if out is not None else torch.empty((S, self.T), dtype=torch.float32, device=near.device)
        if x4_override is not None and (x4_override.dtype != torch.float32 or not x4_override.is_contiguous()
                                        or x4_override.numel() != S * self.T_b * self.cfg.n_bins_b * 4):
            raise ValueError("run_batch: x4_override must be contiguous fp32 [S*T_b*F, 4]")
        self._e.forward([near, far, x4_override], [probs, aec_out], [], S, L, stream)
        return probs

    def echo_estimate(self, near, far=None, x4_override=None):
        """near/far CUDA int16 [S, L] -> aec fp32 [S, L]: the echo estimator's output (NET, Export_DFSMN_VAD.py:209-284)"""
        aec = self._torch.empty((near.shape[0], self.chunk_len), dtype=self._torch.float32, device=near.device)
        self.run_batch(near, far, aec_out=aec, x4_override=x4_override)
        return aec

    def run_batch_graph(self, near, far=None):
        """run_batch through a CUDA graph: one call is about a thousand small launches, so at small batch sizes the
        eager path is bound by host-side launch cost; the call is captured once per batch size into static buffers and
        replayed.  Returns a view of the static output (copy it if it must survive the next call)."""
        torch = self._torch
        S = near.shape[0]
        runners = self.__dict__.setdefault("_graph_runners", {})
        r = runners.get(S)
        if r is None:
            r = {"near": torch.empty_like(near), "far": None if far is None else torch.empty_like(far)}
            r["near"].copy_(near)
            if far is not None:
                r["far"].copy_(far)
            self.run_batch(r["near"], r["far"])          # eager once: constants are uploaded before the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                r["out"] = self.run_batch(r["near"], r["far"])
            r["graph"] = g
            runners[S] = r
        r["near"].copy_(near)
        if far is not None:
            r["far"].copy_(far)
        r["graph"].replay()
        return r["out"]

    # ------------------------------------------------------------------ ORT surface
    def get_inputs(self):
        return list(self._inputs_meta)

    def get_outputs(self):
        return list(self._outputs_meta)

    def get_providers(self):
        return ["B200ExecutionProvider"]

    def run(self, output_names, input_feed: dict):
        torch = self._torch
        if output_names is not None and any(n != "vad_results" for n in output_names):
            raise ValueError(f"InvalidArgument: unknown output name in {output_names}")
        names = [i.name for i in self._inputs_meta]
        if set(input_feed) != set(names):
            raise ValueError(f"InvalidArgument: inputs must be {' and '.join(names)}, got {sorted(input_feed)}")
        arrs = []
        for k in names:
            a = input_feed[k]
            if not isinstance(a, np.ndarray) or a.dtype != np.int16 or a.shape != (1, 1, self.chunk_len):
                raise ValueError(f"InvalidArgument: '{k}' must be int16 of shape (1, 1, {self.chunk_len})")
            arrs.append(torch.from_numpy(np.ascontiguousarray(a[0])).cuda())
        return [self.run_batch(arrs[0], arrs[1] if len(arrs) > 1 else None)[0].cpu().numpy()]


@dataclass
class AecVadResult:
    timestamps: list
    saved: np.ndarray
    lines_second: list
    lines_indices: list
    probs: list


def run_streams(session: DfsmnAecSession, near_aligned, far_aligned, stride: int, look_backward_s: float = LOOK_BACKWARD,
                speaking_score: float = SPEAKING_SCORE, silence_score: float = SILENCE_SCORE, keep_trace: bool = False):
    """near/far: CUDA int16 [S, n] chunk-aligned; every window advances by `stride` samples.
    The look-ahead hysteresis (probability mode) and all stream state stay on the device."""
    S, n = near_aligned.shape
    L, T = session.chunk_len, session.T
    lb = int(look_backward_s * SAMPLE_RATE // OUTPUT_FRAME_LENGTH)
    n_windows = (n - L) // stride + 1
    state = PP.HysteresisState(S, n_windows * (T - lb) + lb, near_aligned.device)
    trace = []
    for wdx in range(n_windows):
        s0 = wdx * stride
        probs = session.run_batch(near_aligned[:, s0:s0 + L].contiguous(),
                                  None if far_aligned is None else far_aligned[:, s0:s0 + L].contiguous())
        PP.lookahead_hysteresis(probs, state, lb, speaking_score, silence_score, is_final=(wdx == n_windows - 1))
        if keep_trace:
            trace.append(probs)
    return state, trace


def run_vad(near, far, session: DfsmnAecSession, look_backward_s: float = LOOK_BACKWARD, rng=None,
            save_timestamps_second: str | None = None, save_timestamps_indices: str | None = None,
            keep_trace: bool = False) -> AecVadResult:
    """One near/far pair like the reference script (:124-163,229-298): truncate to the common length,
    peak-normalise each, overlapping windows with RMS-noise tail padding, hysteresis, timestamps.
    far=None with a near-end-only session follows DFSMN/only_near_end_audio/Inference_DFSMN_VAD_ONNX.py
    (:120-145,210-214): the same loop over one recording."""
    import torch
    if isinstance(near, str):
        near = audio_io.load_wav_int16(near, SAMPLE_RATE)
    if isinstance(far, str):
        far = audio_io.load_wav_int16(far, SAMPLE_RATE)
    n = len(near) if far is None else min(len(near), len(far))
    near16 = audio_io.normalize_to_int16(np.asarray(near[:n], np.float32))
    lb = int(look_backward_s * SAMPLE_RATE // OUTPUT_FRAME_LENGTH)
    na, stride, _ = audio_io.align_overlapping(near16, session.chunk_len, lb, OUTPUT_FRAME_LENGTH, rng)
    d_far = None
    if far is not None:
        far16 = audio_io.normalize_to_int16(np.asarray(far[:n], np.float32))
        fa, _, _ = audio_io.align_overlapping(far16, session.chunk_len, lb, OUTPUT_FRAME_LENGTH, rng)
        d_far = torch.from_numpy(fa).cuda().unsqueeze(0)
    state, trace = run_streams(session, torch.from_numpy(na).cuda().unsqueeze(0), d_far, stride, look_backward_s,
                               keep_trace=keep_trace)
    cnt, seg = state.segments()
    n_flags = int(state.n_saved[0].item())
    pairs = PP.take_segments(cnt, seg, 0)
    frame_d = OUTPUT_FRAME_LENGTH / SAMPLE_RATE
    ts = PP.process_timestamps(PP.runs_to_timestamps(pairs, n_flags, frame_d), FUSION_THRESHOLD, MIN_SPEECH_DURATION)
    sec, idx = PP.timestamp_lines(ts, SAMPLE_RATE)
    if save_timestamps_second and save_timestamps_indices:
        PP.write_timestamp_files(ts, save_timestamps_second, save_timestamps_indices, SAMPLE_RATE)
    saved = state.saved[0, :n_flags].cpu().numpy().astype(bool)
    return AecVadResult(ts, saved, sec, idx, [t[0].cpu().numpy() for t in trace])
