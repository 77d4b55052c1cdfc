"""Oracle (TEST INFRASTRUCTURE): torch-CPU fp32 restatement of the FSMN-VAD graph and chunk loop.

Pinned by tests/golden/fsmn.npz, produced by oracle/make_golden.py from (a) the reference's own
FSMN_VAD wrapper + FunASR FSMN encoder modules with the same seeded weights and (b) the reference's
UNMODIFIED Inference_FSMN_VAD_ONNX.py run under oracle/ref_runner.py.

Follows:
  FSMN_VAD.forward ............ FSMN/Export_FSMN_VAD.py:75-101
  FSMN / BasicBlock / FSMNBlock FSMN/modeling_modified/encoder.py:78-83,108-110,208-217
  chunk loop + hysteresis ..... FSMN/Inference_FSMN_VAD_ONNX.py:79-99,156-234
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import frontend as fe
from . import postproc as OP


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32)))


class FsmnOracle:
    def __init__(self, weights: dict, cfg, input_audio_len: int):
        self.cfg, self.L = cfg, input_audio_len
        self.w = {k: _t(v) for k, v in weights.items()}
        self.kernel = fe.stft_kernel(cfg.n_fft, cfg.win_length, cfg.window, "v1")
        self.bank = _t(fe.torchaudio_bank(cfg.n_fft // 2 + 1, 20, 8000, cfg.n_mels, 16000, None, "htk"))
        self.T = input_audio_len // cfg.hop + 1
        self.inv_ref = float(1.0 / (math.sqrt(input_audio_len) * 2e-5))

    def features(self, audio_i16: torch.Tensor):
        """[S,L] int16 -> (y [S,L] DC-removed pre-emphasised fp32, lfr-cmvn features [S,T,400])."""
        c = self.cfg
        x = audio_i16.float()
        x = x - x.mean(dim=1, keepdim=True)
        y = torch.cat([x[:, :1], x[:, 1:] - c.pre_emphasis * x[:, :-1]], dim=1)
        p = fe.stft_power(y.unsqueeze(1), self.kernel, c.hop, center_pad=True)          # [S,F,T]
        mel = torch.matmul(self.bank.unsqueeze(0), p).transpose(1, 2).clamp(min=c.log_floor).log()  # [S,T,80]
        half = (c.lfr_m - 1) // 2
        padded = torch.cat([mel[:, :1].expand(-1, half, -1), mel], dim=1)
        idx = (torch.arange(0, self.T * c.lfr_n, c.lfr_n).unsqueeze(1) + torch.arange(c.lfr_m)).clamp(max=self.T + half - 1)
        feat = padded[:, idx].reshape(x.shape[0], self.T, -1)
        return y, (feat + self.w["cmvn_means"]) * self.w["cmvn_vars"]

    def encoder(self, feat, caches):
        """feat [S,T,400], caches list of 4 x [S,128,19] -> (P(silence) [S,T], new caches)."""
        w, c = self.w, self.cfg
        h = F.linear(F.linear(feat, w["in_linear1.linear.weight"], w["in_linear1.linear.bias"]),
                     w["in_linear2.linear.weight"], w["in_linear2.linear.bias"]).relu()
        new = []
        for i in range(c.fsmn_layers):
            p = F.linear(h, w[f"fsmn.{i}.linear.linear.weight"])                      # [S,T,128]
            xin = torch.cat([caches[i], p.transpose(1, 2)], dim=2)                    # [S,128,19+T]
            new.append(xin[:, :, -caches[i].shape[2]:])
            mem = p.transpose(1, 2) + F.conv1d(xin, w[f"fsmn.{i}.fsmn_block.conv_left.weight"].squeeze(-1),
                                               dilation=c.lstride, groups=c.proj_dim)
            h = F.linear(mem.transpose(1, 2), w[f"fsmn.{i}.affine.linear.weight"], w[f"fsmn.{i}.affine.linear.bias"]).relu()
        o = F.linear(F.linear(h, w["out_linear1.linear.weight"], w["out_linear1.linear.bias"]),
                     w["out_linear2.linear.weight"], w["out_linear2.linear.bias"])
        return torch.softmax(o, dim=-1)[..., 0], new

    @torch.inference_mode()
    def forward(self, audio_i16, caches, one_minus_thr, noise_db):
        """One chunk for S streams.  audio [S,L] int16; caches 4 x [S,128,19]; noise_db [S].
        -> dict(score u8 [S,T], caches, noisy_dB [S], p_sil [S,T], power_dB [S,T])"""
        a = audio_i16 if torch.is_tensor(audio_i16) else torch.from_numpy(audio_i16)
        c = self.cfg
        y, feat = self.features(a)
        p_sil, new = self.encoder(feat, caches)
        r = c.speech_2_noise_ratio
        score = p_sil + (p_sil.pow(r) if r > 1.0 else (1.0 if r < 1.0 else p_sil))
        frames = (y * self.inv_ref).unfold(1, c.n_fft, c.hop)                          # [S,nE,512]
        pw = torch.log10((frames * frames).sum(-1) + 0.00002)
        pw = torch.cat([pw, pw[:, -1:].expand(-1, self.T - pw.shape[1])], dim=1)
        noise = torch.as_tensor(noise_db, dtype=torch.float32).reshape(-1, 1)
        cond = (score <= float(one_minus_thr)) & (pw >= noise)
        noisy = torch.stack([pw[s][~cond[s]].mean() for s in range(a.shape[0])])
        return {"score": cond.to(torch.uint8), "caches": new, "noisy_dB": noisy, "p_sil": p_sil, "power_dB": pw}


def align_overlapping(audio_i16: np.ndarray, chunk_len: int, look_backward_frames: int, frame_len: int, noise=None):
    """FSMN/Inference_FSMN_VAD_ONNX.py:79-99: overlapping windows, RMS-matched Gaussian tail pad.
    `noise`: optional pre-drawn standard-normal samples (the reference draws from np.random)."""
    a = np.asarray(audio_i16, np.int16).reshape(-1)
    n = a.shape[0]
    stride = chunk_len - (look_backward_frames + 1) * frame_len
    if n > chunk_len:
        num = int(np.ceil((n - chunk_len) / stride)) + 1
        pad = (num - 1) * stride + chunk_len - n
        tail = a[-pad:].astype(np.float32)
    elif n < chunk_len:
        pad = chunk_len - n
        tail = a.astype(np.float32)
    else:
        pad = 0
    if pad > 0:
        z = noise[:pad] if noise is not None else np.random.normal(loc=0.0, scale=1.0, size=(pad,))
        a = np.concatenate((a, (np.sqrt(np.mean(tail * tail)) * z).astype(np.int16)))
    return a, stride


def run_stream(oracle: FsmnOracle, audio_i16: np.ndarray, look_backward_s=0.3, one_minus_thr=1.0, snr_db=10.0,
               noise_init_db=30.0, speaking=0.5, silence_score=0.5, fusion=0.3, min_dur=0.2, noise=None):
    """The reference chunk loop for ONE stream -> dict(saved flags, timestamps, per-chunk traces)."""
    L, hop = oracle.L, 160
    lb = int(look_backward_s * 16000 // hop)
    a, stride = align_overlapping(audio_i16, L, lb, hop, noise)
    caches = [torch.zeros(1, 128, 19) for _ in range(4)]
    noise_avg = np.float32(noise_init_db + snr_db) * np.float32(0.1)
    snr = snr_db * 0.1
    chunks, trace = [], []
    s0 = 0
    while s0 + L <= a.shape[0]:
        r = oracle.forward(a[None, s0:s0 + L], caches, one_minus_thr, [noise_avg])
        caches = r["caches"]
        chunks.append(r["score"][0].numpy())
        nd = r["noisy_dB"][0].numpy()
        trace.append((r["p_sil"][0].numpy(), r["power_dB"][0].numpy(), np.float32(noise_avg), nd))
        if nd > 0.0:
            noise_avg = 0.5 * (noise_avg + nd + snr)
        s0 += stride
    saved = OP.lookahead_hysteresis_flags(chunks, lb, speaking, silence_score)
    ts = OP.fuse_timestamps(OP.runs_to_timestamps(saved, hop / 16000), fusion, min_dur)
    return {"saved": saved, "timestamps": ts, "chunks": chunks, "trace": trace, "aligned": a, "stride": stride}
