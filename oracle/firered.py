"""Oracle (TEST INFRASTRUCTURE): torch-CPU fp32 restatement of the FireRedVAD graph.

Pinned by tests/golden/firered_*.npz, which oracle/make_golden.py produced by running the
reference's own FireRedVAD_ONNX / DetectModel modules (AST-loaded from /root/reference)
on the same seeded weights and inputs.

Follows FireRedVAD/Export_FireRedVAD.py:
  FSMN.forward ............ :213-236   (lookback left-pad (N1-1)*S1; lookahead right-pad N2*S2, drop S2)
  DFSMNBlock.forward ...... :253-263
  DFSMN.forward ........... :290-302
  DetectModel.forward ..... :318-326
  FireRedVAD_ONNX.forward . :420-467   (cast, pad(1,0)+conv[-a,1], STFT no-pad, power, mel conv1x1, clamp, log)
  streaming twins ......... :496-515, :596-622
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import frontend as fe


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class FireRedOracle:
    def __init__(self, weights: dict, cfg, in_sample_rate: int = 16000):
        self.cfg = cfg
        # in-graph resampler of the wrapper (:389-393): down before the pre-emphasis, up after it
        self.rate_scale = 1.0 / (in_sample_rate / 16000.0)
        self.w = {k: _t(np.asarray(v, np.float32)) for k, v in weights.items()}
        self.kernel = fe.stft_kernel(cfg.n_fft, cfg.win_length, cfg.window, "v2")
        self.bank = _t(fe.kaldi_like_bank(cfg.n_fft, cfg.n_mels, 16000)).unsqueeze(-1)
        self.pre = torch.tensor([[[-cfg.pre_emphasis, 1.0]]], dtype=torch.float32)

    # -- frontend ---------------------------------------------------------------
    def logmel(self, audio_i16: torch.Tensor) -> torch.Tensor:
        """[N,1,L] int16 -> [N,n_mels,T] log-mel."""
        x = audio_i16.float()
        if self.rate_scale < 1.0:
            x = F.interpolate(x, scale_factor=self.rate_scale, mode="linear", align_corners=False)
        x = F.conv1d(F.pad(x, (1, 0)), self.pre)
        if self.rate_scale > 1.0:
            x = F.interpolate(x, scale_factor=self.rate_scale, mode="linear", align_corners=False)
        p = fe.stft_power(x, self.kernel, self.cfg.hop, center_pad=False)
        m = F.conv1d(p, self.bank)
        return torch.clamp(m, min=self.cfg.log_floor).log()

    # -- memory block -------------------------------------------------------------
    def _memory(self, x, prefix, cache=None):
        c = self.cfg
        wl = self.w[prefix + "lookback_filter.weight"]
        P = x.shape[1]
        pad = (c.N1 - 1) * c.S1
        if cache is None:
            xin = F.pad(x, (pad, 0))
            new_cache = None
        else:
            xin = torch.cat([cache, x], dim=2)
            new_cache = xin[:, :, -pad:]
        mem = x + F.conv1d(xin, wl, dilation=c.S1, groups=P)
        key = prefix + "lookahead_filter.weight"
        if key in self.w and x.shape[2] > 1:
            la = F.conv1d(F.pad(x, (0, c.N2 * c.S2)), self.w[key], dilation=c.S2, groups=P)
            mem = mem + la[:, :, c.S2:]
        return mem, new_cache

    def detect(self, feat: torch.Tensor, caches=None):
        """[N,idim,T] -> probs [N,odim,T] (and new caches [R,N,P,pad] when streaming)."""
        w, c = self.w, self.cfg
        h = F.relu(F.conv1d(feat, w["dfsmn.fc1.0.weight"], w["dfsmn.fc1.0.bias"]))
        p = F.relu(F.conv1d(h, w["dfsmn.fc2.0.weight"], w["dfsmn.fc2.0.bias"]))
        outc = []
        mem, nc = self._memory(p, "dfsmn.fsmn1.", None if caches is None else caches[0])
        outc.append(nc)
        for i in range(c.R - 1):
            pre = f"dfsmn.fsmns.{i}."
            h = F.relu(F.conv1d(mem, w[pre + "fc1.0.weight"], w[pre + "fc1.0.bias"]))
            p = F.conv1d(h, w[pre + "fc2.weight"])
            m2, nc = self._memory(p, pre + "fsmn.", None if caches is None else caches[i + 1])
            outc.append(nc)
            mem = m2 + mem
        h = F.relu(F.conv1d(mem, w["dfsmn.dnns.0.weight"], w["dfsmn.dnns.0.bias"]))
        for j in range(1, c.M):
            h = F.relu(F.conv1d(h, w[f"dfsmn.dnns.{2 * j}.weight"], w[f"dfsmn.dnns.{2 * j}.bias"]))
        probs = torch.sigmoid(F.conv1d(h, w["out.weight"], w["out.bias"]))
        if caches is None:
            return probs
        return probs, torch.stack(outc, 0)

    @torch.inference_mode()
    def forward(self, audio_i16, caches=None):
        a = audio_i16 if torch.is_tensor(audio_i16) else _t(audio_i16)
        if a.dim() == 2:
            a = a.unsqueeze(1)
        feat = self.logmel(a)
        return self.detect(feat, caches)
