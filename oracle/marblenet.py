"""Oracle (TEST INFRASTRUCTURE): NeMo-shaped stand-in for the MarbleNet encoder/decoder + the
restated wrapper graph.

PARITY UNPINNED for the network arithmetic: `nemo_toolkit` (ConvASREncoder / JasperBlock /
MaskedConv1d / MultiLayerPerceptron, unpinned dependency of the reference) is not in
/root/reference, so the block semantics below are restated from the public NeMo source
(jasper.py: separable = depthwise MaskedConv1d + pointwise MaskedConv1d + BatchNorm1d(eps=1e-3);
ReLU+Dropout between repeats; residual = 1x1 MaskedConv1d + BN added before the block's final ReLU;
'same' padding (dilation*(k-1))//2; seq_len = (L + 2p - d(k-1) - 1)//s + 1).
What IS pinned: the stand-in exposes exactly the attributes the reference touches
(`.encoder.encoder[b].mconv / .res`, `MaskedConv1d.conv`, `.use_mask`, `.decoder`), so
oracle/make_golden.py runs the reference's OWN NVIDIA_VAD_Optimized / NVIDIA_VAD_Reference wrappers
and its OWN BatchNorm-folding code on it (NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:58-335) and
freezes their outputs; tests check this restatement and the CUDA path against those.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import frontend as fe


class MaskedConv1d(nn.Module):
    def __init__(self, cin, cout, k, stride=1, dilation=1, groups=1):
        super().__init__()
        self.use_mask = True
        self.conv = nn.Conv1d(cin, cout, k, stride=stride, padding=(dilation * (k - 1)) // 2, dilation=dilation,
                              groups=groups, bias=False)

    def get_seq_len(self, lens):
        c = self.conv
        return torch.div(lens + 2 * c.padding[0] - c.dilation[0] * (c.kernel_size[0] - 1) - 1, c.stride[0],
                         rounding_mode="trunc") + 1

    def forward(self, x, lens):
        if self.use_mask:
            mask = torch.arange(x.shape[-1])[None, :] < lens[:, None]
            x = x * mask[:, None, :]
        return self.conv(x), self.get_seq_len(lens)


class JasperBlock(nn.Module):
    def __init__(self, cin, blk, bn_eps):
        super().__init__()
        layers, c = [], cin
        for r in range(blk.repeat):
            layers += [MaskedConv1d(c, c, blk.kernel, blk.stride, blk.dilation, groups=c),
                       MaskedConv1d(c, blk.filters, 1), nn.BatchNorm1d(blk.filters, eps=bn_eps)]
            if r < blk.repeat - 1:
                layers += [nn.ReLU(), nn.Dropout(0.0)]
            c = blk.filters
        self.mconv = nn.ModuleList(layers)
        self.res = nn.ModuleList([nn.ModuleList([MaskedConv1d(cin, blk.filters, 1),
                                                 nn.BatchNorm1d(blk.filters, eps=bn_eps)])]) if blk.residual else None
        self.mout = nn.Sequential(nn.ReLU(), nn.Dropout(0.0))

    def forward(self, inp):
        xs, lens_orig = inp
        out, lens = xs[-1], lens_orig
        for l in self.mconv:
            if isinstance(l, MaskedConv1d):
                out, lens = l(out, lens)
            else:
                out = l(out)
        if self.res is not None:
            r = xs[0]
            for l in self.res[0]:
                if isinstance(l, MaskedConv1d):
                    r, _ = l(r, lens_orig)
                else:
                    r = l(r)
            out = out + r
        return [self.mout(out)], lens


class Encoder(nn.Module):
    """forward takes the ([features], length) tuple the reference passes through its patched typecheck
    (NVIDIA_*/modeling_modified/common.py:1136-1145)."""

    def __init__(self, cfg):
        super().__init__()
        blocks, c = [], cfg.feat_in
        for blk in cfg.blocks:
            blocks.append(JasperBlock(c, blk, cfg.bn_eps))
            c = blk.filters
        self.encoder = nn.Sequential(*blocks)

    def forward(self, inp):
        xs, lens = inp
        out, lens = self.encoder((xs, lens))
        return out[-1], lens


class Decoder(nn.Module):
    def __init__(self, hidden, n_classes):
        super().__init__()
        self.layer0 = nn.Linear(hidden, n_classes)

    def forward(self, x):
        return self.layer0(x)


class StandIn(nn.Module):
    def __init__(self, cfg, weights):
        super().__init__()
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg.blocks[-1].filters, cfg.num_classes)
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
        missing, unexpected = self.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
        self.eval()


class MarbleNetOracle:
    """Restated wrapper graph (NVIDIA_VAD_Reference.forward, :312-335) on the un-folded stand-in."""

    def __init__(self, weights, cfg, in_sample_rate: int = 16000):
        self.cfg = cfg
        self.rate_scale = 1.0 / (in_sample_rate / 16000.0)   # in-graph resampler (:180-183, :308-330)
        self.net = StandIn(cfg, weights)
        for b in self.net.encoder.encoder:
            for l in b.mconv:
                l.use_mask = False
            if b.res is not None:
                for l in b.res[0]:
                    l.use_mask = False
        self.kernel = fe.stft_kernel(cfg.n_fft, cfg.win_length, cfg.window, "v2")
        self.bank = torch.from_numpy(fe.torchaudio_bank(cfg.n_fft // 2 + 1, 0, 8000, cfg.n_mels, 16000, "slaney", "slaney"))

    @torch.inference_mode()
    def forward(self, audio_i16):
        """[S,L] int16 -> (score_silence [S,T',1], score_active [S,T',1], signal_len int)"""
        a = audio_i16 if torch.is_tensor(audio_i16) else torch.from_numpy(audio_i16)
        c = self.cfg
        x = a.float()
        if self.rate_scale < 1.0:
            x = torch.nn.functional.interpolate(x.unsqueeze(1), scale_factor=self.rate_scale, mode="linear", align_corners=False)[:, 0]
        x = x * float(1.0 / 32768.0)
        x = torch.cat([x[:, :1], x[:, 1:] - c.pre_emphasis * x[:, :-1]], dim=1)
        if self.rate_scale > 1.0:
            x = torch.nn.functional.interpolate(x.unsqueeze(1), scale_factor=self.rate_scale, mode="linear", align_corners=False)[:, 0]
        p = fe.stft_power(x.unsqueeze(1), self.kernel, c.hop, center_pad=True)
        mel = (torch.matmul(self.bank.unsqueeze(0), p) + c.log_eps).log()
        enc, lens = self.net.encoder(([mel], torch.tensor([mel.shape[-1]] * a.shape[0], dtype=torch.long)))
        score = torch.softmax(self.net.decoder(enc.transpose(1, 2)), dim=-1)
        return score[..., :1], score[..., 1:], int(lens[0]) - 1
