"""Oracle (TEST INFRASTRUCTURE): torch-CPU fp32 restatement of the dual-input DFSMN AEC-VAD graph.

Pinned by tests/golden/dfsmn_aec.npz, which oracle/make_golden.py produced by running the
reference's own NET (ICCRN) / AlphaPredictor / DFSMN_VAD modules (AST-loaded from
DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py, seeded weights) -- the ICCRN architecture is
vendored in the reference, so its arithmetic IS pinned.  The modelscope mask-net shell (attribute
names .linear1/.relu/.deepfsmn/.linear3 and UniDeepFsmn.compute1 are in the reference; layer sizes
are not) uses the declared default sizes: its sizes are an assumption, its arithmetic is pinned.

Follows DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py:
  CFB.forward :87-93, CepsUnit :96-154, LayerNorm :163-167, NET.istft/forward :226-249,
  CH_LSTM_T :261-267, CH_LSTM_F :278-284, DFSMN_VAD.forward :317-354;
  UniDeepFsmn.compute1: modeling_modified/uni_deep_fsmn.py:313-329.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import frontend as fe


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32)))


def ceps_bases(n_fft: int):
    """CepsUnit.__init__ (:104-131): rectangular-window DFT kernels and the pinv inverse basis."""
    half = n_fft // 2
    t = torch.arange(n_fft, dtype=torch.float32).unsqueeze(0)
    f = torch.arange(half + 1, dtype=torch.float32).unsqueeze(1)
    omega = 2 * torch.pi * f * t / n_fft
    cos_k, sin_k = torch.cos(omega), -torch.sin(omega)
    fb = torch.fft.fft(torch.eye(n_fft, dtype=torch.float32))
    fb_ri = torch.vstack([torch.real(fb[:half + 1]), torch.imag(fb[:half + 1])]).float()
    inv = torch.linalg.pinv(fb_ri).T
    return cos_k, sin_k, inv          # [81,160], [81,160], [162,160]


def istft_tables(n_fft: int, hop: int, max_frames: int):
    """NET.__init__ (:186-209)."""
    half = n_fft // 2
    window = torch.hamming_window(n_fft)
    fb = torch.fft.fft(torch.eye(n_fft, dtype=torch.float32))
    fb_ri = torch.vstack([torch.real(fb[:half + 1]), torch.imag(fb[:half + 1])]).float()
    inv = torch.linalg.pinv((fb_ri * n_fft) / hop).T * window.view(1, -1)      # [2*(half+1), n_fft]
    out_len = (max_frames - 1) * hop + n_fft
    wsum = torch.zeros(out_len)
    wsq = window ** 2
    for i in range(max_frames):
        s = i * hop
        n = min(n_fft, out_len - s)
        if n <= 0:
            break
        wsum[s:s + n] += wsq[:n]
    return inv, n_fft / (wsum * hop + 1e-6)


class _Lstm:
    def __init__(self, w, prefix, n_in, hidden, layers=1, bi=False):
        self.m = torch.nn.LSTM(n_in, hidden, num_layers=layers, batch_first=True, bidirectional=bi)
        sd = {k[len(prefix) + 1:]: v for k, v in w.items() if k.startswith(prefix + ".")}
        self.m.load_state_dict(sd, strict=True)
        self.m.eval()

    def __call__(self, x):
        return self.m(x)[0]


class DfsmnAecOracle:
    def __init__(self, weights, cfg):
        self.cfg = cfg
        self.w = w = {k: _t(v) for k, v in weights.items()}
        c = cfg.channels
        self.cos_k, self.sin_k, self.ceps_inv = ceps_bases(cfg.n_bins_b)
        self.inv_basis, self.wsum_inv = istft_tables(cfg.n_fft_b, cfg.hop_b, cfg.max_frames)
        self.kernel_b = fe.stft_kernel(cfg.n_fft_b, cfg.n_fft_b, "hamming", "v1")
        self.kernel_a = fe.stft_kernel(cfg.n_fft_a, cfg.win_a, "hamming", "v1")
        self.bank = _t(fe.torchaudio_bank(cfg.n_fft_a // 2 + 1, 20, 8000, cfg.n_mels, 16000, None, "htk"))
        self.lstm = {"in": _Lstm(w, "iccrn.in_ch_lstm.lstm2", 4, c, bi=True),
                     "mid": _Lstm(w, "iccrn.ch_lstm.lstm2", c, 2 * c, layers=2),
                     "out": _Lstm(w, "iccrn.out_ch_lstm.lstm2", 2 * c, c)}
        for n in [f"cfb_e{i}" for i in range(1, 6)] + [f"cfb_d{i}" for i in range(1, 6)]:
            self.lstm[n] = _Lstm(w, f"iccrn.{n}.ceps_unit.ch_lstm_f.lstm2", 2 * c, c, bi=True)

    # ---- ICCRN pieces; tensors are [1, C, F, T] like the reference ---------------------------
    def _ln(self, x, p):
        mean = x.mean([1, 2], keepdim=True)
        std = x.std([1, 2], keepdim=True)
        return (x - mean) / (std + 1e-6) * self.w[p + ".w"] + self.w[p + ".b"]

    def _lstm_f(self, x, lstm, lw, lb, f):
        y = x.permute(0, 3, 2, 1).contiguous().view(-1, f, x.shape[1])
        y = F.linear(lstm(y), lw, lb)
        return y.view(1, -1, f, lw.shape[0]).permute(0, 3, 2, 1).contiguous()

    def _lstm_t(self, x, lstm, lw, lb, f=160):
        y = x.permute(0, 2, 3, 1).contiguous().view(f, -1, x.shape[1])
        y = F.linear(lstm(y), lw, lb)
        return y.view(1, f, -1, lw.shape[0]).permute(0, 3, 1, 2).contiguous()

    def _ceps(self, x0, name):
        c, nf, cb = self.cfg.channels, self.cfg.n_bins_b, self.cfg.ceps_bins
        p = f"iccrn.{name}.ceps_unit"
        xr = x0.permute(0, 1, 3, 2).contiguous().view(-1, 1, nf)
        re = F.conv1d(xr, self.cos_k.unsqueeze(1), stride=nf).view(1, c, -1, cb).permute(0, 1, 3, 2).contiguous()
        im = F.conv1d(xr, self.sin_k.unsqueeze(1), stride=nf).view(1, c, -1, cb).permute(0, 1, 3, 2).contiguous()
        li = torch.cat([re, im], 1)
        lo = self._lstm_f(self._ln(li, p + ".LN"), self.lstm[name], self.w[p + ".ch_lstm_f.linear.weight"],
                          self.w[p + ".ch_lstm_f.linear.bias"], cb)
        pr, pi = lo[:, :c], lo[:, c:]
        o_re, o_im = pr * re - pi * im, pr * im + pi * re
        inp = torch.cat((o_re.permute(0, 1, 3, 2).contiguous().view(-1, cb, 1),
                         o_im.permute(0, 1, 3, 2).contiguous().view(-1, cb, 1)), dim=1)
        inv = F.conv_transpose1d(inp, self.ceps_inv.unsqueeze(1), stride=nf)
        return inv.view(1, c, -1, nf).permute(0, 1, 3, 2).contiguous()

    def _cfb(self, x, name):
        p = f"iccrn.{name}"
        w = self.w
        g = torch.sigmoid(F.conv2d(self._ln(x, p + ".LN0"), w[p + ".conv_gate.weight"], w[p + ".conv_gate.bias"]))
        xi = F.conv2d(x, w[p + ".conv_input.weight"], w[p + ".conv_input.bias"])
        gx = g * xi
        y = F.conv2d(self._ln(gx, p + ".LN1"), w[p + ".conv.weight"], w[p + ".conv.bias"], padding=(1, 0))
        return y + self._ceps(self._ln(xi - gx, p + ".LN2"), name)

    def iccrn(self, x, trace=None):
        """x [1,4,160,T] -> (echo-estimate waveform [1,1,n], n)"""
        w, cfg = self.w, self.cfg
        e0 = self._lstm_f(x, self.lstm["in"], w["iccrn.in_ch_lstm.linear.weight"], w["iccrn.in_ch_lstm.linear.bias"], 160)
        e0 = F.conv2d(torch.cat([e0, x], 1), w["iccrn.in_conv.weight"], w["iccrn.in_conv.bias"])
        e = [e0]
        for i in range(1, 6):
            e.append(self._cfb(e[-1], f"cfb_e{i}"))
        lo = self._lstm_t(self._ln(e[5], "iccrn.ln"), self.lstm["mid"], w["iccrn.ch_lstm.linear.weight"],
                          w["iccrn.ch_lstm.linear.bias"])
        d = self._cfb(e[5] * lo, "cfb_d5")
        for i in (4, 3, 2, 1):
            d = self._cfb(torch.cat([e[i], d], 1), f"cfb_d{i}")
        d0 = self._lstm_t(torch.cat([e[0], d], 1), self.lstm["out"], w["iccrn.out_ch_lstm.linear.weight"],
                          w["iccrn.out_ch_lstm.linear.bias"])
        out = F.conv2d(torch.cat([d0, d], 1), w["iccrn.out_conv.weight"], w["iccrn.out_conv.bias"])
        if trace is not None:
            trace.update(e0=e[0], e1=e[1], e5=e[5], lstm_out=lo, d1=d, out=out)
        half = cfg.n_fft_b // 2
        inv = F.conv_transpose1d(out.reshape(1, 2 * cfg.n_bins_b, -1), self.inv_basis.unsqueeze(1), stride=cfg.hop_b)
        end = inv.size(-1) - half
        return inv[..., half:end] * self.wsum_inv[half:end], end - half

    # ---- whole graph --------------------------------------------------------------------------
    def _stft_ri(self, x, kernel, hop):
        k = torch.from_numpy(kernel).unsqueeze(1)
        nb = kernel.shape[0] // 2
        y = F.conv1d(F.pad(x, (kernel.shape[1] // 2, kernel.shape[1] // 2)), k, stride=hop)
        return y[:, :nb], y[:, nb:]

    @torch.inference_mode()
    def forward(self, near_i16, far_i16=None, trace=None, far_noise=None):
        """near/far int16 [L] (one stream) -> vad probabilities [T_A].
        Near-end-only variant (DFSMN/only_near_end_audio/Export_DFSMN_VAD.py:319-352): far_i16 is None and
        far_noise = (pow_far [160, max_len, k], far_comp [2, 160, max_len]) -- the two constant white-noise
        buffers the wrapper draws at construction (:309-310) stand in for the far end's power and spectrum."""
        cfg, w = self.cfg, self.w
        near = torch.as_tensor(near_i16).view(1, 1, -1).float() * float(1.0 / 32768.0)
        near = near - near.mean()
        nre, nim = self._stft_ri(near, self.kernel_b, cfg.hop_b)
        mix = torch.cat([nre, nim], 0).unsqueeze(0)                    # [1,2,160,T]
        k = cfg.alpha_k
        T = mix.shape[-1]
        idx = torch.arange(T).unsqueeze(1) + torch.arange(k).unsqueeze(0)
        pad = torch.zeros(1, 2, cfg.n_bins_b, k - 1)
        pm = (torch.cat([pad, mix], -1)[..., idx] ** 2).sum(1, keepdim=True)
        if far_i16 is None:
            pf = torch.as_tensor(far_noise[0]).float()[:, :T].view(1, 1, cfg.n_bins_b, T, k)
            farc = torch.as_tensor(far_noise[1]).float()[..., :T].view(1, 1, 2, cfg.n_bins_b, T)
        else:
            far = torch.as_tensor(far_i16).view(1, 1, -1).float() * float(1.0 / 32768.0)
            far = far - far.mean()
            fre, fim = self._stft_ri(far, self.kernel_b, cfg.hop_b)
            farc = torch.cat([fre, fim], 0).unsqueeze(0)
            pf = (torch.cat([pad, farc], -1)[..., idx] ** 2).sum(1, keepdim=True)
        ci = torch.stack([pf, pm], -1).unsqueeze(1)
        a = F.linear(ci.sum(2, keepdim=True), w["alpha.linear1.weight"], w["alpha.linear1.bias"]).squeeze(-1)
        a = F.linear(a, w["alpha.linear2.weight"], w["alpha.linear2.bias"]).squeeze(-1)
        farc = farc * torch.abs(a)
        x = torch.cat([mix, farc.squeeze(1) if farc.dim() == 5 else farc], 1)
        if trace is not None:
            trace.update(x4=x, alpha=a)
        aec, n = self.iccrn(x, trace)
        near = near[..., :n]
        c = cfg.pre_emphasis
        near = torch.cat([near[:, :, :1], near[:, :, 1:] - c * near[:, :, :-1]], -1)
        aec = torch.cat([aec[:, :, :1], aec[:, :, 1:] - c * aec[:, :, :-1]], -1)
        echo = near - float(cfg.echo_factor) * aec
        feats = []
        for sig in (near, aec, echo):
            re, im = self._stft_ri(sig, self.kernel_a, cfg.hop_a)
            feats.append(torch.matmul(self.bank.unsqueeze(0), re * re + im * im))
        feat = torch.cat(feats, 1).transpose(1, 2).clamp(cfg.log_floor).log()
        shift = (w["shift"] + torch.log(torch.tensor(32768 ** 2, dtype=torch.float32))).view(1, 1, -1)
        feat = (feat + shift) * w["scale"].view(1, 1, -1)
        h = F.relu(F.linear(feat, w["mask.linear1.weight"], w["mask.linear1.bias"]))
        for i in range(cfg.mask_layers):
            p = f"mask.deepfsmn.{i}."
            z = F.linear(F.relu(F.linear(h, w[p + "linear.weight"], w[p + "linear.bias"])), w[p + "project.weight"])
            zt = z.transpose(1, 2)
            conv = F.conv1d(F.pad(zt, (cfg.mask_lorder - 1, 0)), w[p + "conv1.weight"].squeeze(-1), groups=cfg.mask_hidden)
            h = h + (conv + zt).transpose(1, 2)
        if trace is not None:
            trace.update(aec=aec, feat=feat)
        return torch.sigmoid(F.linear(h, w["mask.linear3.weight"], w["mask.linear3.bias"])).reshape(-1)
