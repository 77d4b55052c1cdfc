"""Oracle (TEST INFRASTRUCTURE): Silero VAD v5 (16 kHz) network restatement + call-contract wrapper.

PARITY UNPINNED for the network arithmetic: the reference runs the opaque `silero_vad.onnx` from
the unpinned `silero_vad` pip package (Silero/Export_Silero_VAD.py:34,91,143); nothing under
/root/reference contains the graph.  The layout below restates the public v5 16 kHz graph:
right reflect-pad 64 -> conv-STFT (256/128, hann) -> magnitude [129 x 4] -> 4 x (Conv1d k3 + ReLU),
strides 1,2,2,1 -> LSTMCell(128) -> ReLU -> Conv1d 1x1 -> sigmoid.
What the reference DOES pin, and this file follows:
  OnnxWrapper.__call__ / reset_states / audio_forward .. Silero/modeling_modified/utils_vad.py:87-146
  (64-sample context carried between 512-sample windows, state (2,B,128), zero padding of the tail)
get_speech_timestamps (:247-491) is pure Python in the reference and is run UNMODIFIED by
oracle/make_golden.py on recorded probabilities; tests compare the product's port against it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32)))


class SileroNetOracle:
    def __init__(self, weights, cfg):
        self.cfg = cfg
        self.w = {k: _t(v) for k, v in weights.items()}

    @torch.inference_mode()
    def step(self, x, state):
        """x [B, 576] fp32 (context + window), state [2, B, 128] -> (out [B, 1], new state)"""
        c, w = self.cfg, self.w
        x = F.pad(x.unsqueeze(1), (0, c.reflect_pad), mode="reflect")
        y = F.conv1d(x, w["stft.forward_basis_buffer"], stride=c.hop)
        mag = torch.sqrt(y[:, :c.n_bins] ** 2 + y[:, c.n_bins:] ** 2)
        h = mag
        for i, st in enumerate(c.enc_strides):
            h = F.relu(F.conv1d(h, w[f"encoder.{i}.reparam_conv.weight"], w[f"encoder.{i}.reparam_conv.bias"], stride=st,
                                padding=c.enc_kernel // 2))
        xt = h[:, :, 0]
        gates = F.linear(xt, w["decoder.rnn.weight_ih"], w["decoder.rnn.bias_ih"]) + \
            F.linear(state[0], w["decoder.rnn.weight_hh"], w["decoder.rnn.bias_hh"])
        i_, f_, g_, o_ = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(f_) * state[1] + torch.sigmoid(i_) * torch.tanh(g_)
        h_new = torch.sigmoid(o_) * torch.tanh(c_new)
        out = torch.sigmoid(F.conv1d(F.relu(h_new).unsqueeze(-1), w["decoder.decoder.2.weight"], w["decoder.decoder.2.bias"]))
        return out[:, :, 0], torch.stack([h_new, c_new])


class OnnxWrapperOracle:
    """utils_vad.OnnxWrapper with the ORT session replaced by SileroNetOracle.step (16 kHz only)."""

    def __init__(self, net: SileroNetOracle):
        self.net = net
        self.reset_states()

    def reset_states(self, batch_size=1):
        self._state = torch.zeros((2, batch_size, 128)).float()
        self._context = torch.zeros(0)
        self._last_batch_size = 0

    def __call__(self, x, sr: int = 16000):
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.shape[-1] != 512:
            raise ValueError("Provided number of samples is not 512")
        b = x.shape[0]
        if not self._last_batch_size or self._last_batch_size != b:
            self.reset_states(b)
        if not len(self._context):
            self._context = torch.zeros(b, 64)
        x = torch.cat([self._context, x], dim=1)
        out, self._state = self.net.step(x, self._state)
        self._context = x[..., -64:]
        self._last_batch_size = b
        return out

    def audio_forward(self, x, sr: int = 16000):
        if x.dim() == 1:
            x = x.unsqueeze(0)
        self.reset_states()
        if x.shape[1] % 512:
            x = F.pad(x, (0, 512 - x.shape[1] % 512), "constant", value=0.0)
        return torch.cat([self(x[:, i:i + 512], sr) for i in range(0, x.shape[1], 512)], dim=1)
