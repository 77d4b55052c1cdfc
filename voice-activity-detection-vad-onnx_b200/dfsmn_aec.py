"""DFSMN AEC-VAD (near-end + far-end input) -- the B200 twin of
DFSMN/near_and_far_end_audio/{Export,Inference}_DFSMN_VAD*.py.

Graph (Export_DFSMN_VAD.py:317-354): both inputs scaled + DC-removed -> STFT-B (319/160) ->
AlphaPredictor scaling of the far end -> ICCRN echo estimator (freq bi-LSTM, 10 gated conv blocks
with cepstral bi-LSTMs, 2-layer time LSTM bottleneck, time LSTM, ISTFT) -> pre-emphasis, echo =
near - 1.15*aec -> 3 x STFT-A (1024/640/320) power -> HTK mel -> log -> (x+shift)*scale ->
mask-net (linear1, relu, N x UniDeepFsmn, linear3, sigmoid) -> one probability per 20 ms frame.

Every arithmetic step runs in a libvadx kernel through the C ABI; this module only sequences the
calls for S streams at once on [stream][frame][bin][channel] activations and
owns the re-laid constants.  Post-processing = the look-ahead hysteresis in probability mode
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
    int16 (1,1,L) -> vad_results fp32 (T,), T = L // 320 + 1 (100 for L = 31841)."""

    def __init__(self, weights: dict, cfg: W.DfsmnAecConfig = W.DfsmnAecConfig(), chunk_len: int = 31841,
                 tensor_cores: bool = True, far_noise=None):
        """far_noise = (pow_far [F, max_frames, k], far_comp [2, F, max_frames]) selects the near-end-only graph
        (DFSMN/only_near_end_audio/Export_DFSMN_VAD.py:291-352): one input `audio`, the far end replaced by those
        two constant buffers (see weights.dfsmn_near_noise)."""
        import torch
        self._torch = torch
        self._l = lib.load()
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
        self.use_tc = bool(tensor_cores)
        # The echo estimator is a ~60-layer chain of tiny (20-60 channel) contractions: it stays on the exact
        # fp32 FFMA kernels; the mask-net (128/256 wide) runs on tcgen05.
        self.tc_everywhere = False
        self._c = {}
        self._build_constants({k: _t(v) for k, v in weights.items()})
        self.near_only = far_noise is not None
        if self.near_only:
            pf, fc = (np.ascontiguousarray(np.asarray(a, np.float32)) for a in far_noise)
            want_pf, want_fc = (cfg.n_bins_b, cfg.max_frames, cfg.alpha_k), (2, cfg.n_bins_b, cfg.max_frames)
            if pf.shape != want_pf or fc.shape != want_fc:
                raise ValueError(f"DfsmnAecSession: far_noise shapes {pf.shape}, {fc.shape}; expected {want_pf}, {want_fc}")
            self._c["far.pow"] = torch.from_numpy(pf).to(self.dev)
            self._c["far.comp"] = torch.from_numpy(fc).to(self.dev)
            self._inputs_meta = [NodeArg("audio", [1, 1, self.chunk_len], "tensor(int16)")]
        else:
            self._inputs_meta = [NodeArg("near_end_audio", [1, 1, self.chunk_len], "tensor(int16)"),
                                 NodeArg("far_end_audio", [1, 1, self.chunk_len], "tensor(int16)")]
        self._outputs_meta = [NodeArg("vad_results", [self.T], "tensor(float)")]

    # ------------------------------------------------------------------ constants
    def _put(self, name, arr):
        self._c[name] = self._torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)

    def _put_linear(self, name, w, bias=None):
        """w [out, in] -> transposed, padded device weight (+ tensor-core image when the shape fits)."""
        w = _t(w)
        n_out, n_in = w.shape
        ldw = (n_out + 3) // 4 * 4
        wt = np.zeros((n_in, ldw), np.float32)
        wt[:, :n_out] = w.T
        self._put(name + "#T", wt)
        if self.use_tc and self._l.vadx_tc_supported(n_in, n_out):
            self._put(name + "#TC", lib.pack_weight_tc(w))
        if bias is not None:
            self._put(name + "#b", _t(bias))
        self._c[name + "#shape"] = (n_in, n_out, ldw)

    def _put_lstm(self, name, w, prefix, layers=1, bi=False):
        for l in range(layers):
            for suf in ([""] + (["_reverse"] if bi else [])):
                for part in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    self._put(f"{name}.{part}_l{l}{suf}", w[f"{prefix}.{part}_l{l}{suf}"])

    def _build_constants(self, w):
        cfg = self.cfg
        c, F, cb = cfg.channels, cfg.n_bins_b, cfg.ceps_bins
        basis_b, first_b, _ = tables.interleaved_basis(cfg.n_fft_b, cfg.n_fft_b, "hamming", "v1")
        basis_a, first_a, _ = tables.interleaved_basis(cfg.n_fft_a, cfg.win_a, "hamming", "v1")
        self.first_a = first_a
        self._put("basis_b", basis_b)
        self._put("basis_a", basis_a)
        bank = constants.torchaudio_mel_bank(cfg.n_fft_a // 2 + 1, 20.0, 8000.0, cfg.n_mels, 16000, None, "htk").numpy()
        st, ln, mw = tables.sparse_bank(bank)
        self._put("mel_start", st)
        self._put("mel_len", ln)
        self._put("mel_w", mw)
        self.mel_max = mw.shape[1]
        cos_k, sin_k, ceps_inv = constants.ceps_bases(F)
        self._put_linear("ceps.dft", np.concatenate([cos_k.numpy(), sin_k.numpy()], 0))           # [162, 160]
        self._put_linear("ceps.idft", ceps_inv.numpy().T)                                           # [160, 162]
        inv_basis, wsum_inv = constants.istft_tables(cfg.n_fft_b, cfg.hop_b, cfg.max_frames)
        self._put_linear("istft", inv_basis.numpy().T)                                              # [319, 320]
        self._put("wsum_inv", wsum_inv.numpy())
        self._put("alpha.w2", w["alpha.linear2.weight"].reshape(-1))
        self.alpha = (float(w["alpha.linear1.weight"][0, 0]), float(w["alpha.linear1.weight"][0, 1]),
                      float(w["alpha.linear1.bias"][0]), float(w["alpha.linear2.bias"][0]))
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

    # ------------------------------------------------------------------ kernel wrappers
    def _s(self):
        return lib.stream_ptr()

    def _new(self, *shape):
        return self._torch.empty(shape, dtype=self._torch.float32, device=self.dev)

    def _lin(self, name, x, ldx, rows, y=None, ldy=None, act=lib.ACT_NONE, res=None, ldr=0):
        n_in, n_out, ldw = self._c[name + "#shape"]
        if y is None:
            y = self._new(rows, n_out)
            ldy = n_out
        bias = self._c.get(name + "#b")
        img = self._c.get(name + "#TC")
        if img is not None and n_out > 8 and (self.tc_everywhere or name.startswith("mask.")):
            lib.check(self._l.vadx_linear_tc_f32(x.data_ptr(), ldx, img.data_ptr(), lib.ptr(bias), lib.ptr(res), ldr,
                                                 y.data_ptr(), ldy, rows, n_in, n_out, act, self._s()))
        else:
            lib.check(self._l.vadx_linear_f32(x.data_ptr(), ldx, self._c[name + "#T"].data_ptr(), ldw, lib.ptr(bias),
                                              lib.ptr(res), ldr, y.data_ptr(), ldy, rows, n_in, n_out, act, self._s()))
        return y

    def _ln(self, x, rows, D, wname):
        out = self._new(rows * D)
        lib.check(self._l.vadx_layernorm_f32(x.data_ptr(), rows, D, self._c[wname + ".w"].data_ptr(),
                                             self._c[wname + ".b"].data_ptr(), 1e-6, out.data_ptr(), self._s()))
        return out

    def _perm(self, x, dims, perm):
        out = self._new(int(np.prod(dims)))
        lib.check(self._l.vadx_permute4_f32(x.data_ptr(), out.data_ptr(), *[int(d) for d in dims], *perm, self._s()))
        return out

    def _ew(self, op, a, lda, b, ldb, rows, cols, out=None, ldo=None, out2=None, ldo2=0, scalar=0.0):
        if out is None:
            out = self._new(rows, cols)
            ldo = cols
        lib.check(self._l.vadx_ew2_f32(op, a.data_ptr(), lda, lib.ptr(b), ldb, out.data_ptr(), ldo, lib.ptr(out2), ldo2,
                                       rows, cols, float(scalar), self._s()))
        return out

    def _lstm(self, name, layer, suffix, x, xo, xi, xs, y, y_off, yo, yi, ys, n_seq, n_inner, L, n_in, H, reverse):
        g = lambda part: self._c[f"{name}.{part}_l{layer}{suffix}"].data_ptr()
        lib.check(self._l.vadx_lstm_seq_f32(x.data_ptr(), xo, xi, xs, y.data_ptr() + 4 * y_off, yo, yi, ys, g("weight_ih"),
                                            g("weight_hh"), g("bias_ih"), g("bias_hh"), n_seq, n_inner, L, n_in, H,
                                            1 if reverse else 0, self._s()))

    def _bilstm_rows(self, name, x, n_seq, L, n_in, H):
        """bi-LSTM over L consecutive rows of n_in features per sequence -> [n_seq*L][2H]"""
        y = self._new(n_seq * L, 2 * H)
        self._lstm(name, 0, "", x, L * n_in, 0, n_in, y, 0, L * 2 * H, 0, 2 * H, n_seq, 1, L, n_in, H, False)
        self._lstm(name, 0, "_reverse", x, L * n_in, 0, n_in, y, H, L * 2 * H, 0, 2 * H, n_seq, 1, L, n_in, H, True)
        return y

    # ------------------------------------------------------------------ ICCRN
    def _cfb(self, x, cin, name, S):
        cfg = self.cfg
        c, F, cb, T = cfg.channels, cfg.n_bins_b, cfg.ceps_bins, self.T_b
        R, B = S * T * F, S * T
        g = self._lin(name + ".gate", self._ln(x, B, F * cin, name + ".LN0"), cin, R, act=lib.ACT_SIGMOID)
        xi = self._lin(name + ".input", x, cin, R)
        d = self._new(R, c)
        gx = self._ew(4, g, c, xi, c, R, c, out2=d, ldo2=c)
        col = self._new(R, 3 * c)
        lib.check(self._l.vadx_im2col_f3_f32(self._ln(gx, B, F * c, name + ".LN1").data_ptr(), col.data_ptr(), B, F, c,
                                             self._s()))
        y1 = self._lin(name + ".conv", col, 3 * c, R)
        # cepstral unit on LN2(xi - gx)
        z = self._perm(self._ln(d, B, F * c, name + ".LN2"), (B, F, c, 1), (0, 2, 1, 3))          # [B][c][F]
        spec = self._lin("ceps.dft", z, F, B * c)                                                  # [B*c][2*cb]
        P = self._perm(spec, (B, c, 2, cb), (0, 3, 2, 1))                                          # [B][cb][2][c]
        Pn = self._ln(P, B, cb * 2 * c, name + ".cLN")
        hseq = self._bilstm_rows(name + ".clstm", Pn, B, cb, 2 * c, c)                             # [B*cb][2c]
        Q = self._lin(name + ".clstm.linear", hseq, 2 * c, B * cb)
        O = self._new(B * cb, 2 * c)
        lib.check(self._l.vadx_ceps_cmul_f32(Q.data_ptr(), P.data_ptr(), O.data_ptr(), B * cb, c, self._s()))
        Ot = self._perm(O, (B, cb, 2, c), (0, 3, 2, 1))                                            # [B][c][2][cb]
        inv = self._lin("ceps.idft", Ot, 2 * cb, B * c)                                            # [B*c][F]
        ceps = self._perm(inv, (B, c, F, 1), (0, 2, 1, 3))                                         # [B][F][c]
        return self._ew(0, y1, c, ceps, c, R, c)

    def _cat(self, a, ca, b, cb_, R):
        out = self._new(R, ca + cb_)
        self._ew(3, a, ca, None, 0, R, ca, out=out, ldo=ca + cb_)
        lib.check(self._l.vadx_ew2_f32(3, b.data_ptr(), cb_, None, 0, out.data_ptr() + 4 * ca, ca + cb_, None, 0, R, cb_,
                                       0.0, self._s()))
        return out

    def echo_estimate(self, near, far, trace=None, x4_override=None):
        """near/far CUDA int16 [S, L] -> aec fp32 [S, L] (x4_override: inject the 4-channel ICCRN input
        [S*T*F][4] directly -- stage-level parity tests)"""
        cfg, l, torch = self.cfg, self._l, self._torch
        S, L = near.shape
        c, F, T = cfg.channels, cfg.n_bins_b, self.T_b
        R, B = S * T * F, S * T
        half = cfg.n_fft_b // 2
        Lp = (half + L + half + cfg.n_fft_b + 3) // 4 * 4
        ri = []
        for a in ((near,) if self.near_only else (near, far)):
            sig = self._new(S, Lp)
            lib.check(l.vadx_prep_audio(a.data_ptr(), lib.DT_I16, S, L, L, 1.0 / 32768.0, 1, 0, 0.0, half, sig.data_ptr(),
                                        Lp, self._s()))
            o = self._new(B, 2 * F)
            lib.check(l.vadx_stft_complex_f32(sig.data_ptr(), Lp, S, T, cfg.hop_b, cfg.n_fft_b, self._c["basis_b"].data_ptr(),
                                              self._c["basis_b"].shape[1], F, o.data_ptr(), 2 * F, self._s()))
            ri.append(o)
        x4 = self._new(R, 4)
        w1f, w1m, b1, b2 = self.alpha
        if self.near_only:
            lib.check(l.vadx_alpha_x4_const_f32(ri[0].data_ptr(), self._c["far.pow"].data_ptr(), self._c["far.comp"].data_ptr(),
                                                cfg.max_frames, S, T, F, cfg.alpha_k, w1f, w1m, b1,
                                                self._c["alpha.w2"].data_ptr(), b2, x4.data_ptr(), None, self._s()))
        else:
            lib.check(l.vadx_alpha_x4_f32(ri[0].data_ptr(), ri[1].data_ptr(), S, T, F, cfg.alpha_k, w1f, w1m, b1,
                                          self._c["alpha.w2"].data_ptr(), b2, x4.data_ptr(), None, self._s()))
        if x4_override is not None:
            x4 = x4_override
        h = self._bilstm_rows("in_lstm", x4, B, F, 4, c)                                           # [R][2c]
        cat = self._new(R, c + 4)
        self._lin("in_lstm.linear", h, 2 * c, R, y=cat, ldy=c + 4)
        lib.check(l.vadx_ew2_f32(3, x4.data_ptr(), 4, None, 0, cat.data_ptr() + 4 * c, c + 4, None, 0, R, 4, 0.0, self._s()))
        e = [self._lin("in_conv", cat, c + 4, R)]
        for i in range(1, 6):
            e.append(self._cfb(e[-1], c, f"cfb_e{i}", S))
        # 2-layer time LSTM over T for every (stream, bin)
        ln = self._ln(e[5], B, F * c, "ln")
        H = 2 * c
        y1 = self._new(R, H)
        self._lstm("mid_lstm", 0, "", ln, T * F * c, c, F * c, y1, 0, T * F * H, H, F * H, S * F, F, T, c, H, False)
        y2 = self._new(R, H)
        self._lstm("mid_lstm", 1, "", y1, T * F * H, H, F * H, y2, 0, T * F * H, H, F * H, S * F, F, T, H, H, False)
        lo = self._lin("mid_lstm.linear", y2, H, R)
        d = self._cfb(self._ew(1, e[5], c, lo, c, R, c), c, "cfb_d5", S)
        for i in (4, 3, 2, 1):
            d = self._cfb(self._cat(e[i], c, d, c, R), 2 * c, f"cfb_d{i}", S)
        cat2 = self._cat(e[0], c, d, c, R)
        y3 = self._new(R, c)
        self._lstm("out_lstm", 0, "", cat2, T * F * 2 * c, 2 * c, F * 2 * c, y3, 0, T * F * c, c, F * c, S * F, F, T,
                   2 * c, c, False)
        d0 = self._lin("out_lstm.linear", y3, c, R)                                                # [R][2c]
        out = self._lin("out_conv", self._cat(d0, 2 * c, d, c, R), 3 * c, R)                       # [R][2]
        if trace is not None:
            trace.update(x4=x4, e0=e[0], e1=e[1], e5=e[5], lstm_out=lo, d1=d, out=out)
        Y = self._perm(out, (B, F, 2, 1), (0, 2, 1, 3))                                            # [B][2][F]
        frames = self._lin("istft", Y, 2 * F, B)                                                   # [B][319]
        n_out = (T - 1) * cfg.hop_b + cfg.n_fft_b - 2 * half
        aec = self._new(S, n_out)
        lib.check(l.vadx_istft_ola_f32(frames.data_ptr(), cfg.n_fft_b, S, T, cfg.n_fft_b, cfg.hop_b,
                                       self._c["wsum_inv"].data_ptr(), n_out, aec.data_ptr(), n_out, self._s()))
        return aec

    # ------------------------------------------------------------------ whole graph
    def run_batch(self, near, far=None, trace=None):
        """near, far: CUDA int16 [S, L] -> probabilities CUDA fp32 [S, T] (far is None for the near-end-only graph)"""
        torch, cfg, l = self._torch, self.cfg, self._l
        if self.near_only != (far is None):
            raise ValueError("run_batch: the near-end-only graph takes one input, the near+far graph two")
        for a in ((near,) if far is None else (near, far)):
            if not (torch.is_tensor(a) and a.is_cuda and a.dtype == torch.int16 and a.dim() == 2 and a.is_contiguous()):
                raise ValueError("run_batch: near/far must be contiguous CUDA int16 tensors [S, L]")
        if (far is not None and near.shape != far.shape) or near.shape[1] != self.chunk_len:
            raise ValueError(f"InvalidArgument: inputs must both have shape [S, {self.chunk_len}]")
        S, L = near.shape
        aec = self.echo_estimate(near, far, trace)
        n = aec.shape[1]
        assert n == L
        T = self.T
        pad_left = cfg.n_fft_a // 2 - self.first_a
        n_taps = min(cfg.win_a, cfg.n_fft_a)
        Lp = (pad_left + n + n_taps + cfg.hop_a + 3) // 4 * 4
        near_sig, aec_sig = self._new(S, Lp), self._new(S, Lp)
        lib.check(l.vadx_prep_audio(near.data_ptr(), lib.DT_I16, S, L, L, 1.0 / 32768.0, 1, lib.PREEMPH_KEEP_FIRST,
                                    cfg.pre_emphasis, pad_left, near_sig.data_ptr(), Lp, self._s()))
        lib.check(l.vadx_prep_audio(aec.data_ptr(), lib.DT_F32, S, n, n, 1.0, 0, lib.PREEMPH_KEEP_FIRST, cfg.pre_emphasis,
                                    pad_left, aec_sig.data_ptr(), Lp, self._s()))
        echo_sig = self._ew(2, near_sig, Lp, aec_sig, Lp, S, Lp, scalar=cfg.echo_factor)
        rows = S * T
        nb = cfg.n_fft_a // 2 + 1
        ldp = (nb + 1) // 2 * 2
        feat = self._new(rows, 3 * cfg.n_mels)
        power = self._new(rows, ldp)
        for j, sig in enumerate((near_sig, aec_sig, echo_sig)):
            lib.check(l.vadx_stft_power_f32(sig.data_ptr(), Lp, S, T, cfg.hop_a, n_taps, self._c["basis_a"].data_ptr(),
                                            self._c["basis_a"].shape[1], nb, power.data_ptr(), ldp, self._s()))
            lib.check(l.vadx_mel_log_f32(power.data_ptr(), ldp, rows, nb, cfg.n_mels, self._c["mel_start"].data_ptr(),
                                         self._c["mel_len"].data_ptr(), self._c["mel_w"].data_ptr(), self.mel_max,
                                         lib.FLOOR_CLAMP, cfg.log_floor, feat.data_ptr() + 4 * j * cfg.n_mels,
                                         3 * cfg.n_mels, self._s()))
        H = cfg.mask_hidden
        h = self._lin("mask.linear1", feat, 3 * cfg.n_mels, rows, act=lib.ACT_RELU)
        for i in range(cfg.mask_layers):
            z = self._lin(f"mask.{i}.linear", h, H, rows, act=lib.ACT_RELU)
            p = self._lin(f"mask.{i}.project", z, cfg.mask_inner, rows)
            hn = self._new(rows, H)
            lib.check(l.vadx_fsmn_memory_f32(p.data_ptr(), H, self._c[f"mask.{i}.conv"].data_ptr(), cfg.mask_lorder, 1,
                                             None, 0, 1, h.data_ptr(), H, hn.data_ptr(), H, S, T, H, None, None, self._s()))
            h = hn
        probs = self._new(S, T)
        self._lin("mask.linear3", h, H, rows, y=probs, ldy=1, act=lib.ACT_SIGMOID)
        if trace is not None:
            trace.update(aec=aec, feat=feat)
        return probs

    def run_batch_graph(self, near, far=None):
        """run_batch through a CUDA graph: the graph is a ~60-layer chain of small kernels (about a thousand
        launches per call), so the eager path is bound by host-side launch cost; the whole call is captured once
        per batch size into static buffers and replayed.  Returns a view of the static output (copy it if it
        must survive the next call)."""
        torch = self._torch
        S = near.shape[0]
        runners = self.__dict__.setdefault("_graph_runners", {})
        r = runners.get(S)
        if r is None:
            r = {"near": torch.empty_like(near), "far": None if far is None else torch.empty_like(far)}
            r["near"].copy_(near)
            if far is not None:
                r["far"].copy_(far)
            self.run_batch(r["near"], r["far"])          # eager once: every lazily built constant exists before capture
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
