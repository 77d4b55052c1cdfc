"""Weight dictionaries: names/shapes follow the reference modules' state_dict keys, so a
user holding the pretrained checkpoint can pass `module.state_dict()` (as numpy) straight in.

There is no network in the build environment, so benchmarks and parity tests use
seeded random-init weights of the named architecture (`random_init`).  The generator is
numpy's legacy RandomState (bit-stable across numpy versions) so the same weights can be
regenerated on any box from (model, hyper-parameters, seed).

Reference anchors for the key names:
  FireRed ... FireRedVAD/Export_FireRedVAD.py:185-326 (DetectModel.state_dict())
  FSMN ...... FSMN/modeling_modified/encoder.py:159-217 (funasr FSMN encoder)
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict

import numpy as np


# ----------------------------------------------------------------------------- FireRed
@dataclasses.dataclass(frozen=True)
class FireRedConfig:
    """DetectModel hyper-parameters (`package["args"]`, FireRedVAD/Export_FireRedVAD.py:310-316)."""
    idim: int = 80
    R: int = 8
    M: int = 1
    H: int = 256
    P: int = 128
    N1: int = 20
    S1: int = 1
    N2: int = 20
    S2: int = 1
    odim: int = 1
    # frontend (FireRedVAD/Export_FireRedVAD.py:38-49)
    n_fft: int = 400
    win_length: int = 400
    hop: int = 160
    n_mels: int = 80
    window: str = "povey"
    pre_emphasis: float = 0.97
    log_floor: float = 1e-7
    streaming: bool = False      # Stream-VAD twin: lookback only, caches carried


def firered_spec(cfg: FireRedConfig) -> "OrderedDict[str, tuple]":
    s: OrderedDict[str, tuple] = OrderedDict()
    s["dfsmn.fc1.0.weight"] = (cfg.H, cfg.idim, 1)
    s["dfsmn.fc1.0.bias"] = (cfg.H,)
    s["dfsmn.fc2.0.weight"] = (cfg.P, cfg.H, 1)
    s["dfsmn.fc2.0.bias"] = (cfg.P,)
    s["dfsmn.fsmn1.lookback_filter.weight"] = (cfg.P, 1, cfg.N1)
    if cfg.N2 > 0 and not cfg.streaming:
        s["dfsmn.fsmn1.lookahead_filter.weight"] = (cfg.P, 1, cfg.N2)
    for i in range(cfg.R - 1):
        p = f"dfsmn.fsmns.{i}."
        s[p + "fc1.0.weight"] = (cfg.H, cfg.P, 1)
        s[p + "fc1.0.bias"] = (cfg.H,)
        s[p + "fc2.weight"] = (cfg.P, cfg.H, 1)
        s[p + "fsmn.lookback_filter.weight"] = (cfg.P, 1, cfg.N1)
        if cfg.N2 > 0 and not cfg.streaming:
            s[p + "fsmn.lookahead_filter.weight"] = (cfg.P, 1, cfg.N2)
    s["dfsmn.dnns.0.weight"] = (cfg.H, cfg.P, 1)
    s["dfsmn.dnns.0.bias"] = (cfg.H,)
    for j in range(1, cfg.M):
        s[f"dfsmn.dnns.{2 * j}.weight"] = (cfg.H, cfg.H, 1)
        s[f"dfsmn.dnns.{2 * j}.bias"] = (cfg.H,)
    s["out.weight"] = (cfg.odim, cfg.H, 1)
    s["out.bias"] = (cfg.odim,)
    return s


def _uniform(rs: np.random.RandomState, shape, bound: float) -> np.ndarray:
    return rs.uniform(-bound, bound, size=shape).astype(np.float32)


def _fan_in(shape) -> int:
    if len(shape) == 1:
        return shape[0]
    n = 1
    for d in shape[1:]:
        n *= d
    return n


def _default_init(rs: np.random.RandomState, spec, gain: float = 1.0) -> "OrderedDict[str, np.ndarray]":
    """Variance-preserving uniform init: weights U(+-gain*sqrt(3/fan_in)) (unit gain, so a deep
    random stack neither dies nor explodes), biases U(+-1/sqrt(fan_in of the matching weight))."""
    out: OrderedDict[str, np.ndarray] = OrderedDict()
    last_fan = 1
    for name, shape in spec.items():
        if name.endswith("bias"):
            out[name] = _uniform(rs, shape, 1.0 / np.sqrt(last_fan))
        else:
            last_fan = _fan_in(shape)
            out[name] = _uniform(rs, shape, gain * np.sqrt(3.0 / last_fan))
    return out


def firered_random_init(cfg: FireRedConfig = FireRedConfig(), seed: int = 0):
    """Seeded random weights + a random CMVN folded into the first layer the way
    DetectModel.from_pretrained does (FireRedVAD/Export_FireRedVAD.py:350-360), so that the
    fold is non-trivial.  The head is rescaled so frame probabilities spread over (0,1) and the
    post-processing state machines actually switch."""
    rs = np.random.RandomState(seed)
    w = _default_init(rs, firered_spec(cfg))
    # CMVN stats typical of log-mel of int16-scaled audio
    means = rs.uniform(13.0, 17.0, size=(cfg.idim,)).astype(np.float32)
    inv_std = rs.uniform(0.35, 0.65, size=(cfg.idim,)).astype(np.float32)
    W = w["dfsmn.fc1.0.weight"][:, :, 0].astype(np.float32)
    b = w["dfsmn.fc1.0.bias"]
    w["dfsmn.fc1.0.weight"] = (W * inv_std[None, :])[:, :, None].astype(np.float32)
    w["dfsmn.fc1.0.bias"] = (b - W @ (means * inv_std)).astype(np.float32)
    # head calibrated (offline, default architecture) so logits straddle the 0.4 threshold
    w["out.weight"] = (w["out.weight"] * 0.3).astype(np.float32)
    w["out.bias"] = (w["out.bias"] + 2.9).astype(np.float32)
    return w


# ----------------------------------------------------------------------------- FSMN (FunASR)
@dataclasses.dataclass(frozen=True)
class FsmnConfig:
    """FunASR `speech_fsmn_vad_zh-cn-16k-common` encoder hyper-parameters.  Only proj_dim=128 and
    lorder=20 are pinned by the reference (cache shape (1,128,19,1), FSMN/Export_FSMN_VAD.py:116);
    the others follow the public FunASR config."""
    input_dim: int = 400
    input_affine_dim: int = 140
    fsmn_layers: int = 4
    linear_dim: int = 250
    proj_dim: int = 128
    lorder: int = 20
    rorder: int = 0
    lstride: int = 1
    rstride: int = 1
    output_affine_dim: int = 140
    output_dim: int = 248
    # frontend (FSMN/Export_FSMN_VAD.py:24-34)
    n_fft: int = 512
    win_length: int = 400
    hop: int = 160
    n_mels: int = 80
    window: str = "hamming"
    pre_emphasis: float = 0.97
    log_floor: float = 1e-5
    lfr_m: int = 5
    lfr_n: int = 1
    speech_2_noise_ratio: float = 1.0


def fsmn_spec(cfg: FsmnConfig) -> "OrderedDict[str, tuple]":
    """state_dict keys of FSMN/modeling_modified/encoder.py:159-217 plus the frontend CMVN
    (`cmvn_means` / `cmvn_vars`, FSMN/Export_FSMN_VAD.py:109-110)."""
    s: OrderedDict[str, tuple] = OrderedDict()
    s["in_linear1.linear.weight"] = (cfg.input_affine_dim, cfg.input_dim)
    s["in_linear1.linear.bias"] = (cfg.input_affine_dim,)
    s["in_linear2.linear.weight"] = (cfg.linear_dim, cfg.input_affine_dim)
    s["in_linear2.linear.bias"] = (cfg.linear_dim,)
    for i in range(cfg.fsmn_layers):
        p = f"fsmn.{i}."
        s[p + "linear.linear.weight"] = (cfg.proj_dim, cfg.linear_dim)
        s[p + "fsmn_block.conv_left.weight"] = (cfg.proj_dim, 1, cfg.lorder, 1)
        s[p + "affine.linear.weight"] = (cfg.linear_dim, cfg.proj_dim)
        s[p + "affine.linear.bias"] = (cfg.linear_dim,)
    s["out_linear1.linear.weight"] = (cfg.output_affine_dim, cfg.linear_dim)
    s["out_linear1.linear.bias"] = (cfg.output_affine_dim,)
    s["out_linear2.linear.weight"] = (cfg.output_dim, cfg.output_affine_dim)
    s["out_linear2.linear.bias"] = (cfg.output_dim,)
    s["cmvn_means"] = (cfg.n_mels * cfg.lfr_m,)
    s["cmvn_vars"] = (cfg.n_mels * cfg.lfr_m,)
    return s


def fsmn_random_init(cfg: FsmnConfig = FsmnConfig(), seed: int = 0):
    """Seeded random weights; CMVN typical of log-mel of peak-normalised int16 audio; the silence
    class (index 0) of the softmax head is boosted so that 2*P(silence) straddles the default
    threshold 1.0 and the look-ahead state machine switches."""
    rs = np.random.RandomState(seed)
    spec = fsmn_spec(cfg)
    net = OrderedDict((k, v) for k, v in spec.items() if not k.startswith("cmvn"))
    w = _default_init(rs, net)
    d = cfg.n_mels * cfg.lfr_m
    w["cmvn_means"] = (-rs.uniform(12.0, 17.0, size=(d,))).astype(np.float32)
    w["cmvn_vars"] = rs.uniform(0.25, 0.45, size=(d,)).astype(np.float32)
    w["out_linear2.linear.weight"][0] *= 4.0
    w["out_linear2.linear.bias"][0] += np.float32(np.log(cfg.output_dim - 1.0))
    return w


# ----------------------------------------------------------------------------- MarbleNet (NeMo)
@dataclasses.dataclass(frozen=True)
class JasperBlockCfg:
    filters: int
    repeat: int
    kernel: int
    stride: int = 1
    dilation: int = 1
    residual: bool = False
    separable: bool = True


@dataclasses.dataclass(frozen=True)
class MarbleNetConfig:
    """NVIDIA Frame-VAD Multilingual MarbleNet v2.0.  The encoder/decoder live in nemo_toolkit, which
    is NOT part of the reference checkout; this is the public `marblenet_3x2x64_20ms` layout
     -- a declared assumption, parity for the network arithmetic is unpinned."""
    feat_in: int = 80
    blocks: tuple = (
        JasperBlockCfg(128, 1, 11, stride=2),
        JasperBlockCfg(64, 2, 13, residual=True),
        JasperBlockCfg(64, 2, 15, residual=True),
        JasperBlockCfg(64, 2, 17, residual=True),
        JasperBlockCfg(128, 1, 29, dilation=2),
        JasperBlockCfg(128, 1, 1),
    )
    num_classes: int = 2
    bn_eps: float = 1e-3
    # frontend (NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:25-47)
    n_fft: int = 512
    win_length: int = 400
    hop: int = 160
    n_mels: int = 80
    window: str = "hann_sym"
    pre_emphasis: float = 0.97
    log_eps: float = 1e-7


def marblenet_spec(cfg: MarbleNetConfig) -> "OrderedDict[str, tuple]":
    """NeMo state_dict keys (un-folded): encoder.encoder.{b}.mconv.{i}.conv.weight for MaskedConv1d,
    .mconv.{i}.{weight,bias,running_mean,running_var} for BatchNorm1d, .res.0.{0,1} for the residual
    1x1 conv + BN, decoder.layer0.{weight,bias}."""
    s: OrderedDict[str, tuple] = OrderedDict()
    c_in = cfg.feat_in
    for b, blk in enumerate(cfg.blocks):
        pre = f"encoder.encoder.{b}."
        idx, cin_r = 0, c_in
        for r in range(blk.repeat):
            stride_r = blk.stride if r == 0 else 1  # NeMo applies the stride in every repeat; repeat==1 where stride>1
            del stride_r
            if blk.separable:
                s[pre + f"mconv.{idx}.conv.weight"] = (cin_r, 1, blk.kernel)
                s[pre + f"mconv.{idx + 1}.conv.weight"] = (blk.filters, cin_r, 1)
                bn = idx + 2
            else:
                s[pre + f"mconv.{idx}.conv.weight"] = (blk.filters, cin_r, blk.kernel)
                bn = idx + 1
            for p in ("weight", "bias", "running_mean", "running_var"):
                s[pre + f"mconv.{bn}.{p}"] = (blk.filters,)
            idx = bn + 1
            if r < blk.repeat - 1:
                idx += 2  # ReLU + Dropout between repeats
            cin_r = blk.filters
        if blk.residual:
            s[pre + "res.0.0.conv.weight"] = (blk.filters, c_in, 1)
            for p in ("weight", "bias", "running_mean", "running_var"):
                s[pre + f"res.0.1.{p}"] = (blk.filters,)
        c_in = blk.filters
    s["decoder.layer0.weight"] = (cfg.num_classes, c_in)
    s["decoder.layer0.bias"] = (cfg.num_classes,)
    return s


def marblenet_random_init(cfg: MarbleNetConfig = MarbleNetConfig(), seed: int = 0):
    """Seeded random weights with non-trivial BatchNorm statistics (so that folding matters)."""
    rs = np.random.RandomState(seed)
    w: OrderedDict[str, np.ndarray] = OrderedDict()
    for name, shape in marblenet_spec(cfg).items():
        if name.endswith("running_var"):
            w[name] = rs.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif name.endswith("running_mean"):
            w[name] = rs.uniform(-0.3, 0.3, size=shape).astype(np.float32)
        elif name.endswith(".bias") and len(shape) == 1 and "decoder" not in name:
            w[name] = rs.uniform(-0.2, 0.2, size=shape).astype(np.float32)
        elif len(shape) == 1 and "decoder" not in name:          # BN gamma
            w[name] = rs.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif len(shape) == 1:
            w[name] = rs.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        else:
            w[name] = _uniform(rs, shape, np.sqrt(3.0 / _fan_in(shape)))
    # calibrated offline (default layout): centres logit(active) - logit(silence) on the burst/gap mix
    w["decoder.layer0.bias"] = (w["decoder.layer0.bias"] + np.array([3.6, -3.6], np.float32)[:cfg.num_classes]).astype(np.float32)
    return w


def marblenet_fold(cfg: MarbleNetConfig, w: dict) -> "OrderedDict[str, np.ndarray]":
    """BatchNorm folding exactly as fold_bn_into_conv1d does it
    (NVIDIA_*/Export_NVIDIA_MarbleNet_VAD.py:58-101): scale = gamma * rsqrt(var + eps) goes into the
    conv that precedes the BN (the pointwise conv of a separable pair), bias = -mean*scale + beta.
    Returns per layer: dw (depthwise [C,k]) / pw ([out,in]) / pw_bias, res / res_bias, decoder."""
    import torch
    out: OrderedDict[str, np.ndarray] = OrderedDict()
    c_in = cfg.feat_in

    def fold(conv_w, pre):
        g, b = torch.from_numpy(w[pre + "weight"]), torch.from_numpy(w[pre + "bias"])
        mu, var = torch.from_numpy(w[pre + "running_mean"]), torch.from_numpy(w[pre + "running_var"])
        scale = g * torch.rsqrt(var + cfg.bn_eps)
        nw = torch.from_numpy(conv_w) * scale.reshape(-1, 1, 1)
        nb = (-mu) * scale + b
        return nw.numpy(), nb.numpy()

    for bi, blk in enumerate(cfg.blocks):
        pre = f"encoder.encoder.{bi}."
        idx = 0
        for r in range(blk.repeat):
            if not blk.separable:
                raise NotImplementedError("non-separable Jasper blocks")
            out[f"b{bi}.r{r}.dw"] = w[pre + f"mconv.{idx}.conv.weight"][:, 0, :].copy()
            pw, pb = fold(w[pre + f"mconv.{idx + 1}.conv.weight"], pre + f"mconv.{idx + 2}.")
            out[f"b{bi}.r{r}.pw"], out[f"b{bi}.r{r}.pw_bias"] = pw[:, :, 0].copy(), pb
            idx += 3 + (2 if r < blk.repeat - 1 else 0)
        if blk.residual:
            rw, rb = fold(w[pre + "res.0.0.conv.weight"], pre + "res.0.1.")
            out[f"b{bi}.res"], out[f"b{bi}.res_bias"] = rw[:, :, 0].copy(), rb
        c_in = blk.filters
    out["decoder.weight"] = w["decoder.layer0.weight"]
    out["decoder.bias"] = w["decoder.layer0.bias"]
    del c_in
    return out


# ----------------------------------------------------------------------------- Silero VAD v5 (16 kHz)
@dataclasses.dataclass(frozen=True)
class SileroConfig:
    """Silero VAD v5, 16 kHz branch.  The network ships as an opaque silero_vad.onnx inside the
    unpinned `silero_vad` pip package (Silero/Export_Silero_VAD.py:91), not in the reference
    checkout: this layout restates the public v5 graph -- parity for the network is unpinned; the
    call contract (Silero/modeling_modified/utils_vad.py:93-128) is what the reference pins."""
    window: int = 512
    context: int = 64
    reflect_pad: int = 64
    n_fft: int = 256
    hop: int = 128
    enc_channels: tuple = (128, 64, 64, 128)
    enc_strides: tuple = (1, 2, 2, 1)
    enc_kernel: int = 3
    hidden: int = 128

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def n_frames(self) -> int:
        return (self.context + self.window + self.reflect_pad - self.n_fft) // self.hop + 1


def silero_spec(cfg: SileroConfig) -> "OrderedDict[str, tuple]":
    s: OrderedDict[str, tuple] = OrderedDict()
    s["stft.forward_basis_buffer"] = (2 * cfg.n_bins, 1, cfg.n_fft)
    c = cfg.n_bins
    for i, co in enumerate(cfg.enc_channels):
        s[f"encoder.{i}.reparam_conv.weight"] = (co, c, cfg.enc_kernel)
        s[f"encoder.{i}.reparam_conv.bias"] = (co,)
        c = co
    s["decoder.rnn.weight_ih"] = (4 * cfg.hidden, c)
    s["decoder.rnn.weight_hh"] = (4 * cfg.hidden, cfg.hidden)
    s["decoder.rnn.bias_ih"] = (4 * cfg.hidden,)
    s["decoder.rnn.bias_hh"] = (4 * cfg.hidden,)
    s["decoder.decoder.2.weight"] = (1, cfg.hidden, 1)
    s["decoder.decoder.2.bias"] = (1,)
    return s


def silero_random_init(cfg: SileroConfig = SileroConfig(), seed: int = 0):
    """Seeded random network; the STFT basis is the real (hann-windowed) DFT basis the model carries."""
    rs = np.random.RandomState(seed)
    spec = silero_spec(cfg)
    w = _default_init(rs, OrderedDict((k, v) for k, v in spec.items() if k != "stft.forward_basis_buffer"))
    n = np.arange(cfg.n_fft, dtype=np.float64)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / cfg.n_fft)
    f = np.arange(cfg.n_bins, dtype=np.float64)[:, None]
    ang = 2 * np.pi * f * n[None, :] / cfg.n_fft
    basis = np.concatenate([np.cos(ang) * win, -np.sin(ang) * win], 0).astype(np.float32)
    w["stft.forward_basis_buffer"] = basis[:, None, :]
    # audio is in [-1, 1): lift the first layer so the stack is alive, and spread the head
    w["encoder.0.reparam_conv.weight"] = (w["encoder.0.reparam_conv.weight"] * 4.0).astype(np.float32)
    w["decoder.decoder.2.weight"] = (w["decoder.decoder.2.weight"] * 12.0).astype(np.float32)
    return w


def silero_dense_layers(cfg: SileroConfig, w: dict):
    """The k=3 conv encoder acts on only 4 -> 4 -> 2 -> 1 -> 1 frames per window, so each layer is
    re-expressed as ONE dense matrix on the flattened [frame][channel] vector (block-banded, zero
    where the taps fall outside or on the padding): rows of the batch stay independent and the
    layer runs on the generic dense kernel.  Returns [(W [out, in], b [out])] for the 4 layers."""
    out = []
    t_in, c_in = cfg.n_frames, cfg.n_bins
    for i, (co, st) in enumerate(zip(cfg.enc_channels, cfg.enc_strides)):
        k = cfg.enc_kernel
        pad = k // 2
        t_out = (t_in + 2 * pad - k) // st + 1
        Wc = np.asarray(w[f"encoder.{i}.reparam_conv.weight"], np.float32)      # [co, c_in, k]
        bc = np.asarray(w[f"encoder.{i}.reparam_conv.bias"], np.float32)
        D = np.zeros((t_out * co, t_in * c_in), np.float32)
        for to in range(t_out):
            for j in range(k):
                ti = to * st - pad + j
                if 0 <= ti < t_in:
                    D[to * co:(to + 1) * co, ti * c_in:(ti + 1) * c_in] = Wc[:, :, j]
        out.append((D, np.tile(bc, t_out)))
        t_in, c_in = t_out, co
    assert t_in == 1
    return out


# ----------------------------------------------------------------------------- DFSMN AEC-VAD (dual input)
@dataclasses.dataclass(frozen=True)
class DfsmnAecConfig:
    """DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py: SDAEC AlphaPredictor + ICCRN echo estimator
    (architecture vendored in the Export script, :65-284) followed by the modelscope
    `speech_dfsmn_aec_psm_16k` mask-net (linear1 -> relu -> deepfsmn -> linear3, :350-353).  The
    mask-net's sizes are not in the reference (modelscope is un-vendored): hidden 128, 9 UniDeepFsmn
    layers (128 -> 256 -> 128, lorder 20) is a declared assumption."""
    channels: int = 20
    n_fft_b: int = 319
    hop_b: int = 160
    alpha_k: int = 10
    n_fft_a: int = 1024
    win_a: int = 640
    hop_a: int = 320
    n_mels: int = 80
    pre_emphasis: float = 0.97
    echo_factor: float = 1.15
    log_floor: float = 1e-6
    mask_hidden: int = 128
    mask_layers: int = 9
    mask_inner: int = 256
    mask_lorder: int = 20
    max_frames: int = 200

    @property
    def n_bins_b(self) -> int:
        return self.n_fft_b // 2 + 1   # 160

    @property
    def ceps_bins(self) -> int:
        return self.n_bins_b // 2 + 1  # 81


def _lstm_spec(s, prefix, n_in, hidden, layers=1, bi=False):
    for l in range(layers):
        for suf in ([""] + (["_reverse"] if bi else [])):
            i = n_in if l == 0 else hidden * (2 if bi else 1)
            s[f"{prefix}.weight_ih_l{l}{suf}"] = (4 * hidden, i)
            s[f"{prefix}.weight_hh_l{l}{suf}"] = (4 * hidden, hidden)
            s[f"{prefix}.bias_ih_l{l}{suf}"] = (4 * hidden,)
            s[f"{prefix}.bias_hh_l{l}{suf}"] = (4 * hidden,)


def dfsmn_aec_spec(cfg: DfsmnAecConfig) -> "OrderedDict[str, tuple]":
    """Learnable tensors: `iccrn.*` = NET.state_dict() parameter names (Export_DFSMN_VAD.py:170-249),
    `alpha.*` = AlphaPredictor, `mask.*` = the mask-net shell, `shift` / `scale` = preprocessor.feature."""
    c, F, cb = cfg.channels, cfg.n_bins_b, cfg.ceps_bins
    s: OrderedDict[str, tuple] = OrderedDict()
    _lstm_spec(s, "iccrn.in_ch_lstm.lstm2", 4, c, bi=True)
    s["iccrn.in_ch_lstm.linear.weight"] = (c, 2 * c)
    s["iccrn.in_ch_lstm.linear.bias"] = (c,)
    s["iccrn.in_conv.weight"] = (c, 4 + c, 1, 1)
    s["iccrn.in_conv.bias"] = (c,)

    def cfb(name, cin):
        p = f"iccrn.{name}."
        for conv, shape in (("conv_gate", (c, cin, 1, 1)), ("conv_input", (c, cin, 1, 1)), ("conv", (c, c, 3, 1))):
            s[p + conv + ".weight"] = shape
            s[p + conv + ".bias"] = (c,)
        _lstm_spec(s, p + "ceps_unit.ch_lstm_f.lstm2", 2 * c, c, bi=True)
        s[p + "ceps_unit.ch_lstm_f.linear.weight"] = (2 * c, 2 * c)
        s[p + "ceps_unit.ch_lstm_f.linear.bias"] = (2 * c,)
        s[p + "ceps_unit.LN.w"] = (1, 2 * c, cb, 1)
        s[p + "ceps_unit.LN.b"] = (1, 2 * c, cb, 1)
        for ln, ch in (("LN0", cin), ("LN1", c), ("LN2", c)):
            s[p + ln + ".w"] = (1, ch, F, 1)
            s[p + ln + ".b"] = (1, ch, F, 1)

    for i in range(1, 6):
        cfb(f"cfb_e{i}", c)
    s["iccrn.ln.w"] = (1, c, F, 1)
    s["iccrn.ln.b"] = (1, c, F, 1)
    _lstm_spec(s, "iccrn.ch_lstm.lstm2", c, 2 * c, layers=2)
    s["iccrn.ch_lstm.linear.weight"] = (c, 2 * c)
    s["iccrn.ch_lstm.linear.bias"] = (c,)
    cfb("cfb_d5", c)
    for i in (4, 3, 2, 1):
        cfb(f"cfb_d{i}", 2 * c)
    _lstm_spec(s, "iccrn.out_ch_lstm.lstm2", 2 * c, c)
    s["iccrn.out_ch_lstm.linear.weight"] = (2 * c, c)
    s["iccrn.out_ch_lstm.linear.bias"] = (2 * c,)
    s["iccrn.out_conv.weight"] = (2, 3 * c, 1, 1)
    s["iccrn.out_conv.bias"] = (2,)
    s["alpha.linear1.weight"] = (1, 2)
    s["alpha.linear1.bias"] = (1,)
    s["alpha.linear2.weight"] = (1, cfg.alpha_k)
    s["alpha.linear2.bias"] = (1,)
    H = cfg.mask_hidden
    s["mask.linear1.weight"] = (H, 3 * cfg.n_mels)
    s["mask.linear1.bias"] = (H,)
    for i in range(cfg.mask_layers):
        s[f"mask.deepfsmn.{i}.linear.weight"] = (cfg.mask_inner, H)
        s[f"mask.deepfsmn.{i}.linear.bias"] = (cfg.mask_inner,)
        s[f"mask.deepfsmn.{i}.project.weight"] = (H, cfg.mask_inner)
        s[f"mask.deepfsmn.{i}.conv1.weight"] = (H, 1, cfg.mask_lorder, 1)
    s["mask.linear3.weight"] = (1, H)
    s["mask.linear3.bias"] = (1,)
    s["shift"] = (3 * cfg.n_mels,)
    s["scale"] = (3 * cfg.n_mels,)
    return s


def dfsmn_aec_random_init(cfg: DfsmnAecConfig = DfsmnAecConfig(), seed: int = 0):
    rs = np.random.RandomState(seed)
    spec = dfsmn_aec_spec(cfg)
    w: OrderedDict[str, np.ndarray] = OrderedDict()
    last_fan = 1
    for name, shape in spec.items():
        if name in ("shift", "scale"):
            continue
        if name.endswith(".w"):                      # LayerNorm gain (reference init: ones)
            w[name] = rs.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif name.endswith(".b") and len(shape) == 4:  # LayerNorm bias (reference init: rand * 1e-4)
            w[name] = (rs.uniform(0, 1, size=shape) * 1e-1).astype(np.float32)
        elif "bias" in name:
            w[name] = _uniform(rs, shape, 1.0 / np.sqrt(max(last_fan, 1)))
        else:
            last_fan = _fan_in(shape)
            w[name] = _uniform(rs, shape, np.sqrt(3.0 / last_fan))
    # log-mel of 1/32768-scaled audio sits around -20..-5; shift already includes + ln(32768^2) in the
    # reference wrapper (Export_DFSMN_VAD.py:291), so `shift` here is the raw preprocessor value
    w["shift"] = (-rs.uniform(19.0, 23.0, size=(3 * cfg.n_mels,))).astype(np.float32)
    w["scale"] = rs.uniform(0.2, 0.4, size=(3 * cfg.n_mels,)).astype(np.float32)
    w["alpha.linear1.weight"] = np.array([[0.6, 0.4]], np.float32)
    w["alpha.linear1.bias"] = np.array([0.05], np.float32)
    w["alpha.linear2.weight"] = (rs.uniform(0.02, 0.2, size=(1, cfg.alpha_k))).astype(np.float32)
    w["alpha.linear2.bias"] = np.array([0.3], np.float32)
    for i in range(cfg.mask_layers):      # keep the 9-deep residual stack from blowing up under random init
        w[f"mask.deepfsmn.{i}.project.weight"] = (w[f"mask.deepfsmn.{i}.project.weight"] * 0.35).astype(np.float32)
    w["mask.linear3.weight"] = (w["mask.linear3.weight"] * 2.0).astype(np.float32)
    w["mask.linear3.bias"] = (w["mask.linear3.bias"] - 1.1).astype(np.float32)
    return w


def dfsmn_near_noise(cfg: DfsmnAecConfig = DfsmnAecConfig(), seed: int = 77):
    """The two constant white-noise buffers of the near-end-only DFSMN graph
    (DFSMN/only_near_end_audio/Export_DFSMN_VAD.py:309-310): pow_far = randn(n_bins, max_frames, k).half() ** 2 and
    far_comp = randn(2, n_bins, max_frames).half().  The reference draws them unseeded at export time, so they are
    model constants like the weights; here they come from a numpy seed.  -> (float16 [F, max_len, k], float16 [2, F, max_len])"""
    rs = np.random.RandomState(seed)
    a = rs.standard_normal((cfg.n_bins_b, cfg.max_frames, cfg.alpha_k)).astype(np.float16)
    b = rs.standard_normal((2, cfg.n_bins_b, cfg.max_frames)).astype(np.float16)
    return (a * a).astype(np.float16), b
