"""CPU-side checks: the C-ABI library loads and exports every symbol include/vadx.h declares, fails
loudly without a device, and the host-side logic (tables, loader, formatting) matches the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import vadx
from vadx import audio_io, constants, lib, postprocess as PP, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vadx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vadx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = lib.load()
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(l, n), f"libvadx.so does not export {n}"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in vadx/lib.py"
    assert l.vadx_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly():
    l = lib.load()
    assert l.vadx_device_count() == 0
    h = C.c_void_p()
    hp = (C.c_int32 * 14)(80, 8, 1, 256, 128, 20, 1, 20, 1, 1, 400, 400, 160, 80)
    rc = l.vadx_create(b"firered", hp, 14, C.byref(h))
    assert rc == -2 and b"no CUDA device" in l.vadx_last_error()
    with pytest.raises(RuntimeError):
        vadx.FireRedSession(vadx.weights.firered_random_init())


def test_argument_validation_without_device():
    l = lib.load()
    assert l.vadx_linear_f32(None, 8, None, 8, None, None, 0, None, 8, 4, 8, 8, 0, None) == -1
    assert b"null pointer" in l.vadx_last_error()
    h = C.c_void_p()
    assert l.vadx_create(b"nonsense", None, 0, C.byref(h)) == -1


def test_interleaved_basis_matches_reference_layout():
    b, first, nb = tables.interleaved_basis(400, 400, "povey", "v2")
    ref, _ = constants.dft_basis(400, 400, "povey", "v2")
    assert first == 0 and nb == 201 and b.shape == (400, 404)
    assert np.array_equal(b[:, 0:402:2], ref[:201].numpy().T)
    assert np.array_equal(b[:, 1:402:2], ref[201:].numpy().T)
    assert not b[:, 402:].any()
    b2, first2, _ = tables.interleaved_basis(512, 400, "hamming", "v1")
    assert first2 == 56 and b2.shape == (400, 516)


def test_sparse_bank_is_lossless():
    for bank in (constants.kaldi_like_mel_bank(400, 80, 16000).numpy(),
                 constants.torchaudio_mel_bank(257, 20, 8000, 80, 16000, None, "htk").numpy(),
                 constants.torchaudio_mel_bank(257, 0, 8000, 80, 16000, "slaney", "slaney").numpy()):
        st, ln, w = tables.sparse_bank(bank)
        dense = np.zeros_like(bank)
        for m in range(bank.shape[0]):
            dense[m, st[m]:st[m] + ln[m]] = w[m, :ln[m]]
        assert np.array_equal(dense, bank)


def test_torchaudio_banks_bit_equal():
    ta = pytest.importorskip("torchaudio")
    for args in ((257, 20.0, 8000.0, 80, 16000, None, "htk"), (257, 0.0, 8000.0, 80, 16000, "slaney", "slaney"),
                 (513, 20.0, 8000.0, 80, 16000, None, "htk")):
        ours = constants.torchaudio_mel_bank(*args)
        ref = ta.functional.melscale_fbanks(*args).transpose(0, 1)
        assert torch.equal(ours, ref)


def test_format_time_and_lines(golden_dir):
    g = np.load(os.path.join(golden_dir, "postproc.npz"))
    assert [PP.format_time(float(t)) for t in g["clock_in"]] == [str(s) for s in g["clock_out"]]
    sec, idx = PP.timestamp_lines([(2.28, 3.5)])
    assert sec == ["00:00:02.279 --> 00:00:03.500\n"] and idx == ["36480 --> 56000\n"]


def test_segments_to_seconds_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "postproc.npz"))
    for i in range(7):
        dec = g[f"fr{i}_dec"]
        padded = np.concatenate(([0], dec, [0])).astype(np.int8)
        d = np.diff(padded)
        pairs = np.stack([np.flatnonzero(d == 1), np.flatnonzero(d == -1)], 1)
        cfg = PP.FramePostConfig()
        got = PP.segments_to_seconds(pairs, len(dec), cfg, float(g[f"fr{i}_dur"]))
        assert np.array_equal(np.array(got, np.float64).reshape(-1, 2), g[f"fr{i}_seg"])
        cfgn = PP.FramePostConfig(frame_shift_s=0.02, tail_adds_frame_length=False)
        gotn = PP.segments_to_seconds(pairs, len(dec), cfgn, len(dec) * 0.02 + 0.012)
        assert np.array_equal(np.array(gotn, np.float64).reshape(-1, 2), g[f"nv{i}_seg"])


def test_chunker_pads_with_rms_noise():
    a = (np.arange(40000) % 1000 - 500).astype(np.int16)
    rs = np.random.RandomState(7)
    chunks, n = audio_io.align_non_overlapping(a, 16000, rs)
    assert n == 40000 and chunks.shape == (3, 16000)
    assert np.array_equal(chunks.reshape(-1)[:40000], a)
    pad = 8000
    tail = a[-pad:].astype(np.float32)
    rs2 = np.random.RandomState(7)
    expect = (np.sqrt(np.mean(tail * tail)) * rs2.normal(0.0, 1.0, size=(pad,))).astype(np.int16)
    assert np.array_equal(chunks.reshape(-1)[40000:], expect)
    short, n2 = audio_io.align_non_overlapping(a[:1000], 16000, np.random.RandomState(1))
    assert short.shape == (1, 16000) and n2 == 1000


def test_vad_sample_fixture(golden_dir):
    a = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"]
    assert a.shape == (89431,) and a.dtype == np.int16 and int(np.abs(a).max()) == 4220
