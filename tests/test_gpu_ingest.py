"""On-device ingest (stereo -> mono, rate conversion) through the C ABI against the loader chain
(oracle/ingest.py, pinned to audioop) -- bit-exact."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import audio_io
from oracle import ingest as OI

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("rates", [(48000, 16000), (44100, 16000), (8000, 16000), (22050, 16000), (16000, 16000),
                                   (11025, 16000), (16000, 8000)])
def test_ingest_matches_loader_chain(cuda, ch, rates):
    rs = np.random.RandomState(ch * 7 + rates[0] // 100)
    for n in (1, 2, 3, 1000, 48017):
        pcm = rs.randint(-32768, 32768, size=n * ch).astype(np.int16)
        out, cnt = audio_io.ingest_pcm16_device(torch.from_numpy(pcm).to(cuda), ch, *rates)
        ref = OI.closed_form(pcm, ch, *rates)
        assert int(cnt[0]) == len(ref) == out.shape[1]
        assert np.array_equal(out[0].cpu().numpy(), ref), (n, ch, rates)


def test_ingest_ragged_streams(cuda):
    rs = np.random.RandomState(5)
    S, n = 37, 9000
    pcm = rs.randint(-32768, 32768, size=(S, n * 2)).astype(np.int16)
    lens = rs.randint(0, n + 1, size=S).astype(np.int64)
    lens[:3] = (0, 1, n)
    out, cnt = audio_io.ingest_pcm16_device(torch.from_numpy(pcm).to(cuda), 2, 44100, 16000,
                                            n_frames=torch.from_numpy(lens).to(cuda))
    out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
    for s in range(S):
        ref = OI.closed_form(pcm[s, :lens[s] * 2], 2, 44100, 16000)
        assert cnt[s] == len(ref)
        assert np.array_equal(out[s, :cnt[s]], ref), s
        assert not out[s, cnt[s]:].any()


def test_vad_sample_head(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "vad_sample_pcm_head.npz"))
    for r in (16000, 8000, 22050):
        out, _ = audio_io.ingest_pcm16_device(torch.from_numpy(g["pcm"].copy()).to(cuda), int(g["channels"]), int(g["rate"]), r)
        assert np.array_equal(out[0].cpu().numpy(), g[f"mono_{r}"])


def test_ingest_rejects_bad_arguments(cuda):
    x = torch.zeros(10, dtype=torch.int16, device=cuda)
    with pytest.raises(ValueError):
        audio_io.ingest_pcm16_device(x, 3, 48000, 16000)
    with pytest.raises(ValueError):
        audio_io.ingest_pcm16_device(x[:9], 2, 48000, 16000)
    with pytest.raises(ValueError):
        audio_io.ingest_pcm16_device(x.float(), 1, 48000, 16000)


def test_device_wav_loader_equals_host_loader(cuda, tmp_path):
    """load_wav_int16_device (raw PCM to the GPU, down-mix + rate conversion there) == load_wav_int16 (wave + audioop)."""
    import wave
    rs = np.random.RandomState(3)
    pcm = (rs.randint(-20000, 20000, size=44100 * 2 * 2)).astype(np.int16)     # 2 s, stereo, 44.1 kHz
    path = str(tmp_path / "x.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    host = audio_io.load_wav_int16(path, 16000)
    dev = audio_io.load_wav_int16_device(path, 16000).cpu().numpy()
    assert np.array_equal(host, dev)
