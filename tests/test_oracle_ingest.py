"""The closed form of audioop.ratecv (+ tomono) against the stdlib calls pydub makes, and the frozen
head of vad_sample.wav (tests/golden/vad_sample_pcm_head.npz)."""
import os
import warnings

import numpy as np
import pytest

from oracle import ingest as OI

warnings.filterwarnings("ignore", category=DeprecationWarning)


@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("rates", [(48000, 16000), (44100, 16000), (8000, 16000), (22050, 16000), (16000, 16000),
                                   (24000, 16000), (11025, 16000), (96000, 16000), (16000, 8000)])
def test_closed_form_equals_audioop(ch, rates):
    rs = np.random.RandomState(ch * 100 + rates[0] // 1000)
    for n in (0, 1, 2, 3, 1000, 4801, 30011):
        pcm = rs.randint(-32768, 32768, size=n * ch).astype(np.int16)
        if n >= 1000:
            pcm[:50] = 32767
            pcm[50:100] = -32768
        assert np.array_equal(OI.closed_form(pcm, ch, *rates), OI.pydub_chain(pcm, ch, *rates)), (n, ch, rates)


def test_vad_sample_head_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "vad_sample_pcm_head.npz"))
    for r in (16000, 8000, 22050):
        assert np.array_equal(OI.closed_form(g["pcm"], int(g["channels"]), int(g["rate"]), r), g[f"mono_{r}"])
