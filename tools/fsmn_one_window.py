"""A few 512-sample FSMN windows of one stream (for an ncu launch list of the latency-bound case)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vadx
from vadx import fsmn_vad, weights as W

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "vad_sample_16k.npz"))
audio = g["audio"].astype(np.int16)[:512 + 352 * 5]
cfg = W.FsmnConfig()
sess = vadx.FsmnSession(W.fsmn_random_init(cfg, seed=0), cfg, chunk_len=512)
fsmn_vad.run_vad(audio, sess, 0.0, rng=np.random.RandomState(1))
torch.cuda.synchronize()
