"""Per-stage device time (the library's own CUDA-event timers) of the FSMN / Silero / FireRed-stream family runs
of bench.py -- where the time of each family goes."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vadx
from vadx import firered_vad, fsmn_vad, lib, silero_vad, synth, weights as W

dev = torch.device("cuda:0")
out = {}


def staged(name, fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.profile_enable(True)
    lib.profile_collect()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    st = lib.profile_collect()
    lib.profile_enable(False)
    out[name] = {"wall_ms": round(wall, 3), "stages_ms": {k: (round(v[0] / reps, 3), v[1] // reps) for k, v in st.items() if v[1]}}


cfg = W.FsmnConfig()
sess = vadx.FsmnSession(W.fsmn_random_init(cfg, 0), cfg, chunk_len=16000)
S, stride = 1024, 16000 - 31 * 160
a = torch.from_numpy(synth.synth_chunks_fast(S, 16000 + 3 * stride, seed=11)).to(dev)
staged("fsmn", lambda: fsmn_vad.run_streams(sess, a, stride))
del sess, a

cfg = W.SileroConfig()
sess = vadx.SileroSession(W.silero_random_init(cfg, 0), cfg)
audio = torch.from_numpy(synth.synth_chunks_fast(4096, 32 * 512, seed=13)).to(dev).float() * 0.000030517578
staged("silero", lambda: sess.speech_probs(audio))
del sess, audio

cfg = W.FireRedConfig(N2=0, S2=0, streaming=True)
sess = vadx.FireRedStreamSession(W.firered_random_init(cfg, 5), cfg)
a = torch.from_numpy(synth.synth_chunks_fast(4096, 25 * 2560, seed=16)).to(dev)
staged("firered_stream", lambda: firered_vad.run_stream_vad_streams(sess, a, [25 * 2560] * 4096), reps=2)
print(json.dumps(out, indent=1))
