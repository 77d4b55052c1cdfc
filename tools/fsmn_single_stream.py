"""Single-stream FSMN timing (BASELINE C1-style): chunk 512 / 16000, CUDA graph on/off."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vadx
from vadx import fsmn_vad, weights as W


def main():
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "vad_sample_16k.npz"))
    audio = g["audio"].astype(np.int16)
    cfg = W.FsmnConfig()
    w = W.fsmn_random_init(cfg, seed=0)
    out = {}
    for chunk in (512, 16000):
        sess = vadx.FsmnSession(w, cfg, chunk_len=chunk)
        lb = 0.0 if chunk == 512 else 0.3
        for graph in (False, True):
            fsmn_vad.run_vad(audio, sess, lb, rng=np.random.RandomState(1), graph=graph)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                fsmn_vad.run_vad(audio, sess, lb, rng=np.random.RandomState(1), graph=graph)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            out[f"chunk{chunk}_graph{int(graph)}"] = {
                "sec": dt, "rtf": dt / (len(audio) / 16000.0)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
