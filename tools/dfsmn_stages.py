"""DFSMN AEC-VAD: per-stage device time and throughput vs batch size."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vadx
from vadx import lib, synth, weights as W

dev = torch.device("cuda:0")
cfg = W.DfsmnAecConfig()
sess = vadx.DfsmnAecSession(W.dfsmn_aec_random_init(cfg, 0), cfg, chunk_len=31841)
out = {}
for S in (32, 128):
    far = torch.from_numpy(synth.synth_chunks_fast(S, 31841, seed=14)).to(dev)
    near = torch.from_numpy(synth.synth_chunks_fast(S, 31841, seed=15)).to(dev)
    for _ in range(2):
        sess.run_batch(near, far)
    torch.cuda.synchronize()
    lib.profile_enable(True); lib.profile_collect()
    t0 = time.perf_counter()
    sess.run_batch(near, far)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    st = lib.profile_collect(); lib.profile_enable(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sess.run_batch(near, far); e1.record(); torch.cuda.synchronize()
    out[S] = {"wall_profiled_ms": round(wall, 2), "device_ms": round(e0.elapsed_time(e1), 2),
              "audio_h_per_s": S * 31841 / 16000 / (e0.elapsed_time(e1) / 1e3) / 3600,
              "stages": {k: (round(v[0], 2), v[1]) for k, v in st.items() if v[1]}}
    del far, near
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
