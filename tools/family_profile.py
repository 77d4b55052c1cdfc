"""Per-kernel records of one or more bench families (bench.family_workloads): step time and the library's kernel table.
    python tools/family_profile.py dfsmn_aec [silero ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vadx import lib

dev = torch.device("cuda:0")
pk = bench.peaks()
for name, w in bench.family_workloads(dev, only=set(sys.argv[1:])):
    for _ in range(w["warm"]):
        w["step"]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(w["steps"]):
        w["step"]()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / w["steps"]
    print(name, "ms", round(ms, 3), "audio-h/s", round(w["audio_s"] / (ms / 1e3) / 3600, 2), "|", w["config"][:90])
    lib.profile_enable(True)
    lib.profile_collect_kernels()
    (w["eager"] or w["step"])()
    torch.cuda.synchronize()
    r = bench.kernel_roofline(name, lib.profile_collect_kernels(), pk)
    lib.profile_enable(False)
    for k in r["kernels_by_share"]:
        print("   ", k)
