"""One bench step of one model family (the workloads of bench.py's `families`), for ncu captures:

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 4 -o gpurun_out/<name> \
        python tools/family_step.py <family> [eager]

Two warm-up steps, then one step; `eager` runs the family's eager twin (no CUDA-graph replay)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

name = sys.argv[1]
eager = len(sys.argv) > 2 and sys.argv[2] == "eager"
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
for n, w in bench.family_workloads(dev, only={name}):
    fn = (w["eager"] or w["step"]) if eager else w["step"]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(n, "ok", w["config"])
