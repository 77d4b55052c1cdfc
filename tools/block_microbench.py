"""Micro-benchmark of the default FireRed block pair at the bench size (8192 chunks x 98 frames): fc1 writing operand stages
(vadx_linear_tc_stream_stages_f32) and the fused tail (vadx_fc2_memory_stages_f32), each timed alone with CUDA events.
Environment switches (VADX_BS_*) only act in a `make AB=1` build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vadx
from vadx import lib
l = lib.load()
S, T, K1, H, C, n1, n2 = int(os.environ.get("S", 8192)), 98, 128, 256, 128, 20, 20
dev = torch.device("cuda")
g = torch.Generator().manual_seed(1)
x = torch.randn((S * T, K1), generator=g).to(dev)
img1 = torch.from_numpy(lib.pack_weight_tc((torch.randn((H, K1), generator=g) / K1 ** 0.5).numpy())).to(dev)
img2 = torch.from_numpy(lib.pack_weight_tc((torch.randn((C, H), generator=g) / H ** 0.5).numpy())).to(dev)
b1 = torch.zeros((H,), device=dev); b2 = torch.zeros((C,), device=dev)
wl = (torch.randn((C, n1), generator=g) * 0.2).to(dev); wr = (torch.randn((C, n2), generator=g) * 0.2).to(dev)
res = torch.randn((S * T, C), generator=g).to(dev)
stage_bytes = int(l.vadx_fc2_memory_stages_stream_bytes(H, T))
himg = torch.zeros((S * stage_bytes // 4,), device=dev)
out = torch.empty((S * T, C), device=dev)
def fc1():
    lib.check(l.vadx_linear_tc_stream_stages_f32(x.data_ptr(), img1.data_ptr(), b1.data_ptr(), himg.data_ptr(), S * T, T, K1, H, 1, lib.stream_ptr()))
def tail():
    lib.check(l.vadx_fc2_memory_stages_f32(himg.data_ptr(), H, img2.data_ptr(), b2.data_ptr(), 0, wl.data_ptr(), n1, wr.data_ptr(), n2,
                                           res.data_ptr(), out.data_ptr(), S, T, lib.stream_ptr()))
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fc1()
t1, t2 = timeit(fc1), timeit(tail)
gb1 = S * T * (K1 + H) * 4 / 1e9; gb2 = S * T * (H + 2 * C) * 4 / 1e9
env = {k: v for k, v in os.environ.items() if k.startswith("VADX_")}
print(f"fc1->stages {t1:.4f} ms ({gb1 / t1 * 1e3:.0f} GB/s)  tail {t2:.4f} ms ({gb2 / t2 * 1e3:.0f} GB/s)  {env}")
