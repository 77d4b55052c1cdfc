"""Micro-benchmark of vadx_linear_tc_f32 on the FireRed layer shapes (perf experiments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vadx
from vadx import lib
l = lib.load()
rows = 8192 * 98
dev = torch.device("cuda")
for (K, N) in [(128, 256), (256, 128), (80, 256)]:
    x = torch.randn((rows, K), device=dev)
    w = torch.from_numpy(lib.pack_weight_tc(np.random.randn(N, K).astype(np.float32))).to(dev)
    b = torch.zeros((N,), device=dev)
    y = torch.empty((rows, N), device=dev)
    def run():
        lib.check(l.vadx_linear_tc_f32(x.data_ptr(), K, w.data_ptr(), b.data_ptr(), None, 0, y.data_ptr(), N, rows, K, N, 1, lib.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = rows * (K + N) * 4 / 1e9
    print(f"K={K} N={N}: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s algorithmic, debug={os.environ.get('VADX_TC_DEBUG', '0')}")
